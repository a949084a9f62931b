"""Shared helpers for the parity tests (golden fixtures, tolerances)."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "cases")

# north_star tolerances: fp32 within 1e-5 relative, fp64 within 1e-12 relative
RTOL = {"float32": 1e-5, "float64": 1e-12}


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    args = [z["arg%d" % k] for k in range(meta["nargs"])]
    return meta, args, z["out"]


def default_flags(formula):
    head = formula.split("(")[0]
    sumlike = head in ("Sum_Reduction", "Max_SumShiftExp_Reduction", "Max_SumShiftExpWeight_Reduction",
                       "Grad_WithSavedForward")
    return dict(dtype_acc="auto", sum_scheme="block_sum" if sumlike else "direct_sum", use_fast_math=1,
                enable_chunks=1, multVar_highdim=0)


def is_index_reduction(formula):
    head = formula.split("(")[0]
    return head in ("ArgMin_Reduction", "ArgMax_Reduction", "ArgKMin_Reduction")


def mixed_index_columns(formula, dimout):
    """Boolean mask of output columns that hold indices (must be bit-exact)."""
    head = formula.split("(")[0]
    m = np.zeros(dimout, bool)
    if head in ("ArgMin_Reduction", "ArgMax_Reduction", "ArgKMin_Reduction"):
        m[:] = True
    elif head in ("Min_ArgMin_Reduction", "Max_ArgMax_Reduction"):
        m[dimout // 2:] = True
    elif head == "KMin_ArgKMin_Reduction":
        # layout [K][2][dimF]
        return None
    return m


def rtol_for(meta):
    """north-star tolerance, except for fp64 ``Powf``: the reference evaluates it with the C float
    function ``powf`` whatever the dtype (keopscore/utils/math_functions.py:97-99), so only float32
    accuracy is defined for it."""
    if meta["dtype"] == "float64" and "Powf" in meta["formula"]:
        return 2e-6
    return RTOL[meta["dtype"]]


_scale_cache = {}


def golden_scale(meta, args):
    """``dict(scale=A, floor=f)`` for a golden case: the oracle's condition scale (float64 evaluation of ``sum_j |F_ij|``,
    ``oracle.keops_oracle.abs_scale``) and the noise floor, ready to be splatted into :func:`assert_close`."""
    key = meta["name"]
    hit = _scale_cache.get(key)
    if hit is None:
        from oracle import keops_oracle as ko
        A = ko.abs_scale(meta["formula"], meta["aliases"], args, "float64", sizes=(meta["nx"], meta["ny"]))
        ncol = meta["ny"] if meta["tagIJ"] == 0 else meta["nx"]
        hit = dict(scale=A, floor=noise_floor(args, meta["dtype"], ncol))
        _scale_cache[key] = hit
    return hit


def gpu_abs_scale(formula, nargs, dtype, device, nx, ny, *targs):
    """Condition scale ``sum_j |F_ij|`` of a Sum-type reduction string (``Grad_WithSavedForward`` included) evaluated on
    the GPU by the engine itself, for problems too large for the numpy oracle: the scale of a parity gate only has to be
    right to a few percent.  ``Sum_Reduction(F, t)`` -> ``Sum_Reduction(Abs(F), t)``."""
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.binder import get_binder
    red = parse_reduction(formula, [])
    assert red.kind == "Sum", "gpu_abs_scale: Sum-type reductions only"
    f_abs = "Sum_Reduction(Abs(%r),%d)" % (red.formula, red.tagI)
    out = get_binder(f_abs, [], nargs, dtype).genred_pytorch(device, None, nx, ny, 0, None, *targs)
    return out


ACHIEVED = {}     # what -> achieved max |got - ref| / scale (dumped at session end by conftest.py)


def noise_floor(args, dtype, ncol):
    """Cancellation floor for formulas whose *terms* cancel internally (e.g. gradients that are analytically zero): the
    rounding noise of ``ncol`` accumulated terms that each carry about one ulp of the largest input magnitude ``S``:
    ``4 * eps(dtype) * sqrt(ncol) * max(1, S)``.  Two orders of magnitude below ``rtol * scale`` for O(1) outputs."""
    eps = 2.0 ** -23 if dtype == "float32" else 2.0 ** -52
    S = max([1.0] + [float(np.abs(np.asarray(a, dtype=np.float64)).max()) for a in args if np.asarray(a).size])
    return 4.0 * eps * np.sqrt(max(1, ncol)) * S


def assert_close(got, ref, dtype, what="", rtol=None, scale=None, floor=0.0):
    """Element-wise parity gate: ``|got - ref| <= rtol * scale + floor`` for EVERY entry.

    ``scale`` is the entry's condition scale (``oracle.keops_oracle.abs_scale``): ``sum_j |F_ij|`` for a sum-type
    reduction (= ``|ref|`` itself when the terms have one sign; the stated cancellation floor when they do not: a
    reduction whose every term is ``rtol``-accurate cannot do better), ``|ref|`` for values.  ``floor``: absolute noise
    floor (:func:`noise_floor`), for terms that cancel *inside* the formula.  ``scale=None`` means pure
    element-wise relative error.  The achieved maximum of ``|got - ref| / scale`` is recorded per ``what``."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    rtol = RTOL[dtype] if rtol is None else rtol
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(got)), "%s: non-finite pattern differs" % what
    assert np.array_equal(got[~fin], ref[~fin]) or np.all(np.isnan(ref[~fin]) == np.isnan(got[~fin])), what
    if not fin.any():
        return
    sc = np.abs(ref) if scale is None else np.maximum(np.abs(np.asarray(scale, dtype=np.float64)), np.abs(ref))
    assert sc.shape == ref.shape, "%s: scale shape %s vs %s" % (what, sc.shape, ref.shape)
    tiny = np.finfo(np.float32 if dtype == "float32" else np.float64).tiny
    err = np.abs(got[fin] - ref[fin])
    rel = err / np.maximum(sc[fin] + floor / rtol, tiny / rtol)     # entries below the denormal range are compared absolutely
    worst = float(rel.max())
    ACHIEVED[what] = max(ACHIEVED.get(what, 0.0), worst)
    k = int(rel.argmax())
    assert worst <= rtol, "%s: element-wise error %.3e of the condition scale exceeds rtol %g (|err| %.3e, ref %.6e, scale %.3e)" \
        % (what, worst, rtol, err[k], ref[fin][k], sc[fin][k])
