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


def assert_close(got, ref, dtype, what="", rtol=None):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    rtol = RTOL[dtype] if rtol is None else rtol
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(got)), "%s: non-finite pattern differs" % what
    assert np.array_equal(got[~fin], ref[~fin]) or np.all(np.isnan(ref[~fin]) == np.isnan(got[~fin])), what
    if not fin.any():
        return
    scale = max(np.abs(ref[fin]).max(), 1e-3)   # absolute floor: outputs that are pure rounding noise
    err = np.abs(got[fin] - ref[fin])
    tol = rtol * np.abs(ref[fin]) + rtol * scale
    worst = (err / np.maximum(tol, 1e-300)).max()
    assert worst <= 1.0, "%s: max err %.3e (scale %.3e, rtol %g, worst ratio %.2f)" % (what, err.max(), scale, rtol, worst)
