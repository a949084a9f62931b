"""GPU: the CUDA path, called through the C-ABI (libkb200.so) behind the reference's binder
interface, against golden outputs of the real reference (tests/golden/make_golden.py).

Tolerances (BASELINE.json north_star): fp32 within 1e-5 relative, fp64 within 1e-12 relative,
indices bit-exact except genuine near-ties (checked explicitly against dense fp64 values)."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def run_binder(meta, args, device, flags=None):
    from keops_b200.binder import LoadKeOpsB200
    f = util.default_flags(meta["formula"])
    if flags:
        f.update(flags)
    b = LoadKeOpsB200(1, 0, 1, False, -1, meta["formula"], meta["aliases"], meta["nargs"], meta["dtype"], "torch", f)
    targs = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in args]
    out = b.import_module().genred_pytorch(device, None, meta["nx"], meta["ny"], 0, None, *targs)
    torch.cuda.synchronize()
    return b, out.cpu().numpy()


def check_indices(meta, args, got_idx, ref_idx, slot_to_comp, what, ulps=1.0, max_frac=1e-3):
    """Indices must match bit-exactly.  The one admissible difference: two columns whose exact (fp64) values agree to
    within ``ulps`` fp32 ulps (1 by default; ~sqrt(D) for a D-term sum, whose rounding error grows like that), where the golden CPU code (no FMA contraction, x86-64 baseline ISA) and a GPU kernel (FMA, like
    the reference's own nvcc build) may legitimately order them differently -- float32 fixtures only, never for float64,
    and the number of such entries is recorded next to the achieved parity errors (``index_ties:<case>``)."""
    nbad = int((got_idx != ref_idx).sum())
    util.ACHIEVED["index_ties:" + what] = float(nbad)
    if nbad == 0:
        return
    assert meta["dtype"] == "float32", "%s: %d index mismatches in float64" % (what, nbad)
    from oracle import keops_oracle as ko
    F = ko.dense(meta["formula"], meta["aliases"], args, "float64")
    bad = np.argwhere(got_idx != ref_idx)
    assert len(bad) <= max(1, int(got_idx.size * max_frac)), "%s: %d index mismatches" % (what, len(bad))
    eps = ulps * 2.0 ** -23
    for r, s in bad:
        c = slot_to_comp(s)
        j1, j2 = int(got_idx[r, s]), int(ref_idx[r, s])
        v1, v2 = F[r, j1, c], F[r, j2, c]
        assert abs(v1 - v2) <= eps * max(abs(v1), abs(v2), 1e-30), \
            "%s: row %d slot %d: got j=%d (%.9g) ref j=%d (%.9g) is not a tie" % (what, r, s, j1, v1, j2, v2)


@pytest.mark.parametrize("name", util.golden_names())
def test_cuda_matches_reference(name, cuda_device):
    meta, args, ref = util.load_case(name)
    b, got = run_binder(meta, args, cuda_device)
    assert b.dimout == meta["dimout"] and b.tagIJ == meta["tagIJ"]
    assert got.shape == ref.shape
    head = meta["formula"].split("(")[0]
    d = b.red.dimF
    if head in ("ArgMin_Reduction", "ArgMax_Reduction"):
        check_indices(meta, args, got, ref, lambda s: s, name)
    elif head == "ArgKMin_Reduction":
        check_indices(meta, args, got, ref, lambda s: s % d, name)
    elif head in ("Min_ArgMin_Reduction", "Max_ArgMax_Reduction"):
        util.assert_close(got[:, :d], ref[:, :d], meta["dtype"], name)      # values: pure element-wise relative error
        check_indices(meta, args, got[:, d:], ref[:, d:], lambda s: s, name)
    elif head == "KMin_ArgKMin_Reduction":
        g3, r3 = got.reshape(len(got), -1, 2, d), ref.reshape(len(ref), -1, 2, d)
        util.assert_close(g3[:, :, 0], r3[:, :, 0], meta["dtype"], name)
        check_indices(meta, args, g3[:, :, 1].reshape(len(got), -1), r3[:, :, 1].reshape(len(ref), -1),
                      lambda s: s % d, name)
    else:
        util.assert_close(got, ref, meta["dtype"], name, rtol=util.rtol_for(meta), **util.golden_scale(meta, args))


@pytest.mark.parametrize("name", [n for n in util.golden_names() if n.startswith(("gauss_sum_float", "lse_softmin", "argkmin_K10_D3"))])
@pytest.mark.parametrize("flags", [dict(use_fast_math=0), dict(sum_scheme="direct_sum"), dict(sum_scheme="kahan_scheme")])
def test_option_variants(name, flags, cuda_device):
    meta, args, ref = util.load_case(name)
    head = meta["formula"].split("(")[0]
    if "sum_scheme" in flags and head not in ("Sum_Reduction", "Max_SumShiftExp_Reduction"):
        pytest.skip("sum schemes only apply to sum-type reductions")
    b, got = run_binder(meta, args, cuda_device, flags)
    tag = name + "/" + ",".join("%s=%s" % kv for kv in flags.items())
    if head == "ArgKMin_Reduction":
        check_indices(meta, args, got, ref, lambda s: s % b.red.dimF, tag)
    else:
        util.assert_close(got, ref, meta["dtype"], tag, **util.golden_scale(meta, args))


@pytest.mark.gpu
def test_high_dimensional_gaussian_vs_oracle(cuda_device):
    """D = 300 (beyond the tensor-core path's K <= 128 and the reference's chunk-mode threshold of 100): the generic
    CUDA-core kernel must still be correct."""
    import torch
    from keops_b200.binder import get_binder
    from oracle import keops_oracle as ko
    rng = np.random.default_rng(11)
    D, M, N = 300, 70, 90
    x = (rng.random((M, D)) / np.sqrt(D)).astype(np.float32)
    y = (rng.random((N, D)) / np.sqrt(D)).astype(np.float32)
    b = rng.standard_normal((N, 2)).astype(np.float32)
    g = np.array([3.0], np.float32)
    f = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,2,1),0)" % (D, D)
    targs = [torch.from_numpy(a).to(cuda_device) for a in (x, y, b, g)]
    out = get_binder(f, [], 4, "float32").genred_pytorch(cuda_device, None, M, N, 0, None, *targs)
    torch.cuda.synchronize()
    util.assert_close(out.cpu().numpy(), ko.genred(f, [], [x, y, b, g], "float32"), "float32", "gaussian D=300",
                      scale=ko.abs_scale(f, [], [x, y, b, g]))


@pytest.mark.gpu
@pytest.mark.parametrize("tagI", [0, 1])
def test_wide_streamed_factor_is_chunked(tagI, cuda_device):
    """dim b = 300 (the reference's final-chunk mode, GpuReduc1D_finalchunks.py): evaluated in chunks of 32 components,
    must agree with the unchunked kernel and with the oracle; also with a user-provided ``out``."""
    import os
    import torch
    from keops_b200.binder import get_binder
    from oracle import keops_oracle as ko
    rng = np.random.default_rng(5)
    M, N, DV = 333, 257, 300
    x = rng.random((M, 3)).astype(np.float32)
    y = rng.random((N, 3)).astype(np.float32)
    b = rng.standard_normal((N if tagI == 0 else M, DV)).astype(np.float32)
    f = "Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,%d,%d),%d)" % (DV, 1 - tagI, tagI)
    targs = [torch.from_numpy(a).to(cuda_device) for a in (x, y, b)]
    bd = get_binder(f, [], 3, "float32")
    out = bd.genred_pytorch(cuda_device, None, M, N, 0, None, *targs)
    assert bd.last_launch.get("wide_chunks") == 10, bd.last_launch
    ref = ko.genred(f, [], [x, y, b], "float32")
    A = ko.abs_scale(f, [], [x, y, b])
    util.assert_close(out.cpu().numpy(), ref, "float32", "wide b, chunked", scale=A)
    os.environ["KB200_WIDE_CHUNKS"] = "0"
    try:
        plain = bd.genred_pytorch(cuda_device, None, M, N, 0, None, *targs)
        assert bd.last_launch.get("wide_chunks") is None
    finally:
        os.environ.pop("KB200_WIDE_CHUNKS", None)
    util.assert_close(out.cpu().numpy(), plain.cpu().numpy(), "float32", "wide b, chunked vs unchunked", scale=A)
    dst = torch.empty_like(out)
    res = bd.genred_pytorch(cuda_device, None, M, N, 0, dst, *targs)
    assert res.data_ptr() == dst.data_ptr() and torch.equal(dst, out)


@pytest.mark.gpu
def test_high_dimensional_argkmin_vs_oracle(cuda_device):
    """ArgKMin in D = 200 (beyond the tensor-core filter's D <= 141): the generic kernel, indices against the oracle."""
    import torch
    from keops_b200.binder import get_binder
    from oracle import keops_oracle as ko
    rng = np.random.default_rng(13)
    D, M, N, K = 200, 150, 400, 5
    x = rng.standard_normal((M, D)).astype(np.float32)
    y = rng.standard_normal((N, D)).astype(np.float32)
    f = "ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,0)" % (D, D, K)
    out = get_binder(f, [], 2, "float32").genred_pytorch(cuda_device, None, M, N, 0, None,
                                                         torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device))
    torch.cuda.synchronize()
    ref = ko.genred(f, [], [x, y], "float32")
    # fp32 near-ties may swap between FMA and non-FMA evaluation orders of a 200-term sum: every differing entry must be
    # such a tie (values equal to ~sqrt(D) ulps in exact arithmetic); the count is recorded with the parity errors
    meta = dict(formula=f, aliases=[], dtype="float32")
    check_indices(meta, [x, y], out.cpu().numpy(), ref, lambda s: 0, "argkmin D=200 vs oracle", ulps=2.0 * np.sqrt(D),
                  max_frac=0.01)
