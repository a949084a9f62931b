"""Tensor-core K-min path (tcgen05 filter + exact re-evaluation, ``keops_b200/csrc/kb200_knn_tc.cu``):
its output must be *bit-identical* to the CUDA-core kernel's, whose parity with the reference is pinned
by the golden fixtures (``tests/test_golden_gpu.py``: argkmin_K10_D64, kmin_argkmin, ...)."""
import os

import numpy as np
import pytest

from oracle import keops_oracle as ko


def _run(formula, nargs, args, M, N, dev, tc):
    import torch
    from keops_b200.binder import get_binder
    old = os.environ.get("KB200_KNN_TC")
    os.environ["KB200_KNN_TC"] = "1" if tc else "0"
    try:
        bd = get_binder(formula, [], nargs, "float32")
        out = bd.genred_pytorch(dev, None, M, N, 0, None, *args)
        torch.cuda.synchronize()
        return out, dict(bd.last_launch)
    finally:
        if old is None:
            os.environ.pop("KB200_KNN_TC", None)
        else:
            os.environ["KB200_KNN_TC"] = old


CASES = [
    # kind, D, K, M, N, tagI   (remainder block 16 / 64 / 32 / 64 columns; two main blocks for D = 128)
    ("ArgKMin", 64, 10, 9000, 7001, 0),
    ("KMin_ArgKMin", 100, 5, 8192, 6000, 0),
    ("KMin", 16, 3, 8500, 5000, 0),
    ("ArgKMin", 32, 1, 8200, 4100, 0),
    ("ArgKMin", 128, 10, 8192, 4500, 0),
    ("ArgKMin", 61, 4, 5000, 9000, 1),      # reduction over i: the output (and the tensor-core rows) follow j
    ("ArgKMin", 64, 10, 1500, 70000, 0),    # few rows, many columns: the filter splits the column range (gridDim.y)
    ("KMin_ArgKMin", 32, 5, 2500, 100001, 0),
    ("KMin", 100, 3, 1100, 66000, 0),
    # D + 3 > 128: three ring stages; D + 3 > 160: both operands streamed through the ring in 64-column k-blocks
    ("ArgKMin", 157, 10, 8192, 4100, 0),
    ("ArgKMin", 200, 10, 8192, 5000, 0),
    ("KMin_ArgKMin", 784, 5, 8192, 4500, 0),  # MNIST-sized rows
    ("ArgKMin", 300, 3, 1200, 66000, 0),      # streamed k-blocks + column splits
    # 176 row blocks = one full wave of 148 + a tail of 28: the tail's column range is split so that it fills the machine too
    ("ArgKMin", 64, 10, 45000, 20000, 0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,D,K,M,N,tagI", CASES)
def test_tensor_core_kmin_identical_to_cuda_core(kind, D, K, M, N, tagI, cuda_device):
    import torch
    torch.manual_seed(D * 1000 + K)
    x, y = torch.randn(M, D, device=cuda_device), torch.randn(N, D, device=cuda_device)
    f = "%s_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,%d)" % (kind, D, D, K, tagI)
    o_tc, ll = _run(f, 2, (x, y), M, N, cuda_device, True)
    assert ll.get("engine") == "tcgen05", ll          # the tensor-core kernels really ran
    if (M if tagI == 0 else N) < 8192:
        assert ll["nsplit"] > 1, ll                   # few rows: column splits + per-split lists + merge
    if (M if tagI == 0 else N) == 45000:
        assert [sg[2] for sg in ll["segments"]] == [1, 2] and ll["segments"][0][1] == 148 * 256, ll
    o_cc, ll2 = _run(f, 2, (x, y), M, N, cuda_device, False)
    assert ll2.get("engine") in (None, "cuda-core/chunked"), ll2      # the comparator ran on the CUDA cores
    assert torch.equal(o_tc, o_cc)
    # and a sample of rows against the CPU oracle (reference semantics: ties -> smallest index)
    rows = slice(0, 64)
    xa, ya = x.cpu().numpy(), y.cpu().numpy()
    if tagI == 0:
        ref = ko.genred(f, [], [xa[rows], ya], "float32")
    else:
        ref = ko.genred(f, [], [xa, ya[rows]], "float32")
    got = o_tc[rows].cpu().numpy()
    if kind == "ArgKMin":
        agree = (got == ref).all(axis=1).mean()
        assert agree > 0.9          # the numpy oracle does not contract FMAs: near-ties may swap
    else:
        np.testing.assert_allclose(got[:, ::2] if kind == "KMin_ArgKMin" else got,
                                   ref[:, ::2] if kind == "KMin_ArgKMin" else ref, rtol=1e-5)


@pytest.mark.gpu
def test_tensor_core_kmin_duplicates_and_outliers(cuda_device):
    """Adversarial inputs for the filter: exact duplicate columns (ties must go to the smallest index), a huge
    outlier column and rows, and clustered data whose neighbours are much closer than the norms."""
    import torch
    torch.manual_seed(5)
    M, N, D, K = 8192, 6000, 64, 10
    base = torch.randn(50, D, device=cuda_device) * 3
    y = base[torch.randint(0, 50, (N,), device=cuda_device)] + 0.01 * torch.randn(N, D, device=cuda_device)
    y[100:140] = y[60:100]                      # 40 exact duplicates of earlier columns
    y[7] = 1e3                                   # outlier column
    x = base[torch.randint(0, 50, (M,), device=cuda_device)] + 0.01 * torch.randn(M, D, device=cuda_device)
    x[3] = -5e2                                  # outlier row
    x[4] = y[60]                                 # a query that coincides with duplicated columns
    f = "ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,0)" % (D, D, K)
    o_tc, ll = _run(f, 2, (x, y), M, N, cuda_device, True)
    o_cc, _ = _run(f, 2, (x, y), M, N, cuda_device, False)
    assert ll.get("engine") == "tcgen05"
    assert torch.equal(o_tc, o_cc), "overflow rows: %s" % ll.get("overflow_rows")
    row = o_tc[4].long().tolist()
    assert row[0] == 60 and row[1] == 100       # distance 0 twice: the smaller index first


# ---------------------------------------------------------------------------------------------
# seeded split launches of the CUDA-core K-min kernel (binder._launch_kmin_seeded)
# ---------------------------------------------------------------------------------------------
def _run_seed(formula, dtype, args, M, N, dev, seed):
    import torch
    from keops_b200.binder import get_binder
    old = {k: os.environ.get(k) for k in ("KB200_KMIN_SEED", "KB200_KNN_TC")}
    os.environ["KB200_KMIN_SEED"] = "1" if seed else "0"
    os.environ["KB200_KNN_TC"] = "0"
    try:
        bd = get_binder(formula, [], len(args), dtype)
        out = bd.genred_pytorch(dev, None, M, N, 0, None, *args)
        torch.cuda.synchronize()
        return out, dict(bd.last_launch)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("formula,dtype,M,N,D", [
    ("ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)", "float32", 20000, 300000, 3),
    ("KMin_ArgKMin_Reduction(Square((Var(0,2,0) - Var(1,2,1))),3,0)", "float32", 9000, 140001, 2),   # dimF = 2
    ("KMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),5,1)", "float64", 100000, 7000, 3),       # reduction over i
    ("ArgKMin_Reduction((Var(0,3,0) | Var(1,3,1)),4,0)", "float32", 5000, 70000, 3),                  # negative values
])
def test_seeded_kmin_identical_to_unseeded(formula, dtype, M, N, D, cuda_device):
    """Column splits that share per-row thresholds through global memory must return exactly what
    independent splits return (values and indices, ties included)."""
    import torch
    tdt = torch.float32 if dtype == "float32" else torch.float64
    torch.manual_seed(M + N)
    x, y = torch.randn(M, D, device=cuda_device, dtype=tdt), torch.randn(N, D, device=cuda_device, dtype=tdt)
    tagI = int(formula.rstrip(")").split(",")[-1])
    streamed = y if tagI == 0 else x
    # exact duplicates, inside and outside the sample, so that ties at the seed value occur
    streamed[5000:5400] = streamed[100:500]
    streamed[-300:] = streamed[200:500]
    o_s, ll = _run_seed(formula, dtype, (x, y), M, N, cuda_device, True)
    assert ll.get("kmin_seed_columns", 0) > 0, ll        # the seeded path really ran
    o_u, ll2 = _run_seed(formula, dtype, (x, y), M, N, cuda_device, False)
    assert ll2.get("kmin_seed_columns", 0) == 0 and max(s[2] for s in ll2["segments"]) > 1
    assert torch.equal(o_s, o_u)
