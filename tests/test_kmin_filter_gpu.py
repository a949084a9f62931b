"""Filtered K-min kernel for squared distances in dimension <= 3 (``keops_b200/codegen/emit_kmin_filter.py``): the pair loop
evaluates a 3-FMA filter and only survivors are evaluated exactly, so the output must be *bit-identical* to the generic
kernel's, whose parity with the reference is pinned by the golden fixtures (``argkmin_*``, ``kmin_*``, ``ties_*``)."""
import os

import numpy as np
import pytest

from keops_b200.codegen import emit_kmin_filter
from keops_b200.formulas.parser import parse_reduction


def test_match_and_cross_compile():
    """Host logic + the generated source compiles for sm_100a (NVRTC, no GPU needed)."""
    from keops_b200.binder import get_binder
    from keops_b200.codegen.emit import KernelConfig
    from keops_b200.runtime import plan
    assert emit_kmin_filter.match(parse_reduction("ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)")) == (0, 1, 3)
    assert emit_kmin_filter.match(parse_reduction("KMin_ArgKMin_Reduction(SqDist(Var(1,2,1),Var(0,2,0)),5,1)")) == (1, 0, 2)
    assert emit_kmin_filter.match(parse_reduction("ArgMin_Reduction(SqDist(Var(0,3,0),Var(1,3,1)),0)")) == (0, 1, 3)
    assert emit_kmin_filter.match(parse_reduction("ArgKMin_Reduction(SqDist(Var(0,15,0),Var(1,15,1)),10,0)")) == (0, 1, 15)
    for f in ("ArgKMin_Reduction(Sum(Square((Var(0,16,0) - Var(1,16,1)))),10,0)",          # D = 16: the tensor-core filter's
              "ArgKMin_Reduction(Sum(Abs((Var(0,3,0) - Var(1,3,1)))),10,0)",               # not a squared distance
              "Max_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)",                   # not a minimum
              "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,0)))),10,0)"):           # both operands on one side
        assert emit_kmin_filter.match(parse_reduction(f)) is None
    bd = get_binder("KMin_ArgKMin_Reduction(Sum(Square((Var(1,2,1) - Var(0,2,0)))),5,1)", [], 2, "float32")
    c0 = bd.ks.cfg
    ks = emit_kmin_filter.generate_filtered(bd.red, 2, KernelConfig(T=128, R=c0.R, TJ=c0.TJ, unroll=c0.unroll, min_blocks=3,
                                                                    group=c0.group), bd.ks)
    assert ks is not None and ks.nargs == 4 and ks.dimpart == 10 and "kb200_kfpack" in ks.source
    assert "kb200_reduce_seeded" in ks.source and ks.combine_name in ks.source
    assert len(plan.get_cubin(ks, True)) > 1000


def _run(formula, args, M, N, dev, filtered):
    import torch
    from keops_b200.binder import get_binder
    old = os.environ.get("KB200_KMIN_FILTER")
    os.environ["KB200_KMIN_FILTER"] = "force" if filtered else "0"      # "force": the test shapes are below the size heuristics
    try:
        bd = get_binder(formula, [], 2, "float32")
        out = bd.genred_pytorch(dev, None, M, N, 0, None, *args)
        torch.cuda.synchronize()
        return out, dict(bd.last_launch)
    finally:
        if old is None:
            os.environ.pop("KB200_KMIN_FILTER", None)
        else:
            os.environ["KB200_KMIN_FILTER"] = old


CASES = [
    # kind, D, K, nx, ny, tagI, subtraction order, data
    ("ArgKMin", 3, 10, 20000, 30000, 0, "xy", "randn"),
    ("KMin_ArgKMin", 2, 5, 15000, 25001, 1, "yx", "randn"),        # reduction over i, ragged tile, y - x in the formula
    ("KMin", 3, 3, 12000, 40000, 0, "xy", "uniform"),
    ("ArgKMin", 3, 10, 1500, 200000, 0, "xy", "randn"),            # few rows: column splits + shared thresholds + combine
    ("ArgKMin", 3, 16, 30000, 20000, 0, "xy", "offset"),           # clouds far from the origin: the common shift
    ("ArgKMin", 3, 10, 20000, 20000, 0, "xy", "grid"),             # integer-valued points: exact ties -> smallest index
    ("KMin_ArgKMin", 3, 4, 20000, 20000, 0, "xy", "nonfinite"),    # NaN / inf columns are never selected, as in the generic kernel
    ("ArgKMin", 2, 1, 40000, 10000, 0, "xy", "outlier"),           # one huge point inflates the margin: still exact
    # nearest neighbour / K-means assignment: the K = 1 case of the same kernel (K is ignored in the formula string)
    ("ArgMin", 3, 0, 300000, 1000, 0, "xy", "randn"),
    ("Min_ArgMin", 2, 0, 20000, 30000, 1, "yx", "grid"),
    ("Min", 3, 0, 1500, 200000, 0, "xy", "uniform"),               # few rows: column splits + combine
    # 4 <= D <= 15: two rows per thread, several 16-byte reads per column (filter row / coordinates: 2+2, 3+2, 4+3, 4+4 float4)
    ("ArgKMin", 5, 10, 20000, 30000, 0, "xy", "randn"),
    ("KMin_ArgKMin", 8, 4, 15000, 25001, 1, "yx", "offset"),
    ("ArgKMin", 12, 16, 12000, 20000, 0, "xy", "grid"),
    ("KMin", 15, 10, 20000, 20000, 0, "xy", "nonfinite"),
    ("ArgMin", 10, 0, 1500, 100000, 0, "xy", "outlier"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,D,K,nx,ny,tagI,order,data", CASES)
def test_filtered_kmin_identical_to_generic_kernel(kind, D, K, nx, ny, tagI, order, data, cuda_device):
    import torch
    dev = cuda_device
    torch.manual_seed(D * 100 + K + len(kind))
    if data == "uniform":
        x, y = torch.rand(nx, D, device=dev), torch.rand(ny, D, device=dev)
    elif data == "grid":
        x = torch.randint(0, 12, (nx, D), device=dev).float()
        y = torch.randint(0, 12, (ny, D), device=dev).float()
    else:
        x, y = torch.randn(nx, D, device=dev), torch.randn(ny, D, device=dev)
    if data == "offset":
        x, y = x + 1000.0, y + 1000.0
    streamed = y if tagI == 0 else x
    if data == "nonfinite":
        streamed[5, 0] = float("nan")
        streamed[77, 1] = float("inf")
        streamed[1234, 2] = -float("inf")
    if data == "outlier":
        streamed[3] = 1.0e6
    sub = "(Var(0,%d,0) - Var(1,%d,1))" % (D, D) if order == "xy" else "(Var(1,%d,1) - Var(0,%d,0))" % (D, D)
    if K:
        f = "%s_Reduction(Sum(Square(%s)),%d,%d)" % (kind, sub, K, tagI)
    else:
        f = "%s_Reduction(Sum(Square(%s)),%d)" % (kind, sub, tagI)
    o_f, ll = _run(f, (x, y), nx, ny, dev, True)
    assert ll.get("engine") == "cuda-core/kmin-filter", ll
    if (nx if tagI == 0 else ny) < 2000:
        assert ll["nsplit"] > 1 and ll["kmin_seed_columns"] == (1 if K else 0), ll
    o_g, ll2 = _run(f, (x, y), nx, ny, dev, False)
    assert ll2.get("engine") is None, ll2
    if data == "nonfinite":                                   # NaN != NaN: compare bit patterns
        assert torch.equal(o_f.view(torch.int32), o_g.view(torch.int32))
    else:
        assert torch.equal(o_f, o_g)
    if kind == "ArgKMin" and data in ("randn", "uniform"):    # and a sample of rows against dense torch (float64)
        rows = slice(0, 256)
        a, b = (x, y) if tagI == 0 else (y, x)
        d2 = torch.cdist(a[rows].double(), b.double()) ** 2
        ref = d2.topk(K, largest=False).indices
        agree = (ref.sort(1).values == o_f[rows].long().sort(1).values).all(1).float().mean().item()
        assert agree > 0.98                                   # fp32 near-ties may swap the last neighbour
