"""Chunked-dimension kernels (``keops_b200/codegen/emit_chunked.py``; reference chunk mode:
``keopscore/keopscore/mapreduce/gpu/GpuReduc1D_chunks.py:19-118,259-330``, thresholds ``config/chunks.py:19-21``, test shapes
``pykeops/pykeops/test/test_chunks.py:5-41`` and ``test_torch.py:585-630``).  CPU: the decomposition and the generated code
(cross-compiled, no local-memory traffic); GPU: parity with the oracle for every reduction family."""
import os
import subprocess

import numpy as np
import pytest

from keops_b200.codegen import emit_chunked as ec
from keops_b200.formulas.parser import parse_reduction

MATERN = "Sum_Reduction(Exp(-Sqrt(SqDist(Var(0,256,0),Var(1,256,1))))*Var(2,1,1),0)"
CASES = [
    # formula, nargs, dtype, (nx, ny)
    ("Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,300,0),Var(1,300,1)))*Var(2,2,1),0)", 4, "float32", (333, 517)),
    (MATERN, 3, "float32", (257, 400)),
    ("Min_ArgMin_Reduction(Sum(Exp(Minus(Square((Var(0,990,0) + Var(1,990,1)))))),1)", 2, "float32", (70, 90)),
    ("ArgKMin_Reduction(Sum(Square((Var(0,200,0) - Var(1,200,1)))),5,0)", 2, "float32", (150, 400)),
    ("Max_SumShiftExp_Reduction((Var(0,300,0)|Var(1,300,1)),0)", 2, "float64", (120, 333)),
    ("Sum_Reduction(Exp(-SqDist(Var(0,301,0),Var(1,301,1)))*Var(2,1,1),1)", 3, "float32", (211, 190)),
    # two chunked sums in one formula (cosine similarity) + a row-only sum
    ("Sum_Reduction(((Var(0,128,0)|Var(1,128,1))*Rsqrt(SqNorm2(Var(0,128,0))*SqNorm2(Var(1,128,1))))*Var(2,3,1),0)", 3, "float64", (90, 130)),
    ("KMin_ArgKMin_Reduction(Sum(Abs((Var(0,150,0) - Var(1,150,1)))),3,0)", 2, "float64", (64, 301)),
]


def test_decomposition_into_chunked_and_post_formulas():
    red = parse_reduction(MATERN)
    pl = ec.plan_chunks(red)
    assert pl is not None and pl.D == 256 and len(pl.roots) == 1 and pl.roots[0].op == "Sum"
    assert pl.hd_vars == [(0, 0), (1, 1)]
    assert [v.params for v in pl.post.variables()] == [(2, 1, 1), (ec.PH_BASE, 1, 3)]      # b_j and the placeholder
    assert repr(pl.bodies[0]) == "Square(Subtract(Var(0,1,4),Var(1,1,5)))"
    # low dimension: nothing to chunk; a high-dimensional variable used outside the sums (a gradient): not chunkable
    assert ec.plan_chunks(parse_reduction(MATERN.replace("256", "64"))) is None
    g = "Grad_WithSavedForward(%s,Var(0,256,0),Var(3,1,0),Var(4,1,0))" % MATERN
    assert ec.plan_chunks(parse_reduction(g)) is None
    # the reference's own chunk-mode test formula (test_torch.py:585-630): a non-bilinear element-wise body
    pl = ec.plan_chunks(parse_reduction(CASES[2][0]))
    assert pl is not None and repr(pl.bodies[0]) == "Exp(Minus(Square(Add(Var(0,1,4),Var(1,1,5)))))"


@pytest.mark.parametrize("formula,nargs,dtype,sizes", CASES)
def test_chunked_kernels_compile_without_local_memory(formula, nargs, dtype, sizes):
    from keops_b200.runtime import capi, plan
    red = parse_reduction(formula)
    pl = ec.plan_chunks(red)
    assert pl is not None
    ks = ec.generate_chunked(red, pl, dtype, nargs)
    assert ks is not None and ks.dynamic_smem <= 227 * 1024 and len(ks.source) < 60000      # no D-wide unrolling
    cubin = capi.compile_to_cubin(ks.source, plan.nvrtc_options(dtype == "float32"))
    path = "/tmp/kb200_chunk_%d.cubin" % os.getpid()
    open(path, "wb").write(cubin)
    try:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "kb200_reduce", path], capture_output=True, text=True).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    finally:
        os.remove(path)
    assert "LDL" not in sass and "STL" not in sass, "chunked kernel spills to local memory"
    assert "LDGSTS" in sass and ("LDS.128" in sass or dtype == "float64")
    if dtype == "float32":
        assert "FFMA2" in sass or "FADD2" in sass


def _inputs(formula, sizes, dtype, seed=0, integer=False):
    red = parse_reduction(formula)
    rng = np.random.default_rng(seed)
    nx, ny = sizes
    args = {}
    for v in red.variables():
        ind, dim, cat = v.params
        n = {0: nx, 1: ny}.get(cat)
        shape = (dim,) if cat == 2 else (n, dim)
        if integer:
            a = rng.integers(0, 2, shape).astype(np.float64)
        elif cat == 2:
            a = np.full(shape, 0.7)
        elif "SqDist" in formula or "Abs" in formula or "|" in formula or "Square((Var(0,200" in formula or "301" in formula:
            a = rng.random(shape) / np.sqrt(dim) if dim >= 100 else rng.standard_normal(shape)
        else:
            a = rng.standard_normal(shape)        # well-separated values (no fp32-level near-ties between columns)
        args[ind] = a.astype(dtype)
    return [args[k] for k in sorted(args)]


@pytest.mark.gpu
@pytest.mark.parametrize("formula,nargs,dtype,sizes", CASES)
def test_chunked_kernel_matches_oracle(formula, nargs, dtype, sizes, cuda_device):
    import torch
    from keops_b200.binder import get_binder
    from oracle import keops_oracle as ko
    from tests import util
    args = _inputs(formula, sizes, dtype)
    nx, ny = sizes
    bd = get_binder(formula, [], nargs, dtype)
    os.environ["KB200_CONV_TC"] = "0"
    try:
        out = bd.genred_pytorch(cuda_device, None, nx, ny, 0, None, *[torch.from_numpy(a).to(cuda_device) for a in args])
        torch.cuda.synchronize()
    finally:
        os.environ.pop("KB200_CONV_TC", None)
    assert bd.last_launch.get("engine") == "cuda-core/chunked", bd.last_launch
    got = out.cpu().numpy()
    ref = ko.genred(formula, [], args, dtype)
    head = formula.split("(")[0]
    d = bd.red.dimF
    name = "chunked " + formula[:48]
    if head == "ArgKMin_Reduction":
        assert np.array_equal(got, ref), name
    elif head == "Min_ArgMin_Reduction":
        util.assert_close(got[:, :d], ref[:, :d], dtype, name)
        assert np.array_equal(got[:, d:], ref[:, d:]), name
    elif head == "KMin_ArgKMin_Reduction":
        g3, r3 = got.reshape(len(got), -1, 2, d), ref.reshape(len(ref), -1, 2, d)
        util.assert_close(g3[:, :, 0], r3[:, :, 0], dtype, name)
        assert np.array_equal(g3[:, :, 1], r3[:, :, 1]), name
    else:
        util.assert_close(got, ref, dtype, name, scale=ko.abs_scale(formula, [], args))


@pytest.mark.gpu
def test_chunked_ties_resolve_to_the_smallest_index(cuda_device):
    """Binary-valued points in D = 128: many exactly equal distances.  The column groups of a CTA see interleaved columns, so
    the partial lists are merged with an explicit index comparison -- the result must be the oracle's (smallest index)."""
    import torch
    from keops_b200.binder import get_binder
    from oracle import keops_oracle as ko
    for formula in ("ArgKMin_Reduction(Sum(Square((Var(0,128,0) - Var(1,128,1)))),7,0)",
                    "ArgMin_Reduction(Sum(Abs((Var(0,128,0) - Var(1,128,1)))),1)"):
        args = _inputs(formula, (90, 1200), "float32", seed=3, integer=True)
        args = [a[:, :128] * (np.arange(128) < 6) for a in args]          # only 6 informative coordinates: massive ties
        args = [np.ascontiguousarray(a, dtype=np.float32) for a in args]
        bd = get_binder(formula, [], 2, "float32")
        out = bd.genred_pytorch(cuda_device, None, 90, 1200, 0, None, *[torch.from_numpy(a).to(cuda_device) for a in args])
        torch.cuda.synchronize()
        assert bd.last_launch.get("engine") == "cuda-core/chunked"
        assert np.array_equal(out.cpu().numpy(), ko.genred(formula, [], args, "float32")), formula


@pytest.mark.gpu
def test_chunked_kernel_with_batch_dimensions_and_column_splits(cuda_device):
    """Batch dimensions (gridDim.z), few rows against many columns (column splits x column groups partials) and a
    user-provided output buffer."""
    import torch
    from keops_b200.torch import Genred
    torch.manual_seed(0)
    dev = cuda_device
    B, M, N, D = 3, 40, 9000, 160
    x, y = torch.rand(B, M, D, device=dev) / D ** 0.5, torch.rand(B, N, D, device=dev) / D ** 0.5
    b = torch.randn(B, N, 2, device=dev)
    conv = Genred("Exp(-Sqrt(SqDist(x,y)))*b", ["x=Vi(%d)" % D, "y=Vj(%d)" % D, "b=Vj(2)"], reduction_op="Sum", axis=1)
    got = conv(x, y, b)
    ref = torch.einsum("bmn,bnd->bmd", torch.exp(-torch.cdist(x.double(), y.double())), b.double())
    assert got.shape == (B, M, 2)
    assert torch.allclose(got.double(), ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())
    from keops_b200.torch import genred as gm
    assert gm.LAST_BINDER.last_launch.get("engine") == "cuda-core/chunked", gm.LAST_BINDER.last_launch
    out = torch.empty(M, 2, device=dev)
    ret = conv(x[0], y[0], b[0], out=out)
    assert ret.data_ptr() == out.data_ptr() and torch.allclose(out, got[0], rtol=1e-5, atol=1e-6)
