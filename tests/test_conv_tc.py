"""Tensor-core Gaussian Sum reduction (``keops_b200/csrc/kb200_conv_tc.cu``, ``codegen/conv_tc.py``).

CPU part: pattern recognition, operand layout, eligibility.  GPU part: the tcgen05 path against the CUDA-core kernel
(whose parity with the reference is pinned by the golden fixtures), against the reference's own generated CUDA and CPU
code for the KernelSolve formula (``oracle/_ref``: gauss16_pm_f32) and against the numpy oracle; the accuracy guard
and its fallback.  Tolerance: the north-star fp32 bar, 1e-5 relative (``tests/util.py``)."""
import os

import numpy as np
import pytest

from keops_b200.codegen import conv_tc
from keops_b200.formulas.parser import parse_reduction
from tests import util

KS = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)"


# ---------------------------------------------------------------------------------------------
# host logic (no GPU)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("formula,expect", [
    (KS, (0, 1, 16, 2, 1)),
    ("Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)", (0, 1, 3, 2, 1)),
    ("Sum_Reduction(Var(2,3,1)*Exp(-SqDist(Var(0,8,0),Var(1,8,1))/Var(3,1,2)),0)", (0, 1, 8, 2, 3)),
    ("Sum_Reduction(Exp(-IntCst(2)*SqNorm2(Var(0,8,0)-Var(1,8,1)))*Var(2,1,1),0)", (0, 1, 8, 2, 1)),
    # reduction over i: the output follows j, x_i and b_i are streamed
    ("Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,2,0),1)", (1, 0, 16, 2, 2)),
    ("Sum_Reduction(Exp(-(Var(3,1,2)*Var(3,1,2))*SqDist(Var(1,16,1),Var(0,16,0)))*Var(2,1,1),0)", (0, 1, 16, 2, 1)),
])
def test_match_recognises_gaussian_convolutions(formula, expect):
    m = conv_tc.match(parse_reduction(formula))
    assert m is not None and m[:5] == expect


@pytest.mark.parametrize("formula", [
    "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,0),0)",      # b is output-indexed
    "Sum_Reduction(Exp(-Var(3,1,0)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)",      # per-row bandwidth
    "Sum_Reduction(Exp(-Sqrt(SqDist(Var(0,16,0),Var(1,16,1))))*Var(2,1,1),0)",           # Laplacian kernel
    "Sum_Reduction(Exp(-SqDist(Var(0,16,0),Var(1,16,0)))*Var(2,1,1),0)",                 # both operands on one side
    "Max_SumShiftExp_Reduction(-SqDist(Var(0,16,0),Var(1,16,1)),0)",
    "Sum_Reduction(Exp(-SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,9,1),0)",                 # dim b > MAX_DB
])
def test_match_rejects_other_formulas(formula):
    assert conv_tc.match(parse_reduction(formula)) is None


def test_match_grad_recognises_position_gradients_structurally():
    """``Grad_WithSavedForward`` of the Gaussian sum w.r.t. x / y (what the backward of ``Genred`` / ``KernelSolve`` requests,
    pykeops/pykeops/torch/generic/generic_red.py:190-214) is recognised from the provenance the parser records; gradients
    w.r.t. b already have the forward's shape (transposed) and match the plain pattern; a parameter gradient does not."""
    g = "Grad_WithSavedForward(" + KS + ",%s,Var(4,1,0),Var(5,1,0))"
    mx = conv_tc.match_grad(parse_reduction(g % "Var(0,16,0)"))
    assert mx[:7] == ("x", 0, 1, 16, 2, 4, 5)
    assert conv_tc.match_grad(parse_reduction(g % "Var(1,16,1)")) is None      # d/dy: measured no faster, not routed
    os.environ["KB200_CONV_TC_GRAD_Y"] = "1"
    try:
        my = conv_tc.match_grad(parse_reduction(g % "Var(1,16,1)"))
    finally:
        del os.environ["KB200_CONV_TC_GRAD_Y"]
    assert my[:7] == ("y", 0, 1, 16, 2, 4, 5)
    rb = parse_reduction(g % "Var(2,1,1)")
    assert conv_tc.match_grad(rb) is None and conv_tc.match(rb) is not None and rb.tagI == 1
    assert conv_tc.match_grad(parse_reduction(g % "Var(3,1,2)")) is None
    assert conv_tc.match_grad(parse_reduction(KS)) is None
    # dim b = 2 (the upstream gradient is a vector), D = 3: outside the path
    g2 = "Grad_WithSavedForward(Sum_Reduction(Exp(-SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,2,1),0),Var(0,16,0),Var(3,2,0),Var(4,2,0))"
    g3 = "Grad_WithSavedForward(Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,1,1),0),Var(0,3,0),Var(3,1,0),Var(4,1,0))"
    assert conv_tc.match_grad(parse_reduction(g2)) is None and conv_tc.match_grad(parse_reduction(g3)) is None
    # provenance is not part of a reduction's identity
    assert parse_reduction(g % "Var(2,1,1)") == parse_reduction(str_of_grad_b())


def str_of_grad_b():
    return "Sum_Reduction(Var(4,1,0)*Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1))),1)"


def test_layout_and_eligibility():
    assert conv_tc.layout(16) == (1, 64, 1)        # 96 product + 3 penalty columns, padded to 128
    assert conv_tc.layout(21) == (1, 64, 0)        # 126 product columns; |y|^2 goes to the epilogue
    assert conv_tc.layout(8) == (0, 64, 1)         # 48 + 3 columns in one 64-column block
    assert conv_tc.layout(20) == (1, 64, 1)
    red = parse_reduction(KS)
    assert conv_tc.eligible(red, "float32", 100000, 100000, 1, True) is not None
    assert conv_tc.eligible(red, "float64", 100000, 100000, 1, True) is None       # fp64: exact path only
    assert conv_tc.eligible(red, "float32", 100000, 100000, 1, False) is None      # no fast math: exact expf
    assert conv_tc.eligible(red, "float32", 100000, 100000, 2, True) is None       # batch dimensions
    assert conv_tc.eligible(red, "float32", 1000, 100000, 1, True) is None         # too few rows to fill the machine
    assert conv_tc.eligible(red, "float32", 2000, 100000, 1, True) is not None     # ... unless the columns can be split
    assert conv_tc.eligible(red, "float32", 2000, 50000, 1, True) is None
    assert conv_tc.eligible(red, "float32", 100000, 100000, 1, True, "kahan_scheme") is None
    g3 = parse_reduction("Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)")
    assert conv_tc.eligible(g3, "float32", 100000, 100000, 1, True) is None        # D = 3 stays on the CUDA cores
    g64 = parse_reduction(KS.replace("16", "64"))
    assert conv_tc.is_hd(64) and conv_tc.layout(64) == (6, 64, 1)                  # 387 columns -> 7 k-blocks of 64
    assert conv_tc.eligible(g64, "float32", 100000, 100000, 1, True) is not None   # high-dimensional kernel
    g96 = parse_reduction(KS.replace("16", "96"))
    assert conv_tc.eligible(g96, "float32", 100000, 100000, 1, True) is not None   # K > 576: both operands stream
    assert conv_tc.smem_bytes(96, 1) < conv_tc.smem_bytes(95, 1)
    g4k = parse_reduction(KS.replace("16", "4096"))
    assert conv_tc.eligible(g4k, "float32", 100000, 100000, 1, True) is None
    assert conv_tc.eligible(parse_reduction(KS.replace("Var(2,1,1)", "Var(2,8,1)")), "float32", 100000, 100000, 1, True) is None  # b ring too wide for smem
    for D, DB in ((8, 1), (16, 1), (16, 4), (21, 4)):
        assert conv_tc.smem_bytes(D, DB) <= 227 * 1024


def test_kernels_compile_for_sm100a():
    eng = conv_tc.ConvTensorCore(16, 1, device=-1)            # NVRTC, --gpu-architecture=sm_100a, no GPU needed
    assert eng.cubin[:4] == b"\x7fELF"


# ---------------------------------------------------------------------------------------------
# GPU parity
# ---------------------------------------------------------------------------------------------
class _env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(formula, nargs, args, nx, ny, dev, tc, **env):
    import torch
    from keops_b200.binder import get_binder
    with _env(KB200_CONV_TC="1" if tc else "0", **env):
        bd = get_binder(formula, [], nargs, "float32")
        out = bd.genred_pytorch(dev, None, nx, ny, 0, None, *args)
        torch.cuda.synchronize()
        return out, dict(bd.last_launch)


def _scale(formula, nargs, args, nx, ny, dev):
    """Condition scale sum_j |F_ij| (CUDA-core evaluation; tests/util.py::gpu_abs_scale) as a numpy array."""
    with _env(KB200_CONV_TC="0"):
        return util.gpu_abs_scale(formula, nargs, "float32", dev, nx, ny, *args).cpu().numpy()


def _ms_scale(o):
    """(m, s) outputs of a LogSumExp reduction with positive weights: s is compared relatively, the running maximum m
    (an exponent, possibly close to 0) with an absolute floor of 1."""
    sc = np.abs(o)
    sc[:, 0] = np.maximum(sc[:, 0], 1.0)
    return sc


CASES = [
    # D, DB, nx, ny, tagI, coefficient spelling
    (16, 1, 9000, 7001, 0, "-Var(3,1,2)*"),           # KernelSolve formula, ragged last tile
    (8, 3, 8192, 4096, 0, "-Var(3,1,2)*"),            # penalty columns inside K, vector b
    (5, 1, 9000, 5000, 0, "-Var(3,1,2)*"),            # smallest dimension on the tensor cores (K = 33 -> 64)
    (6, 2, 3000, 8192, 1, "-IntCst(3)*"),             # reduction over i: 8192 output rows follow j
    (13, 2, 8300, 5000, 0, "-IntCst(2)*"),            # constant coefficient, K = 64 + 16
    (20, 1, 8192, 3000, 0, "-Var(3,1,2)*Var(3,1,2)*"),
    (16, 2, 5000, 9000, 1, "-Var(3,1,2)*"),           # reduction over i
    (21, 4, 8192, 2000, 0, "-Var(3,1,2)*"),           # widest operand rows of the 256-row kernel
    (22, 1, 8200, 3001, 0, "-Var(3,1,2)*"),           # high-dimensional kernel (kb200_conv_tc_hd): 3 k-blocks
    (64, 2, 8192, 2500, 0, "-Var(3,1,2)*"),           # 7 k-blocks, vector b
    (95, 1, 4000, 8300, 1, "-Var(3,1,2)*"),           # 9 k-blocks (the resident maximum), reduction over i
    (128, 2, 8192, 1300, 0, "-Var(3,1,2)*"),          # 13 k-blocks: both operands stream through the ring
    (300, 1, 8192, 1100, 0, "-Var(3,1,2)*"),          # 29 k-blocks
    (16, 1, 1500, 70000, 0, "-Var(3,1,2)*"),          # few rows, many columns: column chunks fill the machine
    (64, 2, 1100, 66000, 0, "-Var(3,1,2)*"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("D,DB,nx,ny,tagI,coef", CASES)
def test_tensor_core_conv_matches_cuda_core(D, DB, nx, ny, tagI, coef, cuda_device):
    import torch
    torch.manual_seed(D * 100 + DB)
    dev = cuda_device
    x, y = torch.rand(nx, D, device=dev), torch.rand(ny, D, device=dev)
    b = torch.randn((ny if tagI == 0 else nx), DB, device=dev)
    g = torch.tensor([0.7 if D <= 21 else 8.0 / D], device=dev)      # keep exp(-g |x-y|^2) away from underflow
    f = "Sum_Reduction(Exp(%sSqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,%d,%d),%d)" % (coef, D, D, DB, 1 - tagI, tagI)
    o_tc, ll = _run(f, 4, (x, y, b, g), nx, ny, dev, True)
    assert ll.get("engine") == "tcgen05-conv", ll         # the tensor-core kernels really ran
    assert ll["error_bound"] <= 5e-6
    o_cc, ll2 = _run(f, 4, (x, y, b, g), nx, ny, dev, False)
    assert ll2.get("engine") in (None, "cuda-core/chunked"), ll2      # the comparator ran on the CUDA cores
    sc = _scale(f, 4, (x, y, b, g), nx, ny, dev)
    tag = " D=%d DB=%d %dx%d tagI=%d" % (D, DB, nx, ny, tagI)
    util.assert_close(o_tc.cpu().numpy(), o_cc.cpu().numpy(), "float32", "tcgen05 conv vs CUDA cores" + tag, scale=sc)
    # several column chunks (forced) give the same sums up to fp32 re-association
    o_ch, ll3 = _run(f, 4, (x, y, b, g), nx, ny, dev, True, KB200_CONV_TC_CHUNK_MB=1)
    if (ny if tagI == 0 else nx) * 2 * (64 * conv_tc.layout(D)[0] + conv_tc.layout(D)[1]) > (1 << 20):
        assert ll3["nchunks"] > 1
    util.assert_close(o_ch.cpu().numpy(), o_cc.cpu().numpy(), "float32", "chunked tcgen05 conv vs CUDA cores" + tag, scale=sc)
    # sampled rows against the numpy oracle (reference operation order, exact exp)
    from oracle import keops_oracle as ko
    rows = slice(0, 48)
    xa, ya, ba, ga = (t.cpu().numpy() for t in (x, y, b, g))
    if tagI == 0:
        ref = ko.genred(f, [], [xa[rows], ya, ba, ga], "float32")
    else:
        ref = ko.genred(f, [], [xa, ya[rows], ba, ga], "float32")
    util.assert_close(o_tc[rows].cpu().numpy(), ref, "float32", "tcgen05 conv vs oracle" + tag, scale=sc[rows])


@pytest.mark.gpu
def test_tensor_core_conv_vs_reference_code(cuda_device):
    """KernelSolve matvec formula against the reference's own generated CUDA kernel (all rows) and CPU code (sample)."""
    import torch
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built")
    dev = cuda_device
    torch.manual_seed(3)
    nx, ny = 40000, 30011
    x, y, b = torch.rand(nx, 16, device=dev), torch.rand(ny, 16, device=dev), torch.randn(ny, 1, device=dev)
    g = torch.tensor([1.0], device=dev)
    o_tc, ll = _run(KS, 4, (x, y, b, g), nx, ny, dev, True)
    assert ll.get("engine") == "tcgen05-conv"
    ref = ref_runner.RefCuda("gauss16_pm_f32", 0)([x, y, b, g], nx, ny)
    torch.cuda.synchronize()
    sc = _scale(KS, 4, (x, y, b, g), nx, ny, dev)
    util.assert_close(o_tc.cpu().numpy(), ref.cpu().numpy(), "float32", "tcgen05 conv D=16 vs reference CUDA", scale=sc)
    rows = np.arange(0, nx, 157)
    cpu = ref_runner.run_cpu("gauss16_pm_f32", [x.cpu().numpy()[rows], y.cpu().numpy(), b.cpu().numpy(), g.cpu().numpy()],
                             len(rows), ny)
    util.assert_close(o_tc.cpu().numpy()[rows], cpu, "float32", "tcgen05 conv D=16 vs reference CPU", scale=sc[rows])


@pytest.mark.gpu
def test_accuracy_guard_falls_back_to_cuda_cores(cuda_device):
    """A cloud much wider than the kernel bandwidth would lose digits in |x|^2 + |y|^2 - 2 x.y: the guard must send it
    to the CUDA-core kernel (which forms x - y directly), and the result must be that kernel's, bit for bit."""
    import torch
    dev = cuda_device
    torch.manual_seed(4)
    nx, ny = 8192, 4096
    x, y = 50.0 * torch.rand(nx, 16, device=dev), 50.0 * torch.rand(ny, 16, device=dev)
    y[:nx // 2] = x[:ny][:nx // 2] + 0.05 * torch.randn(nx // 2, 16, device=dev)     # close pairs carry the sums
    b, g = torch.randn(ny, 1, device=dev), torch.tensor([2.0], device=dev)
    o, ll = _run(KS, 4, (x, y, b, g), nx, ny, dev, True)
    assert ll.get("engine") is None, ll
    o_cc, _ = _run(KS, 4, (x, y, b, g), nx, ny, dev, False)
    assert torch.equal(o, o_cc)
    # a shifted (not wider) cloud is fine: the common shift removes the offset
    x2, y2 = torch.rand(nx, 16, device=dev) + 1000.0, torch.rand(ny, 16, device=dev) + 1000.0
    o2, ll2 = _run(KS, 4, (x2, y2, b, g), nx, ny, dev, True)
    assert ll2.get("engine") == "tcgen05-conv" and ll2["error_bound"] <= 5e-6
    o2_cc, _ = _run(KS, 4, (x2, y2, b, g), nx, ny, dev, False)
    util.assert_close(o2.cpu().numpy(), o2_cc.cpu().numpy(), "float32", "tcgen05 conv, shifted cloud",
                      scale=_scale(KS, 4, (x2, y2, b, g), nx, ny, dev))


@pytest.mark.gpu
def test_tensor_core_conv_is_deterministic_and_linear(cuda_device):
    import torch
    dev = cuda_device
    torch.manual_seed(5)
    nx, ny = 10000, 12000
    x, y = torch.rand(nx, 16, device=dev), torch.rand(ny, 16, device=dev)
    b1, b2, g = torch.randn(ny, 1, device=dev), torch.randn(ny, 1, device=dev), torch.tensor([1.3], device=dev)
    o1, _ = _run(KS, 4, (x, y, b1, g), nx, ny, dev, True)
    o1b, _ = _run(KS, 4, (x, y, b1, g), nx, ny, dev, True)
    assert torch.equal(o1, o1b)
    o2, _ = _run(KS, 4, (x, y, b2, g), nx, ny, dev, True)
    o12, _ = _run(KS, 4, (x, y, b1 + 2.0 * b2, g), nx, ny, dev, True)
    sc = _scale(KS, 4, (x, y, b1, g), nx, ny, dev) + 2.0 * _scale(KS, 4, (x, y, b2, g), nx, ny, dev)
    util.assert_close(o12.cpu().numpy(), (o1 + 2.0 * o2).cpu().numpy(), "float32", "tcgen05 conv, linearity in b", scale=sc,
                      rtol=2e-5)


# ---------------------------------------------------------------------------------------------
# LogSumExp on the tensor cores (kb200_conv_tc / kb200_conv_tc_hd in LSE mode)
# ---------------------------------------------------------------------------------------------
def test_match_lse_patterns():
    soft = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,16,0) - Var(2,16,1)))) / IntCst(2)) / Var(3,1,2))),0)"
    m = conv_tc.match_lse(parse_reduction(soft))
    assert m is not None and m[:6] == (1, 2, 16, 0, None, 1)
    w = "Max_SumShiftExpWeight_Reduction(-SqDist(Var(0,16,0),Var(1,16,1))*Var(3,1,2),0,Var(2,2,1))"
    assert conv_tc.match_lse(parse_reduction(w))[:6] == (0, 1, 16, None, 2, 2)
    assert conv_tc.eligible_lse(parse_reduction(soft), "float32", 100000, 100000, 1, True) is not None
    assert conv_tc.eligible_lse(parse_reduction(soft.replace("16", "3")), "float32", 100000, 100000, 1, True) is None
    assert conv_tc.eligible_lse(parse_reduction(soft.replace("16", "21")), "float32", 100000, 100000, 1, True) is None
    # a row-indexed offset or a non-distance exponent is not this pattern
    assert conv_tc.match_lse(parse_reduction(soft.replace("Var(0,1,1)", "Var(0,1,0)"))) is None
    assert conv_tc.match_lse(parse_reduction("Max_SumShiftExp_Reduction((Var(0,16,0) | Var(1,16,1)),0)")) is None
    assert conv_tc.match_lse(parse_reduction(KS)) is None


LSE_CASES = [
    # D, weight dim (0 = plain LogSumExp), nx, ny, tagI
    (16, 0, 9000, 7001, 0),
    (8, 2, 8192, 4096, 0),
    (16, 0, 5000, 9000, 1),
    (64, 0, 8192, 2500, 0),       # high-dimensional kernel
    (128, 3, 8192, 1300, 0),      # both operands streamed
    (16, 2, 1500, 70000, 0),      # few rows, many columns: column chunks + (m, s) combine
]


@pytest.mark.gpu
@pytest.mark.parametrize("D,DW,nx,ny,tagI", LSE_CASES)
def test_tensor_core_lse_matches_cuda_core(D, DW, nx, ny, tagI, cuda_device):
    import torch
    torch.manual_seed(D * 10 + DW)
    dev = cuda_device
    x, y = torch.rand(nx, D, device=dev), torch.rand(ny, D, device=dev)
    ns = ny if tagI == 0 else nx                      # streamed side
    a = torch.rand(ns, 1, device=dev)
    eps = torch.tensor([D / 24.0], device=dev)        # temperature ~ typical squared distance / 8
    cs = 1 - tagI
    if DW:
        w = torch.rand(ns, DW, device=dev) + 0.1
        f = ("Max_SumShiftExpWeight_Reduction((Var(2,1,%d) - ((Sum(Square((Var(0,%d,0) - Var(1,%d,1)))) / IntCst(2)) / Var(3,1,2))),%d,Var(4,%d,%d))"
             % (cs, D, D, tagI, DW, cs))
        args = (x, y, a, eps, w)
    else:
        f = ("Max_SumShiftExp_Reduction((Var(2,1,%d) - ((Sum(Square((Var(0,%d,0) - Var(1,%d,1)))) / IntCst(2)) / Var(3,1,2))),%d)"
             % (cs, D, D, tagI))
        args = (x, y, a, eps)
    o_tc, ll = _run(f, len(args), args, nx, ny, dev, True)
    assert ll.get("engine") == "tcgen05-lse", ll
    o_cc, ll2 = _run(f, len(args), args, nx, ny, dev, False)
    assert ll2.get("engine") in (None, "cuda-core/chunked"), ll2      # the comparator ran on the CUDA cores
    # raw (m, s) output: m = max_j F_ij, s = sum_j exp(F_ij - m) b_j
    tag = " D=%d DW=%d %dx%d tagI=%d" % (D, DW, nx, ny, tagI)
    util.assert_close(o_tc.cpu().numpy(), o_cc.cpu().numpy(), "float32", "tcgen05 LSE (m, s) vs CUDA cores" + tag,
                      scale=_ms_scale(o_cc.cpu().numpy()))
    # forced column chunks: the (m, s) partials are merged by kb200_conv_combine_lse
    o_ch, ll3 = _run(f, len(args), args, nx, ny, dev, True, KB200_CONV_TC_CHUNK_MB=1)
    util.assert_close(o_ch.cpu().numpy(), o_cc.cpu().numpy(), "float32", "chunked tcgen05 LSE vs CUDA cores" + tag,
                      scale=_ms_scale(o_cc.cpu().numpy()))
    lse_tc = o_tc[:, 0] + torch.log(o_tc[:, 1])
    lse_cc = o_cc[:, 0] + torch.log(o_cc[:, 1])
    util.assert_close(lse_tc.cpu().numpy(), lse_cc.cpu().numpy(), "float32", "tcgen05 log-sum-exp vs CUDA cores" + tag,
                      scale=np.maximum(np.abs(lse_cc.cpu().numpy()), 1.0))
    from oracle import keops_oracle as ko
    rows = slice(0, 40)
    host = [t.cpu().numpy() for t in args]
    if tagI == 0:
        host[0] = host[0][rows]
    else:
        host[1] = host[1][rows]
    ref = ko.genred(f, [], host, "float32")
    util.assert_close(o_tc[rows].cpu().numpy(), ref, "float32", "tcgen05 LSE vs oracle" + tag, scale=_ms_scale(ref))


@pytest.mark.gpu
def test_cluster_multicast_variant_is_identical(cuda_device):
    """Opt-in 2-CTA clusters that TMA-multicast the streamed k-blocks (``KB200_CONV_TC_MC=1``): same MMAs, same
    epilogue, so the output must be bit-identical to the single-CTA kernel -- also with an odd number of row blocks."""
    import torch
    dev = cuda_device
    torch.manual_seed(9)
    nx, ny, D = 128 * 65 + 17, 3001, 64                 # 66 row blocks + a ragged one -> 67, padded to 68 CTAs
    x, y, b = torch.rand(nx, D, device=dev), torch.rand(ny, D, device=dev), torch.randn(ny, 1, device=dev)
    g = torch.tensor([0.1], device=dev)
    f = KS.replace("16", "64")
    o1, ll1 = _run(f, 4, (x, y, b, g), nx, ny, dev, True, KB200_CONV_TC_MC=1)
    o0, ll0 = _run(f, 4, (x, y, b, g), nx, ny, dev, True, KB200_CONV_TC_MC=0)
    assert ll1.get("engine") == ll0.get("engine") == "tcgen05-conv"
    assert torch.equal(o1, o0)


@pytest.mark.gpu
@pytest.mark.parametrize("D,var,nx,ny", [
    (16, "x", 9000, 7001),                              # the KernelSolve backward, ragged tiles, two weight chunks
    (8, "x", 8192, 4096), (21, "x", 8300, 5000),        # one weight chunk; widest operand rows
    (21, "y", 5000, 8300), (12, "y", 4000, 9000),       # opt-in d/dy (KB200_CONV_TC_GRAD_Y): reduction over i
])
def test_tensor_core_position_gradient_matches_cuda_core(D, var, nx, ny, cuda_device):
    """Gradient of the Gaussian sum w.r.t. x / y assembled from tensor-core sums against the CUDA-core gradient kernel
    (generated from the differentiated formula, pinned against the reference by the golden gradient fixtures): inside the
    fp32 parity bar element-wise, relative to sum_j |F_ij| of every component."""
    import torch
    torch.manual_seed(D + len(var))
    dev = cuda_device
    x, y = torch.rand(nx, D, device=dev), torch.rand(ny, D, device=dev)
    b, eta = torch.randn(ny, 1, device=dev), torch.randn(nx, 1, device=dev)
    g = torch.tensor([0.7 if D <= 21 else 8.0 / D], device=dev)
    F = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,1,1),0)" % (D, D)
    fwd, _ = _run(F, 4, (x, y, b, g), nx, ny, dev, True)
    v = "Var(0,%d,0)" % D if var == "x" else "Var(1,%d,1)" % D
    f = "Grad_WithSavedForward(%s,%s,Var(4,1,0),Var(5,1,0))" % (F, v)
    args = (x, y, b, g, eta, fwd)
    if var == "y":
        os.environ["KB200_CONV_TC_GRAD_Y"] = "1"
    try:
        o_tc, ll = _run(f, 6, args, nx, ny, dev, True)
    finally:
        os.environ.pop("KB200_CONV_TC_GRAD_Y", None)
    assert ll.get("engine") == "tcgen05-conv-grad", ll
    assert ll["weight_columns"] == (D if var == "x" else D + 1)      # the saved forward spares the scalar sum for d/dx
    old = os.environ.get("KB200_CONV_TC_GRAD")
    os.environ["KB200_CONV_TC_GRAD"] = "0"
    try:
        o_cc, ll2 = _run(f, 6, args, nx, ny, dev, True)
    finally:
        if old is None:
            os.environ.pop("KB200_CONV_TC_GRAD", None)
        else:
            os.environ["KB200_CONV_TC_GRAD"] = old
    assert ll2.get("engine") in (None, "cuda-core/chunked"), ll2
    assert o_tc.shape == o_cc.shape == ((nx if var == "x" else ny), D)
    scale = util.gpu_abs_scale(f, 6, "float32", dev, nx, ny, *args).cpu().numpy()
    util.assert_close(o_tc.cpu().numpy(), o_cc.cpu().numpy(), "float32",
                      "tcgen05 gradient d/d%s vs CUDA cores D=%d %dx%d" % (var, D, nx, ny), scale=scale)
