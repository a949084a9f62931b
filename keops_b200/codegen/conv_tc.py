"""Tensor-core path for Gaussian-kernel Sum reductions in medium / large dimension.

``Sum_Reduction(Exp(c * |x_i - y_j|^2) * b_j, tagI)`` -- the Gaussian convolution / kernel matrix-vector product
behind ``KernelSolve`` (``pykeops/pykeops/torch/operations.py:29-87``) -- costs the reference's generic
``GpuReduc1D`` kernel 2 D + 2 FP32 instructions and one ``ex2`` per pair
(``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:70-90``).  For ``D >= 5`` the squared distance *is* a dense
contraction, so here it runs as a tcgen05 GEMM on three-term bf16 splits of the fp32 coordinates (six
products: fp32-grade accuracy, see ``keops_b200/csrc/kb200_conv_tc.cu`` for the bound) and the CUDA cores only
do ``ex2`` + the weighted accumulation: the kernel moves from the FP32-pipe roofline to the MUFU one.

This module recognises the pattern, JIT-compiles the two kernels for ``(D, dim b)`` and drives them:
``prep`` (bf16 operand rows, squared norms, padded b) x 2 -> ``conv_tc`` (TMA + tcgen05 + TMEM + epilogue).
The scalar coefficient ``c`` (constants and ``Pm`` parameters) is evaluated on the device with a few
one-element torch ops.

Stated tolerance: the relative error of one term ``exp(c |x_i - y_j|^2)`` is about
``2^-22 |c| (max_i |x_i - m|^2 + max_j |y_j - m|^2)`` with ``m`` the mean of the streamed cloud (cancellation in
``|x|^2 + |y|^2 - 2 x.y``).  ``run`` evaluates that bound on the device and refuses the launch when it exceeds
``TOL`` (5e-6, half of the fp32 parity bar of 1e-5); the binder then runs the CUDA-core kernel, which forms
``x - y`` directly.  Measured at the BASELINE KernelSolve shape (D = 16, M = N = 1e6) the result agrees with the
CUDA-core kernel to 4e-6 of the largest output.  ``KB200_CONV_TC=0`` disables the path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

from ..formulas.ir import Node
from ..formulas.reductions import Reduction
from ..runtime import capi, plan

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
_SRC_PATH = os.path.join(_CSRC, "kb200_conv_tc.cu")
_COMMON_PATH = os.path.join(_CSRC, "kb200_tc_common.cuh")
VERSION = "conv-tc-14"
BM, BN, STAGES, HS = 256, 128, 4, 8
THREADS = 64 + 32 * 16       # TMA warp + MMA warp + 16 epilogue warps (CT_THREADS in the kernel)
MAX_DB = 8
# smallest dimension sent to the tensor cores.  Measured at M = N = 1e6 (tensor-core / CUDA-core kernel): D = 4: 253 / 319 ms,
# D = 5: 251 / 380 ms, D = 6: 251 / 447 ms, D = 7: 251 / 508 ms, D = 8: 251 / 575 ms; D = 3 stays on the CUDA cores (259 ms)
DMIN = 5
ATMEM = 1                  # kb200_conv_tc_hd: resident operand in TMEM (KB200_CONV_TC_ATMEM=0: in shared memory)
MULTICAST = 0              # kb200_conv_tc_hd: 2-CTA clusters multicasting the streamed k-blocks (KB200_CONV_TC_MC=1 enables).
                           # Correct (tests pass with it on) but no faster: D=64, M=N=1e6 717 ms without, 726 ms with -- the
                           # bound is the shared-memory port (operand reads of M=128 x N=128 MMAs + TMA fills), not L2
POLY = 0                   # experiment (see kb200_conv_tc.cu): of every 8 pairs, exp2 of this many on the FMA pipe
CHUNK_MB = 48              # bf16 operand rows of one column chunk (must stay L2-resident)
# stated tolerance of the path: bound on the relative error of one term exp(c |x_i - y_j|^2) (0 disables the guard)
TOL = 5e-6                 # default; the environment variable KB200_CONV_TC_TOL overrides it per call
_lock = threading.RLock()


# ---------------------------------------------------------------------------------------------
# pattern matching
# ---------------------------------------------------------------------------------------------
def _sqdist(node: Node) -> Optional[Tuple[Node, Node]]:
    """``(a, b)`` if node is ``Sum(Square(Subtract(Var a, Var b)))`` (what SqDist / SqNorm2(x-y) lower to)."""
    if node.op != "Sum" or len(node.args) != 1 or node.args[0].op != "Square":
        return None
    sub = node.args[0].args[0]
    if sub.op != "Subtract" or len(sub.args) != 2 or any(a.op != "Var" for a in sub.args):
        return None
    return sub.args[0], sub.args[1]


def _is_scalar_param_expr(node: Node) -> bool:
    """dimension-1 expression of constants and parameters (category 2 variables) only"""
    if node.dim != 1:
        return False
    if node.op == "Var":
        return node.params[2] == 2
    if node.op in ("IntCst", "RatCst", "FloatCst"):
        return True
    if node.op in ("Minus", "Mult", "Divide", "Inv", "Square", "Add", "Subtract"):
        return all(_is_scalar_param_expr(a) for a in node.args)
    return False


def _scaled_sqdist(node: Node):
    """``(coef_terms, (a, b))`` when node = (product of scalar parameter expressions, possibly negated / divided)
    times a squared distance between two variables.  ``coef_terms`` is a list of ``(node, power)`` with power +-1
    and a leading sign as ``("sign", s)``."""
    sd = _sqdist(node)
    if sd is not None:
        return [], sd
    if node.op == "Minus":
        r = _scaled_sqdist(node.args[0])
        return None if r is None else (r[0] + [(None, -1)], r[1])
    if node.op == "Mult":
        a, b = node.args
        for s, o in ((a, b), (b, a)):
            if _is_scalar_param_expr(s):
                r = _scaled_sqdist(o)
                if r is not None:
                    return r[0] + [(s, 1)], r[1]
        return None
    if node.op == "Divide":
        a, b = node.args
        if _is_scalar_param_expr(b):
            r = _scaled_sqdist(a)
            if r is not None:
                return r[0] + [(b, -1)], r[1]
        return None
    return None


def match(red: Reduction):
    """``(row_var, col_var, D, b_var, DB, coef_terms)`` for ``Sum_Reduction(Exp(c |x - y|^2) * b, tagI)`` with x the
    output-side variable, y and b streamed variables; ``None`` otherwise."""
    if red.kind != "Sum":
        return None
    f = red.formula
    if f.op != "Mult" or len(f.args) != 2:
        return None
    for e, b in ((f.args[0], f.args[1]), (f.args[1], f.args[0])):
        if e.op != "Exp" or b.op != "Var":
            continue
        r = _scaled_sqdist(e.args[0])
        if r is None:
            continue
        terms, (va, vb) = r
        (ia, da, ca), (ib, db, cb) = va.params, vb.params
        if da != db or {ca, cb} != {0, 1}:
            continue
        row, col = (ia, ib) if ca == red.tagI else (ib, ia)
        bi, bdim, bcat = b.params
        if bcat != 1 - red.tagI or bdim > MAX_DB or e.dim != 1:
            continue
        return row, col, da, bi, bdim, terms
    return None


def match_lse(red: Reduction):
    """``(row_var, col_var, D, a_var or None, weight_var or None, DB, coef_terms)`` for
    ``Max_SumShiftExp[Weight]_Reduction(a_j + c |x_i - y_j|^2, tagI[, b_j])`` -- the soft-min of a Sinkhorn iteration
    (BASELINE configs[3] in dimension D >= 5): ``a`` an optional streamed scalar variable (added or the minuend of a
    subtraction), ``b`` absent (weight 1) or a streamed variable; ``None`` otherwise."""
    if red.kind != "LSE" or red.formula.dim != 1:
        return None
    f = red.formula
    cands = [(None, f, False)]
    if f.op == "Add" and len(f.args) == 2:
        cands += [(f.args[0], f.args[1], False), (f.args[1], f.args[0], False)]
    if f.op == "Subtract" and len(f.args) == 2:
        cands += [(f.args[0], f.args[1], True)]                    # a - E: the coefficient of E changes sign
    for a, e, negate in cands:
        if a is not None and (a.op != "Var" or a.params[1] != 1 or a.params[2] != 1 - red.tagI):
            continue
        r = _scaled_sqdist(e)
        if r is None:
            continue
        terms, (va, vb) = r
        if negate:
            terms = terms + [(None, -1)]
        (ia, da, ca), (ib, db, cb) = va.params, vb.params
        if da != db or {ca, cb} != {0, 1}:
            continue
        row, col = (ia, ib) if ca == red.tagI else (ib, ia)
        w = red.weight
        if w is None or (w.op == "IntCst" and w.params[0] == 1):
            wvar, DB = None, 1
        elif w.op == "Var" and w.params[2] == 1 - red.tagI and w.params[1] <= MAX_DB:
            wvar, DB = w.params[0], w.params[1]
        else:
            continue
        return row, col, da, (a.params[0] if a is not None else None), wvar, DB, terms
    return None


# dimensions in which the gradient is assembled from tensor-core sums (see match_grad).  Measured on B200 (experiments/
# conv_grad_check.py, M = N = 1e6): D = 16: d/dx 1321 ms in two 8-column launches vs 1579 ms for the CUDA-core gradient kernel
# (x 1.20), d/dy (17 columns, three launches) 1582 vs 1574 ms; D = 8: d/dx 656 vs 810 ms (x 1.24), d/dy 903 vs 806 ms;
# D = 32, 5e5 points (high-dimensional kernel): 718 vs 863 ms and 822 vs 861 ms.  A launch with 8 weight columns costs 2.5 x a
# launch with one (660 vs 260 ms at D = 16): per pair the epilogue issues 8 scalar FFMA and 2 broadcast LDS.128 wavefronts
# that every one of the 16 epilogue warps repeats -- LSU- and issue-bound, not FMA-bound (9 of 128 lanes per clock).  So
# only d/dx (whose scalar sum is the saved forward) is routed here; d/dy stays on the CUDA cores.
GRAD_DMIN, GRAD_DMAX = 8, 21


def grad_chunk(D: int):
    """``(weight columns per launch, ring stages)`` of the gradient's forward-type sums: the widest chunk that fits shared
    memory, giving up one ring stage before narrowing the chunk (fewer launches win: each one redoes the GEMM)."""
    for db in range(MAX_DB, 0, -1):
        for st in (STAGES, STAGES - 1):
            if smem_bytes(D, db, st) <= 227 * 1024:
                return db, st
    return None


def match_grad(red: Reduction):
    """Gradient of a Gaussian-kernel sum with respect to a position variable, recognised STRUCTURALLY from the
    provenance that ``Grad_WithSavedForward`` records (``Reduction.origin``) -- the differentiated formula itself is a
    D-vector expression that the pattern of :func:`match` does not cover.

    For ``out_i = sum_j exp(c |x_i - y_j|^2) b_j`` (dim b = 1) and upstream gradient ``e_i``
    (``pykeops/pykeops/torch/generic/generic_red.py:190-214`` requests exactly this reduction in every backward pass,
    ``pykeops/pykeops/torch/operations.py:107-147`` for ``KernelSolve``):

    * ``d/dx_i = 2 c e_i [ (x_i - m) S0_i - S1_i ]``, ``S0_i = sum_j k_ij b_j`` (the saved forward) and
      ``S1_i = sum_j k_ij b_j (y_j - m)`` -- a forward-type sum with a D-wide weight;
    * ``d/dy_j = 2 c b_j [ (y_j - m') T0_j - T1_j ]``, ``T0_j = sum_i k_ij e_i``, ``T1_j = sum_i k_ij e_i (x_i - m')``;

    ``m`` / ``m'`` = mean of the streamed cloud (the shift the tensor-core kernel applies anyway).  The sums run on
    ``kb200_conv_tc*`` in chunks of ``MAX_DB`` weight columns instead of 2 D + 2 FP32 instructions per pair on the CUDA cores.
    Cancellation: both terms of a bracket are at most ``R sum_j |k_ij b_j|`` (R = radius of the shifted clouds) while the
    gradient's own scale is ``delta sum_j |k_ij b_j|`` with ``delta ~ |c|^-1/2`` the distance of the pairs that matter; the
    accuracy guard of the sums bounds ``|c| R^2 <= 2^22 TOL ~ 21``, i.e. at most ~4.6 x the sums' own relative error
    (measured 1e-7 .. 1e-6) -- inside the fp32 parity bar; `tests/test_conv_tc.py` pins it against the CUDA-core kernel.

    Returns ``(kind, row_var, col_var, D, b_var, eta_var, fwd_var or None, coef_terms)`` with kind "x" (the output-side
    position of the forward) or "y" (the streamed one), or ``None``."""
    o = getattr(red, "origin", None)
    if not o or o[0] != "grad" or red.kind != "Sum":
        return None
    _, inner, v, gradin, f0 = o
    m = match(inner)
    if m is None:
        return None
    row, col, D, bi, DB, terms = m
    if DB != 1 or inner.dim != 1 or not (GRAD_DMIN <= D <= GRAD_DMAX):
        return None
    if v.op != "Var" or gradin.op != "Var" or gradin.params[1] != 1 or gradin.params[2] != inner.tagI:
        return None
    fwd = None
    if f0 is not None and f0.op == "Var" and f0.params[1] == 1 and f0.params[2] == inner.tagI:
        fwd = f0.params[0]
    kind = "x" if v.params[0] == row else ("y" if v.params[0] == col else None)
    if kind is None or v.params[1] != D:
        return None
    if kind == "y" and os.environ.get("KB200_CONV_TC_GRAD_Y", "0") != "1":
        return None                      # measured no faster than the CUDA-core kernel (see GRAD_DMIN above)
    return kind, row, col, D, bi, gradin.params[0], fwd, terms


def _eval_scalar(node: Node, args):
    """one-element fp32 tensor (or python float) value of a scalar parameter expression -- device ops only"""
    if node.op == "Var":
        return args[node.params[0]].reshape(-1)[:1]
    if node.op == "IntCst":
        return float(node.params[0])
    if node.op == "RatCst":
        return float(node.params[0]) / float(node.params[1])
    if node.op == "FloatCst":
        return float(node.params[0])
    v = [_eval_scalar(a, args) for a in node.args]
    if node.op == "Minus":
        return -v[0]
    if node.op == "Mult":
        return v[0] * v[1]
    if node.op == "Divide":
        return v[0] / v[1]
    if node.op == "Inv":
        return 1.0 / v[0]
    if node.op == "Square":
        return v[0] * v[0]
    if node.op == "Add":
        return v[0] + v[1]
    if node.op == "Subtract":
        return v[0] - v[1]
    raise ValueError(node.op)


def coefficient(terms, args, device):
    """the scalar c of ``exp(c |x - y|^2)`` as a one-element fp32 CUDA tensor"""
    import torch
    c = 1.0
    for node, power in terms:
        if node is None:
            c = -c
        elif power == 1:
            c = c * _eval_scalar(node, args)
        else:
            c = c / _eval_scalar(node, args)
    if not isinstance(c, torch.Tensor):
        c = torch.full((1,), float(c), dtype=torch.float32, device=device)
    return c.to(torch.float32).contiguous()


def eligible(red: Reduction, dtype: str, nrow: int, ncol: int, nbatch: int, use_fast_math: bool,
             sum_scheme: str = "block_sum", dtype_acc: str = "float32"):
    """The tensor-core path pays off once 2 D + 2 FP32 instructions per pair exceed the MUFU bound (measured: from D = 4, used from D = 5) and the
    grid fills the machine with 256-row CTAs; it reproduces the default (block_sum, fp32 accumulation) scheme."""
    if os.environ.get("KB200_CONV_TC", "1") == "0" or dtype != "float32" or nbatch != 1 or not use_fast_math:
        return None
    if sum_scheme not in ("block_sum", "auto") or dtype_acc not in ("float32", "auto"):
        return None
    m = match(red)
    if m is None:
        return None
    D = m[2]
    dmin = int(os.environ.get("KB200_CONV_TC_DMIN", DMIN))
    kbm, rem, _ = layout(D)
    min_rows = 8192 if ncol < 65536 else 1024       # few rows, many columns: column chunks fill the machine
    if D < dmin or D > HD_MAX_D or nrow < min_rows or ncol < 1024 or ncol >= (1 << 31) - BN:
        return None
    if smem_bytes(D, m[4]) > 227 * 1024:
        return None
    return m


def _blocks(total: int) -> Tuple[int, int]:
    """operand rows of ``total`` bf16 columns as KBM 64-column blocks (128-byte swizzle) + one remainder block of
    16 / 32 / 64 columns (swizzled at its own width)"""
    full, r = divmod(total, 64)
    if r == 0:
        return full - 1, 64
    return full, (16 if r <= 16 else (32 if r <= 32 else 64))


HD_MAX_KB = 9              # resident operand rows of the high-dimensional kernel: 9 x 16 KB (D <= 95); beyond, both sides stream
HD_MAX_D = 2048            # operand rows of 6 D + 3 bf16 columns: 24 KB per point at D = 2048


def is_hd(D: int) -> bool:
    """K = 6 D + 3 > 128 bf16 columns (D >= 21 with the penalty columns, D >= 22 always): ``kb200_conv_tc_hd`` -- one
    128-row block per CTA resident for all of K, the streamed side in 64-column k-blocks accumulated in TMEM."""
    return 6 * D > 128


def layout(D: int) -> Tuple[int, int, int]:
    """``(KBM, REM, PEN)``: the six product blocks (6 D columns) plus the three -|y|^2/2 columns (PEN = 1) whenever
    they fit K <= 128; otherwise (PEN = 0, D = 21) |y|^2 is added in the epilogue.  Measured on B200 at D = 16,
    M = N = 1e6: K = 128 with the penalty columns 277 ms, K = 96 with the epilogue FFMA + wider b tile 294 ms -- the
    epilogue's issue slots and LDS bandwidth are scarcer than tensor-core time."""
    with_pen, without = _blocks(6 * D + 3), _blocks(6 * D)
    width = lambda kb: 64 * kb[0] + kb[1]
    if is_hd(D):                          # high-dimensional kernel: uniform 64-column k-blocks, penalty inside K
        return -(-(6 * D + 3) // 64) - 1, 64, 1
    if os.environ.get("KB200_CONV_TC_PEN") in ("0", "1"):
        pen = int(os.environ["KB200_CONV_TC_PEN"])
        return (with_pen if pen else without) + (pen,)
    if width(with_pen) > 128 and width(without) <= 128:
        return without + (0,)
    return with_pen + (1,)


def smem_bytes(D: int, DB: int, stages: int = STAGES, lse: bool = False) -> int:
    """dynamic shared memory of kb200_conv_tc / kb200_conv_tc_hd (CT_SMEM_BYTES / HD_SMEM_BYTES in the kernel)"""
    kbm, rem, pen = layout(D)
    OW = DB + (1 if lse else 0)
    if is_hd(D):                         # HD_SMEM_BYTES
        if kbm + 1 > HD_MAX_KB:          # CT_ASTREAM: both operands go through the ring
            return 1024 + stages * 32768 + HS * BN * DB * 4 + 3 * 128 * OW * 4 + 640
        return 1024 + (kbm + 1) * 16384 + stages * 16384 + HS * BN * DB * 4 + 3 * 128 * OW * 4 + 640
    blk = lambda rows: kbm * rows * 128 + rows * rem * 2       # one operand block, all of K
    return 1024 + 2 * blk(128) + stages * blk(BN) + HS * BN * (DB + 1 - pen) * 4 + BM * OW * 4 + 640


def eligible_lse(red: Reduction, dtype: str, nrow: int, ncol: int, nbatch: int, use_fast_math: bool,
                 sum_scheme: str = "block_sum", dtype_acc: str = "float32"):
    """Same conditions as :func:`eligible` for the LogSumExp family; the column term rides in the penalty columns,
    which the D = 21 layout does not have."""
    if os.environ.get("KB200_CONV_TC", "1") == "0" or dtype != "float32" or nbatch != 1 or not use_fast_math:
        return None
    if sum_scheme not in ("block_sum", "auto") or dtype_acc not in ("float32", "auto"):
        return None
    m = match_lse(red)
    if m is None:
        return None
    D, DB = m[2], m[5]
    dmin = int(os.environ.get("KB200_CONV_TC_DMIN", DMIN))
    min_rows = 8192 if ncol < 65536 else 1024
    if D < dmin or D > HD_MAX_D or layout(D)[2] != 1 or nrow < min_rows or ncol < 1024 or ncol >= (1 << 31) - BN:
        return None
    if smem_bytes(D, DB, lse=True) > 227 * 1024:
        return None
    return m


class ConvTensorCore:
    """Compiled kernels for one (D, dim b, Sum | LSE) combination on one device."""

    def __init__(self, D: int, DB: int, device: int, lse: bool = False, stages: Optional[int] = None):
        self.D, self.DB, self.lse = D, DB, bool(lse)
        self.OUTW = DB + (1 if lse else 0)
        self.KBM, self.REM, self.PEN = layout(D)
        self.hd = is_hd(D)
        # CTA pairs sharing the streamed k-blocks by TMA multicast (high-dimensional kernel only)
        self.mc = 1 if (self.hd and os.environ.get("KB200_CONV_TC_MC", str(MULTICAST)) != "0") else 0
        # resident operand in tensor memory: 32 columns per k-block next to the 2 x 128 accumulator columns (K <= 512)
        self.atmem = 1 if (self.hd and self.KBM + 1 <= 8 and os.environ.get("KB200_CONV_TC_ATMEM", str(ATMEM)) != "0") else 0
        # 256-row kernel: the same with single-buffered accumulators, when every k-block is 64 columns wide (8 <= D <= 20)
        self.atmem2 = 1 if (not self.hd and self.REM == 64 and self.PEN == 1
                            and os.environ.get("KB200_CONV_TC_ATMEM", str(ATMEM)) != "0") else 0
        self.DPG = 64 * self.KBM + self.REM
        self.W = DB + 1 - self.PEN                   # floats per column of the b tile
        with open(_COMMON_PATH) as f:
            body = f.read()
        with open(_SRC_PATH) as f:
            body += f.read()
        self.STAGES = int(os.environ.get("CT_STAGES", stages or STAGES))
        defs = ("#define CT_D %d\n#define CT_DB %d\n#define CT_KBM %d\n#define CT_REM %d\n#define CT_STAGES %d\n"
                "#define CT_HS %d\n#define CT_DBG %d\n#define CT_PEN %d\n#define CT_POLY %d\n#define CT_ASTREAM %d\n"
                "#define CT_LSE %d\n#define CT_MC %d\n#define CT_ATMEM %d\n#define CT_ATMEM2 %d\n"
                % (D, DB, self.KBM, self.REM, self.STAGES, HS, int(os.environ.get("CT_DBG", 0)), self.PEN,
                   0 if lse else int(os.environ.get("CT_POLY", POLY)),
                   1 if (self.hd and self.KBM + 1 > HD_MAX_KB) else 0, 1 if lse else 0, self.mc, self.atmem, self.atmem2))
        src = defs + body
        self.cubin = cubin = plan.compile_cached(src, plan.nvrtc_options(True), "kb200_convtc", VERSION)
        self.smem = smem_bytes(D, DB, self.STAGES, lse)
        if self.smem > 227 * 1024:
            raise ValueError("conv_tc: D = %d, dim b = %d needs %d bytes of shared memory" % (D, DB, self.smem))
        self.smem = max(self.smem, 120 * 1024)     # one CTA per SM: each CTA allocates all 512 TMEM columns
        self.module = None
        self.last_bound = None
        if device >= 0:
            self.module = capi.Module(cubin, device)
            self.fn_prep = self.module.function("kb200_conv_prep")
            self.fn_conv = self.module.function("kb200_conv_tc_hd" if self.hd else "kb200_conv_tc")
            self.fn_combine = self.module.function("kb200_conv_combine_lse" if lse else "kb200_conv_combine")
            self.sm_count = capi.device_info(device)["sm_count"]
            capi.set_max_dynamic_smem(self.fn_conv, self.smem)

    def _maps(self, abf, n, box_rows):
        pitch = self.DPG * 2
        rem = capi.tensor_map_2d(abf.data_ptr() + 128 * self.KBM, 2, self.REM, n, pitch, self.REM, box_rows, self.REM * 2)
        if self.KBM == 0:
            return rem, rem            # the main map is never dereferenced
        return capi.tensor_map_2d(abf.data_ptr(), 2, 64 * self.KBM, n, pitch, 64, box_rows, 128), rem

    def run(self, x, y, b, coef, out, stream_obj, clock_probe=None, acol=None):
        """x (M, D), y (N, D), b (N, DB) fp32 contiguous CUDA tensors, coef a one-element fp32 tensor;
        out (M, DB) fp32.  Returns False (nothing written) when the accuracy guard rejects the inputs.
        ``clock_probe``: optional int64 CUDA tensor of 2 elements receiving (SM cycles, nanoseconds) of CTA 0.
        LSE engines: ``b`` = weights (ones for the plain LogSumExp), ``acol`` = the column term a_j (N,) or None,
        ``out`` (M, 1 + DB) receives (m_i, s_i)."""
        import torch
        M, N = x.shape[0], y.shape[0]
        dev = x.device
        stream = stream_obj.cuda_stream
        i64, ptr, i32 = C.c_longlong, C.c_void_p, C.c_int
        ntiles = -(-N // BN)
        npad = ntiles * BN
        # the bf16 operand copies take 2 * CT_DPG bytes per point (24 KB at D = 2048): leave the job to the CUDA-core
        # kernel rather than exhaust the device
        free_bytes, _ = torch.cuda.mem_get_info(dev)
        if (M + N) * self.DPG * 2 + (M + 2 * npad) * 4 * (self.W + 2) > 0.8 * free_bytes:
            self.last_bound = None
            return False
        shift = y.mean(dim=0, dtype=torch.float32).contiguous()            # common shift: x - y is unchanged
        ybf = torch.empty((N, self.DPG), dtype=torch.int16, device=dev)
        xbf = torch.empty((M, self.DPG), dtype=torch.int16, device=dev)
        ysq = torch.empty(N, dtype=torch.float32, device=dev)
        xsq = torch.empty(M, dtype=torch.float32, device=dev)
        bpad = torch.empty(npad * self.W, dtype=torch.float32, device=dev)
        capi.launch(self.fn_prep, stream, (-(-npad // 8),), (256,), 0,            # one warp per row
                    [i64(N), i64(npad), ptr(y.data_ptr()), ptr(shift.data_ptr()), ptr(ybf.data_ptr()), ptr(ysq.data_ptr()),
                     ptr(b.data_ptr()), ptr(bpad.data_ptr()), i32(0), ptr(acol.data_ptr() if acol is not None else 0),
                     ptr(coef.data_ptr())])
        capi.launch(self.fn_prep, stream, (-(-M // 8),), (256,), 0,
                    [i64(M), i64(M), ptr(x.data_ptr()), ptr(shift.data_ptr()), ptr(xbf.data_ptr()), ptr(xsq.data_ptr()),
                     ptr(0), ptr(0), i32(1), ptr(0), ptr(coef.data_ptr())])
        tol = float(os.environ.get("KB200_CONV_TC_TOL", TOL))
        self.last_bound = None
        if tol > 0.0:
            # accuracy guard: the exponent carries an absolute error ~ 2^-22 |c| log2(e) (max |x-m|^2 + max |y-m|^2); when
            # that exceeds the stated tolerance (clouds much wider than the kernel bandwidth) the caller falls back to
            # the CUDA-core kernel, which forms x - y directly.  One host read, like the reference's per-call sync.
            bound = coef.abs() * (xsq.max() + ysq.max()) * (2.0 ** -22)
            if acol is not None:                   # the column term shares the accumulation: + 2^-22 * 2 max |a_j|
                bound = bound + acol.abs().max() * (2.0 ** -21)
            self.last_bound, c_abs = torch.stack([bound.reshape(()), coef.abs().reshape(())]).tolist()
            if not self.last_bound <= tol:          # also catches NaN / inf inputs
                return False
            if self.lse and acol is not None and not c_abs > 0.0:
                return False                        # a_j / (2 c) in the penalty columns needs c != 0
        if self.hd:
            pitch = self.DPG * 2
            map_y = capi.tensor_map_2d(ybf.data_ptr(), 2, self.DPG, N, pitch, 64, BN // 2 if self.mc else BN, 128)
            map_x = capi.tensor_map_2d(xbf.data_ptr(), 2, self.DPG, M, pitch, 64, 128, 128)
        else:
            map_y, map_yr = self._maps(ybf, N, BN)
            map_x, map_xr = self._maps(xbf, M, 128)
        # column chunks: every CTA streams all operand rows of its chunk, so a chunk must stay L2-resident
        # (126 MB L2; measured at M = N = 1e6, D = 16: one 256 MB chunk 288 ms, L2-sized chunks %s)
        chunk_bytes = int(os.environ.get("KB200_CONV_TC_CHUNK_MB", CHUNK_MB)) << 20
        tiles_per_chunk = max(1, chunk_bytes // (BN * self.DPG * 2))
        nchunks = -(-ntiles // tiles_per_chunk)
        nblocks = -(-M // (128 if self.hd else BM))
        if nblocks < self.sm_count:
            # few output rows against many columns: the column chunks also serve to fill the machine
            nchunks = max(nchunks, min(self.sm_count // nblocks, max(1, ntiles // 16)))
        tiles_per_chunk = -(-ntiles // nchunks)                   # balanced
        nchunks = -(-ntiles // tiles_per_chunk)
        dst = out if nchunks == 1 else torch.empty((nchunks, M, self.OUTW), dtype=torch.float32, device=dev)
        if self.hd:
            gx = -(-M // 128)
            if self.mc:
                gx += gx & 1               # whole CTA pairs; the odd block out computes rows >= M and stores nothing
            capi.launch(self.fn_conv, stream, (gx, nchunks), (THREADS,), self.smem,
                        [map_x, map_y, i64(M), i64(N), i64(tiles_per_chunk * BN), ptr(bpad.data_ptr()),
                         ptr(xsq.data_ptr()), ptr(coef.data_ptr()), ptr(dst.data_ptr()), ptr(xbf.data_ptr())])
        else:
            capi.launch(self.fn_conv, stream, (-(-M // BM), nchunks), (THREADS,), self.smem,
                        [map_x, map_xr, map_y, map_yr, i64(M), i64(N), i64(tiles_per_chunk * BN), ptr(bpad.data_ptr()),
                         ptr(xsq.data_ptr()), ptr(coef.data_ptr()), ptr(dst.data_ptr()),
                         ptr(clock_probe.data_ptr() if clock_probe is not None else 0), ptr(xbf.data_ptr())])
        if nchunks > 1:
            nel = M if self.lse else M * self.DB          # LSE: one thread per row; Sum: one per output element
            capi.launch(self.fn_combine, stream, (-(-nel // 256),), (256,), 0,
                        [i64(nel), i32(nchunks), ptr(dst.data_ptr()), ptr(out.data_ptr())])
        for t in (shift, ybf, xbf, ysq, xsq, bpad, coef, dst):
            t.record_stream(stream_obj)
        self.last = dict(ntiles=ntiles, ctas=-(-M // (128 if self.hd else BM)) * nchunks, nchunks=nchunks)
        return True


_cache: Dict[tuple, ConvTensorCore] = {}


def get(D: int, DB: int, device: int, lse: bool = False, stages: Optional[int] = None) -> ConvTensorCore:
    key = (D, DB, device, bool(lse), os.environ.get("KB200_CONV_TC_MC", str(MULTICAST)), os.environ.get("CT_POLY", str(POLY)),
           os.environ.get("KB200_CONV_TC_ATMEM", str(ATMEM)), stages)
    with _lock:
        k = _cache.get(key)
        if k is None:
            k = ConvTensorCore(D, DB, device, lse, stages)
            _cache[key] = k
        return k
