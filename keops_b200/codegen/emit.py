"""CUDA source emitter: reduction descriptor -> one sm_100a translation unit.

Replaces the reference's kernel skeletons ``GpuReduc1D`` / ``GpuReduc2D``
(``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:47-95``, ``GpuReduc2D.py:85-199``).  The design is
different:

* a thread owns ``R`` output rows held as packed f32x2 register pairs, so one FADD2/FMUL2/FFMA2
  issue slot serves two (i,j) pairs and one shared-memory read of ``y_j`` serves ``R`` pairs;
* the streamed operand is staged with ``cp.async`` into a double-buffered tile (one barrier per
  tile instead of two), in one padded row per ``j`` so it is read back with broadcast ``LDS.128``;
* everything that does not depend on both indices is hoisted out of the pair loop;
* the column range can be split over ``gridDim.y`` (bounded temporary, merged by a second tiny
  kernel) so that the grid can be shaped to the 148 SMs whatever ``M`` is.
"""
from __future__ import annotations

import hashlib
import sys
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from ..formulas.ir import FormulaError
from ..formulas.reductions import Reduction
from .lower import Lowerer
from .scalar import Graph, SNode, SV, C, P, I, J, IJ, topo_order

GENERATOR_VERSION = "kb200-gen-5"
MAX_SMEM_BYTES = 227 * 1024        # per-CTA shared memory on sm_100a (opt-in maximum)

_PRELUDE_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "kb200_device.cuh")
_prelude_cache: Optional[str] = None


def prelude() -> str:
    global _prelude_cache
    if _prelude_cache is None:
        with open(_PRELUDE_PATH) as f:
            _prelude_cache = f.read()
    return _prelude_cache


@dataclass(frozen=True)
class KernelConfig:
    T: int = 256        # threads per CTA
    R: int = 4          # output rows per thread
    TJ: int = 256       # streamed rows per shared-memory tile
    unroll: int = 8     # unroll factor of the pair loop
    min_blocks: int = 1  # __launch_bounds__ second argument
    group: int = 1      # Min / KMin families: streamed rows evaluated branch-free before one threshold test

    def key(self) -> str:
        k = "T%d_R%d_TJ%d_U%d_B%d" % (self.T, self.R, self.TJ, self.unroll, self.min_blocks)
        return k if self.group == 1 else k + "_G%d" % self.group


@dataclass
class KernelSource:
    source: str
    kernel_name: str
    combine_name: str
    cfg: KernelConfig
    dtype: str
    tagI: int
    dim: int               # output reals per row
    dimred: int            # accumulator size reported to callers (reference convention)
    dimpart: int           # reals per row in a split-j partial
    nargs: int
    row_vars: List[Tuple[int, int]]   # (arg index, dim) of output-side variables
    col_vars: List[Tuple[int, int]]   # streamed variables
    par_vars: List[Tuple[int, int]]
    stride: int            # reals per staged row
    smem_bytes: int
    is_zero: bool = False
    hash: str = ""
    dynamic_smem: int = 0  # > 0: the staged tile is `extern __shared__` (beyond the 48 KB static limit), launch with this many bytes


# ------------------------------------------------------------------------------------------
# fusion / sign-resolution pass on the scalar DAG
# ------------------------------------------------------------------------------------------
class _Rewriter:
    """Contracts add/sub of products into fma and resolves pending negations where they are free."""

    def __init__(self, g: Graph):
        self.g = g
        self.memo: Dict[int, SNode] = {}

    def vmode(self, n: SNode) -> bool:
        return n.cls in (I, IJ)

    def cheap_neg(self, operand: SNode, user_cls: int, user_v: bool) -> bool:
        if not user_v:
            return True  # scalar SASS has negate modifiers on every FP operand
        return operand.op == "const" or operand.cls < user_cls

    def neg_of(self, n: SNode) -> SNode:
        """Rewritten node for -(n), n being an *original* node."""
        g = self.g
        if n.op == "const":
            return g._mk("const", (), -n.val, C)
        if n.op == "neg":
            return self.rw(n.args[0])
        if n.op == "sub":
            p, q = n.args
            return self._sub(q, p, n.cls)
        v = self.vmode(n)
        if n.op == "mul":
            a, b = n.args
            for x, y in ((a, b), (b, a)):
                if x.op == "const" or (x.cls < n.cls) or not v:
                    return g._mk("mul", (self.neg_of(x), self.rw(y)), None, n.cls)
        return g._mk("neg", (self.rw(n),), None, n.cls)

    def _add(self, x: SNode, y: SNode, cls: int) -> SNode:
        g = self.g
        cands = [a for a in (x, y) if a.op == "mul"]
        if cands:
            m = max(cands, key=lambda a: a.uid)
            other = y if m is x else x
            return g._mk("fma", (self.rw(m.args[0]), self.rw(m.args[1]), self.rw(other)), None, cls)
        return g._mk("add", (self.rw(x), self.rw(y)), None, cls)

    def _sub(self, p: SNode, q: SNode, cls: int) -> SNode:
        """p - q with original nodes p, q."""
        g = self.g
        v = cls in (I, IJ)
        if q.op == "mul":
            c, d = q.args
            for x, y in ((c, d), (d, c)):
                if self.cheap_neg(x, cls, v):
                    return g._mk("fma", (self.neg_of(x), self.rw(y), self.rw(p)), None, cls)
        if p.op == "mul" and self.cheap_neg(q, cls, v):
            return g._mk("fma", (self.rw(p.args[0]), self.rw(p.args[1]), self.neg_of(q)), None, cls)
        return g._mk("sub", (self.rw(p), self.rw(q)), None, cls)

    def rw(self, n: SNode) -> SNode:
        out = self.memo.get(n.uid)
        if out is not None:
            return out
        g = self.g
        if n.op in ("const", "load"):
            out = n
        elif n.op == "add":
            out = self._add(n.args[0], n.args[1], n.cls)
        elif n.op == "sub":
            out = self._sub(n.args[0], n.args[1], n.cls)
        elif n.op == "neg":
            out = self.neg_of(n.args[0])
        elif n.op == "ex2" and n.args[0].op == "neg" and n.args[0].args[0].op in ("add", "fma") and self.g.fast:
            # 2^(-s) of a sum: the negation is a free source modifier of MUFU.EX2 (pre-scaled squared distances)
            out = g._mk("ex2n", (self.rw(n.args[0].args[0]),), None, n.cls)
        else:
            out = g._mk(n.op, tuple(self.rw(a) for a in n.args), n.val, n.cls)
        self.memo[n.uid] = out
        return out

    def signed(self, sv: SV) -> SNode:
        return self.neg_of(sv.n) if sv.neg else self.rw(sv.n)


# ------------------------------------------------------------------------------------------
# expression printing
# ------------------------------------------------------------------------------------------
def _lit(v: float, f64: bool) -> str:
    if v != v:
        return "(real)NAN"
    if v in (float("inf"), float("-inf")):
        return ("-" if v < 0 else "") + "KB_INF"
    if f64:
        return repr(float(v))
    s = "%.9g" % v
    if "." not in s and "e" not in s and "inf" not in s:
        s += ".0"
    return s + "f"


_FN1 = {"rcp", "exp", "ex2", "ex2n", "sqrt_rn", "keep", "log", "lg2", "sqrt", "rsqrt", "sin", "cos", "asin", "acos", "atan", "abs",
        "sign", "step", "relu", "floor", "round0", "xlogx", "sinxdivx"}
_FN2 = {"div", "atan2", "powf", "fmax", "fmin", "lt", "le", "eq", "ne"}
_FN3 = {"sel_ge0", "clamp", "inrange", "mod"}


class _Printer:
    def __init__(self, f64: bool):
        self.f64 = f64
        self.names: Dict[int, str] = {}

    def ref(self, n: SNode) -> str:
        if n.op == "const":
            return _lit(n.val, self.f64)
        return self.names[n.uid]

    def expr(self, n: SNode) -> str:
        v = n.cls in (I, IJ)
        sfx = "_v" if v else ""
        a = [self.ref(x) for x in n.args]
        op = n.op
        if op in ("add", "sub", "mul"):
            return "kb_%s%s(%s, %s)" % (op, sfx, a[0], a[1])
        if op == "fma":
            return "kb_fma%s(%s, %s, %s)" % (sfx, a[0], a[1], a[2])
        if op == "neg":
            return "kb_neg%s(%s)" % (sfx, a[0])
        if op in _FN1 or op in _FN2 or op in _FN3:
            return "kb_%s%s(%s)" % (op, sfx, ", ".join(a))
        raise NotImplementedError("emit: unknown scalar op " + op)


# ------------------------------------------------------------------------------------------
# the generator
# ------------------------------------------------------------------------------------------
def _collect_vars(red: Reduction):
    vs = {}
    for v in red.variables():
        ind, dim, cat = v.params
        if ind in vs and vs[ind] != (dim, cat):
            raise FormulaError("variable %d is used with inconsistent (dim, cat)" % ind)
        vs[ind] = (dim, cat)
    return vs


def generate(red: Reduction, dtype: str, nargs: int, cfg: KernelConfig, use_fast_math: bool = True,
             sum_scheme: str = "block_sum", dtype_acc: str = "auto", ranges: bool = False,
             prescale: bool = False) -> Optional[KernelSource]:
    """``ranges=True`` emits the block-sparse variant ``kb200_reduce_ranges`` (reference:
    ``keopscore/keopscore/mapreduce/gpu/GpuReduc1D_ranges.py:77-172``) instead of the dense pair
    ``kb200_reduce`` / ``kb200_combine``.

    ``prescale=True`` emits the *pre-scaled, packed* variant for Gaussian-type sums in low dimension
    (``_prescale_rewrite``): a third kernel ``kb200_pack`` writes all streamed variables of a column -- the distance
    coordinates shifted and scaled -- into one padded row of ``ypk``, which ``kb200_reduce`` stages with one 16-byte
    ``cp.async`` per 4 reals (instead of one 4-byte request per element from several arrays).  Returns None when the formula
    does not have that shape."""
    if dtype not in ("float32", "float64"):
        raise ValueError("dtype %s is not implemented (float32 / float64 only)" % dtype)
    # the scalar DAG of a D-dimensional left-to-right sum is D levels deep and the rewriters below are recursive:
    # make room for the high-dimensional formulas the reference handles in chunk mode (GpuReduc1D_chunks.py)
    maxdim = max([v.params[1] for v in red.variables()] + [1])
    need = 2000 + 40 * maxdim
    if sys.getrecursionlimit() < need:
        sys.setrecursionlimit(need)
    f64 = dtype == "float64"
    fast = bool(use_fast_math) and not f64
    tagI = red.tagI
    lpp = 1 if f64 else 2
    if cfg.R % lpp != 0:
        raise ValueError("R must be even for float32 kernels")

    varinfo = _collect_vars(red)
    for ind in varinfo:
        if ind >= nargs:
            raise FormulaError("formula uses Var(%d,...) but only %d arguments are declared" % (ind, nargs))
    row_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == tagI)
    col_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == 1 - tagI)
    par_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == 2)

    kind = red.kind
    base = dict(cfg=cfg, dtype=dtype, tagI=tagI, dim=red.dim, dimred=red.dimred, nargs=nargs,
                row_vars=row_vars, col_vars=col_vars, par_vars=par_vars)
    if kind == "Zero":
        return KernelSource(source="", kernel_name="", combine_name="", dimpart=0, stride=0, smem_bytes=0,
                            is_zero=True, **base)

    if prescale and (f64 or not fast or ranges or kind != "Sum"):
        return None
    low = lower_reduction(red, fast, prescale=prescale)
    if low is None:
        if prescale:
            return None
        return KernelSource(source="", kernel_name="", combine_name="", dimpart=0, stride=0, smem_bytes=0,
                            is_zero=True, **base)
    g, acc_terms, value_nodes, order = low.graph, low.acc_terms, low.value_nodes, low.order
    row_scales = low.row_scales

    pr = _Printer(f64)
    stride = sum(d for _, d in col_vars)
    stride = max(4, (stride + 3) // 4 * 4) if not f64 else max(2, (stride + 1) // 2 * 2)
    col_off: Dict[int, int] = {}
    off = 0
    for ind, d in col_vars:
        col_off[ind] = off
        off += d
    # packed mode: one slot per streamed value the formula reads -- (index, component, scaled copy?) -- in first-use order
    slots: List[Tuple[int, int, bool]] = []
    if prescale:
        for n in order:
            if n.op == "load" and n.cls == J:
                key = (n.val[0], n.val[1], n.val[2] == 7)
                if key not in slots:
                    slots.append(key)
        stride = max(4, (len(slots) + 3) // 4 * 4)

    # --- statements per region ----------------------------------------------------------------
    pre: List[str] = []      # before the loop: P scalars, I vectors
    body_j: List[str] = []   # per streamed row: J scalars
    body_ij: List[str] = []  # per pair
    used_col: Dict[Tuple[int, int], str] = {}
    used_pos: Dict[int, str] = {}     # packed mode: slot position -> name
    pre_p: List[str] = []             # the constant / parameter-only statements (also needed by kb200_pack)
    for n in order:
        if n.op == "const":
            continue
        if n.op == "load":
            ind, comp, cat = n.val
            if cat == 6:
                nm = "sh%d" % ind
                pre.append("const real %s = shf[%d];" % (nm, ind))
                pre_p.append(pre[-1])
            elif n.cls == P:
                nm = "p%d_%d" % (ind, comp)
                pre.append("const real %s = a%d[%d];" % (nm, ind, comp))
                pre_p.append(pre[-1])
            elif n.cls == I:
                nm = "x%d_%d" % (ind, comp)
                d = varinfo[ind][0]
                pre.append("V %s; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) %s.l[q] = a%d[rowc[q] * %d + %d];"
                           % (nm, nm, ind, d, comp))
            elif prescale:
                nm = ("ys%d_%d" if cat == 7 else "y%d_%d") % (ind, comp)
                used_pos[slots.index((ind, comp, cat == 7))] = nm
            else:
                nm = "y%d_%d" % (ind, comp)
                used_col[(ind, comp)] = nm
            pr.names[n.uid] = nm
            continue
        prefix = {P: "p", I: "h", J: "s", IJ: "t", C: "c"}[n.cls]
        nm = "%s%d" % (prefix, n.uid)
        pr.names[n.uid] = nm
        ty = "V" if n.cls in (I, IJ) else "real"
        stmt = "const %s %s = %s;" % (ty, nm, pr.expr(n))
        if n.cls in (C, P, I):
            pre.append(stmt)
            if n.cls in (C, P):
                pre_p.append(stmt)
        elif n.cls == J:
            body_j.append(stmt)
        else:
            body_ij.append(stmt)

    # loads of the staged row: vector loads of the padded row, then named components
    vec = 2 if f64 else 4
    vty = "double2" if f64 else "float4"
    comp_names = ["x", "y", "z", "w"]
    positions = dict(used_pos) if prescale else {col_off[ind] + comp: nm for (ind, comp), nm in used_col.items()}
    need_chunks = sorted({pos // vec for pos in positions})
    def row_loads_from(ptr: str) -> List[str]:
        out = ["const %s q%d = %s[%d];" % (vty, c, ptr, c) for c in need_chunks]
        for pos, nm in sorted(positions.items()):
            out.append("const real %s = q%d.%s;" % (nm, pos // vec, comp_names[pos % vec]))
        return out

    row_loads = row_loads_from("yrow")

    W = _Writer()
    hdr = _kernel_header(W, red, cfg, f64, fast, stride, nargs, varinfo)
    kname, cname = "kb200_reduce", "kb200_combine"

    red_code = _ReductionCode(red, pr, acc_terms, value_nodes, sum_scheme, dtype_acc, f64, fast, row_scales)

    def stage_tile(jbase: str, buf: str):
        if not prescale:
            _stage(W, col_vars, col_off, jbase, buf)
            return
        # packed rows: KB_STRIDE / 4 aligned 16-byte requests per streamed row, coalesced, no bank conflicts
        W("for (int e = tid; e < KB_TJ * (KB_STRIDE / 4); e += KB_T) {")
        W("  const int jr_ = e / (KB_STRIDE / 4), c_ = (e - jr_ * (KB_STRIDE / 4)) * 4;")
        W("  const long long j_ = (%s) + jr_; const bool ok_ = j_ < jend;" % jbase)
        W("  kb_cp_async16(&sm[%s][jr_ * KB_STRIDE + c_], ypk + (ok_ ? (j_ * KB_STRIDE + c_) : 0), ok_);" % buf)
        W("}")

    def tile_loop():
        """Stage + stream the column range [jbeg, jend) through the double-buffered tile."""
        W("// ---- stage the first tile")
        stage_tile("jbeg", "0")
        W("kb_cp_commit(); kb_cp_wait_all(); __syncthreads();")
        W("int buf = 0;")
        W("for (long long jt = jbeg; jt < jend; jt += KB_TJ) {")
        W.indent()
        W("if (jt + KB_TJ < jend) {")
        W.indent()
        stage_tile("jt + KB_TJ", "buf ^ 1")
        W.dedent()
        W("}")
        W("kb_cp_commit();")
        W("const int jn = (int)((jend - jt < KB_TJ) ? (jend - jt) : KB_TJ);")
        red_code.tile_begin(W)
        W("const %s* yrow = reinterpret_cast<const %s*>(&sm[buf][0]);" % (vty, vty))
        G = cfg.group if red_code.supports_groups() else 1
        W("int jr = 0;")
        if G > 1:
            # branch-free evaluation of G streamed rows, then ONE combined threshold test; the rare
            # path replays the group in column order so that ties still go to the smallest index
            W("#pragma unroll 1")
            W("for (; jr + %d <= jn; jr += %d, yrow += %d * (KB_STRIDE / %d)) {" % (G, G, G, vec))
            W.indent()
            W("V fg_[%d][%d];" % (G, len(value_nodes)))
            W("KB_UNROLL_ALL for (int g_ = 0; g_ < %d; g_++) {" % G)
            W.indent()
            W("const %s* yr_ = yrow + g_ * (KB_STRIDE / %d);" % (vty, vec))
            for s_ in row_loads_from("yr_"):
                W(s_)
            for s_ in body_j:
                W(s_)
            for s_ in body_ij:
                W(s_)
            red_code.group_store(W)
            W.dedent()
            W("}")
            red_code.group_update(W, G)
            W.dedent()
            W("}")
        W("#pragma unroll %d" % (cfg.unroll if G == 1 else 1))
        W("for (; jr < jn; jr++, yrow += KB_STRIDE / %d) {" % vec)
        W.indent()
        for s_ in row_loads:
            W(s_)
        for s_ in body_j:
            W(s_)
        for s_ in body_ij:
            W(s_)
        red_code.update(W)
        W.dedent()
        W("}")
        red_code.tile_end(W)
        W("kb_cp_wait_all(); __syncthreads();")
        W("buf ^= 1;")
        W.dedent()
        W("}")

    arg_decl = "".join(", const real* __restrict__ a%d_" % k for k in range(nargs))
    smem_bytes = 2 * cfg.TJ * stride * (8 if f64 else 4)
    dynamic_smem = smem_bytes if smem_bytes > 48 * 1024 else 0
    if dynamic_smem > MAX_SMEM_BYTES:
        raise FormulaError("the streamed variables have %d components in total: one double-buffered tile of %d rows does not "
                           "fit the %d KB of shared memory of an SM" % (stride, cfg.TJ, MAX_SMEM_BYTES // 1024))
    if dynamic_smem:
        # beyond the 48 KB static limit the tile is dynamic shared memory (opt-in up to 227 KB per CTA on sm_100a)
        sm_decl = ("extern __shared__ __align__(16) real kb_sm_dyn[]; "
                   "real (*sm)[KB_TJ * KB_STRIDE] = reinterpret_cast<real (*)[KB_TJ * KB_STRIDE]>(kb_sm_dyn);")
    else:
        sm_decl = "__shared__ __align__(16) real sm[2][KB_TJ * KB_STRIDE];"
    if not ranges:
        # ---------------- dense kernel ----------------
        stride_decl = "".join(", const long long bs%d" % k for k in range(nargs))
        if prescale:
            stride_decl += ", const real* __restrict__ ypk, const real* __restrict__ shf"
        kfam = kind in ("KMin", "ArgKMin", "KMin_ArgKMin")
        red_code.seeded = kfam
        if kfam:
            # K-min kernels: one device body, two entry points -- kb200_reduce (the C-ABI signature of
            # kb200_genred_launch) and kb200_reduce_seeded (+ per-row thresholds shared by the column splits)
            W("__device__ __forceinline__ void kb200_reduce_body(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, real* __restrict__ partial%s%s, real* __restrict__ shthr, const int force_partial)"
              % (arg_decl, stride_decl))
        else:
            W("extern \"C\" __global__ void __launch_bounds__(KB_T, %d)" % cfg.min_blocks)
            W("%s(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, real* __restrict__ partial%s%s)"
              % (kname, arg_decl, stride_decl))
        W("{")
        W.indent()
        W(sm_decl)
        W("const int tid = threadIdx.x;")
        W("// batch index = blockIdx.z; bs<k> is the element stride of argument k between batches (0 = shared)")
        for k in range(nargs):
            W("const real* __restrict__ a%d = a%d_ + (long long)blockIdx.z * bs%d;" % (k, k, k))
        W("out += (long long)blockIdx.z * nrow * %d;" % red.dim)
        W("partial += (long long)blockIdx.z * gridDim.y * nrow * %d;" % red_code.dimpart)
        W("const long long row0 = (long long)blockIdx.x * (KB_T * KB_R) + tid;")
        W("long long rowc[KB_R];")
        W("KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) { long long r_ = row0 + (long long)q * KB_T; rowc[q] = r_ < nrow ? r_ : nrow - 1; }")
        W("const long long jbeg = (long long)blockIdx.y * col_chunk;")
        W("const long long jend = (jbeg + col_chunk < ncol) ? (jbeg + col_chunk) : ncol;")
        W("// ---- hoisted: parameters and output-side (row) values")
        for s_ in pre:
            W(s_)
        W("// ---- accumulators")
        red_code.declare(W)
        tile_loop()
        W("// ---- write back")
        red_code.finalize(W)
        W.dedent()
        W("}")
        W("")
        if kfam:
            call_args = "nrow, ncol, col_chunk, out, partial" + "".join(", a%d_" % k for k in range(nargs)) \
                + "".join(", bs%d" % k for k in range(nargs))
            W("extern \"C\" __global__ void __launch_bounds__(KB_T, %d)" % cfg.min_blocks)
            W("%s(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, real* __restrict__ partial%s%s)"
              % (kname, arg_decl, stride_decl))
            W("{ kb200_reduce_body(%s, nullptr, 0); }" % call_args)
            W("extern \"C\" __global__ void __launch_bounds__(KB_T, %d)" % cfg.min_blocks)
            W("%s_seeded(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, real* __restrict__ partial%s%s, real* __restrict__ shthr, const int force_partial)"
              % (kname, arg_decl, stride_decl))
            W("{ kb200_reduce_body(%s, shthr, force_partial); }" % call_args)
            W("")
        # ---------------- combine kernel (merges split-j partials) ----------------
        red_code.combine_kernel(W, cname)
        if prescale:
            # ---------------- pack kernel: streamed variables -> padded rows (shifted / scaled distance coordinates) ----------
            pinfo = low.prescale
            W("")
            W("// One thread per streamed row: slot s of ypk[j] holds the s-th streamed value the formula reads; the distance")
            W("// coordinates are stored as (y - shift) * sigma.  Thread 0 also reports K (the exponent's scale, sigma^2).")
            W("extern \"C\" __global__ void __launch_bounds__(256)")
            W("kb200_pack(const long long ncol, real* __restrict__ ypk_out, const real* __restrict__ shf, real* __restrict__ info%s)"
              % "".join(", const real* __restrict__ a%d" % k for k in range(nargs)))
            W("{")
            W.indent()
            W("const long long j_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
            for s_ in pre_p:
                W(s_)
            W("if (j_ == 0) { info[0] = %s; info[1] = %s; }" % (pr.ref(pinfo["K"]), pr.ref(pinfo["sigma"])))
            W("if (j_ >= ncol) return;")
            W("real row_[KB_STRIDE];")
            W("KB_UNROLL_ALL for (int s_ = 0; s_ < KB_STRIDE; s_++) row_[s_] = (real)0;")
            for pos, (ind, comp, scaled) in enumerate(slots):
                d = varinfo[ind][0]
                if scaled:
                    k = [(q[2], q[3]) for q in pinfo["pairs"]].index((ind, comp))
                    W("row_[%d] = (a%d[j_ * %d + %d] - shf[%d]) * %s;" % (pos, ind, d, comp, k, pr.ref(pinfo["sigma"])))
                else:
                    W("row_[%d] = a%d[j_ * %d + %d];" % (pos, ind, d, comp))
            W("float4* dst_ = reinterpret_cast<float4*>(ypk_out + j_ * KB_STRIDE);")
            W("KB_UNROLL_ALL for (int s_ = 0; s_ < KB_STRIDE / 4; s_++) dst_[s_] = make_float4(row_[4 * s_], row_[4 * s_ + 1], row_[4 * s_ + 2], row_[4 * s_ + 3]);")
            W.dedent()
            W("}")
    else:
        # ---------------- block-sparse kernel ----------------
        kname, cname = "kb200_reduce_ranges", ""
        W("// One CTA = one chunk of KB_T*KB_R rows of one row-range; blk_end[k] is the inclusive prefix sum of")
        W("// the number of chunks of ranges 0..k (the reference builds a (range, start, end) lookup table per")
        W("// block on the host, keops_nvrtc.cpp:210-237; here each CTA finds its range by bisection).")
        W("extern \"C\" __global__ void __launch_bounds__(KB_T, %d)" % cfg.min_blocks)
        W("%s(const long long nrow, const long long ncol, real* __restrict__ out%s, const long long* __restrict__ rng_x, const long long* __restrict__ slc_x, const long long* __restrict__ rng_y, const long long* __restrict__ blk_end, const int nranges)"
          % (kname, arg_decl))
        W("{")
        W.indent()
        W(sm_decl)
        W("const int tid = threadIdx.x;")
        for k in range(nargs):
            W("const real* __restrict__ a%d = a%d_;" % (k, k))
        W("int lo_ = 0, hi_ = nranges;")
        W("while (lo_ < hi_) { const int mid_ = (lo_ + hi_) >> 1; if (blk_end[mid_] > (long long)blockIdx.x) hi_ = mid_; else lo_ = mid_ + 1; }")
        W("if (lo_ >= nranges) return;")
        W("const long long first_ = lo_ ? blk_end[lo_ - 1] : 0;")
        W("const long long start_x = rng_x[2 * lo_] + ((long long)blockIdx.x - first_) * (KB_T * KB_R);")
        W("const long long end_x = (start_x + KB_T * KB_R < rng_x[2 * lo_ + 1]) ? (start_x + KB_T * KB_R) : rng_x[2 * lo_ + 1];")
        W("const long long row0 = start_x + tid;")
        W("long long rowc[KB_R];")
        W("KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) { long long r_ = row0 + (long long)q * KB_T; rowc[q] = r_ < end_x ? r_ : end_x - 1; }")
        W("// ---- hoisted: parameters and output-side (row) values")
        for s_ in pre:
            W(s_)
        W("// ---- accumulators")
        red_code.declare(W)
        W("// ---- the column ranges of this row-range; contiguous ones are merged (GpuReduc1D_ranges.py:129-131)")
        W("long long sl_ = lo_ ? slc_x[lo_ - 1] : 0;")
        W("const long long sl_end_ = slc_x[lo_];")
        W("while (sl_ < sl_end_) {")
        W.indent()
        W("const long long jbeg = rng_y[2 * sl_];")
        W("long long jend = rng_y[2 * sl_ + 1];")
        W("sl_++;")
        W("while (sl_ < sl_end_ && rng_y[2 * sl_] == jend) { jend = rng_y[2 * sl_ + 1]; sl_++; }")
        tile_loop()
        W.dedent()
        W("}")
        W("// ---- write back")
        red_code.finalize(W, row_limit="end_x", allow_partial=False)
        W.dedent()
        W("}")

    src = hdr + W.text()
    h = hashlib.sha256((GENERATOR_VERSION + src).encode()).hexdigest()[:16]
    ks = KernelSource(source=src, kernel_name=kname, combine_name=cname, dimpart=red_code.dimpart,
                      stride=stride, smem_bytes=smem_bytes, hash=h, dynamic_smem=dynamic_smem, **base)
    if prescale:
        ks.prescale = dict(pairs=low.prescale["pairs"], slots=slots, stride=stride)
    return ks


@dataclass
class LoweredReduction:
    """Scalar-level program of a reduction: what is evaluated per pair and how it is accumulated."""
    graph: Graph
    acc_terms: List[Tuple[str, Tuple[SNode, ...], bool]]   # Sum / LSE weights: (mode, nodes, negative)
    value_nodes: List[SNode]                                # LSE exponent / Min / KMin values
    order: List[SNode]                                      # topological order of everything needed
    row_scales: List[Optional[SNode]] = field(default_factory=list)  # Sum: per-term row-side factor applied at write-back
    prescale: Optional[dict] = None     # pre-scaled squared distance (``_prescale_rewrite``): sigma node, coordinate pairs


def lower_reduction(red: Reduction, fast: bool, g: Optional[Graph] = None, L: Optional[Lowerer] = None,
                    prescale: bool = False) -> Optional[LoweredReduction]:
    """Vector IR -> optimised scalar DAG.  Returns None when the reduction is structurally zero.

    Sum terms and LSE weights accumulate the *unsigned* product / value and the pulled-out sign is
    applied once at write-back (negation is free there, not in the packed FFMA2 loop)."""
    if g is None:
        g = Graph(fast=fast, tagI=red.tagI)
        L = Lowerer(g)
    F = L(red.formula)
    G = L(red.weight) if red.weight is not None else []
    kind = red.kind
    if kind == "Sum" and all(sv.const() == 0.0 for sv in F):
        return None
    pre_info = None
    if prescale:
        hit = _prescale_rewrite(g, F) if (kind == "Sum" and fast) else None
        if hit is None:
            return None              # the caller asked for the pre-scaled variant and it does not apply
        F, pre_info = hit
    rwt = _Rewriter(g)
    acc_terms: List[Tuple[str, Tuple[SNode, ...], bool]] = []
    value_nodes: List[SNode] = []
    row_scales: List[Optional[SNode]] = []
    if kind == "Sum":
        for sv in F:
            acc_terms.append(_acc_term(rwt, sv))
        if fast:
            alt = _pull_row_factors(g, F)
            if alt is not None:
                rwt2 = _Rewriter(g)
                terms2 = [_acc_term(rwt2, sv) for sv, _ in alt]
                scales2 = [None if h is None else rwt2.rw(h) for _, h in alt]
                if _pair_cost(terms2) < _pair_cost(acc_terms):
                    acc_terms, row_scales = terms2, scales2
    elif kind == "LSE":
        value_nodes = [rwt.signed(F[0])]
        for sv in G:
            acc_terms.append(_acc_term(rwt, sv))
    else:
        value_nodes = [rwt.signed(sv) for sv in F]
    roots: List[SNode] = list(value_nodes)
    for mode, nodes, _ in acc_terms:
        roots.extend(nodes)
    roots.extend(h for h in row_scales if h is not None)
    if pre_info is not None:
        pre_info["sigma"] = rwt.rw(pre_info["sigma"])
        pre_info["K"] = rwt.rw(pre_info["K"])
        roots.extend([pre_info["sigma"], pre_info["K"]])
    return LoweredReduction(g, acc_terms, value_nodes, topo_order(roots), row_scales, pre_info)


PRESCALE_MAX_DIM = 8


def _prescale_rewrite(g: Graph, F: Sequence[SV]):
    """``2^(-K |x_i - y_j|^2)`` with the scale folded into the coordinates: ``x' = (x - m) sigma``, ``y' = (y - m) sigma``
    with ``sigma = sqrt(K)``, so that the exponent is ``-(sum_k (x'_k - y'_k)^2)`` -- the ``K *`` multiply of the pair loop
    (one of 8 FP32-pipe operations per pair of the Gaussian convolution in D = 3, which co-limits with the MUFU pipe)
    disappears; the negation is a free source modifier of MUFU.EX2 (``ex2n``).  ``m`` is a common shift (the streamed
    cloud's mean, passed at run time) that keeps the rounding of the scaled coordinates relative to the cloud's extent.
    Other uses of the differences ``d_k`` (gradients) become ``(1 / sigma) d'_k``, a hoistable factor.

    The scaled streamed coordinates are produced once per call by ``kb200_pack`` (cat-7 loads); the scaled output-side
    coordinates are hoisted per thread.  Rounding: each scaled coordinate carries a relative error <= 2^-23, i.e. the
    exponent an absolute error <= 4 * 2^-24 |d'| (|x'| + |y'|) -- the caller evaluates this bound on the data and falls
    back to the exact-order kernel when it is not small (``binder._prescale_guard``).

    Matches: exactly one ``(K, S)`` pair over all ``ex2`` nodes, ``ex2(neg(mul(K, S)))`` with ``K`` a positive constant or a
    parameter-only value (its sign is checked at run time) and ``S = sum_k (a_k - b_k)^2``, ``a_k`` / ``b_k`` plain components
    of an output-side / a streamed variable, at most ``PRESCALE_MAX_DIM`` terms.  Returns ``(F', info)`` or None."""
    roots = [sv.n for sv in F]
    found = None
    for n in topo_order(roots):
        if n.op != "ex2":
            continue
        a = n.args[0]
        if a.op != "neg" or a.args[0].op != "mul":
            return None              # an exponential of something else: leave the formula alone
        p, q = a.args[0].args
        K, S = (p, q) if p.cls <= P else (q, p)
        if K.cls > P or S.cls != IJ:
            return None
        if found is not None and (found[0] is not K or found[1] is not S):
            return None
        found = (K, S)
    if found is None:
        return None
    K, S = found
    # S: nested adds of squares of differences
    leaves, stack = [], [S]
    while stack:
        t = stack.pop()
        if t.op == "add":
            stack.extend(t.args)
        else:
            leaves.append(t)
    if not (1 <= len(leaves) <= PRESCALE_MAX_DIM):
        return None
    pairs = []          # (d node, a node (row side), b node (streamed side), a_first)
    for t in leaves:
        if t.op != "mul" or t.args[0] is not t.args[1] or t.args[0].op != "sub":
            return None
        d = t.args[0]
        u, v = d.args
        if u.op != "load" or v.op != "load" or {u.cls, v.cls} != {I, J} or u.val[2] not in (0, 1) or v.val[2] not in (0, 1):
            return None
        pairs.append((d, u, v) if u.cls == I else (d, v, u))
    if len({id(p[0]) for p in pairs}) != len(pairs):
        return None
    pairs.sort(key=lambda p: (p[1].val[0], p[1].val[1]))
    sigma = g.sqrt_rn(SV(K))
    inv_sigma = g.rcp(sigma) if sigma.const() is None else g.const(1.0 / sigma.const())
    dmap: Dict[int, SV] = {}
    sq: Dict[int, SV] = {}
    for k, (d, a, b) in enumerate(pairs):
        shift = g.load(k, 0, 6)
        # hoisted per thread; "keep" stops the FMA contraction from pulling this product back into the pair loop
        a2 = SV(g._mk("keep", (g.materialize(g.mul(g.sub(SV(a), shift), sigma)),)))
        b2 = g.load(b.val[0], b.val[1], 7)                           # (b - shift) * sigma, from the packed tile
        d2 = g.sub(a2, b2) if d.args[0] is a else g.sub(b2, a2)
        dmap[d.uid] = d2
        sq[d.uid] = g.mul(d2, d2)
    # S' with the original association
    smemo: Dict[int, SV] = {}

    def rebuild_S(t: SNode) -> SV:
        r = smemo.get(t.uid)
        if r is None:
            r = g.add(rebuild_S(t.args[0]), rebuild_S(t.args[1])) if t.op == "add" else sq[t.args[0].uid]
            smemo[t.uid] = r
        return r

    S2 = rebuild_S(S)
    memo: Dict[int, SV] = {}

    def rebuild(n: SNode) -> SV:
        r = memo.get(n.uid)
        if r is not None:
            return r
        if n.op == "ex2":
            r = SV(g._mk("ex2", (g.materialize(-S2),)))
        elif n.uid in dmap:
            r = g.mul(inv_sigma, dmap[n.uid])
        elif n.op in ("const", "load"):
            r = SV(n)
        else:
            args = [rebuild(x) for x in n.args]
            if n.op == "add":
                r = g.add(args[0], args[1])
            elif n.op == "sub":
                r = g.sub(args[0], args[1])
            elif n.op == "mul":
                r = g.mul(args[0], args[1])
            elif n.op == "neg":
                r = -args[0]
            else:
                r = SV(g._mk(n.op, tuple(g.materialize(x) for x in args), n.val))
        memo[n.uid] = r
        return r

    F2 = []
    for sv in F:
        r = rebuild(sv.n)
        F2.append(-r if sv.neg else r)
    info = dict(sigma=g.materialize(sigma), K=K,
                pairs=[(a.val[0], a.val[1], b.val[0], b.val[1]) for _, a, b in pairs])
    return F2, info


def _pair_cost(acc_terms) -> int:
    """Number of per-pair (class IJ) operations of a Sum program, accumulation included."""
    roots = [n for _, nodes, _ in acc_terms for n in nodes]
    cost = sum(1 for n in topo_order(roots) if n.cls == IJ and n.op not in ("const", "load"))
    return cost + sum(1 for mode, _, _ in acc_terms if mode != "zero")


def _pull_row_factors(g: Graph, F: Sequence[SV]):
    """Sum_j h_i * G_ij = h_i * Sum_j G_ij: factors of a summed product that depend on the output row
    only (class <= I: constants, parameters, row-side values) are taken out of the pair loop and
    applied once at write-back.  E.g. the Gaussian-conv gradient w.r.t. x accumulates
    ``(x-y) * (b_j * e)`` and multiplies by ``2*eta_i`` at the end: 11 instead of 12 FP32 operations
    per pair.  A ``use_fast_math`` re-association, like the reference's nvcc ``-use_fast_math`` build.
    Returns [(stripped SV, row factor node or None)] or None when nothing can be pulled."""
    memo: Dict[int, Tuple[Optional[SV], List[SNode]]] = {}

    def strip(n: SNode) -> Tuple[Optional[SV], List[SNode]]:
        r = memo.get(n.uid)
        if r is not None:
            return r
        if n.cls in (C, P, I):
            r = (None, [n])
        elif n.op == "mul":
            (ra, fa), (rb, fb) = strip(n.args[0]), strip(n.args[1])
            if not fa and not fb:
                r = (SV(n), [])
            elif ra is None:
                r = (rb, fa + fb)
            elif rb is None:
                r = (ra, fa + fb)
            else:
                r = (g.mul(ra, rb), fa + fb)
        else:
            r = (SV(n), [])
        memo[n.uid] = r
        return r

    out, any_pulled = [], False
    for sv in F:
        if sv.const() is not None or sv.n.cls != IJ:
            out.append((sv, None))
            continue
        rest, facs = strip(sv.n)
        facs = [f for f in facs if not (f.op == "const" and f.val == 1.0)]
        if rest is None or not facs or all(f.op == "const" for f in facs):
            out.append((sv, None))
            continue
        h = SV(facs[0])
        for f in facs[1:]:
            h = g.mul(h, SV(f))
        neg = bool(sv.neg) ^ bool(rest.neg) ^ bool(h.neg)
        out.append((SV(rest.n, neg), h.n))
        any_pulled = True
    return out if any_pulled else None


def _acc_term(rwt: _Rewriter, sv: SV):
    """Decompose an accumulated value: ('fma', (a, b), neg) -> acc += a*b ; ('add', (v,), neg)."""
    n = sv.n
    if sv.const() == 0.0:
        return ("zero", (), False)
    if n.op == "mul":
        return ("fma", (rwt.rw(n.args[0]), rwt.rw(n.args[1])), sv.neg)
    return ("add", (rwt.rw(n),), sv.neg)


class _Writer:
    def __init__(self):
        self.lines: List[str] = []
        self.lvl = 0

    def __call__(self, s: str = ""):
        self.lines.append(("  " * self.lvl + s) if s else "")

    def indent(self):
        self.lvl += 1

    def dedent(self):
        self.lvl -= 1

    def text(self) -> str:
        return "\n".join(self.lines) + "\n"


def _kernel_header(W, red, cfg, f64, fast, stride, nargs, varinfo) -> str:
    lines = [
        "// generated by keops_b200 (%s) -- do not edit" % GENERATOR_VERSION,
        "// reduction: %s" % red.canonical(),
        "#define KB_T %d" % cfg.T,
        "#define KB_R %d" % cfg.R,
        "#define KB_TJ %d" % cfg.TJ,
        "#define KB_STRIDE %d" % stride,
        "#define KB_F64 %d" % (1 if f64 else 0),
        "#define KB_FAST %d" % (1 if fast else 0),
        prelude(),
        "",
    ]
    return "\n".join(lines)


def _stage(W: _Writer, col_vars, col_off, jbase: str, buf: str):
    """cp.async one tile of every streamed variable into sm[buf] (row-padded AoS)."""
    for ind, d in col_vars:
        W("for (int e = tid; e < KB_TJ * %d; e += KB_T) {" % d)
        W.indent()
        if d == 1:
            W("const int jr_ = e; const int c_ = 0;")
        else:
            W("const int jr_ = e / %d; const int c_ = e - jr_ * %d;" % (d, d))
        W("const long long j_ = (%s) + jr_;" % jbase)
        W("const bool ok_ = j_ < jend;")
        W("kb_cp_async(&sm[%s][jr_ * KB_STRIDE + %d + c_], a%d + (ok_ ? (j_ * %d + c_) : 0), ok_);"
          % (buf, col_off[ind], ind, d))
        W.dedent()
        W("}")


# ------------------------------------------------------------------------------------------
# reduction-specific code
# ------------------------------------------------------------------------------------------
class _ReductionCode:
    def __init__(self, red: Reduction, pr: _Printer, acc_terms, value_nodes, sum_scheme, dtype_acc, f64, fast,
                 row_scales=()):
        self.red, self.pr = red, pr
        self.row_scales = list(row_scales) if row_scales else [None] * len(acc_terms)
        self.kind = red.kind
        self.acc_terms, self.values = acc_terms, value_nodes
        self.f64, self.fast = f64, fast
        self.scheme = sum_scheme if self.kind in ("Sum", "LSE") else "direct_sum"
        if self.scheme not in ("block_sum", "direct_sum", "kahan_scheme"):
            raise ValueError('invalid value for option sum_scheme : should be one of "auto", "direct_sum", "block_sum" or "kahan_scheme".')
        # fp32 data with fp64 accumulators (reference: MapReduce.py:74 declares ``acc`` in dtypeacc for every reduction;
        # only the sum-like ones have an accumulator whose precision matters)
        self.acc64 = (not f64) and dtype_acc in ("float64", "double") and self.kind in ("Sum", "LSE")
        self.kahan = self.scheme == "kahan_scheme"
        self.dimF = red.dimF
        k = self.kind
        if k == "Sum":
            self.dimpart = red.dim
        elif k == "LSE":
            self.dimpart = red.dim
        elif k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            self.dimpart = 2 * self.dimF
        else:
            self.dimpart = 2 * red.K * self.dimF
        self.is_max = k in ("Max", "ArgMax", "Max_ArgMax")
        self.seeded = False        # set by generate(): dense K-min kernels share per-row thresholds between splits

    # ---- helpers -----------------------------------------------------------------------------
    def _term_expr(self, term, acc: str) -> Optional[str]:
        mode, nodes, _ = term
        if mode == "zero":
            return None
        if mode == "fma":
            return "kb_fma_v(%s, %s, %s)" % (self.pr.ref(nodes[0]), self.pr.ref(nodes[1]), acc)
        return "kb_add_v(%s, %s)" % (acc, self.pr.ref(nodes[0]))

    def _term_value(self, term) -> Optional[str]:
        """Expression of |term| as a V (for Kahan)."""
        mode, nodes, _ = term
        if mode == "zero":
            return None
        if mode == "fma":
            return "kb_mul_v(%s, %s)" % (self.pr.ref(nodes[0]), self.pr.ref(nodes[1]))
        n = nodes[0]
        if n.cls in (I, IJ):
            return self.pr.ref(n)
        return "kb_bcast_v(%s)" % self.pr.ref(n)

    # ---- declarations --------------------------------------------------------------------------
    def declare(self, W):
        k, d = self.kind, self.dimF
        if k == "Sum":
            n = len(self.acc_terms)
            if self.acc64:
                W("double accd[%d][KB_R];" % n)
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) accd[k_][q] = 0.0;" % n)
            else:
                W("V acc[%d];" % n)
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) acc[k_] = kb_bcast_v((real)0);" % n)
            self._declare_comp(W, n)
        elif k == "LSE":
            n = len(self.acc_terms)
            W("V mx = kb_bcast_v(-KB_INF);")
            if self.acc64:
                W("double accd[%d][KB_R];" % n)
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) accd[k_][q] = 0.0;" % n)
            else:
                W("V acc[%d];" % n)
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) acc[k_] = kb_bcast_v((real)0);" % n)
            self._declare_comp(W, n)
        elif k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            init = "-KB_INF" if self.is_max else "KB_INF"
            W("V best[%d]; int bidx[%d][KB_R];" % (d, d))
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) { best[k_] = kb_bcast_v(%s); KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) bidx[k_][q] = 0; }" % (d, init))
        else:
            K = self.red.K
            W("real kval[%d][KB_R][%d]; int kidx[%d][KB_R][%d]; V thr[%d];" % (d, K, d, K, d))
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) { thr[k_] = kb_bcast_v(KB_INF);" % d)
            W("  KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { kval[k_][q][l_] = KB_INF; kidx[k_][q][l_] = 0; } }" % K)
            if self.seeded:
                # Column splits of one row cooperate through shthr[row][k_] (global memory, starts at +inf): the smallest
                # K-th value any split has published so far -- an upper bound of the row's final K-th smallest value.
                # Candidates above it are not inserted, so a split does not replay the whole transient of a list that
                # starts empty (measured: 15 % of the 4-column groups took the insertion path with 7 splits, 3 % unsplit).
                # Non-strict at the bound (kb_next_up), so ties still reach the merge, which orders them by split.
                W("real* const shp = (shthr != nullptr) ? (shthr + row0 * %d) : nullptr;" % d)
                W("if (shp != nullptr) {")
                W("  KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++)" % d)
                W("    if (row0 + (long long)q * KB_T < nrow) thr[k_].l[q] = kb_next_up(__ldcg(shp + q * (KB_T * %d) + k_)); }" % d)

    def _declare_comp(self, W, n):
        """Kahan compensation terms, held in the accumulator's precision (Sum_Reduction.py:31-40)."""
        if not self.kahan:
            return
        if self.acc64:
            W("double compd[%d][KB_R];" % n)
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) compd[k_][q] = 0.0;" % n)
        else:
            W("V comp[%d];" % n)
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) comp[k_] = kb_bcast_v((real)0);" % n)

    def _kahan_add(self, W, i, val):
        """acc[i] (+)= val with compensation; ``val`` is a V expression."""
        if self.acc64:
            W("{ const V v_ = %s; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) { const double a_ = (double)v_.l[q] - compd[%d][q]; "
              "const double b_ = accd[%d][q] + a_; compd[%d][q] = (b_ - accd[%d][q]) - a_; accd[%d][q] = b_; } }" % (val, i, i, i, i, i))
        else:
            W("{ const V a_ = kb_sub_v(%s, comp[%d]); const V b_ = kb_add_v(acc[%d], a_); comp[%d] = kb_sub_v(kb_sub_v(b_, acc[%d]), a_); acc[%d] = b_; }"
              % (val, i, i, i, i, i))

    def tile_begin(self, W):
        if self.kind in ("Sum", "LSE") and self.scheme == "block_sum":
            n = len(self.acc_terms)
            W("V tmp[%d];" % n)
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) tmp[k_] = kb_bcast_v((real)0);" % n)

    def _accname(self):
        return "tmp" if self.scheme == "block_sum" else "acc"

    def update(self, W):
        k = self.kind
        if k == "Sum":
            if self.scheme == "kahan_scheme":
                for i, term in enumerate(self.acc_terms):
                    val = self._term_value(term)
                    if val is None:
                        continue
                    self._kahan_add(W, i, val)
                return
            if self.acc64 and self.scheme != "block_sum":
                for i, term in enumerate(self.acc_terms):
                    val = self._term_value(term)
                    if val is None:
                        continue
                    W("{ const V v_ = %s; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) accd[%d][q] += (double)v_.l[q]; }" % (val, i))
                return
            a = self._accname()
            for i, term in enumerate(self.acc_terms):
                e = self._term_expr(term, "%s[%d]" % (a, i))
                if e is not None:
                    W("%s[%d] = %s;" % (a, i, e))
        elif k == "LSE":
            f = self.pr.ref(self.values[0])
            fv = f if self.values[0].cls in (I, IJ) else "kb_bcast_v(%s)" % f
            a = self._accname()
            n = len(self.acc_terms)
            W("const V f_ = %s;" % fv)
            # rare path: a new running maximum on at least one lane
            W("{ bool any_ = false; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (f_.l[q] > mx.l[q]);")
            W("  if (any_) { KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) if (f_.l[q] > mx.l[q]) {")
            W("      const real sc_ = kb_exp(mx.l[q] - f_.l[q]); mx.l[q] = f_.l[q];")
            resc = "accd[k_][q] *= (double)sc_;" if self.acc64 else "acc[k_].l[q] *= sc_;"
            if a == "tmp":
                resc += " tmp[k_].l[q] *= sc_;"
            if self.kahan:
                resc += " compd[k_][q] *= (double)sc_;" if self.acc64 else " comp[k_].l[q] *= sc_;"
            W("      KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) { %s }" % (n, resc))
            W("  } } }")
            if self.fast:
                W("const V e_ = kb_ex2_v(kb_mul_v(kb_sub_v(f_, mx), %s));" % _lit(1.4426950408889634, self.f64))
            else:
                W("const V e_ = kb_exp_v(kb_sub_v(f_, mx));")
            direct64 = self.acc64 and a == "acc"          # direct_sum into fp64 accumulators: per-lane adds
            for i, term in enumerate(self.acc_terms):
                mode, nodes, _ = term
                if mode == "zero":
                    continue
                if mode == "fma":
                    val = "kb_mul_v(kb_mul_v(%s, %s), e_)" % (self.pr.ref(nodes[0]), self.pr.ref(nodes[1]))
                    upd = "kb_fma_v(kb_mul_v(%s, %s), e_, %s[%d])" % (self.pr.ref(nodes[0]), self.pr.ref(nodes[1]), a, i)
                else:
                    nd = nodes[0]
                    if nd.op == "const" and nd.val == 1.0:
                        val = "e_"
                        upd = "kb_add_v(%s[%d], e_)" % (a, i)
                    else:
                        val = "kb_mul_v(%s, e_)" % self.pr.ref(nd)
                        upd = "kb_fma_v(%s, e_, %s[%d])" % (self.pr.ref(nd), a, i)
                if self.kahan:
                    self._kahan_add(W, i, val)
                elif direct64:
                    W("{ const V v_ = %s; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) accd[%d][q] += (double)v_.l[q]; }" % (val, i))
                else:
                    W("%s[%d] = %s;" % (a, i, upd))
        elif k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            cmp = ">" if self.is_max else "<"
            for i, nd in enumerate(self.values):
                f = self.pr.ref(nd)
                fv = f if nd.cls in (I, IJ) else "kb_bcast_v(%s)" % f
                W("{ const V f_ = %s; bool any_ = false; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (f_.l[q] %s best[%d].l[q]);" % (fv, cmp, i))
                W("  if (any_) { KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) if (f_.l[q] %s best[%d].l[q]) { best[%d].l[q] = f_.l[q]; bidx[%d][q] = (int)(jt + jr); } } }"
                  % (cmp, i, i, i))
        else:
            K = self.red.K
            for i, nd in enumerate(self.values):
                f = self.pr.ref(nd)
                fv = f if nd.cls in (I, IJ) else "kb_bcast_v(%s)" % f
                W("{ const V f_ = %s; bool any_ = false; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (f_.l[q] < thr[%d].l[q]);" % (fv, i))
                W("  if (any_) { KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) if (f_.l[q] < thr[%d].l[q]) {" % i)
                # seeded: thr = min(local K-th, shared bound); the local K-th only decreases, so min with the old thr keeps it
                newthr = ("kb_fmin(kval[%d][q][%d], thr[%d].l[q])" % (i, K - 1, i)) if self.seeded else ("kval[%d][q][%d]" % (i, K - 1))
                W("      kb_kmin_insert<%d, int>(kval[%d][q], kidx[%d][q], f_.l[q], (int)(jt + jr)); thr[%d].l[q] = %s; } } }"
                  % (K, i, i, i, newthr))

    # ---- grouped evaluation (Min / KMin families) ----------------------------------------------
    def supports_groups(self) -> bool:
        return self.kind not in ("Sum", "LSE")

    def _value_v(self, nd) -> str:
        f = self.pr.ref(nd)
        return f if nd.cls in (I, IJ) else "kb_bcast_v(%s)" % f

    def group_store(self, W):
        for i, nd in enumerate(self.values):
            W("fg_[g_][%d] = %s;" % (i, self._value_v(nd)))

    def group_update(self, W, G: int):
        k = self.kind
        nv = len(self.values)
        minlike = k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax")
        cmp = ">" if self.is_max else "<"
        ref = "best" if minlike else "thr"
        W("bool any_ = false;")
        W("KB_UNROLL_ALL for (int g_ = 0; g_ < %d; g_++) KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (fg_[g_][k_].l[q] %s %s[k_].l[q]);"
          % (G, nv, cmp, ref))
        W("if (any_) {")
        W.indent()
        W("KB_UNROLL_ALL for (int g_ = 0; g_ < %d; g_++) KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) {" % (G, nv))
        W.indent()
        W("const real f1_ = fg_[g_][k_].l[q];")
        if minlike:
            W("if (f1_ %s best[k_].l[q]) { best[k_].l[q] = f1_; bidx[k_][q] = (int)(jt + jr + g_); }" % cmp)
        else:
            K = self.red.K
            newthr = ("kb_fmin(kval[k_][q][%d], thr[k_].l[q])" % (K - 1)) if self.seeded else ("kval[k_][q][%d]" % (K - 1))
            W("if (f1_ < thr[k_].l[q]) { kb_kmin_insert<%d, int>(kval[k_][q], kidx[k_][q], f1_, (int)(jt + jr + g_)); thr[k_].l[q] = %s; }"
              % (K, newthr))
        W.dedent()
        W("}")
        W.dedent()
        W("}")

    def tile_end(self, W):
        if self.seeded:
            # once per tile: publish this split's K-th value if it beats the shared bound, and tighten the local
            # threshold with the bound (invariant: thr = min(local K-th value, next_up(bound seen so far)))
            K, d = self.red.K, self.dimF
            W("if (shp != nullptr) {")
            W("  KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++)" % d)
            W("    if (row0 + (long long)q * KB_T < nrow) {")
            W("      real* a_ = shp + q * (KB_T * %d) + k_; real s_ = __ldcg(a_);" % d)
            W("      if (kval[k_][q][%d] < s_) { kb_atomic_min(a_, kval[k_][q][%d]); s_ = kval[k_][q][%d]; }" % (K - 1, K - 1, K - 1))
            W("      thr[k_].l[q] = kb_fmin(thr[k_].l[q], kb_next_up(s_)); } }")
        if self.kind in ("Sum", "LSE") and self.scheme == "block_sum":
            n = len(self.acc_terms)
            if self.acc64:
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) accd[k_][q] += (double)tmp[k_].l[q];" % n)
            else:
                W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) acc[k_] = kb_add_v(acc[k_], tmp[k_]);" % n)

    # ---- output ----------------------------------------------------------------------------------
    def _signs(self):
        return ["-" if t[2] else "" for t in self.acc_terms]

    def finalize(self, W, row_limit: str = "nrow", allow_partial: bool = True, row_expr: str = "row0 + (long long)q * KB_T",
                 split_expr: str = "(long long)blockIdx.y", final_expr: Optional[str] = None):
        """Write either the final output (gridDim.y == 1) or this split's partial accumulator."""
        k, d = self.kind, self.dimF
        W("KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) {")
        W.indent()
        W("const long long r_ = %s;" % row_expr)
        W("if (r_ < %s) {" % row_limit)
        W.indent()
        final = "(gridDim.y == 1 && !force_partial)" if (self.seeded and allow_partial) else "(gridDim.y == 1)"
        if final_expr is not None:
            final = final_expr
        self._final_cond = final
        if allow_partial:
            W("real* po = %s ? (out + r_ * %d) : (partial + (%s * nrow + r_) * %d);" % (final, self.red.dim, split_expr, self.dimpart))
        else:
            W("real* po = out + r_ * %d;" % self.red.dim)
        if k == "Sum":
            for i, s in enumerate(self._signs()):
                src = "(real)accd[%d][q]" % i if self.acc64 else "acc[%d].l[q]" % i
                h = self.row_scales[i]
                if h is not None:   # row-side factor pulled out of the sum (_pull_row_factors)
                    src += " * " + (self.pr.ref(h) + ".l[q]" if h.cls == I else self.pr.ref(h))
                W("po[%d] = %s%s;" % (i, s, src))
        elif k == "LSE":
            W("po[0] = mx.l[q];")
            for i, s in enumerate(self._signs()):
                W("po[%d] = %s%s;" % (1 + i, s, "(real)accd[%d][q]" % i if self.acc64 else "acc[%d].l[q]" % i))
        elif k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            W("if %s {" % final)
            W.indent()
            for i in range(d):
                if k in ("Min", "Max"):
                    W("po[%d] = best[%d].l[q];" % (i, i))
                elif k in ("ArgMin", "ArgMax"):
                    W("po[%d] = (real)bidx[%d][q];" % (i, i))
                else:
                    W("po[%d] = best[%d].l[q]; po[%d] = (real)bidx[%d][q];" % (i, i, d + i, i))
            W.dedent()
            W("} else {")
            W.indent()
            for i in range(d):
                W("po[%d] = best[%d].l[q]; po[%d] = (real)bidx[%d][q];" % (i, i, d + i, i))
            W.dedent()
            W("}")
        else:
            K = self.red.K
            W("if %s {" % final)
            W.indent()
            for i in range(d):
                W("KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) {" % K)
                if k == "KMin":
                    W("  po[l_ * %d + %d] = kval[%d][q][l_];" % (d, i, i))
                elif k == "ArgKMin":
                    W("  po[l_ * %d + %d] = (real)kidx[%d][q][l_];" % (d, i, i))
                else:
                    W("  po[l_ * %d + %d] = kval[%d][q][l_]; po[l_ * %d + %d] = (real)kidx[%d][q][l_];"
                      % (2 * d, i, i, 2 * d, d + i, i))
                W("}")
            W.dedent()
            W("} else {")
            W.indent()
            for i in range(d):
                W("KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { po[l_ * %d + %d] = kval[%d][q][l_]; po[l_ * %d + %d] = (real)kidx[%d][q][l_]; }"
                  % (K, 2 * d, i, i, 2 * d, d + i, i))
            W.dedent()
            W("}")
        W.dedent()
        W("}")
        W.dedent()
        W("}")

    def combine_kernel(self, W, cname: str, tie_by_index: bool = False):
        """One thread per output row merges the ``nsplit`` partial accumulators in split order
        (so ties still resolve to the smallest index).  ``tie_by_index``: the splits hold *interleaved* column sets (chunked
        kernel: the warp groups of a CTA share every tile), so on equal values the smaller index is chosen explicitly."""
        k, d, dp, dim = self.kind, self.dimF, self.dimpart, self.red.dim
        W("extern \"C\" __global__ void __launch_bounds__(256)")
        W("%s(const long long nrow, const int nsplit, const real* __restrict__ partial, real* __restrict__ out)" % cname)
        W("{")
        W.indent()
        W("const long long r_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
        W("if (r_ >= nrow) return;")
        W("partial += (long long)blockIdx.z * nsplit * nrow * %d;" % dp)
        W("out += (long long)blockIdx.z * nrow * %d;" % dim)
        W("const real* pp = partial + r_ * %d;" % dp)
        W("const long long sstride = nrow * %d;" % dp)
        W("real* po = out + r_ * %d;" % dim)
        if k == "Sum":
            W("real a_[%d];" % dp)
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) a_[k_] = pp[k_];" % dp)
            W("for (int s_ = 1; s_ < nsplit; s_++) { KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) a_[k_] += pp[s_ * sstride + k_]; }" % dp)
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) po[k_] = a_[k_];" % dp)
        elif k == "LSE":
            W("real m_ = pp[0]; real a_[%d];" % (dp - 1))
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) a_[k_] = pp[1 + k_];" % (dp - 1))
            W("for (int s_ = 1; s_ < nsplit; s_++) {")
            W("  const real* ps = pp + s_ * sstride; const real m2_ = ps[0];")
            W("  if (m_ > m2_) { const real e_ = kb_exp(m2_ - m_); KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) a_[k_] += ps[1 + k_] * e_; }" % (dp - 1))
            W("  else { const real e_ = kb_exp(m_ - m2_); KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) a_[k_] = ps[1 + k_] + e_ * a_[k_]; m_ = m2_; }" % (dp - 1))
            W("}")
            W("po[0] = m_; KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) po[1 + k_] = a_[k_];" % (dp - 1))
        elif k in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            cmp = ">" if self.is_max else "<"
            W("real v_[%d], ix_[%d];" % (d, d))
            W("KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) { v_[k_] = pp[k_]; ix_[k_] = pp[%d + k_]; }" % (d, d))
            W("for (int s_ = 1; s_ < nsplit; s_++) { const real* ps = pp + s_ * sstride;")
            tie = (" || (ps[k_] == v_[k_] && ps[%d + k_] < ix_[k_])" % d) if tie_by_index else ""
            W("  KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) if (ps[k_] %s v_[k_]%s) { v_[k_] = ps[k_]; ix_[k_] = ps[%d + k_]; } }" % (d, cmp, tie, d))
            for i in range(d):
                if k in ("Min", "Max"):
                    W("po[%d] = v_[%d];" % (i, i))
                elif k in ("ArgMin", "ArgMax"):
                    W("po[%d] = ix_[%d];" % (i, i))
                else:
                    W("po[%d] = v_[%d]; po[%d] = ix_[%d];" % (i, i, d + i, i))
        else:
            K = self.red.K
            W("for (int k_ = 0; k_ < %d; k_++) {" % d)
            W("  real v_[%d], ix_[%d];" % (K, K))
            W("  KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { v_[l_] = pp[l_ * %d + k_]; ix_[l_] = pp[l_ * %d + %d + k_]; }" % (K, 2 * d, 2 * d, d))
            W("  for (int s_ = 1; s_ < nsplit; s_++) { const real* ps = pp + s_ * sstride;")
            if tie_by_index:
                W("    for (int l_ = 0; l_ < %d; l_++) { const real f_ = ps[l_ * %d + k_], fi_ = ps[l_ * %d + %d + k_];" % (K, 2 * d, 2 * d, d))
                W("      if (!(f_ < v_[%d] || (f_ == v_[%d] && fi_ < ix_[%d]))) break;" % (K - 1, K - 1, K - 1))
                W("      kb_kmin_insert_tie<%d>(v_, ix_, f_, fi_); } }" % K)
            else:
                W("    for (int l_ = 0; l_ < %d; l_++) { const real f_ = ps[l_ * %d + k_]; if (!(f_ < v_[%d])) break;" % (K, 2 * d, K - 1))
                W("      kb_kmin_insert<%d, real>(v_, ix_, f_, ps[l_ * %d + %d + k_]); } }" % (K, 2 * d, d))
            W("  KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) {" % K)
            if k == "KMin":
                W("    po[l_ * %d + k_] = v_[l_];" % d)
            elif k == "ArgKMin":
                W("    po[l_ * %d + k_] = ix_[l_];" % d)
            else:
                W("    po[l_ * %d + k_] = v_[l_]; po[l_ * %d + %d + k_] = ix_[l_];" % (2 * d, 2 * d, d))
            W("  }")
            W("}")
        W.dedent()
        W("}")
