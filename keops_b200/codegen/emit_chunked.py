"""Chunked-dimension kernels: high-dimensional ``Sum`` / ``Scalprod`` sub-formulas on the CUDA cores without spills.

Replaces the reference's *chunk mode* (``keopscore/keopscore/mapreduce/gpu/GpuReduc1D_chunks.py:19-118,259-330``, switched
on by ``keopscore/keopscore/get_keops_dll.py:101-111`` when a ``Chunkable_Op`` has dimension >= 146 or one of the special
dimensions of ``keopscore/keopscore/config/chunks.py:19-21``; decomposition into *chunked* and *post-chunk* formulas:
``keopscore/keopscore/mapreduce/Chunk_Mode_Constants.py:6-120``, ``keopscore/keopscore/formulas/Chunkable_Op.py:6-57``).

Why: the generic kernel keeps ``x_i`` in registers and unrolls the whole formula, so beyond D ~ 128 it spills (D = 256:
660 ms at M = N = 1e5), its staged tile passes 48 KB and the D-wide unrolled source takes minutes to compile (D = 990).

Decomposition (same idea as the reference, on the vector IR):

* *chunk roots*: maximal ``Sum(E)`` / ``Scalprod(A, B)`` nodes whose argument has dimension D >= ``CHUNK_MIN_DIM`` and is
  built from element-wise operations on D-dimensional variables (scalars broadcast).  Each root becomes the scalar
  *chunk body* ``e(x_k, y_k)`` -- one component -- accumulated left to right over k, the reference's summation order;
* the *post formula*: the reduction's formula with every root replaced by a per-pair placeholder ``Var(ind, 1, 3)``; it
  must not use a high-dimensional variable any more (the reference's ``chunk_postchunk_mix`` condition), otherwise the
  generic kernel is used.

Kernel (different from the reference, which keeps a ``blockDim x dimout`` local array per thread):

* a CTA owns ``ROWS`` output rows whose high-dimensional variables stay **resident in shared memory for all of D**
  (transposed ``[k][row]``, so a thread reads its two rows' component as one conflict-free 8-byte word) -- the block is
  fetched from global memory once, whatever N is;
* the streamed side goes through a double-buffered ``TJ x DC`` tile (16-byte ``cp.async``), one barrier per chunk;
* the CTA's warps are ``G`` *column groups*: thread ``(tr, g)`` accumulates the roots of its 2 rows x ``NJ = TJ / G`` columns
  in registers (packed FADD2 / FFMA2, a broadcast ``LDS.128`` of ``y`` feeds 4 components), so the CTA runs 8 warps even
  when only 64-128 rows fit;
* after the last chunk the post formula and the reduction update run per column exactly as in the generic kernel
  (same ``_ReductionCode``); each column group writes its own partial accumulator and ``kb200_combine`` merges them --
  with an explicit index comparison on ties, because the groups' column sets interleave.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from ..formulas.ir import Node
from ..formulas.reductions import Reduction
from .emit import (GENERATOR_VERSION, KernelConfig, KernelSource, MAX_SMEM_BYTES, _Printer, _ReductionCode, _Rewriter, _Writer,
                   _acc_term, _kernel_header, _stage, lower_reduction)
from .lower import Lowerer
from .scalar import Graph, SNode, C, P, I, J, IJ, topo_order

CHUNK_MIN_DIM = 100         # reference: 146, plus hand-picked dimensions from 99 (config/chunks.py:19-21)
PH_BASE = 1 << 20           # variable indices of the per-pair placeholders
DC = 32                     # components of the streamed chunk per pipeline step
NJ = 8                      # columns per thread inside a tile

_ELEMENTWISE = {
    "Add", "Subtract", "Mult", "Divide", "Minus", "Square", "Inv", "Exp", "Log", "Sqrt", "Rsqrt", "Sin", "Cos", "Asin", "Acos",
    "Atan", "Abs", "Sign", "Step", "ReLU", "Floor", "XLogX", "SinXDivX", "Pow", "Powf", "Round", "Atan2", "LessThan",
    "LessOrEqual", "Equal", "NotEqual", "IfElse", "Clamp", "ClampInt", "DiffClampInt", "Mod",
}


@dataclass
class ChunkPlan:
    D: int
    roots: List[Node]               # the chunked Sum / Scalprod nodes of the original formula
    bodies: List[Node]              # one scalar formula per root, over component stand-ins Var(ind, 1, 4 + cat)
    post: Reduction                 # the reduction over the post formula (placeholders Var(PH_BASE + r, 1, 3))
    hd_vars: List[Tuple[int, int]]  # (index, category) of the high-dimensional variables


def _has_hd_var(n: Node) -> bool:
    return any(v.params[1] >= CHUNK_MIN_DIM for v in n.variables())


def _component(e: Node, D: int, memo: Dict[int, Optional[Node]]) -> Optional[Node]:
    """Scalar clone of a D-dimensional element-wise expression (one component), or None if it is not chunkable."""
    key = id(e)
    if key in memo:
        return memo[key]
    out: Optional[Node] = None
    if e.dim == 1:
        out = None if _has_hd_var(e) else e                     # a scalar broadcast against the D components
    elif e.dim == D:
        if e.op == "Var":
            ind, _, cat = e.params
            out = Node("Var", (), (ind, 1, 4 + cat), 1) if cat in (0, 1) else None
        elif e.op == "SumT":
            out = _component(e.args[0], D, memo)
        elif e.op in _ELEMENTWISE:
            args = [_component(a, D, memo) for a in e.args]
            if all(a is not None for a in args):
                out = Node(e.op, tuple(args), e.params, 1)
    memo[key] = out
    return out


def plan_chunks(red: Reduction) -> Optional[ChunkPlan]:
    """Decompose ``red`` for the chunked kernel, or None when it does not apply."""
    if red.kind == "Zero":
        return None
    dims = {v.params[1] for v in red.variables() if v.params[1] >= CHUNK_MIN_DIM}
    if len(dims) != 1:
        return None                  # no high-dimensional variable (or several different dimensions)
    D = dims.pop()
    roots: List[Node] = []
    bodies: List[Node] = []
    index: Dict[int, int] = {}
    memo_c: Dict[int, Optional[Node]] = {}
    memo_r: Dict[int, Node] = {}

    def replace(n: Node) -> Node:
        hit = memo_r.get(id(n))
        if hit is not None:
            return hit
        res = None
        if n.op in ("Sum", "Scalprod") and n.args[0].dim == D:
            if n.op == "Sum":
                body = _component(n.args[0], D, memo_c)
            else:
                a, b = _component(n.args[0], D, memo_c), _component(n.args[1], D, memo_c)
                body = None if (a is None or b is None) else Node("Mult", (a, b), (), 1)
            if body is not None and _has_hd_var(n):
                r = index.get(id(n))
                if r is None:
                    r = len(roots)
                    index[id(n)] = r
                    roots.append(n)
                    bodies.append(body)
                res = Node("Var", (), (PH_BASE + r, 1, 3), 1)
        if res is None:
            res = n if not n.args else Node(n.op, tuple(replace(a) for a in n.args), n.params, n.dim)
        memo_r[id(n)] = res
        return res

    post_f = replace(red.formula)
    post_w = replace(red.weight) if red.weight is not None else None
    if not roots or len(roots) > 4:
        return None
    post = Reduction(red.kind, post_f, red.tagI, red.K, post_w)
    if any(v.params[1] >= CHUNK_MIN_DIM for v in post.variables()):
        return None                  # a high-dimensional variable is also used outside the chunked sums
    hd = sorted({(v.params[0], v.params[2]) for r in roots for v in r.variables() if v.params[1] == D})
    return ChunkPlan(D, roots, bodies, post, hd)


def default_sizes(plan: ChunkPlan, dtype: str, stride_low: int) -> Optional[Dict[str, int]]:
    """CTA shape from the shared-memory budget: ROWS resident rows, TR row threads, G column groups."""
    f64 = dtype == "float64"
    esz, lpp = (8, 1) if f64 else (4, 2)
    tagI = plan.post.tagI
    ni = sum(1 for _, cat in plan.hd_vars if cat == tagI)
    nj = len(plan.hd_vars) - ni
    D = plan.D
    for rows in (256, 128, 64, 32, 16):
        TR = rows // lpp
        G = max(1, min(8, 256 // TR))
        T = TR * G
        TJ = G * NJ
        ys = DC + (4 if not f64 else 2)
        rs = rows + (2 if not f64 else 1)
        xs_elems = -(-(ni * D * rs) // 4) * 4                      # the streamed chunk starts 16-byte aligned
        smem = (xs_elems + 2 * max(nj, 1) * TJ * ys + 2 * TJ * stride_low) * esz
        if smem <= 200 * 1024:
            return dict(ROWS=rows, TR=TR, G=G, T=T, TJ=TJ, YS=ys, RS=rs, NI=ni, NJV=nj, smem=smem, XS=xs_elems)
    return None


def generate_chunked(red: Reduction, plan: ChunkPlan, dtype: str, nargs: int, use_fast_math: bool = True,
                     sum_scheme: str = "block_sum", dtype_acc: str = "auto") -> Optional[KernelSource]:
    """Source of the chunked ``kb200_reduce`` / ``kb200_combine`` pair (same kernel signatures as the generic kernel, so the
    same C-ABI launch applies), or None if the formula needs something this scheme does not cover."""
    f64 = dtype == "float64"
    fast = bool(use_fast_math) and not f64
    post = plan.post
    tagI = post.tagI
    D = plan.D
    vec = 2 if f64 else 4
    vty = "double2" if f64 else "float4"
    lpp = 1 if f64 else 2

    # ---- one scalar DAG for the post formula and the chunk bodies (shared hoisting) -------------------------
    g = Graph(fast=fast, tagI=tagI)
    L = Lowerer(g)
    low = lower_reduction(post, fast, g, L)
    if low is None:
        return None
    rwt = _Rewriter(g)
    body_terms = []
    for bnode in plan.bodies:
        sv = L(bnode)[0]
        body_terms.append(_acc_term(rwt, sv))
    body_roots = [n for _, nodes, _ in body_terms for n in nodes]
    post_roots: List[SNode] = list(low.value_nodes)
    for _, nodes, _ in low.acc_terms:
        post_roots.extend(nodes)
    post_roots.extend(h for h in low.row_scales if h is not None)

    # k-dependence: does a node depend on the current component of a high-dimensional variable?
    kdep: Dict[int, bool] = {}
    order_all = topo_order(post_roots + body_roots)
    for n in order_all:
        if n.op == "load":
            kdep[n.uid] = n.val[2] in (4, 5)
        else:
            kdep[n.uid] = any(kdep[a.uid] for a in n.args)
    if any(kdep[n.uid] for n in topo_order(post_roots)):
        return None
    for n in topo_order(body_roots):
        if not kdep[n.uid] and n.cls in (J, IJ) and n.op != "const":
            return None              # a low-dimensional streamed scalar inside the chunked sum: not covered

    # ---- low-dimensional variables (post formula and scalar factors of the bodies) ------------------------
    varinfo: Dict[int, Tuple[int, int]] = {}
    for n in order_all:
        if n.op == "load" and n.val[2] in (0, 1, 2):
            ind, _, cat = n.val
            varinfo[ind] = (0, cat)
    for f in post.all_formulas() + plan.bodies:
        for v in f.variables():
            ind, dim, cat = v.params
            if cat in (0, 1, 2):
                varinfo[ind] = (dim, cat)
    for ind in list(varinfo):
        if ind >= nargs:
            return None
    row_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == tagI)
    col_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == 1 - tagI)
    par_vars = sorted((ind, d) for ind, (d, c) in varinfo.items() if c == 2)
    stride = sum(d for _, d in col_vars)
    stride = (max(4, (stride + 3) // 4 * 4) if not f64 else max(2, (stride + 1) // 2 * 2)) if col_vars else vec
    col_off: Dict[int, int] = {}
    off = 0
    for ind, d in col_vars:
        col_off[ind] = off
        off += d
    sz = default_sizes(plan, dtype, stride)
    if sz is None:
        return None
    hd_i = [ind for ind, cat in plan.hd_vars if cat == tagI]
    hd_j = [ind for ind, cat in plan.hd_vars if cat != tagI]
    cfg = KernelConfig(T=sz["T"], R=lpp, TJ=sz["TJ"], unroll=1, min_blocks=1)

    # ---- statements per region -----------------------------------------------------------------------------
    pr = _Printer(f64)
    pre: List[str] = []          # hoisted: constants, parameters, row-side values that do not depend on k
    body_j: List[str] = []       # post stage, per column
    body_ij: List[str] = []      # post stage, per pair
    k_i: List[str] = []          # chunk loop, per component: row-side values
    k_ij: List[str] = []         # chunk loop, per (component, column)
    used_col: Dict[Tuple[int, int], str] = {}
    for n in order_all:
        if n.op == "const":
            continue
        if n.op == "load":
            ind, comp, cat = n.val
            if cat == 3:
                nm = "tph%d" % (ind - PH_BASE)
            elif cat == 4 + tagI:
                nm = "xk%d" % ind
            elif cat in (4, 5):
                nm = "yk%d" % ind
            elif n.cls == P:
                nm = "p%d_%d" % (ind, comp)
                pre.append("const real %s = a%d[%d];" % (nm, ind, comp))
            elif n.cls == I:
                nm = "x%d_%d" % (ind, comp)
                pre.append("V %s; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) %s.l[q] = a%d[rowc[q] * %d + %d];"
                           % (nm, nm, ind, varinfo[ind][0], comp))
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
        if kdep[n.uid]:
            (k_i if n.cls == I else k_ij).append(stmt)      # a k-dependent J value is recomputed per column (rare)
        elif n.cls in (C, P, I):
            pre.append(stmt)
        elif n.cls == J:
            body_j.append(stmt)
        else:
            body_ij.append(stmt)

    need_chunks = sorted({(col_off[ind] + comp) // vec for (ind, comp) in used_col})
    comp_names = ["x", "y", "z", "w"]
    row_loads = ["const %s q%d = yrow[%d];" % (vty, c, c) for c in need_chunks]
    for (ind, comp), nm in sorted(used_col.items()):
        pos = col_off[ind] + comp
        row_loads.append("const real %s = q%d.%s;" % (nm, pos // vec, comp_names[pos % vec]))

    red_code = _ReductionCode(post, pr, low.acc_terms, low.value_nodes, sum_scheme, dtype_acc, f64, fast, low.row_scales)
    red_code.seeded = False
    nroots = len(plan.roots)

    def acc_stmt(r: int, jc: str) -> Optional[str]:
        mode, nodes, _ = body_terms[r]
        if mode == "zero":
            return None
        if mode == "fma":
            return "s%d[%s] = kb_fma_v(%s, %s, s%d[%s]);" % (r, jc, pr.ref(nodes[0]), pr.ref(nodes[1]), r, jc)
        v = pr.ref(nodes[0])
        if nodes[0].cls not in (I, IJ):
            v = "kb_bcast_v(%s)" % v
        return "s%d[%s] = kb_add_v(s%d[%s], %s);" % (r, jc, r, jc, v)

    W = _Writer()
    hdr = _kernel_header(W, post, cfg, f64, fast, stride, nargs, varinfo)
    hdr += "\n".join([
        "// chunked mode: reduction %s" % red.canonical(),
        "#define KC_D %d" % D, "#define KC_DC %d" % DC, "#define KC_TR %d" % sz["TR"], "#define KC_G %d" % sz["G"],
        "#define KC_NJ %d" % NJ, "#define KC_ROWS %d" % sz["ROWS"], "#define KC_RS %d" % sz["RS"], "#define KC_YS %d" % sz["YS"],
        "#define KC_NI %d" % max(sz["NI"], 1), "#define KC_NJV %d" % max(sz["NJV"], 1), ""]) + "\n"
    kname, cname = "kb200_reduce", "kb200_combine"
    arg_decl = "".join(", const real* __restrict__ a%d_" % k for k in range(nargs))
    stride_decl = "".join(", const long long bs%d" % k for k in range(nargs))
    aligned16 = (D % vec == 0)

    def stage_chunk(k0: str, buf: str):
        """cp.async the components [k0, k0 + KC_DC) of KB_TJ streamed rows of every high-dimensional streamed variable."""
        for v, ind in enumerate(hd_j):
            if aligned16:
                W("for (int e = tid; e < KB_TJ * (KC_DC / %d); e += KB_T) {" % vec)
                W("  const int jr_ = e / (KC_DC / %d), c_ = (e - jr_ * (KC_DC / %d)) * %d;" % (vec, vec, vec))
                W("  const long long j_ = jt + jr_; const bool ok_ = (j_ < jend) && ((%s) + c_ < KC_D);" % k0)
                W("  kb_cp_async16(&ys[(((%s) * KC_NJV + %d) * KB_TJ + jr_) * KC_YS + c_], a%d + (ok_ ? (j_ * KC_D + (%s) + c_) : 0), ok_);"
                  % (buf, v, ind, k0))
                W("}")
            else:
                W("for (int e = tid; e < KB_TJ * KC_DC; e += KB_T) {")
                W("  const int jr_ = e / KC_DC, c_ = e - jr_ * KC_DC;")
                W("  const long long j_ = jt + jr_; const bool ok_ = (j_ < jend) && ((%s) + c_ < KC_D);" % k0)
                W("  kb_cp_async(&ys[(((%s) * KC_NJV + %d) * KB_TJ + jr_) * KC_YS + c_], a%d + (ok_ ? (j_ * KC_D + (%s) + c_) : 0), ok_);"
                  % (buf, v, ind, k0))
                W("}")

    def chunk_compute(nk: int):
        """Inner code for ``nk`` (4 or 1) consecutive components starting at chunk-relative position kk."""
        W("{")
        W.indent()
        for v, ind in enumerate(hd_i):
            W("V xq%d[%d];" % (ind, nk))
            W("KB_UNROLL_ALL for (int u_ = 0; u_ < %d; u_++) xq%d[u_].p[0] = *reinterpret_cast<const pk_t*>(&xs[(%d * KC_D + k0 + kk + u_) * KC_RS + tr * KB_R]);"
              % (nk, ind, v))
        W("KB_UNROLL_ALL for (int jc = 0; jc < KC_NJ; jc++) {")
        W.indent()
        for v, ind in enumerate(hd_j):
            base = "yb[(%d * KB_TJ + grp * KC_NJ + jc) * KC_YS + kk]" % v
            if nk == 4 and not f64:
                W("const float4 yq%d_ = *reinterpret_cast<const float4*>(&%s);" % (ind, base))
                W("const real yv%d[4] = {yq%d_.x, yq%d_.y, yq%d_.z, yq%d_.w};" % (ind, ind, ind, ind, ind))
            elif nk == 4:
                W("const double2 yq%da_ = *reinterpret_cast<const double2*>(&%s), yq%db_ = *reinterpret_cast<const double2*>(&%s + 2);"
                  % (ind, base, ind, base))
                W("const real yv%d[4] = {yq%da_.x, yq%da_.y, yq%db_.x, yq%db_.y};" % (ind, ind, ind, ind, ind))
            else:
                W("const real yv%d[1] = {%s};" % (ind, base))
        W("KB_UNROLL_ALL for (int u_ = 0; u_ < %d; u_++) {" % nk)
        W.indent()
        for ind in hd_i:
            W("const V xk%d = xq%d[u_];" % (ind, ind))
        for ind in hd_j:
            W("const real yk%d = yv%d[u_];" % (ind, ind))
        for s_ in k_i:
            W(s_)
        for s_ in k_ij:
            W(s_)
        for r in range(nroots):
            st = acc_stmt(r, "jc")
            if st:
                W(st)
        W.dedent()
        W("}")
        W.dedent()
        W("}")
        W.dedent()
        W("}")

    W("extern \"C\" __global__ void __launch_bounds__(KB_T, 1)")
    W("%s(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, real* __restrict__ partial%s%s)"
      % (kname, arg_decl, stride_decl))
    W("{")
    W.indent()
    W("extern __shared__ __align__(16) unsigned char kc_raw[];")
    W("real* const xs = reinterpret_cast<real*>(kc_raw);                         // [KC_NI][KC_D][KC_RS]: resident rows, transposed")
    W("real* const ys = xs + %d;                                                 // [2][KC_NJV][KB_TJ][KC_YS]: streamed chunk" % sz["XS"])
    W("real (*sm)[KB_TJ * KB_STRIDE] = reinterpret_cast<real (*)[KB_TJ * KB_STRIDE]>(ys + 2 * KC_NJV * KB_TJ * KC_YS);   // low-dim tile")
    W("const int tid = threadIdx.x, tr = tid % KC_TR, grp = tid / KC_TR;")
    for k in range(nargs):
        W("const real* __restrict__ a%d = a%d_ + (long long)blockIdx.z * bs%d;" % (k, k, k))
    W("out += (long long)blockIdx.z * nrow * %d;" % post.dim)
    W("partial += (long long)blockIdx.z * gridDim.y * KC_G * nrow * %d;" % red_code.dimpart)
    W("const long long rowbase = (long long)blockIdx.x * KC_ROWS;")
    W("const long long row0 = rowbase + tr * KB_R;")
    W("long long rowc[KB_R];")
    W("KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) { long long r_ = row0 + q; rowc[q] = r_ < nrow ? r_ : nrow - 1; }")
    W("const long long jbeg = (long long)blockIdx.y * col_chunk;")
    W("const long long jend = (jbeg + col_chunk < ncol) ? (jbeg + col_chunk) : ncol;")
    W("// ---- the CTA's rows of the high-dimensional output-side variables: global -> shared, once")
    for v, ind in enumerate(hd_i):
        W("for (int e = tid; e < KC_ROWS * KC_D; e += KB_T) {")
        W("  const int r_ = e / KC_D, k_ = e - r_ * KC_D;")
        W("  const long long rr_ = (rowbase + r_ < nrow) ? (rowbase + r_) : (nrow - 1);")
        W("  xs[(%d * KC_D + k_) * KC_RS + r_] = a%d[rr_ * KC_D + k_];" % (v, ind))
        W("}")
    W("__syncthreads();")
    W("// ---- hoisted: parameters and output-side (row) values")
    for s_ in pre:
        W(s_)
    W("// ---- accumulators")
    red_code.declare(W)
    W("for (long long jt = jbeg; jt < jend; jt += KB_TJ) {")
    W.indent()
    if col_vars:
        _stage(W, col_vars, col_off, "jt", "0")
    for r in range(nroots):
        W("V s%d[KC_NJ];" % r)
        W("KB_UNROLL_ALL for (int jc = 0; jc < KC_NJ; jc++) s%d[jc] = kb_bcast_v((real)0);" % r)
    stage_chunk("0", "0")
    W("kb_cp_commit();")
    W("for (int ci_ = 0, k0 = 0; k0 < KC_D; ci_++, k0 += KC_DC) {")
    W.indent()
    W("kb_cp_wait_all(); __syncthreads();        // chunk ci_ landed and everybody is done with chunk ci_ - 1")
    W("if (k0 + KC_DC < KC_D) {")
    W.indent()
    stage_chunk("k0 + KC_DC", "(ci_ + 1) & 1")
    W.dedent()
    W("}")
    W("kb_cp_commit();")
    W("const int kn = (KC_D - k0 < KC_DC) ? (KC_D - k0) : KC_DC;")
    W("const real* const yb = ys + (ci_ & 1) * (KC_NJV * KB_TJ * KC_YS);")
    W("int kk = 0;")
    W("#pragma unroll 2")
    W("for (; kk + 4 <= kn; kk += 4)")
    chunk_compute(4)
    W("for (; kk < kn; kk++)")
    chunk_compute(1)
    W.dedent()
    W("}")
    W("// ---- post formula + reduction update for this thread's columns of the tile")
    red_code.tile_begin(W)
    W("KB_UNROLL_ALL for (int jc = 0; jc < KC_NJ; jc++) {")
    W.indent()
    W("const int jr = grp * KC_NJ + jc;")
    W("if (jt + jr < jend) {")
    W.indent()
    for r in range(nroots):
        neg = body_terms[r][2]
        W("const V tph%d = %s;" % (r, ("kb_neg_v(s%d[jc])" % r) if neg else ("s%d[jc]" % r)))
    if col_vars:
        W("const %s* yrow = reinterpret_cast<const %s*>(&sm[0][jr * KB_STRIDE]);" % (vty, vty))
        for s_ in row_loads:
            W(s_)
    for s_ in body_j:
        W(s_)
    for s_ in body_ij:
        W(s_)
    red_code.update(W)
    W.dedent()
    W("}")
    W.dedent()
    W("}")
    red_code.tile_end(W)
    W("__syncthreads();                              // the next tile restages ys[0] and the low-dimensional tile")
    W.dedent()
    W("}")
    W("// ---- write back: one partial accumulator per (column split, column group)")
    red_code.finalize(W, row_expr="row0 + q", split_expr="((long long)blockIdx.y * KC_G + grp)",
                      final_expr="(gridDim.y * KC_G == 1)")
    W.dedent()
    W("}")
    W("")
    red_code.combine_kernel(W, cname, tie_by_index=True)

    src = hdr + W.text()
    h = hashlib.sha256((GENERATOR_VERSION + src).encode()).hexdigest()[:16]
    ks = KernelSource(source=src, kernel_name=kname, combine_name=cname, cfg=cfg, dtype=dtype, tagI=tagI, dim=post.dim,
                      dimred=post.dimred, dimpart=red_code.dimpart, nargs=nargs, row_vars=row_vars, col_vars=col_vars,
                      par_vars=par_vars, stride=stride, smem_bytes=sz["smem"], hash=h, dynamic_smem=sz["smem"])
    ks.chunk = dict(sz, D=D, hd_i=hd_i, hd_j=hd_j, align16=aligned16)        # launch-shape information for the binder
    assert sz["smem"] <= MAX_SMEM_BYTES
    return ks
