"""Filtered K-min / Min / ArgMin kernels for squared Euclidean distances in dimension 2..15 (the K-nearest-neighbour search of
BASELINE configs[2]: ``ArgKMin_Reduction(SqDist(x_i, y_j), K)`` in D = 3; ``ArgMin``: the assignment step of K-means).

The reference evaluates ``|x_i - y_j|^2`` as D subtractions, one multiplication and D - 1 FMAs per pair and tests every
value against the row's K-th smallest (``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:70-90``,
``keopscore/keopscore/formulas/reductions/KMin_ArgKMin_Reduction.py:71-100``); the generic kernel of this package does the same
with packed instructions: 2 D packed operations + 1 compare per pair, 0.85 of that roofline (ncu: FMA pipe 75 %, issue 70 %).

Here the pair loop evaluates a *filter* instead: with both clouds shifted by the streamed cloud's mean ``m``,

    s_ij = |y'_j|^2 - 2 x'_i . y'_j   ( = |x_i - y_j|^2 - |x'_i|^2 ),      x' = x - m,  y' = y - m

costs D FMAs per pair -- the streamed row ``(y'_0 .. y'_{D-1}, |y'|^2)`` is one 16-byte shared-memory read, written once per
call by ``kb200_kfpack`` -- and only the smallest value of a row over a group of 8 columns is compared with
``thr_i - |x'_i|^2 + margin_i`` (FMNMX3: two columns per instruction; the next group's FMAs are issued before the test).  Only pairs that pass (a few hundred per
row over the whole stream) are evaluated exactly, with the generic kernel's own packed operations in the reference's order, and
go through the same insertion: **the output is bit-identical to the generic kernel's** (``tests/test_kmin_filter_gpu.py``),
hence to the reference's wherever that one is.

Margin (fp32, round to nearest, flush-to-zero harmless): let ``A = (|x'_i| + max_j |y'_j|)^2``.  The FMA chain and
``|y'|^2`` carry at most 6 * 2^-24 A, ``|x'|^2`` 3 * 2^-24 A, the shift of the coordinates 2 * 2^-24 A, the exact value itself
5 * 2^-24 A against the real ``|x - y|^2``, forming the threshold (``thr_i + (margin_i - |x'_i|^2)``) at most 12 * 2^-24 A: 28 *
2^-24 A in all.  ``margin_i = 2^-18 (|x'_i|^2 + max_j |y'_j|^2) + 1e-30 >= 32 * 2^-24 A``, with the
maximum taken over the streamed tile at hand (1024 columns), so that an outlier loosens the filter of its own tile only.  A pair with ``d_ij < thr_i`` therefore always passes; columns with non-finite
coordinates never pass, exactly as they never satisfy ``d_ij < thr_i``.  For O(1) clouds of 1e6 points the margin widens the
tested radius by a few percent: a handful of extra exact evaluations per row.  In dimension D the same count gives
``(4 D + 18) 2^-24 A`` (D roundings each in the FMA chain, the two squared norms and the exact value, the rest as above with
the threshold counted at its worst); the constant used is ``max(2^-18, (4 D + 20) 2^-23)``.

Column splits, shared thresholds between the splits (``kb200_reduce_seeded``), the partial format and ``kb200_combine`` are the
generic kernel's: the binder launches this source through the same code path with two extra pointer arguments (packed rows,
``[m_0 .. m_{D-1}, pad, max |y'|^2 per block of 256 columns]``).
"""
from __future__ import annotations

import hashlib
from typing import Optional, Tuple

from ..formulas.reductions import Reduction
from .emit import GENERATOR_VERSION, KernelConfig, KernelSource, prelude

VERSION = "kmin-filter-11"
MAX_D = 15                   # from D = 16 the tensor-core filter takes over (codegen/knn_tc.py)
# reduction kind -> (what a row of the output holds, is it a list of K).  Min / ArgMin / Min_ArgMin (nearest neighbour, the
# assignment step of K-means) are the K = 1 case: same strict '<' update, same (value, index) partials.
KINDS = {"KMin": ("val", True), "ArgKMin": ("idx", True), "KMin_ArgKMin": ("both", True),
         "Min": ("val", False), "ArgMin": ("idx", False), "Min_ArgMin": ("both", False)}


def match(red: Reduction) -> Optional[Tuple[int, int, int]]:
    """``(row_var, col_var, D)`` for a K-min family reduction of ``Sum(Square(Var - Var))`` between an output-side and a
    streamed variable of dimension <= 3 (either order of the subtraction: the squares are the same)."""
    if red.kind not in KINDS or red.dimF != 1 or not (1 <= red.K <= 16):
        return None
    f = red.formula
    if f.op != "Sum" or len(f.args) != 1 or f.args[0].op != "Square":
        return None
    sub = f.args[0].args[0]
    if sub.op != "Subtract" or len(sub.args) != 2 or any(a.op != "Var" for a in sub.args):
        return None
    (ia, da, ca), (ib, db, cb) = sub.args[0].params, sub.args[1].params
    if da != db or {ca, cb} != {0, 1} or not (1 <= da <= MAX_D):
        return None
    row, col = (ia, ib) if ca == red.tagI else (ib, ia)
    return row, col, da


def generate_filtered(red: Reduction, nargs: int, cfg: KernelConfig, generic: KernelSource) -> Optional[KernelSource]:
    """The filtered kernel for ``red`` (fp32, fast math), same entry points and calling convention as ``generic`` plus two
    pointer arguments; ``None`` when the reduction does not have the shape."""
    m = match(red)
    if m is None or generic.dtype != "float32" or generic.is_zero:
        return None
    ix, iy, D = m
    first_is_row = red.formula.args[0].args[0].args[0].params[0] == ix      # x - y or y - x in the formula
    what, is_list = KINDS[red.kind]
    K, R, T, TJ = (red.K if is_list else 1), cfg.R, cfg.T, cfg.TJ
    assert R % 2 == 0 and TJ % 256 == 0
    F4, C4 = -(-(D + 1) // 4), -(-D // 4)          # float4 per column: filter row (y', |y'|^2), coordinates
    S4 = F4 + C4
    GC = 8 if D <= 3 else 4                        # columns per group: one compare per row and group
    # margin constant: in units of 2^-24 (|x'| + |y'|)^2 the filter and the exact value differ by at most D (FMA chain) + D (|y'|^2)
    # + D (|x'|^2) + 4 (shifted coordinates) + D + 2 (the exact value against the real |x - y|^2) + 12 (forming thr + tbase + ctmax
    # when thr <= 4 (|x'| + |y'|)^2; a larger thr passes trivially) = 4 D + 18; (|x'|^2 + max |y'|^2) >= (|x'| + |y'|)^2 / 2, so
    # (4 D + 20) 2^-23 covers it.  D <= 3: 2^-18 = 64 2^-24 (>= 2 x 30).
    cm = max(2.0 ** -18, (4 * D + 20) * 2.0 ** -23)
    CM = "%.9ef" % cm
    comp = lambda src, k: "%s%d.%s" % (src, k // 4, "xyzw"[k % 4])      # component k of a group of float4 registers
    NA = nargs + 2                         # a<nargs> = packed rows, a<nargs + 1> = [m_0 .. m_{D-1}, pad, block maxima of |y'|^2]
    dimout = 2 * K if what == "both" else K
    cname = generic.combine_name
    marker = 'extern "C" __global__ void __launch_bounds__(256)\n' + cname
    if marker not in generic.source:
        return None
    combine_src = generic.source[generic.source.index(marker):]          # identical partial format -> identical combine

    ptr_args = ", ".join("const real* __restrict__ a%d_" % k for k in range(NA))
    str_args = ", ".join("const long long bs%d" % k for k in range(NA))
    call_args = ", ".join(["a%d_" % k for k in range(NA)] + ["bs%d" % k for k in range(NA)])

    def sub(a: str, b: str) -> str:        # the subtraction of the formula, in its own order
        return "kb_sub_v(%s, %s)" % ((a, b) if first_is_row else (b, a))

    L = []
    W = L.append
    W("#define KB_T %d\n#define KB_R %d\n#define KB_TJ %d\n#define KB_STRIDE %d\n#define KB_F64 0\n#define KB_FAST 1" % (T, R, TJ, 4 * S4))
    W(prelude())
    # ---- pack kernel -------------------------------------------------------------------------------------------------
    W("// one thread per streamed point: shifted coordinates + squared norm as one 16-byte row; max |y'|^2 over the finite rows")
    W('extern "C" __global__ void __launch_bounds__(256)')
    W("kb200_kfpack(const long long ncol, const real* __restrict__ y, real* __restrict__ ypk, real* __restrict__ info)")
    W("{")
    W("  // info = [m_0 .. m_{D-1}, pad to 16, then one maximum of |y'|^2 per block of 256 columns (zero-initialised)]")
    W("  const long long j_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
    W("  real s_ = 0.0f;")
    W("  if (j_ < ncol) {")
    W("    real r_[%d], o_[%d];" % (4 * F4, 4 * C4))
    W("    KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) r_[k_] = 0.0f;" % (4 * F4))
    W("    KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) o_[k_] = 0.0f;" % (4 * C4))
    W("    KB_UNROLL_ALL for (int k_ = 0; k_ < %d; k_++) { o_[k_] = y[j_ * %d + k_]; r_[k_] = o_[k_] - info[k_]; }" % (D, D))
    W("    s_ = r_[0] * r_[0];")
    W("    KB_UNROLL_ALL for (int k_ = 1; k_ < %d; k_++) s_ = fmaf(r_[k_], r_[k_], s_);" % D)
    W("    r_[%d] = s_;" % D)
    W("    float4* dst_ = reinterpret_cast<float4*>(ypk + j_ * KB_STRIDE);")
    W("    KB_UNROLL_ALL for (int v_ = 0; v_ < %d; v_++) dst_[v_] = make_float4(r_[4 * v_], r_[4 * v_ + 1], r_[4 * v_ + 2], r_[4 * v_ + 3]);   // the filter's row" % F4)
    W("    KB_UNROLL_ALL for (int v_ = 0; v_ < %d; v_++) dst_[%d + v_] = make_float4(o_[4 * v_], o_[4 * v_ + 1], o_[4 * v_ + 2], o_[4 * v_ + 3]);   // the untouched coordinates (exact path)" % (C4, F4))
    W("    if (!(s_ < KB_INF)) s_ = 0.0f;                 // non-finite rows never pass the filter: keep them out of the margin")
    W("  }")
    W("  KB_UNROLL_ALL for (int o_ = 16; o_ > 0; o_ >>= 1) s_ = fmaxf(s_, __shfl_xor_sync(0xffffffffu, s_, o_));")
    W("  if ((threadIdx.x & 31) == 0 && s_ > 0.0f) atomicMax(reinterpret_cast<int*>(info) + 16 + blockIdx.x, __float_as_int(s_));")
    W("}")
    W("")
    # ---- reduce body -------------------------------------------------------------------------------------------------
    W("__device__ __forceinline__ void kb200_reduce_body(const long long nrow, const long long ncol, const long long col_chunk, "
      "real* __restrict__ out, real* __restrict__ partial, %s, %s, real* __restrict__ shthr, const int force_partial)" % (ptr_args, str_args))
    W("{")
    W("  // two tiles of KB_TJ columns x KB_STRIDE reals: dynamic shared memory (it can pass the 48 KB static limit)")
    W("  extern __shared__ __align__(16) real kb_sm_dyn[];")
    W("  real (*sm)[KB_TJ * KB_STRIDE] = reinterpret_cast<real (*)[KB_TJ * KB_STRIDE]>(kb_sm_dyn);")
    W("  const int tid = threadIdx.x;")
    W("  const real* __restrict__ ax = a%d_ + (long long)blockIdx.z * bs%d;" % (ix, ix))
    W("  const real* __restrict__ ay = a%d_ + (long long)blockIdx.z * bs%d;" % (iy, iy))
    W("  const real* __restrict__ ypk = a%d_;" % nargs)
    W("  const real* __restrict__ info = a%d_;" % (nargs + 1))
    W("  out += (long long)blockIdx.z * nrow * %d;" % dimout)
    W("  partial += (long long)blockIdx.z * gridDim.y * nrow * %d;" % (2 * K))
    W("  const long long row0 = (long long)blockIdx.x * (KB_T * KB_R) + tid;")
    W("#define KF_ROW(q) ((row0 + (long long)(q) * KB_T < nrow) ? (row0 + (long long)(q) * KB_T) : (nrow - 1))   /* recomputed, not kept */")
    W("  const long long jbeg = (long long)blockIdx.y * col_chunk;")
    W("  const long long jend = (jbeg + col_chunk < ncol) ? (jbeg + col_chunk) : ncol;")
    W("  // ---- hoisted row values: -2 x' (filter) and tbase = 2^-18 |x'|^2 + 1e-30 - |x'|^2; the filter threshold of a tile is")
    W("  // thr + tbase + 2^-18 max_{j in tile} |y'_j|^2 (the exact path re-reads x: registers are scarce)")
    W("  V tbase;")
    W("  { V xsq = kb_bcast_v(0.0f);")
    for k in range(D):
        W("    KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) { const real xs_ = ax[KF_ROW(q) * %d + %d] - info[%d]; xsq.l[q] = fmaf(xs_, xs_, xsq.l[q]); }" % (D, k, k))
    W("    KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) tbase.l[q] = fmaf(%s, xsq.l[q], 1e-30f) - xsq.l[q]; }" % CM)
    for k in range(D):
        W("  V m2x%d; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) m2x%d.l[q] = -2.0f * (ax[KF_ROW(q) * %d + %d] - info[%d]);" % (k, k, D, k, k))
    W("  real ctmax = 0.0f;                               // 2^-18 x the largest |y'|^2 of the current tile")
    W("  // ---- accumulators: K smallest values + indices per row, strict threshold thr, filter threshold thrp")
    W("  real kval[KB_R][%d]; int kidx[KB_R][%d]; V thr = kb_bcast_v(KB_INF), thrp;" % (K, K))
    W("  KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { kval[q][l_] = KB_INF; kidx[q][l_] = 0; }" % K)
    W("#define KF_THRP(q) thrp.l[q] = (thr.l[q] + tbase.l[q]) + ctmax")
    W("  real* const shp = (shthr != nullptr) ? (shthr + row0) : nullptr;")
    W("  if (shp != nullptr) {")
    W("    KB_UNROLL_ALL for (int q = 0; q < KB_R; q++)")
    W("      if (row0 + (long long)q * KB_T < nrow) thr.l[q] = kb_next_up(__ldcg(shp + q * KB_T)); }")
    W("  KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) KF_THRP(q);")

    def stage(jbase: str, buf: str):
        W("    for (int e = tid; e < %d * KB_TJ; e += KB_T) {" % S4)
        if S4 & (S4 - 1) == 0:
            W("      const int c_ = e >> %d, v_ = e & %d;" % (S4.bit_length() - 1, S4 - 1))
        else:
            W("      const int c_ = e / %d, v_ = e - c_ * %d;" % (S4, S4))
        W("      const long long j_ = (%s) + c_;" % jbase)
        W("      const bool ok_ = j_ < jend;")
        W("      kb_cp_async16(&sm[%s][e * 4], ypk + (ok_ ? j_ * KB_STRIDE + v_ * 4 : 0), ok_);" % buf)
        W("    }")

    def load_row(pre: str, name: str, col: str, first: int, count: int):
        # `count` float4 of the staged column `col` (tile row index), starting at float4 `first` of the column
        W(pre + " ".join("const float4 %s%d = yrow[%d * (%s) + %d];" % (name, v, S4, col, first + v) for v in range(count)))

    def filt(src: str) -> str:
        e = comp(src, D)                    # |y'|^2
        for k in reversed(range(D)):
            e = "kb_fma_v(m2x%d, %s, %s)" % (k, comp(src, k), e)
        return e

    def load_x(pre: str):
        # output-side coordinates of the thread's rows (L1 / L2): once per rare-path entry, not kept across the pair loop
        for k in range(D):
            W(pre + "V x%d_; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) x%d_.l[q] = __ldg(ax + KF_ROW(q) * %d + %d);" % (k, k, D, k))

    def exact(col: str, trow: str, pre: str):
        # the generic kernel's own evaluation, on all KB_R rows of the thread, then the same test and insertion
        W(pre + "{ const long long jc_ = %s;" % col)
        load_row(pre + "  ", "yo", trow, F4, C4)
        for k in range(D):
            W(pre + "  const V d%d_ = %s;" % (k, sub("x%d_" % k, comp("yo", k))))
        W(pre + "  V f_ = kb_mul_v(d0_, d0_);")
        for k in range(1, D):
            W(pre + "  f_ = kb_fma_v(d%d_, d%d_, f_);" % (k, k))
        W(pre + "  KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) if (f_.l[q] < thr.l[q]) {")
        W(pre + "    kb_kmin_insert<%d, int>(kval[q], kidx[q], f_.l[q], (int)jc_); thr.l[q] = kb_fmin(kval[q][%d], thr.l[q]); KF_THRP(q); } }" % (K, K - 1))

    W("  // smallest filter value of every row over the %d columns starting at tile row `r0_` (FMNMX3: two columns per instruction)" % GC)
    W("#define KF_GROUP_MIN(dst, r0_) { const float4* yg_ = yrow + %d * (r0_); dst = kb_bcast_v(KB_INF); \\" % S4)
    for c in range(0, GC, 2):
        la = " ".join("const float4 qa%d = yg_[%d];" % (v, S4 * c + v) for v in range(F4))
        lb = " ".join("const float4 qb%d = yg_[%d];" % (v, S4 * (c + 1) + v) for v in range(F4))
        W("    { %s %s const V fa_ = %s; const V fb_ = %s; \\" % (la, lb, filt("qa"), filt("qb")))
        W("      KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) dst.l[q] = fminf(fminf(dst.l[q], fa_.l[q]), fb_.l[q]); } \\")
    W("  }")
    W("  // ---- stage the first tile")
    W("  {")
    stage("jbeg", "0")
    W("  }")
    W("  kb_cp_commit(); kb_cp_wait_all(); __syncthreads();")
    W("  int buf = 0;")
    W("  for (long long jt = jbeg; jt < jend; jt += KB_TJ) {")
    W("    if (jt + KB_TJ < jend) {")
    stage("jt + KB_TJ", "buf ^ 1")
    W("    }")
    W("    kb_cp_commit();")
    W("    const int jn = (int)((jend - jt < KB_TJ) ? (jend - jt) : KB_TJ);")
    W("    { real tm_ = 0.0f;                               // per-tile margin: an outlier only loosens the filter of its own tile")
    W("      KB_UNROLL_ALL for (int b_ = 0; b_ < KB_TJ / 256; b_++) { const long long jb_ = jt / 256 + b_; if (jb_ * 256 < ncol) tm_ = fmaxf(tm_, __ldg(info + 16 + jb_)); }")
    W("      ctmax = %s * tm_;" % CM)
    W("      KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) KF_THRP(q); }")
    W("    const float4* yrow = reinterpret_cast<const float4*>(&sm[buf][0]);      // %d float4 per column: filter row, coordinates" % S4)
    W("    const int jg = jn & ~%d;                        // whole groups" % (GC - 1))
    W("    // software pipeline: the next group's filter values are issued before this group's minima are tested, so that the")
    W("    // compare / branch chain overlaps independent FFMA2 work (ncu of the unpipelined loop: issue 54 %, fixed-latency waits)")
    W("    V mn_;")
    W("    if (jg > 0) KF_GROUP_MIN(mn_, 0)")
    W("    #pragma unroll 1")
    W("    for (int jr = 0; jr < jg; jr += %d) {" % GC)
    W("      V nx_ = kb_bcast_v(KB_INF);")
    W("      if (jr + %d < jg) KF_GROUP_MIN(nx_, jr + %d)" % (GC, GC))
    W("      bool any_ = false;")
    W("      KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (mn_.l[q] < thrp.l[q]);")
    W("      if (any_) {")
    W("        // rare (~5 % of the groups): which columns of the group pass?  All eight filter values again, straight-line (the")
    W("        // looped version serialised eight shared-memory round trips), then the exact evaluation of the few that do")
    W("        unsigned hits_ = 0u;")
    W("        KB_UNROLL_ALL for (int g_ = 0; g_ < %d; g_++) {" % GC)
    load_row("          ", "q", "jr + g_", 0, F4)
    W("          const V fv_ = %s;" % filt("q"))
    W("          bool hit_ = false;")
    W("          KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) hit_ |= (fv_.l[q] < thrp.l[q]);")
    W("          hits_ |= hit_ ? (1u << g_) : 0u;")
    W("        }")
    W("        if (hits_) {")
    load_x("          ")
    W("          #pragma unroll 1")
    W("          for (; hits_; hits_ &= hits_ - 1) {")
    W("            const int g_ = __ffs(hits_) - 1;")
    exact("jt + jr + g_", "jr + g_", "            ")
    W("          }")
    W("        }")
    W("      }")
    W("      mn_ = nx_;")
    W("    }")
    W("    #pragma unroll 1")
    W("    for (int jr = jg; jr < jn; jr++) {")
    load_row("      ", "q", "jr", 0, F4)
    W("      const V fv_ = %s;" % filt("q"))
    W("      bool any_ = false;")
    W("      KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) any_ |= (fv_.l[q] < thrp.l[q]);")
    W("      if (any_) {")
    load_x("        ")
    exact("jt + jr", "jr", "        ")
    W("      }")
    W("    }")
    W("    if (shp != nullptr) {")
    W("      KB_UNROLL_ALL for (int q = 0; q < KB_R; q++)")
    W("        if (row0 + (long long)q * KB_T < nrow) {")
    W("          real* a_ = shp + q * KB_T; real s_ = __ldcg(a_);")
    W("          if (kval[q][%d] < s_) { kb_atomic_min(a_, kval[q][%d]); s_ = kval[q][%d]; }" % (K - 1, K - 1, K - 1))
    W("          thr.l[q] = kb_fmin(thr.l[q], kb_next_up(s_)); KF_THRP(q); } }")
    W("    kb_cp_wait_all(); __syncthreads();")
    W("    buf ^= 1;")
    W("  }")
    W("#undef KF_GROUP_MIN")
    W("  // ---- write back")
    W("  KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) {")
    W("    const long long r_ = row0 + (long long)q * KB_T;")
    W("    if (r_ < nrow) {")
    W("      if (gridDim.y == 1 && !force_partial) {")
    W("        real* po = out + r_ * %d;" % dimout)
    if what == "idx":
        W("        KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) po[l_] = (real)kidx[q][l_];" % K)
    elif what == "val":
        W("        KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) po[l_] = kval[q][l_];" % K)
    else:
        W("        KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { po[l_ * 2] = kval[q][l_]; po[l_ * 2 + 1] = (real)kidx[q][l_]; }" % K)
    W("      } else {")
    W("        real* po = partial + ((long long)blockIdx.y * nrow + r_) * %d;" % (2 * K))
    W("        KB_UNROLL_ALL for (int l_ = 0; l_ < %d; l_++) { po[l_ * 2] = kval[q][l_]; po[l_ * 2 + 1] = (real)kidx[q][l_]; }" % K)
    W("      }")
    W("    }")
    W("  }")
    W("#undef KF_THRP")
    W("#undef KF_ROW")
    W("}")
    W("")
    sig = ("(const long long nrow, const long long ncol, const long long col_chunk, real* __restrict__ out, "
           "real* __restrict__ partial, %s, %s" % (ptr_args, str_args))
    W('extern "C" __global__ void __launch_bounds__(KB_T, %d)' % cfg.min_blocks)
    W("%s%s)" % (generic.kernel_name, sig))
    W("{ kb200_reduce_body(nrow, ncol, col_chunk, out, partial, %s, nullptr, 0); }" % call_args)
    W('extern "C" __global__ void __launch_bounds__(KB_T, %d)' % cfg.min_blocks)
    W("kb200_reduce_seeded%s, real* __restrict__ shthr, const int force_partial)" % sig)
    W("{ kb200_reduce_body(nrow, ncol, col_chunk, out, partial, %s, shthr, force_partial); }" % call_args)
    W("")
    W(combine_src)
    src = "\n".join(L)
    h = hashlib.sha256((GENERATOR_VERSION + VERSION + src).encode()).hexdigest()[:16]
    ks = KernelSource(source=src, kernel_name=generic.kernel_name, combine_name=cname, cfg=cfg, dtype="float32",
                      tagI=red.tagI, dim=generic.dim, dimred=generic.dimred, dimpart=2 * K, nargs=NA,
                      row_vars=generic.row_vars, col_vars=generic.col_vars, par_vars=generic.par_vars, stride=4 * S4,
                      smem_bytes=2 * TJ * 4 * S4 * 4, hash=h, dynamic_smem=2 * TJ * 4 * S4 * 4)
    ks.kfilter = dict(row=ix, col=iy, D=D, stride=4 * S4)
    return ks
