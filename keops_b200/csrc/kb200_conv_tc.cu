// kb200_conv_tc.cu -- Gaussian-kernel Sum reductions  out_i = sum_j exp(c |x_i - y_j|^2) b_j  in medium / large
// dimension on the 5th-generation tensor cores of sm_100a (tcgen05.mma, TMEM accumulators, TMA operand tiles).
//
// Reference semantics: Sum_Reduction (keopscore/keopscore/formulas/reductions/Sum_Reduction.py:10-50) of
// Exp(c * Sum(Square(x - y))) * b evaluated pair by pair on CUDA cores by GpuReduc1D
// (keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:70-90): 2 D + 2 FP32 instructions + 1 MUFU.EX2 per pair
// (34 + 1 at D = 16, the KernelSolve configuration).  Here the squared distance is a dense contraction
//
//     |x - y|^2 = |x|^2 + |y|^2 - 2 x.y ,      acc_ij = x_i . y_j - |y_j|^2 / 2      (one tcgen05 GEMM)
//
// so a pair costs 1 FFMA (exponent) + 1 MUFU.EX2 + DB FFMA (accumulation) on the CUDA cores: the kernel is bound by
// the MUFU pipe (16 ex2 / clk / SM) instead of the FP32 pipe.
//
// Precision (the "stated tolerance" of the tensor-core path): every fp32 coordinate is split
// into three bf16 terms v = v1 + v2 + v3 (8 + 8 + 8 significant bits: exact), and the GEMM accumulates the six
// products  x1 y1, x1 y2, x2 y1, x1 y3, x2 y2, x3 y1  in fp32 -- everything down to 2^-24 |x_k y_k|; the dropped
// products (x2 y3, x3 y2, x3 y3) are <= 2^-23 |x_k y_k|.  -|y|^2/2 enters the same accumulation through three
// more K columns (a three-term bf16 split against 1, 1, 1 on the x side).  Both clouds are shifted by a common
// vector (the caller passes the mean of the streamed cloud), which leaves x - y unchanged and keeps |x|^2 + |y|^2
// -- hence the cancellation error  ~2^-22 (|x|^2 + |y|^2) |c| log2(e)  of the exponent -- small.
//
// kb200_conv_prep : fp32 rows -> the 6 D + 3 bf16 operand columns (zero padded to CT_DPG) + |x - shift|^2;
//                   on the streamed side also a zero-padded copy of b (whole 128-column tiles).
// kb200_conv_tc   : one CTA = 256 output rows (two 128-row accumulators) x all streamed columns in 128-column
//                   tiles.  Warp 0 = TMA producer (operand tiles + the b tile), warp 1 = MMA issuer, warps 2..17 =
//                   epilogue: a thread owns one output row and one 64-column half of every tile (two threads per
//                   row; 4 epilogue warps per scheduler hide the TMEM-load and barrier latencies behind the other
//                   warps' ex2 stream): tcgen05.ld of its 64 accumulators, ex2, FFMA with b_j (broadcast LDS.128),
//                   per-tile partial sums added to the running sum (the reference's block_sum scheme,
//                   sum_schemes.py:55-71); the two halves of a row are added through shared memory at the end.
//                   4-stage smem ring, double-buffered TMEM accumulators.
//
// Compile-time parameters (keops_b200/codegen/conv_tc.py): CT_D, CT_DB (dimension of b), CT_KBM / CT_REM (operand
// rows = CT_KBM 128-byte-swizzled 64-column blocks + one remainder block of CT_REM columns), CT_STAGES, CT_HS.
// The PTX wrappers (kt_*) come from kb200_tc_common.cuh, prepended by the host code.

#ifndef CT_D
#define CT_D 16
#define CT_DB 1
#define CT_KBM 1
#define CT_REM 64
#endif
#ifndef CT_STAGES
#define CT_STAGES 4
#endif
#ifndef CT_LSE
#define CT_LSE 0                              /* 1: Max_SumShiftExpWeight reduction  (m_i, s_i) = (max_j F_ij, sum_j exp(F_ij - m_i) b_j)
                                                 of F_ij = a_j + c |x_i - y_j|^2  (reference: Max_SumShiftExpWeight_Reduction.py:16-123);
                                                 a_j rides in the penalty columns, the running max lives in base-2 units   */
#endif
#define CT_OUTW (CT_DB + CT_LSE)              /* output floats per row: [m ,] s                                   */
#ifndef CT_HS
#define CT_HS 8                               /* slots of the b-tile ring                                        */
#endif
#ifndef CT_DBG
#define CT_DBG 0                              /* timing experiments: 1 no ex2, 2 no TMEM loads                   */
#endif
#define CT_BM 256                             /* rows per CTA: two 128-row accumulators                          */
#define CT_BN 128                             /* columns per tile (= N of one tcgen05.mma)                       */
#ifndef CT_ATMEM2
#define CT_ATMEM2 0                           /* kb200_conv_tc (256 rows), CT_REM == 64 only: both resident row blocks in
                                                 tensor memory (TS-form MMAs), accumulators single-buffered: 2 x 128
                                                 accumulator + 2 x CT_DPG / 2 operand columns                            */
#endif
#define CT_NBUF (CT_ATMEM2 ? 1 : 2)           /* TMEM accumulator buffers per row block (2: 2 x 2 x 128 = 512 cols)      */
#define CT_TMEM_A2 256                        /* CT_ATMEM2: first operand column                                        */
#ifndef CT_PEN
#define CT_PEN 1                              /* 1: -|y|^2/2 rides in three extra K columns; 0: |y|^2 is added in the
                                                 epilogue (one more FFMA per pair) when that saves a K block       */
#endif
#define CT_NCOL (6 * CT_D + 3 * CT_PEN)       /* used K columns                                                  */
#define CT_W (CT_DB + 1 - CT_PEN)             /* floats per column of the b tile: [ |y|^2 (CT_PEN == 0) , b ]      */
#define CT_DPG (64 * CT_KBM + CT_REM)
#define CT_TILE_BYTES(rows) ((rows) * 128)
#define CT_RTILE_BYTES(rows) ((rows) * CT_REM * 2)
#define CT_BLK_BYTES(rows) (CT_KBM * CT_TILE_BYTES(rows) + CT_RTILE_BYTES(rows))
#define CT_A_BYTES (CT_ATMEM2 ? 0 : 2 * CT_BLK_BYTES(128))
#define CT_B_BYTES CT_BLK_BYTES(CT_BN)
#define CT_H_BYTES (CT_BN * CT_W * 4)         /* one b tile                                                      */
#define CT_X_BYTES (CT_BM * CT_OUTW * 4)      /* exchange of the second column half's sums at the end            */
#define CT_EPI_WARPS 16
#define CT_THREADS (64 + 32 * CT_EPI_WARPS)
#define CT_HALF (CT_BN / 2)                   /* columns per epilogue thread and tile                            */
#define CT_SMEM_BYTES (1024 + CT_A_BYTES + CT_STAGES * CT_B_BYTES + CT_HS * CT_H_BYTES + CT_X_BYTES + 640) /* mirrored in conv_tc.py */
#define CT_LOG2E 1.4426950408889634f
// bf16 x bf16 -> f32, A and B K-major, M = 128, N = CT_BN
#define CT_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | (((u32)CT_BN >> 3) << 17) | ((128u >> 4) << 24))

// ------------------------------------------------------------------------------------------------
// prep
//   side 0 (streamed, y): [ y1 | y2 | y1 | y3 | y2 | y1 | -h1 -h2 -h3 | 0 .. ],  h = |y - shift|^2 / 2
//                         + bpad (npad x CT_W): [ |y - shift|^2 if CT_PEN == 0 , b ] for rows < n, 0 for the padding
//                           of the last tile
//   side 1 (output,   x): [ x1 | x1 | x2 | x1 | x2 | x3 |  1   1   1  | 0 .. ],  sq[r] = |x - shift|^2
//   (v1, v2, v3 = successive bf16 round-to-nearest remainders of the shifted coordinate)
// ------------------------------------------------------------------------------------------------
// One warp per row: lanes stride over the coordinates, so that the fp32 row is read and each of the six bf16 blocks is
// written with coalesced accesses (a thread per row scatters 2-byte stores 2 * CT_DPG bytes apart: at D = 784 the prep of
// 1e5 points took longer than a tenth of the main kernel).  The squared norm is a fixed-order lane tree (deterministic).
extern "C" __global__ void __launch_bounds__(256)
kb200_conv_prep(const long long n, const long long npad, const float* __restrict__ a, const float* __restrict__ shift,
                unsigned short* __restrict__ abf, float* __restrict__ sq, const float* __restrict__ b,
                float* __restrict__ bpad, const int side, const float* __restrict__ acol,
                const float* __restrict__ coef) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (side == 0 && r < npad) {
    for (int d = lane; d < CT_DB; d += 32) bpad[r * CT_W + (1 - CT_PEN) + d] = (r < n) ? b[r * CT_DB + d] : 0.0f;
    if (CT_PEN == 0 && r >= n && lane == 0) bpad[r * CT_W] = 0.0f;
  }
  if (r >= n) return;
  const float* row = a + r * CT_D;
  unsigned short* orow = abf + r * CT_DPG;
  float s = 0.0f;
  for (int k = lane; k < CT_D; k += 32) {
    const float v = row[k] - (shift ? shift[k] : 0.0f);
    s = fmaf(v, v, s);
    const unsigned short p1 = kt_bf16_rn(v);
    const float r1 = v - kt_bf16_to_f(p1);
    const unsigned short p2 = kt_bf16_rn(r1);
    const unsigned short p3 = kt_bf16_rn(r1 - kt_bf16_to_f(p2));
    if (side == 0) {
      orow[k] = p1; orow[CT_D + k] = p2; orow[2 * CT_D + k] = p1;
      orow[3 * CT_D + k] = p3; orow[4 * CT_D + k] = p2; orow[5 * CT_D + k] = p1;
    } else {
      orow[k] = p1; orow[CT_D + k] = p1; orow[2 * CT_D + k] = p2;
      orow[3 * CT_D + k] = p1; orow[4 * CT_D + k] = p2; orow[5 * CT_D + k] = p3;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int k = CT_NCOL + lane; k < CT_DPG; k += 32) orow[k] = 0;
  if (lane != 0) return;
  if (CT_PEN == 0) {
    if (side == 0) bpad[r * CT_W] = s;
  } else if (side == 0) {
    // LSE: F_ij = a_j + c |x - y|^2 = -2c (x.y - |y|^2/2 - a_j / (2c)) + c |x|^2 : the column term joins the penalty
    const float h = 0.5f * s + ((CT_LSE && acol != nullptr) ? acol[r] / (2.0f * coef[0]) : 0.0f);
    const unsigned short hi = kt_bf16_rn(h);
    const float r1 = h - kt_bf16_to_f(hi);
    const unsigned short mid = kt_bf16_rn(r1);
    const unsigned short lo = kt_bf16_rn(r1 - kt_bf16_to_f(mid));
    orow[6 * CT_D] = hi ^ 0x8000u; orow[6 * CT_D + 1] = mid ^ 0x8000u; orow[6 * CT_D + 2] = lo ^ 0x8000u;   // negated
  } else {
    orow[6 * CT_D] = 0x3F80u; orow[6 * CT_D + 1] = 0x3F80u; orow[6 * CT_D + 2] = 0x3F80u;                   // 1.0
  }
  sq[r] = s;
}

// ------------------------------------------------------------------------------------------------
// epilogue of one tile for one output row: tmp_d = sum_e exp2(fma(u_e, c2, rowc) [+ cl |y_e|^2]) * b[e][d]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ct_ex2(float s) {
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(s));
  return p;
}

// Experiment kept behind CT_POLY (default 0 = off): of every 8 pairs, CT_POLY get exp2 from a polynomial on the FMA
// pipe instead of MUFU.EX2, all arithmetic packed (f32x2).  Measured on B200 (M = N = 1e6, D = 8 / 16): off 258 / 277 ms,
// CT_POLY = 2: 261 / 283 ms with the XU pipe down from 91 % to 67 % busy (ncu) -- with 4 epilogue warps per scheduler
// and 96 registers per thread the longer dependency chains are not hidden, so the freed MUFU time is not converted.
#ifndef CT_POLY
#define CT_POLY 0
#endif
#define CT_FASTPATH (CT_W == 1 && CT_POLY > 0)     /* dim b = 1, penalty inside K: packed f32x2 epilogue           */

// ---- packed (f32x2) helpers: one issue slot feeds two pairs ----
__device__ __forceinline__ u64 ct_pack(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void ct_unpack(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ct_fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 ct_add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// exp2 of two arguments on the FMA pipe (the MUFU pipe is the bottleneck of this kernel; the FMA pipe is 3/4 idle):
// round-to-nearest split s = n + f with the 1.5 * 2^23 trick, degree-5 minimax polynomial of 2^f on [-1/2, 1/2]
// (max relative error 2.1e-7 in fp32 Horner form, the same as ex2.approx's 2 ulp), exponent patched in with an integer
// multiply-add.  Arguments are clamped to [-125, 126]: below, ex2.approx.ftz would return 0 and this returns < 2.4e-38.
__device__ __forceinline__ u64 ct_exp2_poly2(u64 s2) {
  float s0, s1;
  ct_unpack(s2, s0, s1);
  s0 = fminf(fmaxf(s0, -125.0f), 126.0f);
  s1 = fminf(fmaxf(s1, -125.0f), 126.0f);
  const u64 s = ct_pack(s0, s1);
  const u64 magic = ct_pack(12582912.0f, 12582912.0f), nmagic = ct_pack(-12582912.0f, -12582912.0f);
  const u64 t = ct_add2(s, magic);                         // low mantissa bits of t = round(s)
  const u64 n = ct_add2(t, nmagic);
  const u64 f = ct_fma2(n, ct_pack(-1.0f, -1.0f), s);      // exact
  u64 p = ct_pack(1.327647129e-03f, 1.327647129e-03f);
  p = ct_fma2(p, f, ct_pack(9.675540961e-03f, 9.675540961e-03f));
  p = ct_fma2(p, f, ct_pack(5.550713092e-02f, 5.550713092e-02f));
  p = ct_fma2(p, f, ct_pack(2.402212024e-01f, 2.402212024e-01f));
  p = ct_fma2(p, f, ct_pack(6.931469440e-01f, 6.931469440e-01f));
  p = ct_fma2(p, f, ct_pack(1.000000119e+00f, 1.000000119e+00f));
  float p0, p1, t0, t1;
  ct_unpack(p, p0, p1);
  ct_unpack(t, t0, t1);
  const int r0 = __float_as_int(t0) * 0x800000 + __float_as_int(p0);     // IMAD: (n << 23) + bits(p)
  const int r1 = __float_as_int(t1) * 0x800000 + __float_as_int(p1);
  return ct_pack(__int_as_float(r0), __int_as_float(r1));
}

#if CT_FASTPATH
// dim b = 1: eight columns per step -- 4 FFMA2 (exponents), 8 - CT_POLY MUFU.EX2 + CT_POLY / 2 packed polynomials,
// 4 FFMA2 (accumulation, two interleaved packed partial sums), 2 broadcast LDS.128
__device__ __forceinline__ void ct_tile_fast(const float (&u)[CT_HALF], const float* __restrict__ bt, const float c2,
                                             const float rowc, float (&acc)[CT_DB]) {
  const u64 c2p = ct_pack(c2, c2), rcp = ct_pack(rowc, rowc);
  u64 tmp0 = 0ull, tmp1 = 0ull;
#pragma unroll
  for (int e8 = 0; e8 < CT_HALF / 8; e8++) {
    const ulonglong2 ba = *reinterpret_cast<const ulonglong2*>(bt + 8 * e8);         // b of columns 0..3 as two packs
    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(bt + 8 * e8 + 4);     // columns 4..7
    u64 sp[4], pp[4];
#pragma unroll
    for (int k = 0; k < 4; k++) sp[k] = ct_fma2(ct_pack(u[8 * e8 + 2 * k], u[8 * e8 + 2 * k + 1]), c2p, rcp);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k >= 4 - CT_POLY / 2) {
        pp[k] = ct_exp2_poly2(sp[k]);
      } else {
        float a, b;
        ct_unpack(sp[k], a, b);
        pp[k] = ct_pack(ct_ex2(a), ct_ex2(b));
      }
    }
    tmp0 = ct_fma2(pp[0], ba.x, tmp0);
    tmp1 = ct_fma2(pp[1], ba.y, tmp1);
    tmp0 = ct_fma2(pp[2], bb.x, tmp0);
    tmp1 = ct_fma2(pp[3], bb.y, tmp1);
  }
  float a0, a1, b0, b1;
  ct_unpack(tmp0, a0, a1);
  ct_unpack(tmp1, b0, b1);
  acc[0] += (a0 + a1) + (b0 + b1);
}
#endif

template <bool TAIL, int NC>
__device__ __forceinline__ void ct_tile(const float (&u)[NC], const float* __restrict__ bt, const float c2,
                                        const float cl, const float rowc, const long long j0, const long long ncol,
                                        float (&acc)[CT_DB]) {
  float tmp[2][CT_DB];                                // two interleaved partial sums per component
#pragma unroll
  for (int d = 0; d < CT_DB; d++) tmp[0][d] = tmp[1][d] = 0.0f;
#pragma unroll
  for (int e4 = 0; e4 < NC / 4; e4++) {
    // the four b-tile rows of these columns: 4 * CT_W consecutive floats, read with broadcast LDS.128
    float bv[4 * CT_W];
    const float4* b4 = reinterpret_cast<const float4*>(bt + 4 * e4 * CT_W);
#pragma unroll
    for (int v = 0; v < CT_W; v++) {
      const float4 q = b4[v];
      bv[4 * v + 0] = q.x; bv[4 * v + 1] = q.y; bv[4 * v + 2] = q.z; bv[4 * v + 3] = q.w;
    }
    float p[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float sarg = fmaf(u[4 * e4 + e], c2, rowc);
      if (CT_PEN == 0) sarg = fmaf(bv[e * CT_W], cl, sarg);          // + cl |y_j|^2
#if CT_DBG == 1
      p[e] = sarg;
#else
      p[e] = ct_ex2(sarg);
#endif
      if (TAIL && j0 + 4 * e4 + e >= ncol) p[e] = 0.0f;
    }
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
      for (int d = 0; d < CT_DB; d++)
        tmp[e & 1][d] = fmaf(p[e], bv[e * CT_W + (1 - CT_PEN) + d], tmp[e & 1][d]);
  }
#pragma unroll
  for (int d = 0; d < CT_DB; d++) acc[d] += tmp[0][d] + tmp[1][d];
}

#if CT_LSE
// one tile of a Max_SumShiftExpWeight reduction for one output row, base-2 units: v_e = log2(e) F_e;
// pass 1: exponents + tile maximum; rescale the running sums if the maximum moved; pass 2: sums of 2^(v_e - M2) b_e
template <bool TAIL, int NC>
__device__ __forceinline__ void ct_tile_lse(float (&u)[NC], const float* __restrict__ bt, const float c2,
                                            const float rowc, const long long j0, const long long ncol, float& M2,
                                            float (&acc)[CT_DB]) {
  float vmax = -KT_INF;
#pragma unroll
  for (int e = 0; e < NC; e++) {
    float v = fmaf(u[e], c2, rowc);
    if (TAIL && j0 + e >= ncol) v = -KT_INF;
    u[e] = v;
    vmax = fmaxf(vmax, v);
  }
  if (vmax > M2) {                                    // (also the first tile: M2 = -inf, the sums are still 0)
    const float sc = ct_ex2(M2 - vmax);
#pragma unroll
    for (int d = 0; d < CT_DB; d++) acc[d] *= sc;
    M2 = vmax;
  }
  if (M2 == -KT_INF) return;                          // nothing but masked / -inf entries so far
  float tmp[2][CT_DB];
#pragma unroll
  for (int d = 0; d < CT_DB; d++) tmp[0][d] = tmp[1][d] = 0.0f;
#pragma unroll
  for (int e4 = 0; e4 < NC / 4; e4++) {
    float bv[4 * CT_DB];
    const float4* b4 = reinterpret_cast<const float4*>(bt + 4 * e4 * CT_DB);
#pragma unroll
    for (int v = 0; v < CT_DB; v++) {
      const float4 q = b4[v];
      bv[4 * v + 0] = q.x; bv[4 * v + 1] = q.y; bv[4 * v + 2] = q.z; bv[4 * v + 3] = q.w;
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float p = ct_ex2(u[4 * e4 + e] - M2);
#pragma unroll
      for (int d = 0; d < CT_DB; d++) tmp[e & 1][d] = fmaf(p, bv[e * CT_DB + d], tmp[e & 1][d]);
    }
  }
#pragma unroll
  for (int d = 0; d < CT_DB; d++) acc[d] += tmp[0][d] + tmp[1][d];
}

// merge (M2b, Sb) into (M2a, Sa): the pair algebra of the reference's ReducePair, in base-2 units
__device__ __forceinline__ void ct_lse_merge(float& Ma, float (&Sa)[CT_DB], const float Mb, const float* Sb) {
  const float M = fmaxf(Ma, Mb);
  const float fa = (Ma == -KT_INF) ? 0.0f : ct_ex2(Ma - M), fb = (Mb == -KT_INF) ? 0.0f : ct_ex2(Mb - M);
#pragma unroll
  for (int d = 0; d < CT_DB; d++) Sa[d] = Sa[d] * fa + Sb[d] * fb;
  Ma = M;
}
#endif

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(CT_THREADS, 1)
kb200_conv_tc(const __grid_constant__ KtTensorMap map_x, const __grid_constant__ KtTensorMap map_xr,
              const __grid_constant__ KtTensorMap map_y, const __grid_constant__ KtTensorMap map_yr,
              const long long nrow, const long long ncol, const long long chunk_cols,
              const float* __restrict__ bpad, const float* __restrict__ xsq, const float* __restrict__ coef,
              float* __restrict__ out, unsigned long long* __restrict__ clk, const unsigned int* __restrict__ xbf) {
  extern __shared__ unsigned char ct_raw[];
  const u32 raw = kt_smem(ct_raw);
  const u32 base = (raw + 1023u) & ~1023u;                       // swizzle atoms need (up to) 1024-byte alignment
  unsigned char* gbase = ct_raw + (base - raw);
  const u32 sA = base;                                           // 2 row blocks x [KBM main tiles | remainder tile]
  const u32 sB = sA + CT_A_BYTES;                                // STAGES x [KBM main tiles | remainder tile]
  const u32 sH = sB + CT_STAGES * CT_B_BYTES;                    // ring of b tiles (CT_HS x CT_BN x CT_DB floats)
  const float* hb = reinterpret_cast<const float*>(gbase + CT_A_BYTES + CT_STAGES * CT_B_BYTES);
  const u32 sX = sH + CT_HS * CT_H_BYTES;                         // CT_BM x CT_DB floats
  float* xch = reinterpret_cast<float*>(gbase + CT_A_BYTES + CT_STAGES * CT_B_BYTES + CT_HS * CT_H_BYTES);
  const u32 sBar = sX + CT_X_BYTES;
  // barriers (8 bytes each): full[8] @0, empty[8] @64, tfull[buf][rb] @128, tempty[buf][rb] @192,
  //                          hfull[16] @256, hempty[16] @384, afull @512, tmem pointer @520
  const u32 bFull = sBar, bEmpty = sBar + 64, bTFull = sBar + 128, bTEmpty = sBar + 192, bAFull = sBar + 512;
  const u32 bHFull = sBar + 256, bHEmpty = sBar + 384;
  volatile u32* tmem_slot =
      reinterpret_cast<volatile u32*>(gbase + CT_A_BYTES + CT_STAGES * CT_B_BYTES + CT_HS * CT_H_BYTES + CT_X_BYTES + 520);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // optional clock probe (CTA 0): SM cycles and nanoseconds spent by this CTA -> the SM clock this kernel really ran at
  // (nvidia-smi reports 1965 MHz while the tensor pipe is busy; cycles / time says ~1.83 GHz)
  unsigned long long clk_c0 = 0, clk_t0 = 0;
  if (clk != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(clk_c0));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(clk_t0));
  }
  const long long row_base = (long long)blockIdx.x * CT_BM;
  // blockIdx.y = column chunk (chunk_cols is a multiple of CT_BN).  CTAs are dispatched x-fastest, so all row blocks
  // of one chunk run together and the chunk's operand rows stay L2-resident while every CTA streams them.
  const long long col0 = (long long)blockIdx.y * chunk_cols;
  const long long col1 = (col0 + chunk_cols < ncol) ? col0 + chunk_cols : ncol;
  const int t0 = (int)(col0 / CT_BN);
  const int ntiles = (int)((col1 - col0 + CT_BN - 1) / CT_BN);
  out += (long long)blockIdx.y * nrow * CT_OUTW;                 // chunk partial sums (gridDim.y > 1) or the output

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CT_STAGES; s++) { kt_mbar_init(bFull + 8 * s, 1); kt_mbar_init(bEmpty + 8 * s, 1); }
    for (int b = 0; b < 2 * CT_NBUF; b++) { kt_mbar_init(bTFull + 8 * b, 1); kt_mbar_init(bTEmpty + 8 * b, CT_EPI_WARPS / 2); }
    for (int h = 0; h < CT_HS; h++) { kt_mbar_init(bHFull + 8 * h, 1); kt_mbar_init(bHEmpty + 8 * h, CT_EPI_WARPS); }
    kt_mbar_init(bAFull, CT_ATMEM2 ? CT_EPI_WARPS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sBar + 520), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  kt_fence_before();
  __syncthreads();
  kt_fence_after();
  const u32 tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    if (!CT_ATMEM2 && kt_elect()) {
      kt_mbar_expect_tx(bAFull, CT_A_BYTES);
      for (int rb = 0; rb < 2; rb++) {
        for (int kb = 0; kb < CT_KBM; kb++)
          kt_tma_2d(sA + rb * CT_BLK_BYTES(128) + kb * CT_TILE_BYTES(128), &map_x, kb * 64, (int)(row_base + rb * 128), bAFull);
        kt_tma_2d(sA + rb * CT_BLK_BYTES(128) + CT_KBM * CT_TILE_BYTES(128), &map_xr, 0, (int)(row_base + rb * 128), bAFull);
      }
    }
    __syncwarp();
    for (int t = 0; t < ntiles; t++) {
      const int s = t % CT_STAGES, h = t % CT_HS;
      kt_mbar_wait(bEmpty + 8 * s, ((t / CT_STAGES) & 1) ^ 1);
      kt_mbar_wait(bHEmpty + 8 * h, ((t / CT_HS) & 1) ^ 1);
      if (kt_elect()) {
        kt_mbar_expect_tx(bFull + 8 * s, CT_B_BYTES);
        for (int kb = 0; kb < CT_KBM; kb++)
          kt_tma_2d(sB + s * CT_B_BYTES + kb * CT_TILE_BYTES(CT_BN), &map_y, kb * 64, (t0 + t) * CT_BN, bFull + 8 * s);
        kt_tma_2d(sB + s * CT_B_BYTES + CT_KBM * CT_TILE_BYTES(CT_BN), &map_yr, 0, (t0 + t) * CT_BN, bFull + 8 * s);
        kt_mbar_expect_tx(bHFull + 8 * h, CT_H_BYTES);
        kt_bulk_1d(sH + h * CT_H_BYTES, bpad + (long long)(t0 + t) * CT_BN * CT_W, CT_H_BYTES, bHFull + 8 * h);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    kt_mbar_wait(bAFull, 0);
    kt_fence_after();
    for (int t = 0; t < ntiles; t++) {
      const int s = t % CT_STAGES, buf = t % CT_NBUF;
      kt_mbar_wait(bFull + 8 * s, (t / CT_STAGES) & 1);
#pragma unroll
      for (int rb = 0; rb < 2; rb++) {
        kt_mbar_wait(bTEmpty + 8 * (buf * 2 + rb), ((t / CT_NBUF) & 1) ^ 1);
        kt_fence_after();
        if (kt_elect()) {
          const u32 d = tmem + (u32)((buf * 2 + rb) * CT_BN);
#if CT_ATMEM2
#pragma unroll
          for (int kb = 0; kb <= CT_KBM; kb++) {       // CT_REM == 64: every k-block is 64 wide, 128-byte swizzled
            const u64 bd = kt_smem_desc(sB + s * CT_B_BYTES + kb * CT_TILE_BYTES(CT_BN), 128);
#pragma unroll
            for (int k = 0; k < 4; k++)
              kt_mma_bf16_ts(d, tmem + (u32)(CT_TMEM_A2 + rb * (CT_DPG / 2) + kb * 32 + k * 8), bd + (u64)(2 * k), CT_IDESC,
                             (u32)((kb | k) != 0));
          }
#else
#pragma unroll
          for (int kb = 0; kb < CT_KBM; kb++) {
            const u64 ad = kt_smem_desc(sA + rb * CT_BLK_BYTES(128) + kb * CT_TILE_BYTES(128), 128);
            const u64 bd = kt_smem_desc(sB + s * CT_B_BYTES + kb * CT_TILE_BYTES(CT_BN), 128);
#pragma unroll
            for (int k = 0; k < 4; k++)              // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle atom
              kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), CT_IDESC, (u32)((kb | k) != 0));
          }
          {
            const u64 ad = kt_smem_desc(sA + rb * CT_BLK_BYTES(128) + CT_KBM * CT_TILE_BYTES(128), CT_REM * 2);
            const u64 bd = kt_smem_desc(sB + s * CT_B_BYTES + CT_KBM * CT_TILE_BYTES(CT_BN), CT_REM * 2);
#pragma unroll
            for (int k = 0; k < CT_REM / 16; k++)
              kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), CT_IDESC, (u32)((CT_KBM | k) != 0));
          }
#endif
          kt_mma_commit(bTFull + 8 * (buf * 2 + rb));  // this accumulator is complete
          if (rb == 1) kt_mma_commit(bEmpty + 8 * s);  // smem stage may be refilled once these MMAs have read it
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: one thread per (output row, 64-column half) =====================
    // warps 2-5: row block 0, half 0 | 6-9: row block 1, half 0 | 10-13: row block 0, half 1 | 14-17: row block 1,
    // half 1; inside each group the four warps cover the four TMEM lane quarters (a warp may only read quarter warp % 4)
    const int ew = warp - 2, q = warp & 3, rb = (ew >> 2) & 1, half = ew >> 3;
    const int rloc = rb * 128 + q * 32 + lane;        // row inside the CTA
    const long long i = row_base + rloc;
    const bool live = i < nrow;
    const float cl = coef[0] * CT_LOG2E;              // exponent = cl * |x - y|^2  (base 2)
    const float c2 = -2.0f * cl;
    const float rowc = live ? cl * xsq[i] : 0.0f;
#if CT_ATMEM2
    // the resident operand: this thread's row of bf16 pairs, k-blocks half, half + 2, ... -> 32 TMEM columns each
    for (int kb = half; kb <= CT_KBM; kb += 2) {
      u32 av[32];
      const uint4* src = reinterpret_cast<const uint4*>(xbf + (live ? i : 0) * (CT_DPG / 2) + kb * 32);
#pragma unroll
      for (int v = 0; v < 8; v++) {
        const uint4 w4 = live ? __ldg(src + v) : make_uint4(0u, 0u, 0u, 0u);
        av[4 * v + 0] = w4.x; av[4 * v + 1] = w4.y; av[4 * v + 2] = w4.z; av[4 * v + 3] = w4.w;
      }
      kt_tmem_st32(tmem + ((u32)(q * 32) << 16) + (u32)(CT_TMEM_A2 + rb * (CT_DPG / 2) + kb * 32), av);
    }
    kt_tmem_wait_st();
    kt_fence_before();
    __syncwarp();
    if (lane == 0) kt_mbar_arrive(bAFull);
#endif
    float acc[CT_DB];
#pragma unroll
    for (int d = 0; d < CT_DB; d++) acc[d] = 0.0f;
#if CT_LSE
    float M2 = -KT_INF;                               // running maximum of log2(e) F_ij
#endif
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % CT_NBUF, hslot = t % CT_HS;
      kt_mbar_wait(bTFull + 8 * (buf * 2 + rb), (t / CT_NBUF) & 1);
      kt_fence_after();
      const u32 tbase = tmem + ((u32)(q * 32) << 16) + (u32)((buf * 2 + rb) * CT_BN + half * CT_HALF);
      float uu[CT_HALF];
#if CT_DBG == 2
#pragma unroll
      for (int e = 0; e < CT_HALF; e++) uu[e] = -1.0f;
#else
      kt_tmem_ld64_nowait(tbase, uu);
      kt_tmem_wait_ld();
#endif
      kt_fence_before();
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bTEmpty + 8 * (buf * 2 + rb));   // accumulator back to the MMA warp
      kt_mbar_wait(bHFull + 8 * hslot, (t / CT_HS) & 1);
      const float* bt = hb + hslot * (CT_BN * CT_W) + half * (CT_HALF * CT_W);
      const long long j0 = (long long)(t0 + t) * CT_BN + half * CT_HALF;
#if CT_LSE
      if (j0 + CT_HALF > ncol) ct_tile_lse<true, CT_HALF>(uu, bt, c2, rowc, j0, ncol, M2, acc);
      else ct_tile_lse<false, CT_HALF>(uu, bt, c2, rowc, j0, ncol, M2, acc);
#else
      if (j0 + CT_HALF > ncol) ct_tile<true, CT_HALF>(uu, bt, c2, cl, rowc, j0, ncol, acc);
#if CT_FASTPATH && CT_DBG == 0
      else ct_tile_fast(uu, bt, c2, rowc, acc);
#else
      else ct_tile<false, CT_HALF>(uu, bt, c2, cl, rowc, j0, ncol, acc);
#endif
#endif
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bHEmpty + 8 * hslot);
    }
    // row result = first half (+) second half (fixed order: deterministic)
    if (half == 1) {
#if CT_LSE
      xch[rloc * CT_OUTW] = M2;
#endif
#pragma unroll
      for (int d = 0; d < CT_DB; d++) xch[rloc * CT_OUTW + CT_LSE + d] = acc[d];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(32 * CT_EPI_WARPS) : "memory");
    if (half == 0 && live) {
#if CT_LSE
      ct_lse_merge(M2, acc, xch[rloc * CT_OUTW], xch + rloc * CT_OUTW + 1);
      out[i * CT_OUTW] = (gridDim.y > 1) ? M2 : M2 * 0.6931471805599453f;   // natural units: m_i = max_j F_ij (chunk partials stay base 2)
#pragma unroll
      for (int d = 0; d < CT_DB; d++) out[i * CT_OUTW + 1 + d] = acc[d];
#else
#pragma unroll
      for (int d = 0; d < CT_DB; d++) out[i * CT_DB + d] = acc[d] + xch[rloc * CT_DB + d];
#endif
    }
  }
  kt_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (clk != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long c1, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c1));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    clk[0] = c1 - clk_c0;
    clk[1] = t1 - clk_t0;
  }
}

// ------------------------------------------------------------------------------------------------
// high-dimensional variant (K = 6 D + 3 > 128, up to 576 bf16 columns: D <= 95)
//   One CTA = ONE 128-row block whose operand rows stay resident for all of K (CT_KB 64-column blocks, 16 KB each);
//   the streamed side goes through the smem ring one 128-column x 64-K block at a time, and the MMA warp accumulates
//   a tile over the K blocks in TMEM before committing it to the epilogue.  Per pair this moves 2 * CT_KB * 64 / 128
//   bytes from L2 (the 256-row kernel above would have to re-stream its operand rows as well: 1.5 x more).
//   Epilogue: 16 warps, a thread = one row x one 32-column quarter of every tile.
// ------------------------------------------------------------------------------------------------
#ifndef CT_KB
#define CT_KB (CT_KBM + 1)                    /* only meaningful when CT_REM == 64                               */
#endif
#ifndef CT_ASTREAM
#define CT_ASTREAM 0                          /* 1 (K > 576): the row block's k-blocks are streamed through the ring
                                                 next to the streamed side's instead of staying resident: any D, at
                                                 twice the L2 -> smem traffic per pair                              */
#endif
#ifndef CT_MC
#define CT_MC 0                               /* 1: CTA pairs (cluster 2x1x1) share the streamed k-blocks: each CTA loads
                                                 one 64-column half and TMA-multicasts it to both, halving the L2 -> smem
                                                 traffic of the streamed side.  Measured: no gain (717 vs 726 ms at D = 64,
                                                 M = N = 1e6) -- the kernel is bound by the shared-memory port, not by L2  */
#endif
#ifndef CT_ATMEM
#define CT_ATMEM 0                            /* 1 (K <= 512): the resident row block lives in tensor memory (32 columns per
                                                 k-block, written once by the epilogue warps with tcgen05.st) and enters the
                                                 MMAs as a TMEM A operand: the shared-memory port, which bounds this kernel,
                                                 only carries the streamed side                                          */
#endif
#define HD_BLK 16384                          /* 128 rows x 64 bf16                                              */
#define HD_STAGE_BYTES ((1 + CT_ASTREAM) * HD_BLK)
#define HD_A_BYTES ((CT_ASTREAM || CT_ATMEM) ? 0 : CT_KB * HD_BLK)
#define HD_TMEM_A (CT_NBUF * CT_BN)           /* first TMEM column of the resident operand (after the accumulators)   */
#define HD_QUART (CT_BN / 4)
#define HD_X_BYTES (3 * 128 * CT_OUTW * 4)    /* sums of column quarters 1..3, added to quarter 0 at the end      */
#define HD_SMEM_BYTES (1024 + HD_A_BYTES + CT_STAGES * HD_STAGE_BYTES + CT_HS * CT_H_BYTES + HD_X_BYTES + 640) /* mirrored in conv_tc.py */

extern "C" __global__ void
#if CT_MC
__cluster_dims__(2, 1, 1)
#endif
__launch_bounds__(CT_THREADS, 1)
kb200_conv_tc_hd(const __grid_constant__ KtTensorMap map_x, const __grid_constant__ KtTensorMap map_y,
                 const long long nrow, const long long ncol, const long long chunk_cols,
                 const float* __restrict__ bpad, const float* __restrict__ xsq, const float* __restrict__ coef,
                 float* __restrict__ out, const unsigned int* __restrict__ xbf) {
  extern __shared__ unsigned char ct_raw[];
  const u32 raw = kt_smem(ct_raw);
  const u32 base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = ct_raw + (base - raw);
  const u32 sA = base;                                           // CT_KB resident k-blocks of this CTA's 128 rows
  const u32 sB = sA + HD_A_BYTES;                                // ring of streamed k-blocks
  const u32 sH = sB + CT_STAGES * HD_STAGE_BYTES;                        // ring of b tiles
  const float* hb = reinterpret_cast<const float*>(gbase + HD_A_BYTES + CT_STAGES * HD_STAGE_BYTES);
  const u32 sX = sH + CT_HS * CT_H_BYTES;
  float* xch = reinterpret_cast<float*>(gbase + HD_A_BYTES + CT_STAGES * HD_STAGE_BYTES + CT_HS * CT_H_BYTES);
  const u32 sBar = sX + HD_X_BYTES;
  // barriers: full[8] @0, empty[8] @64, tfull[2] @128, tempty[2] @192, hfull[16] @256, hempty[16] @384, afull @512, tmem @520
  const u32 bFull = sBar, bEmpty = sBar + 64, bTFull = sBar + 128, bTEmpty = sBar + 192, bAFull = sBar + 512;
  const u32 bHFull = sBar + 256, bHEmpty = sBar + 384;
  volatile u32* tmem_slot = reinterpret_cast<volatile u32*>(
      gbase + HD_A_BYTES + CT_STAGES * HD_STAGE_BYTES + CT_HS * CT_H_BYTES + HD_X_BYTES + 520);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_base = (long long)blockIdx.x * 128;
  const long long col0 = (long long)blockIdx.y * chunk_cols;
  const long long col1 = (col0 + chunk_cols < ncol) ? col0 + chunk_cols : ncol;
  const int t0 = (int)(col0 / CT_BN);
  const int ntiles = (int)((col1 - col0 + CT_BN - 1) / CT_BN);
  out += (long long)blockIdx.y * nrow * CT_OUTW;

  if (warp == 1 && lane == 0) {
    // CT_MC: a ring slot is written by both CTAs of the pair, so it is free only when both MMA warps have read it
    for (int s = 0; s < CT_STAGES; s++) { kt_mbar_init(bFull + 8 * s, 1); kt_mbar_init(bEmpty + 8 * s, 1 + CT_MC); }
    for (int b = 0; b < CT_NBUF; b++) { kt_mbar_init(bTFull + 8 * b, 1); kt_mbar_init(bTEmpty + 8 * b, CT_EPI_WARPS); }
    for (int h = 0; h < CT_HS; h++) { kt_mbar_init(bHFull + 8 * h, 1); kt_mbar_init(bHEmpty + 8 * h, CT_EPI_WARPS); }
    kt_mbar_init(bAFull, CT_ATMEM ? CT_EPI_WARPS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sBar + 520), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  kt_fence_before();
  __syncthreads();
  kt_fence_after();
  const u32 tmem = *tmem_slot;
#if CT_MC
  kt_cluster_sync();                                             // the peer's barriers exist before anything targets them
  const u32 crank = kt_cluster_ctarank();
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
#if !CT_ASTREAM && !CT_ATMEM
    if (kt_elect()) {
      kt_mbar_expect_tx(bAFull, HD_A_BYTES);
      for (int kb = 0; kb < CT_KB; kb++) kt_tma_2d(sA + kb * HD_BLK, &map_x, kb * 64, (int)row_base, bAFull);
    }
    __syncwarp();
#endif
    int it = 0;
    for (int t = 0; t < ntiles; t++) {
      const int h = t % CT_HS;
      kt_mbar_wait(bHEmpty + 8 * h, ((t / CT_HS) & 1) ^ 1);
      if (kt_elect()) {
        kt_mbar_expect_tx(bHFull + 8 * h, CT_H_BYTES);
        kt_bulk_1d(sH + h * CT_H_BYTES, bpad + (long long)(t0 + t) * CT_BN * CT_W, CT_H_BYTES, bHFull + 8 * h);
      }
      __syncwarp();
      for (int kb = 0; kb < CT_KB; kb++, it++) {
        const int s = it % CT_STAGES;
        kt_mbar_wait(bEmpty + 8 * s, ((it / CT_STAGES) & 1) ^ 1);
        if (kt_elect()) {
          kt_mbar_expect_tx(bFull + 8 * s, HD_STAGE_BYTES);
#if CT_MC
          // this CTA's half (64 of the 128 columns of the tile) goes to both CTAs of the pair
          kt_tma_2d_mc(sB + s * HD_STAGE_BYTES + crank * (HD_BLK / 2), &map_y, kb * 64,
                       (t0 + t) * CT_BN + (int)crank * (CT_BN / 2), bFull + 8 * s, (unsigned short)3);
#else
          kt_tma_2d(sB + s * HD_STAGE_BYTES, &map_y, kb * 64, (t0 + t) * CT_BN, bFull + 8 * s);
#endif
#if CT_ASTREAM
          kt_tma_2d(sB + s * HD_STAGE_BYTES + HD_BLK, &map_x, kb * 64, (int)row_base, bFull + 8 * s);
#endif
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: a tile accumulates over the CT_KB k-blocks =====================
#if !CT_ASTREAM
    kt_mbar_wait(bAFull, 0);
    kt_fence_after();
#endif
    int it = 0;
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % CT_NBUF;
      kt_mbar_wait(bTEmpty + 8 * buf, ((t / CT_NBUF) & 1) ^ 1);
      kt_fence_after();
      for (int kb = 0; kb < CT_KB; kb++, it++) {
        const int s = it % CT_STAGES;
        kt_mbar_wait(bFull + 8 * s, (it / CT_STAGES) & 1);
        kt_fence_after();
        if (kt_elect()) {
          const u32 d = tmem + (u32)(buf * CT_BN);
          const u64 bd = kt_smem_desc(sB + s * HD_STAGE_BYTES, 128);
#if CT_ATMEM
#pragma unroll
          for (int k = 0; k < 4; k++)                      // 16 bf16 of K = 8 TMEM columns per MMA
            kt_mma_bf16_ts(d, tmem + (u32)(HD_TMEM_A + kb * 32 + k * 8), bd + (u64)(2 * k), CT_IDESC, (u32)((kb | k) != 0));
#else
          const u64 ad = kt_smem_desc(CT_ASTREAM ? (sB + s * HD_STAGE_BYTES + HD_BLK) : (sA + kb * HD_BLK), 128);
#pragma unroll
          for (int k = 0; k < 4; k++)
            kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), CT_IDESC, (u32)((kb | k) != 0));
#endif
#if CT_MC
          kt_mma_commit_mc(bEmpty + 8 * s, (unsigned short)3);   // releases the slot in both CTAs of the pair
#else
          kt_mma_commit(bEmpty + 8 * s);                   // the ring slot may be refilled once these MMAs have read it
#endif
          if (kb == CT_KB - 1) kt_mma_commit(bTFull + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: one thread per (row, 32-column quarter) =====================
    const int ew = warp - 2, q = warp & 3, part = ew >> 2;
    const int rloc = q * 32 + lane;
    const long long i = row_base + rloc;
    const bool live = i < nrow;
    const float cl = coef[0] * CT_LOG2E;
    const float c2 = -2.0f * cl;
    const float rowc = live ? cl * xsq[i] : 0.0f;
#if CT_ATMEM
    // the resident operand: this thread's row of bf16 pairs, k-blocks part, part + 4, ... -> 32 TMEM columns each
    for (int kb = part; kb < CT_KB; kb += 4) {
      u32 av[32];
      const uint4* src = reinterpret_cast<const uint4*>(xbf + (live ? i : 0) * (CT_DPG / 2) + kb * 32);
#pragma unroll
      for (int v = 0; v < 8; v++) {
        const uint4 w4 = live ? __ldg(src + v) : make_uint4(0u, 0u, 0u, 0u);
        av[4 * v + 0] = w4.x; av[4 * v + 1] = w4.y; av[4 * v + 2] = w4.z; av[4 * v + 3] = w4.w;
      }
      kt_tmem_st32(tmem + ((u32)(q * 32) << 16) + (u32)(HD_TMEM_A + kb * 32), av);
    }
    kt_tmem_wait_st();
    kt_fence_before();
    __syncwarp();
    if (lane == 0) kt_mbar_arrive(bAFull);
#endif
    float acc[CT_DB];
#pragma unroll
    for (int d = 0; d < CT_DB; d++) acc[d] = 0.0f;
#if CT_LSE
    float M2 = -KT_INF;
#endif
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % CT_NBUF, hslot = t % CT_HS;
      kt_mbar_wait(bTFull + 8 * buf, (t / CT_NBUF) & 1);
      kt_fence_after();
      float uu[HD_QUART];
      kt_tmem_ld32_nowait(tmem + ((u32)(q * 32) << 16) + (u32)(buf * CT_BN + part * HD_QUART), uu);
      kt_tmem_wait_ld();
      kt_fence_before();
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bTEmpty + 8 * buf);
      kt_mbar_wait(bHFull + 8 * hslot, (t / CT_HS) & 1);
      const float* bt = hb + hslot * (CT_BN * CT_W) + part * (HD_QUART * CT_W);
      const long long j0 = (long long)(t0 + t) * CT_BN + part * HD_QUART;
#if CT_LSE
      if (j0 + HD_QUART > ncol) ct_tile_lse<true, HD_QUART>(uu, bt, c2, rowc, j0, ncol, M2, acc);
      else ct_tile_lse<false, HD_QUART>(uu, bt, c2, rowc, j0, ncol, M2, acc);
#else
      if (j0 + HD_QUART > ncol) ct_tile<true, HD_QUART>(uu, bt, c2, cl, rowc, j0, ncol, acc);
      else ct_tile<false, HD_QUART>(uu, bt, c2, cl, rowc, j0, ncol, acc);
#endif
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bHEmpty + 8 * hslot);
    }
    if (part > 0) {
#if CT_LSE
      xch[((part - 1) * 128 + rloc) * CT_OUTW] = M2;
#endif
#pragma unroll
      for (int d = 0; d < CT_DB; d++) xch[((part - 1) * 128 + rloc) * CT_OUTW + CT_LSE + d] = acc[d];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(32 * CT_EPI_WARPS) : "memory");
    if (part == 0 && live) {
#if CT_LSE
#pragma unroll
      for (int p_ = 0; p_ < 3; p_++)
        ct_lse_merge(M2, acc, xch[(p_ * 128 + rloc) * CT_OUTW], xch + (p_ * 128 + rloc) * CT_OUTW + 1);
      out[i * CT_OUTW] = (gridDim.y > 1) ? M2 : M2 * 0.6931471805599453f;
#pragma unroll
      for (int d = 0; d < CT_DB; d++) out[i * CT_OUTW + 1 + d] = acc[d];
#else
#pragma unroll
      for (int d = 0; d < CT_DB; d++)
        out[i * CT_DB + d] = ((acc[d] + xch[rloc * CT_DB + d]) + xch[(128 + rloc) * CT_DB + d]) + xch[(256 + rloc) * CT_DB + d];
#endif
    }
  }
  kt_fence_before();
  __syncthreads();
#if CT_MC
  kt_cluster_sync();                                             // no CTA leaves while its peer may still target it
#endif
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// combine: out = sum over column chunks of the partial sums, in chunk order (deterministic)
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
kb200_conv_combine(const long long n, const int nchunks, const float* __restrict__ partial, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = partial[e];
  for (int c = 1; c < nchunks; c++) s += partial[(long long)c * n + e];
  out[e] = s;
}

#if CT_LSE
// combine of the column chunks of a LogSumExp reduction: (M2, S) partials in base-2 units, merged in chunk order
extern "C" __global__ void __launch_bounds__(256)
kb200_conv_combine_lse(const long long nrow, const int nchunks, const float* __restrict__ partial, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrow) return;
  float M = partial[i * CT_OUTW], S[CT_DB];
#pragma unroll
  for (int d = 0; d < CT_DB; d++) S[d] = partial[i * CT_OUTW + 1 + d];
  for (int c = 1; c < nchunks; c++) {
    const float* p = partial + ((long long)c * nrow + i) * CT_OUTW;
    ct_lse_merge(M, S, p[0], p + 1);
  }
  out[i * CT_OUTW] = M * 0.6931471805599453f;
#pragma unroll
  for (int d = 0; d < CT_DB; d++) out[i * CT_OUTW + 1 + d] = S[d];
}
#endif
