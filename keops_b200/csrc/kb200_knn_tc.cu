// kb200_knn_tc.cu -- K-nearest-neighbour reductions (KMin / ArgKMin / KMin_ArgKMin of a squared
// Euclidean distance, large D) on the 5th-generation tensor cores of sm_100a.
//
// Reference semantics: KMin_ArgKMin_Reduction (keopscore/keopscore/formulas/reductions/
// KMin_ArgKMin_Reduction.py:16-100) applied to Sum(Square(x_i - y_j)); the reference evaluates every
// pair on CUDA cores (GpuReduc1D.py:70-90; 129 FP32 instructions per pair at D = 64).
//
// Design ("tensor-core K-min"): the contraction x_i . y_j runs on tcgen05.mma (bf16
// operands, fp32 accumulation in TMEM) purely as a *filter*; every surviving pair is re-evaluated exactly
// (fp32, the reference's operation order) by kb200_knn_select, so the result is bit-identical to the
// CUDA-core kernel whatever the filter's rounding:
//
//   acc_ij ~ x_i.y_j   with |acc_ij - x_i.y_j| <= (c/2)(|x_i|^2 + |y_j|^2)     (c covers bf16 rounding of both
//                                                                               operands, fp32 accumulation and
//                                                                               the fp32 rounding of the exact path)
//   (acc_ij is produced *including* the column penalty: three extra K columns hold 1 on the x side and
//    the bf16 hi/mid/lo split of -(1-c)|y_j|^2/2 on the y side, so the epilogue is a pure max-reduction)
//   U_ij = acc_ij - (1-c)|y_j|^2/2   (upper bound of  x.y - |y|^2/2  minus (c/2)|x|^2)
//   L_ij = acc_ij - (1+c)|y_j|^2/2   (lower bound of  x.y - |y|^2/2  plus  (c/2)|x|^2)
//   Theta_i = K-th largest L_ij seen so far  ==>  at least K columns are at least that close, hence column j
//   can belong to the final K nearest only if  U_ij >= Theta_i - c|x_i|^2 =: tau_i  (non-strict: ties kept).
//
// kb200_knn_filter : one CTA = 256 rows (two 128-row accumulators) x all columns, streamed in 128-column
//   tiles.  Warp 0 = TMA producer (y tiles, 128B-swizzled, + the per-column penalty), warp 1 = MMA issuer,
//   warps 2..17 = epilogue: TWO threads per row, one per 64-column half of a tile: tcgen05.ld of its 64 accumulators,
//   FMNMX3 max-reduction against tau_i.  The row's state (K best lower bounds, tau_i, candidate count) lives in shared
//   memory and is shared by the row's two threads: a survivor group is appended as one token (8-column group index +
//   8-bit hit mask) and its best lower bound is inserted into the row's list under a per-row lock.
//   4-stage smem ring, double-buffered TMEM accumulators (2 row blocks x 2 buffers x 128 columns = 512).
//   Why this shape (ncu source view of the previous one-thread-per-row epilogue, profiles/r01_ncu_knn_filter_d64.txt:
//   tensor pipe 47 % busy): the 8 epilogue warps were the critical path and latency-bound -- per tile each spent
//   ~160 clk waiting for tcgen05.ld, ~255 clk in the FMNMX3 reduction, ~400 clk in the divergent survivor path and 440 clk
//   waiting for the next accumulator, because the MMA warp in turn waited for the slowest of the four warps of a row
//   block.  Two threads per row halve the per-warp work and give every scheduler four warps; the shared row state keeps
//   ONE threshold per row (two private lists per row had cost 324 instead of ~200 candidates).
//   Where the time goes now (B200, D = 64, K = 10, M = N = 1e6; KT_DBG builds, profiles/r02_knn_filter_experiments.md):
//   with the survivor path disabled the kernel takes 101 ms, with the max-reduction disabled too 96 ms -- tensor pipe 95 %
//   busy at an SM clock that the power cap has pulled down to 1.45 GHz, i.e. 1.67 PFLOP/s, the throughput cuBLAS reaches
//   on this GPU (MEASURED_PEAKS.json).  That is the floor; the full kernel takes 134-138 ms (0.70-0.72 of it), the
//   difference being the survivor handling: ~150 groups per row, each costing its warp ~500 clk, with only two accumulator
//   buffers per row block between the warps and the MMA issuer.  Measured and rejected on the way (all bit-identical):
//   parking survivor groups in registers and handling them in batches (178 ms: fewer but ~3x longer stalls, which the two
//   buffers absorb worse), mailing them to two helper warps that own the lists (152-156 ms: their polling and the extra
//   warps cost clock under the power cap), 32-column TMEM reads overlapped with the reductions (156 ms: the accumulator
//   is released later), private token sub-lists instead of a shared counter (no change), 512-row CTAs with four single-buffered
//   row blocks and one thread per row (144.7 ms: same cycles per row, 13.2 -> 14 waves), L2-sized column chunks (272 ms:
//   every chunk replays the transient of an empty list).
// kb200_knn_prep   : fp32 -> bf16 operand rows (data, penalty columns, zero padding) + per-row penalty terms.
// kb200_knn_select : one warp per row, exact distances of the candidates (tokens decoded) + ordered top-K (ties ->
//   smallest j).
//
// Compile-time parameters (defined by keops_b200/codegen/knn_tc.py): KT_D (true dimension); the bf16 operand
// rows hold D + 3 columns padded to KT_DPG = 64 * KT_KBM + KT_REM (KT_KBM 128-byte-swizzled 64-column blocks
// plus one remainder block of KT_REM = 0 / 16 / 32 / 64 columns swizzled at its own width); KT_STREAM (0: the CTA's
// row operand stays resident in shared memory for all of K; 1: large D -- both operands go through the ring one
// 64-column k-block at a time and a tile accumulates over the k-blocks in TMEM; KT_REM = 0); KT_STAGES; KT_K, KT_CAP
// (candidate tokens per row, multiple of 32), KT_OUT (0 ArgKMin, 1 KMin, 2 KMin_ArgKMin), KT_FAST (1: FTZ arithmetic
// of --use_fast_math, like the CUDA-core kernel), KT_C (the filter's relative margin c, as a float literal).

#ifndef KT_D
#define KT_D 64
#define KT_KBM 1
#define KT_REM 16
#define KT_K 10
#define KT_CAP 320
#define KT_OUT 0
#define KT_FAST 1
#define KT_STREAM 0
#define KT_STAGES 4
#define KT_C 0.008f
#endif
#ifndef KT_DBG
#define KT_DBG 0                              /* timing experiments: 1 = no survivors at all, 2 = TMEM loads + release only */
#endif

#define KT_BM 256                             /* rows per CTA: two 128-row accumulators                         */
#define KT_BN 128                             /* columns per tile (= N of one tcgen05.mma)                      */
#define KT_NBUF 2                             /* TMEM accumulator buffers per row block (2 x 2 x 128 = 512 columns) */
#define KT_DPG (64 * KT_KBM + KT_REM)
#define KT_HS 8                               /* slots of the column-penalty ring                               */
#define KT_KP 16                              /* row stride (floats) of the shared lists of lower bounds: K <= 16 */
#define KT_EPI_WARPS 16                       /* epilogue warps: (lane quarter) x (row block) x (column half)   */
#define KT_THREADS (64 + 32 * KT_EPI_WARPS)
#define KT_TILE_BYTES(rows) ((rows) * 128)                    /* rows x 64 bf16                                 */
#define KT_RTILE_BYTES(rows) ((rows) * KT_REM * 2)            /* rows x KT_REM bf16                             */
#define KT_BLK_BYTES(rows) (KT_KBM * KT_TILE_BYTES(rows) + KT_RTILE_BYTES(rows))   /* one operand block, all of K */
#if KT_STREAM
#define KT_A_BYTES 0
#define KT_B_BYTES KT_TILE_BYTES(KT_BM + KT_BN)               /* one ring stage: a k-block of both operands     */
#else
#define KT_A_BYTES (2 * KT_BLK_BYTES(128))
#define KT_B_BYTES KT_BLK_BYTES(KT_BN)
#endif
#define KT_STATE_BYTES (KT_BM * (KT_KP + 3) * 4)              /* lists + tau + count + lock of the CTA's rows   */
/* mirrored in knn_tc.py */
#define KT_SMEM_BYTES (1024 + KT_A_BYTES + KT_STAGES * KT_B_BYTES + KT_HS * KT_BN * 4 + KT_STATE_BYTES + 640)

// bf16 x bf16 -> f32, A and B K-major, M = 128, N = KT_BN
#define KT_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | (((u32)KT_BN >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ float4 kt_lds128_volatile(u32 addr) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void kt_sts128_volatile(u32 addr, float a, float b, float c, float d) {
  asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// prep: fp32 rows -> bf16 operand rows of KT_DPG columns + fp32 penalty
//   mode 0 (streamed side y): [ y_0 .. y_{D-1}, -h_hi, -h_mid, -h_lo, 0 .. ],  h = (1-c)|y|^2/2 split into three
//                             bf16 terms (24 significant bits);  pen = h
//   mode 1 (output side x)  : [ x_0 .. x_{D-1}, 1, 1, 1, 0 .. ];               pen = c|x|^2
// ------------------------------------------------------------------------------------------------

extern "C" __global__ void __launch_bounds__(256)
kb200_knn_prep(const long long n, const float* __restrict__ a, unsigned short* __restrict__ abf,
               float* __restrict__ pen, const int mode) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* row = a + r * KT_D;
  unsigned short* orow = abf + r * KT_DPG;
  float s = 0.0f;
#pragma unroll 8
  for (int k = 0; k < KT_D; k++) {
    const float v = row[k];
    s = fmaf(v, v, s);
    orow[k] = kt_bf16_rn(v);
  }
  if (mode == 0) {
    const float h = (1.0f - KT_C) * 0.5f * s;
    const unsigned short hi = kt_bf16_rn(h);
    const float r1 = h - kt_bf16_to_f(hi);
    const unsigned short mid = kt_bf16_rn(r1);
    const unsigned short lo = kt_bf16_rn(r1 - kt_bf16_to_f(mid));
    orow[KT_D] = hi ^ 0x8000u; orow[KT_D + 1] = mid ^ 0x8000u; orow[KT_D + 2] = lo ^ 0x8000u;   // negated
    pen[r] = h;
  } else {
    orow[KT_D] = 0x3F80u; orow[KT_D + 1] = 0x3F80u; orow[KT_D + 2] = 0x3F80u;                   // 1.0
    pen[r] = KT_C * s;
  }
  for (int k = KT_D + 3; k < KT_DPG; k++) orow[k] = 0;
}

// ------------------------------------------------------------------------------------------------
// filter
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(KT_THREADS, 1)
kb200_knn_filter(const __grid_constant__ KtTensorMap map_x, const __grid_constant__ KtTensorMap map_xr,
                 const __grid_constant__ KtTensorMap map_y, const __grid_constant__ KtTensorMap map_yr,
                 const long long nrow, const long long ncol, const float* __restrict__ ypen,
                 const float* __restrict__ xpen, int* __restrict__ cand, int* __restrict__ cand_count,
                 const int tiles_per_split) {
  extern __shared__ unsigned char kt_raw[];
  const u32 raw = kt_smem(kt_raw);
  const u32 base = (raw + 1023u) & ~1023u;                       // swizzle atoms need (up to) 1024-byte alignment
  unsigned char* gbase = kt_raw + (base - raw);
  const u32 sA = base;                                           // resident mode: 2 row blocks x [KBM main tiles | remainder tile]
  const u32 sB = sA + KT_A_BYTES;                                // STAGES x stage (see KT_B_BYTES)
  const u32 sH = sB + KT_STAGES * KT_B_BYTES;                    // ring of column penalties (KT_HS x KT_BN floats)
  const float* hy = reinterpret_cast<const float*>(gbase + KT_A_BYTES + KT_STAGES * KT_B_BYTES);
  // per-row state shared by the row's two epilogue threads: K best lower bounds (descending), tau, token count, lock
  const u32 sLst = sH + KT_HS * KT_BN * 4;
  unsigned char* gstate = gbase + KT_A_BYTES + KT_STAGES * KT_B_BYTES + KT_HS * KT_BN * 4;
  volatile float* lst_s = reinterpret_cast<volatile float*>(gstate);
  volatile float* tau_s = reinterpret_cast<volatile float*>(gstate + KT_BM * KT_KP * 4);
  int* cnt_s = reinterpret_cast<int*>(gstate + KT_BM * (KT_KP + 1) * 4);
  int* lock_s = reinterpret_cast<int*>(gstate + KT_BM * (KT_KP + 2) * 4);
  const u32 sBar = sLst + KT_STATE_BYTES;
  // barriers (8 bytes each): full[8] @0, empty[8] @64, tfull[buf][rb] (4) @128, tempty[buf][rb] @192,
  //                          hfull[8] @256, afull @512, tmem pointer @520
  const u32 bFull = sBar, bEmpty = sBar + 64, bTFull = sBar + 128, bTEmpty = sBar + 192, bAFull = sBar + 512;
  const u32 bHFull = sBar + 256;
  volatile u32* tmem_slot = reinterpret_cast<volatile u32*>(gstate + KT_STATE_BYTES + 520);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_base = (long long)blockIdx.x * KT_BM;
  // blockIdx.y = column split (few rows, many columns: the row blocks alone would not fill the machine).  Every split
  // keeps its own candidate list and threshold; kb200_knn_select evaluates each list exactly and kb200_knn_merge
  // merges the per-split top-K lists in column order.
  const int ntiles_all = (int)((ncol + KT_BN - 1) / KT_BN);
  const int t0 = (int)blockIdx.y * tiles_per_split;
  const int ntiles = (ntiles_all - t0 < tiles_per_split) ? (ntiles_all - t0) : tiles_per_split;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KT_STAGES; s++) { kt_mbar_init(bFull + 8 * s, 1); kt_mbar_init(bEmpty + 8 * s, 1); }
    for (int b = 0; b < 2 * KT_NBUF; b++) { kt_mbar_init(bTFull + 8 * b, 1); kt_mbar_init(bTEmpty + 8 * b, KT_EPI_WARPS / 2); }
    for (int h = 0; h < KT_HS; h++) kt_mbar_init(bHFull + 8 * h, 1);
    kt_mbar_init(bAFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sBar + 520), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < KT_BM; r += KT_THREADS) {
#if KT_DBG == 1                                                     /* timing experiment: no survivors at all */
    tau_s[r] = KT_INF;
#else
    tau_s[r] = (row_base + r < nrow) ? -KT_INF : KT_INF;         // rows past the end never produce a candidate
#endif
    cnt_s[r] = 0;
    lock_s[r] = 0;
#pragma unroll
    for (int l = 0; l < KT_KP; l++) lst_s[r * KT_KP + l] = -KT_INF;
  }
  kt_fence_before();
  __syncthreads();
  kt_fence_after();
  const u32 tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
#if !KT_STREAM
    if (kt_elect()) {
      kt_mbar_expect_tx(bAFull, KT_A_BYTES);
      for (int rb = 0; rb < 2; rb++) {
        for (int kb = 0; kb < KT_KBM; kb++)
          kt_tma_2d(sA + rb * KT_BLK_BYTES(128) + kb * KT_TILE_BYTES(128), &map_x, kb * 64, (int)(row_base + rb * 128), bAFull);
        kt_tma_2d(sA + rb * KT_BLK_BYTES(128) + KT_KBM * KT_TILE_BYTES(128), &map_xr, 0, (int)(row_base + rb * 128), bAFull);
      }
    }
    __syncwarp();
#endif
    // The ring of column penalties needs no "empty" barrier: a slot is refilled KT_HS = 8 tiles later, and by then the
    // wait on this tile's operand stage has ordered the producer behind the MMA warp (at most KT_STAGES tiles ahead of
    // it), which is itself at most KT_NBUF tiles ahead of the slowest epilogue warp: KT_STAGES + KT_NBUF + 1 <= KT_HS.
    for (int t = 0; t < ntiles; t++) {
      const int h = t % KT_HS;
#if KT_STREAM
      for (int kb = 0; kb < KT_KBM; kb++) {
        const int it = t * KT_KBM + kb, s = it % KT_STAGES;
        kt_mbar_wait(bEmpty + 8 * s, ((it / KT_STAGES) & 1) ^ 1);
        if (kt_elect()) {
          const u32 st = sB + s * KT_B_BYTES;
          kt_mbar_expect_tx(bFull + 8 * s, KT_B_BYTES);
          kt_tma_2d(st, &map_x, kb * 64, (int)row_base, bFull + 8 * s);
          kt_tma_2d(st + KT_TILE_BYTES(128), &map_x, kb * 64, (int)(row_base + 128), bFull + 8 * s);
          kt_tma_2d(st + KT_TILE_BYTES(256), &map_y, kb * 64, (t0 + t) * KT_BN, bFull + 8 * s);
          if (kb == 0) {       // column penalties of this tile (only the survivor path of the epilogue reads them)
            kt_mbar_expect_tx(bHFull + 8 * h, KT_BN * 4);
            kt_bulk_1d(sH + h * KT_BN * 4, ypen + (long long)(t0 + t) * KT_BN, KT_BN * 4, bHFull + 8 * h);
          }
        }
        __syncwarp();
      }
#else
      const int s = t % KT_STAGES;
      kt_mbar_wait(bEmpty + 8 * s, ((t / KT_STAGES) & 1) ^ 1);
      if (kt_elect()) {
        kt_mbar_expect_tx(bFull + 8 * s, KT_B_BYTES);
        for (int kb = 0; kb < KT_KBM; kb++)
          kt_tma_2d(sB + s * KT_B_BYTES + kb * KT_TILE_BYTES(KT_BN), &map_y, kb * 64, (t0 + t) * KT_BN, bFull + 8 * s);
#if KT_REM
        kt_tma_2d(sB + s * KT_B_BYTES + KT_KBM * KT_TILE_BYTES(KT_BN), &map_yr, 0, (t0 + t) * KT_BN, bFull + 8 * s);
#endif
        // column penalties of this tile (only the survivor path of the epilogue reads them)
        kt_mbar_expect_tx(bHFull + 8 * h, KT_BN * 4);
        kt_bulk_1d(sH + h * KT_BN * 4, ypen + (long long)(t0 + t) * KT_BN, KT_BN * 4, bHFull + 8 * h);
      }
      __syncwarp();
#endif
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole loop =====================
    // (ncu source view of the first version of this kernel: with one warp-wide elect + reconvergence per accumulator and
    // the stage / buffer / parity arithmetic in the tile loop, the issuing warp executed ~200 dependent instructions per
    // tile -- ~880 clk against 640 clk of tensor-pipe work -- and WAS the critical path.  Now the tile loop is unrolled over
    // the stage x buffer period, so every descriptor is a loop-invariant base plus a compile-time constant.)
    if (kt_elect()) {
#if KT_STREAM
      const u64 d0 = kt_smem_desc(sB, 128);
      for (int t = 0; t < ntiles; t++) {
        const int buf = t % KT_NBUF;
        kt_mbar_wait(bTEmpty + 8 * (buf * 2 + 0), ((t / KT_NBUF) & 1) ^ 1);
        kt_mbar_wait(bTEmpty + 8 * (buf * 2 + 1), ((t / KT_NBUF) & 1) ^ 1);
        kt_fence_after();
        const u32 dt = tmem + (u32)(buf * 2 * KT_BN);
        const int it0 = t * KT_KBM;
#pragma unroll 1
        for (int kb0 = 0; kb0 < KT_KBM; kb0 += KT_STAGES) {
#pragma unroll
          for (int u = 0; u < KT_STAGES; u++) {
            const int kb = kb0 + u;
            if (kb < KT_KBM) {
              const int it = it0 + kb, s = it % KT_STAGES;
              kt_mbar_wait(bFull + 8 * s, (it / KT_STAGES) & 1);
              kt_fence_after();
              const u64 st = d0 + (u64)((u32)(s * KT_B_BYTES) >> 4);
#pragma unroll
              for (int rb = 0; rb < 2; rb++) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                  kt_mma_bf16(dt + (u32)(rb * KT_BN), st + (u64)((rb * KT_TILE_BYTES(128)) >> 4) + (u64)(2 * k),
                              st + (u64)(KT_TILE_BYTES(256) >> 4) + (u64)(2 * k), KT_IDESC, (u32)((kb | k) != 0));
              }
              kt_mma_commit(bEmpty + 8 * s);               // the stage may be refilled once these MMAs have read it
            }
          }
        }
        kt_mma_commit(bTFull + 8 * (buf * 2 + 0));         // both accumulators of the tile are complete
        kt_mma_commit(bTFull + 8 * (buf * 2 + 1));
      }
#else
      // period of (stage, buffer): tiles t = KT_PER * it + u use stage u % KT_STAGES and buffer u % KT_NBUF
      constexpr int KT_PER = (KT_STAGES % KT_NBUF == 0) ? KT_STAGES : KT_STAGES * KT_NBUF;
      const u64 a0 = kt_smem_desc(sA, 128), b0 = kt_smem_desc(sB, 128);
#if KT_REM
      const u64 ar0 = kt_smem_desc(sA + KT_KBM * KT_TILE_BYTES(128), KT_REM * 2);
      const u64 br0 = kt_smem_desc(sB + KT_KBM * KT_TILE_BYTES(KT_BN), KT_REM * 2);
#endif
      kt_mbar_wait(bAFull, 0);
#pragma unroll 1
      for (int tt = 0, it = 0; tt < ntiles; tt += KT_PER, it++) {
#pragma unroll
        for (int u = 0; u < KT_PER; u++) {
          if (tt + u < ntiles) {
            const int s = u % KT_STAGES, buf = u % KT_NBUF;
            // (t / KT_STAGES) & 1 and ((t / KT_NBUF) & 1) ^ 1 for t = KT_PER * it + u
            const u32 pf = ((u32)(it * (KT_PER / KT_STAGES)) + (u32)(u / KT_STAGES)) & 1u;
            const u32 pe = (((u32)(it * (KT_PER / KT_NBUF)) + (u32)(u / KT_NBUF)) & 1u) ^ 1u;
            kt_mbar_wait(bFull + 8 * s, pf);
#pragma unroll
            for (int rb = 0; rb < 2; rb++) {
              // each (buffer, row block) accumulator has its own full / empty pair: the other row block's MMAs
              // cover the commit -> epilogue -> release round trip of this one
              kt_mbar_wait(bTEmpty + 8 * (buf * 2 + rb), pe);
              kt_fence_after();
              const u32 d = tmem + (u32)((buf * 2 + rb) * KT_BN);
#pragma unroll
              for (int kb = 0; kb < KT_KBM; kb++) {
#pragma unroll
                for (int k = 0; k < 4; k++)            // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle atom
                  kt_mma_bf16(d, a0 + (u64)((rb * KT_BLK_BYTES(128) + kb * KT_TILE_BYTES(128)) >> 4) + (u64)(2 * k),
                              b0 + (u64)((s * KT_B_BYTES + kb * KT_TILE_BYTES(KT_BN)) >> 4) + (u64)(2 * k), KT_IDESC,
                              (u32)((kb | k) != 0));
              }
#if KT_REM
#pragma unroll
              for (int k = 0; k < KT_REM / 16; k++)
                kt_mma_bf16(d, ar0 + (u64)((rb * KT_BLK_BYTES(128)) >> 4) + (u64)(2 * k),
                            br0 + (u64)((s * KT_B_BYTES) >> 4) + (u64)(2 * k), KT_IDESC, (u32)((KT_KBM | k) != 0));
#endif
              kt_mma_commit(bTFull + 8 * (buf * 2 + rb));  // this accumulator is complete
              if (rb == 1) kt_mma_commit(bEmpty + 8 * s);  // smem stage may be refilled once these MMAs have read it
            }
          }
        }
      }
#endif
    }
    __syncwarp();
  } else {
    // ===================== epilogue: two threads per row, one per 64-column half of every tile =====================
    // (a warp reads the TMEM lanes 32 * (warp % 4) .. + 31)
    const int ew = warp - 2, q = warp & 3, rb = (ew >> 2) & 1, half = ew >> 3;
    const int rl = rb * 128 + q * 32 + lane;                       // row inside the CTA
    const long long i = row_base + rl;
    const bool live = i < nrow;
    const float cx = live ? xpen[i] : 0.0f;
    int* my_cand = cand + ((live ? i : 0) * (long long)gridDim.y + blockIdx.y) * (long long)KT_CAP;
    const float cy_scale = 2.0f * KT_C / (1.0f - KT_C);            // c|y|^2 = cy_scale * (1-c)|y|^2/2
    const u32 my_lst = sLst + (u32)(rl * KT_KP * 4);
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % KT_NBUF, hslot = t % KT_HS;
      kt_mbar_wait(bTFull + 8 * (buf * 2 + rb), (t / KT_NBUF) & 1);
      kt_fence_after();
      float u[64];
      kt_tmem_ld64_nowait(tmem + ((u32)(q * 32) << 16) + (u32)((buf * 2 + rb) * KT_BN + half * 64), u);
      float tau = tau_s[rl];                          // the row's current threshold (only ever grows: a stale value is safe)
      kt_tmem_wait_ld();
      // the accumulator buffer goes back to the MMA warp before any (data-dependent) candidate processing
      kt_fence_before();
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bTEmpty + 8 * (buf * 2 + rb));
#if KT_DBG == 2                                                     /* timing experiment: TMEM loads + release only */
      if (u[0] == 12345.678f && u[63] == 1.0f) tau_s[rl] = u[5];
      continue;
#endif
      float mg[8];                                    // maxima of the eight 8-column groups
#pragma unroll
      for (int g = 0; g < 8; g++) {
        float m = fmaxf(fmaxf(u[8 * g + 0], u[8 * g + 1]), u[8 * g + 2]);
        m = fmaxf(fmaxf(m, u[8 * g + 3]), u[8 * g + 4]);
        m = fmaxf(fmaxf(m, u[8 * g + 5]), u[8 * g + 6]);
        mg[g] = fmaxf(m, u[8 * g + 7]);
      }
      const float m01 = fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3]));
      const float m23 = fmaxf(fmaxf(mg[4], mg[5]), fmaxf(mg[6], mg[7]));
      if (fmaxf(m01, m23) >= tau) {
        // survivor path (a few hundred times per row over the whole stream): everything statically indexed
        const float* hyt = hy + hslot * KT_BN + half * 64;
        bool hy_ready = false;                        // the penalty tile is waited for on first use only
#pragma unroll
        for (int g = 0; g < 8; g++) {
          if (mg[g] >= tau) {
            u32 mask = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) mask |= (u[8 * g + e] >= tau) ? (1u << e) : 0u;
            const long long j0 = (long long)(t0 + t) * KT_BN + half * 64 + g * 8;
            const bool whole = j0 + 8 <= ncol;        // the group lies inside the column range
            if (!whole) mask = (ncol > j0) ? (mask & ((1u << (int)(ncol - j0)) - 1u)) : 0u;
            if (mask) {
              // one token per surviving group: 8-column group index (24 bits) | hit mask (8 bits)
              const int slot = atomicAdd(&cnt_s[rl], 1);
              if (slot < KT_CAP) my_cand[slot] = (int)((((u32)(j0 >> 3)) << 8) | mask);
              if (whole) {
                if (!hy_ready) { kt_mbar_wait(bHFull + 8 * hslot, (t / KT_HS) & 1); hy_ready = true; }
                // lower bound of the group's best column: its U is mg[g], its penalty is at most the largest of the hits'
                float hm = 0.0f;
                for (u32 mm = mask; mm; mm &= mm - 1) hm = fmaxf(hm, hyt[g * 8 + __ffs(mm) - 1]);
                const float lo = mg[g] - cy_scale * hm;
                if (lo > lst_s[rl * KT_KP + KT_K - 1]) {
                  // sorted insertion into the row's shared list under the row's lock (contended only by the thread
                  // that owns the other column half of this row)
                  bool done = false;
                  while (!done) {
                    if (atomicCAS(&lock_s[rl], 0, 1) == 0) {
                      float l[KT_KP];
#pragma unroll
                      for (int v4 = 0; v4 < (KT_K + 3) / 4; v4++) {
                        const float4 w = kt_lds128_volatile(my_lst + 16 * v4);
                        l[4 * v4] = w.x; l[4 * v4 + 1] = w.y; l[4 * v4 + 2] = w.z; l[4 * v4 + 3] = w.w;
                      }
                      float v = lo;
#pragma unroll
                      for (int k = 0; k < KT_K; k++) { const float hi = fmaxf(l[k], v); v = fminf(l[k], v); l[k] = hi; }
#pragma unroll
                      for (int v4 = 0; v4 < (KT_K + 3) / 4; v4++)
                        kt_sts128_volatile(my_lst + 16 * v4, l[4 * v4], l[4 * v4 + 1], l[4 * v4 + 2], l[4 * v4 + 3]);
                      tau = l[KT_K - 1] - cx;
                      tau_s[rl] = tau;
                      __threadfence_block();
                      atomicExch(&lock_s[rl], 0);
                      done = true;
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  kt_fence_before();
  __syncthreads();
  if (threadIdx.x < KT_BM && row_base + threadIdx.x < nrow)
    cand_count[(row_base + threadIdx.x) * (long long)gridDim.y + blockIdx.y] = cnt_s[threadIdx.x];
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// select: exact fp32 distances of the candidates (the CUDA-core kernel's operation order: d0*d0, then
// fma(d_k, d_k, acc) for k = 1..D-1) and the K smallest in (distance, index) order
// ------------------------------------------------------------------------------------------------
#if KT_D <= 160
#define KT_XREG 1                             /* the x row lives in registers (loaded once per warp row)        */
#else
#define KT_XREG 0                             /* large D: x is re-read through L1 (same address in every lane)  */
#endif
// y row streamed with 16-byte loads when D allows it
#if KT_XREG
__device__ __forceinline__ float kt_sqdist(const float (&xr)[KT_D], const float* __restrict__ yr) {
  float acc = 0.0f;
#if KT_D % 4 == 0
  const float4* y4 = reinterpret_cast<const float4*>(yr);
#pragma unroll
  for (int k4 = 0; k4 < KT_D / 4; k4++) {
    const float4 v = __ldg(y4 + k4);
    const float d0 = xr[4 * k4 + 0] - v.x, d1 = xr[4 * k4 + 1] - v.y, d2 = xr[4 * k4 + 2] - v.z, d3 = xr[4 * k4 + 3] - v.w;
    acc = (k4 == 0) ? d0 * d0 : fmaf(d0, d0, acc);
    acc = fmaf(d1, d1, acc);
    acc = fmaf(d2, d2, acc);
    acc = fmaf(d3, d3, acc);
  }
#else
#pragma unroll
  for (int k = 0; k < KT_D; k++) {
    const float d = xr[k] - __ldg(yr + k);
    acc = (k == 0) ? d * d : fmaf(d, d, acc);
  }
#endif
  return acc;
}
#else
__device__ __forceinline__ float kt_sqdist(const float* __restrict__ xr, const float* __restrict__ yr) {
#if KT_D % 4 == 0
  const float4* x4 = reinterpret_cast<const float4*>(xr);
  const float4* y4 = reinterpret_cast<const float4*>(yr);
  float acc;
  {
    const float4 a = __ldg(x4), v = __ldg(y4);
    const float d0 = a.x - v.x, d1 = a.y - v.y, d2 = a.z - v.z, d3 = a.w - v.w;
    acc = d0 * d0;
    acc = fmaf(d1, d1, acc);
    acc = fmaf(d2, d2, acc);
    acc = fmaf(d3, d3, acc);
  }
#pragma unroll 8
  for (int k4 = 1; k4 < KT_D / 4; k4++) {
    const float4 a = __ldg(x4 + k4), v = __ldg(y4 + k4);
    const float d0 = a.x - v.x, d1 = a.y - v.y, d2 = a.z - v.z, d3 = a.w - v.w;
    acc = fmaf(d0, d0, acc);
    acc = fmaf(d1, d1, acc);
    acc = fmaf(d2, d2, acc);
    acc = fmaf(d3, d3, acc);
  }
#else
  const float d0 = __ldg(xr) - __ldg(yr);
  float acc = d0 * d0;
#pragma unroll 8
  for (int k = 1; k < KT_D; k++) {
    const float d = __ldg(xr + k) - __ldg(yr + k);
    acc = fmaf(d, d, acc);
  }
#endif
  return acc;
}
#endif

extern "C" __global__ void __launch_bounds__(256)
kb200_knn_select(const long long nrow, const long long ncol, const float* __restrict__ x, const float* __restrict__ y,
                 const int* __restrict__ cand, const int* __restrict__ cand_count, float* __restrict__ out,
                 int* __restrict__ overflow_count, int* __restrict__ overflow_rows, const int nsplit,
                 float* __restrict__ part, const long long row_offset) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (row, split) pair
  const long long i = w / nsplit;
  if (i >= nrow) return;
  const int* tok = cand + w * (long long)KT_CAP;      // this pair's token list
  const int n = cand_count[w];
  constexpr int OUTW = (KT_OUT == 2) ? 2 * KT_K : KT_K;
  if (n > KT_CAP) {                                   // token list overflowed: the caller recomputes this row
    if (lane == 0) { const int s = atomicAdd(overflow_count, 1); overflow_rows[s] = (int)(i + row_offset); }   // global row index
    return;
  }
  // every lane keeps the K best (distance, index) pairs of the candidates it evaluated, ascending
  float lv[KT_K];
  int lj[KT_K];
#pragma unroll
  for (int l = 0; l < KT_K; l++) { lv[l] = KT_INF; lj[l] = 0x7fffffff; }
#if KT_XREG
  float xr[KT_D];
#pragma unroll
  for (int k = 0; k < KT_D; k++) xr[k] = __ldg(x + i * KT_D + k);
#else
  const float* xr = x + i * KT_D;
#endif
  for (int p = lane; p < n; p += 32) {
    const u32 tk = (u32)tok[p];
    const long long j8 = (long long)(tk >> 8) << 3;
    for (u32 mask = tk & 0xffu; mask; mask &= mask - 1) {
      const int j = (int)(j8 + __ffs(mask) - 1);
      const float d = kt_sqdist(xr, y + (long long)j * KT_D);
      if (d < lv[KT_K - 1] || (d == lv[KT_K - 1] && j < lj[KT_K - 1])) {
        float cv = d;
        int cj = j;
#pragma unroll
        for (int m = 0; m < KT_K; m++) {                // statically indexed insertion
          if (cv < lv[m] || (cv == lv[m] && cj < lj[m])) {
            const float tv = lv[m]; const int tj = lj[m];
            lv[m] = cv; lj[m] = cj; cv = tv; cj = tj;
          }
        }
      }
    }
  }
  for (int l = 0; l < KT_K; l++) {
    float wv = lv[0];
    int wj = lj[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, wj, o);
      if (ov < wv || (ov == wv && oj < wj)) { wv = ov; wj = oj; }
    }
    if (wj == lj[0] && wj != 0x7fffffff) {            // this lane owned the winner (column indices are unique): pop it
#pragma unroll
      for (int m = 0; m + 1 < KT_K; m++) { lv[m] = lv[m + 1]; lj[m] = lj[m + 1]; }
      lv[KT_K - 1] = KT_INF;
      lj[KT_K - 1] = 0x7fffffff;
    }
    if (lane == 0) {
      const bool none = (wj == 0x7fffffff);           // fewer than K columns: (inf, 0) like the reference's init
      const float val = none ? KT_INF : wv;
      const float idx = none ? 0.0f : (float)wj;
      if (nsplit > 1) { part[(w * KT_K + l) * 2] = val; part[(w * KT_K + l) * 2 + 1] = idx; continue; }
      float* po = out + i * OUTW;
      if (KT_OUT == 0) po[l] = idx;
      else if (KT_OUT == 1) po[l] = val;
      else { po[2 * l] = val; po[2 * l + 1] = idx; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// merge: the per-split exact top-K lists of a row (ascending (value, index) order, splits in column order) -> its
// global top-K; ties keep the smallest index because a later split only wins with a strictly smaller value
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
kb200_knn_merge(const long long nrow, const int nsplit, const float* __restrict__ part, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrow) return;
  constexpr int OUTW = (KT_OUT == 2) ? 2 * KT_K : KT_K;
  float v[KT_K], ix[KT_K];
  const float* p0 = part + i * (long long)nsplit * KT_K * 2;
#pragma unroll
  for (int l = 0; l < KT_K; l++) { v[l] = p0[2 * l]; ix[l] = p0[2 * l + 1]; }
  for (int s = 1; s < nsplit; s++) {
    const float* ps = p0 + (long long)s * KT_K * 2;
    for (int l = 0; l < KT_K; l++) {
      const float f = ps[2 * l], fi = ps[2 * l + 1];
      if (!(f < v[KT_K - 1])) break;                  // lists are sorted: nothing further can enter
      // insertion from the tail, strict comparison (the element already present wins ties)
      float cv = f, ci = fi;
#pragma unroll
      for (int m = 0; m < KT_K; m++) {
        if (cv < v[m]) { const float tv = v[m], ti = ix[m]; v[m] = cv; ix[m] = ci; cv = tv; ci = ti; }
      }
    }
  }
  float* po = out + i * OUTW;
#pragma unroll
  for (int l = 0; l < KT_K; l++) {
    if (KT_OUT == 0) po[l] = ix[l];
    else if (KT_OUT == 1) po[l] = v[l];
    else { po[2 * l] = v[l]; po[2 * l + 1] = ix[l]; }
  }
}
