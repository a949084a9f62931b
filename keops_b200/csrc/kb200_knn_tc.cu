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
//   warps 2..9 = epilogue: one thread per row, tcgen05.ld of its 128 accumulators, FMNMX3 max-reduction against
//   tau_i; survivors are appended to the row's candidate list and update Theta_i (register-resident list).
//   4-stage smem ring, double-buffered TMEM accumulators (2 row blocks x 2 buffers x 128 columns = 512).
//   (Measured alternative, not kept: two threads per row, one per 64-column half as in kb200_conv_tc -- ncu shows the
//   one-thread-per-row epilogue latency-bound (tensor pipe 47 %, ALU 40 %, issue 32 % busy) -- but each half then
//   thresholds on its own K-th best bound: 324 instead of ~200 candidates per row at D = 64, K = 10, N = 1e6, the
//   filter 158 -> 180 ms and most rows overflow the candidate slots.)
// kb200_knn_prep   : fp32 -> bf16 operand rows (data, penalty columns, zero padding) + per-row penalty terms.
// kb200_knn_select : one warp per row, exact distances of the candidates + ordered top-K (ties -> smallest j).
//
// Compile-time parameters (defined by keops_b200/codegen/knn_tc.py): KT_D (true dimension); the bf16 operand
// rows hold D + 3 columns padded to KT_DPG = 64 * KT_KBM + KT_REM (KT_KBM 128-byte-swizzled 64-column blocks
// plus one remainder block of KT_REM = 16 / 32 / 64 columns swizzled at its own width); KT_K, KT_CAP (candidate
// slots per row, multiple of 32), KT_OUT (0 ArgKMin, 1 KMin, 2 KMin_ArgKMin), KT_FAST (1: FTZ arithmetic of
// --use_fast_math, like the CUDA-core kernel).

#ifndef KT_D
#define KT_D 64
#define KT_KBM 1
#define KT_REM 16
#define KT_K 10
#define KT_CAP 320
#define KT_OUT 0
#define KT_FAST 1
#endif

#define KT_BM 256                             /* rows per CTA: two 128-row accumulators                         */
#ifndef KT_BN
#define KT_BN 128                             /* columns per tile (= N of one tcgen05.mma): 64 or 128           */
#endif
#ifndef KT_PIPE
#define KT_PIPE 0                             /* 1: software-pipelined TMEM reads in the filter's epilogue (measured
                                                 on B200, D = 64, M = N = 1e6: 183 ms vs 162 ms un-pipelined -- kept as an
                                                 opt-in experiment)                                              */
#endif
#ifndef KT_DBG
#define KT_DBG 0                              /* 1: no rare path, 2: TMEM loads only, 3: no TMEM loads (timing experiments) */
#endif
#define KT_NBUF (512 / (2 * KT_BN))           /* TMEM accumulator buffers per row block (all 512 columns used)   */
#define KT_DPG (64 * KT_KBM + KT_REM)
#ifndef KT_STAGES
#define KT_STAGES 4
#endif
#define KT_HS 16                              /* slots of the column-penalty ring                               */
#define KT_C 0.008f
#define KT_TILE_BYTES(rows) ((rows) * 128)                    /* rows x 64 bf16                                 */
#define KT_RTILE_BYTES(rows) ((rows) * KT_REM * 2)            /* rows x KT_REM bf16                             */
#define KT_BLK_BYTES(rows) (KT_KBM * KT_TILE_BYTES(rows) + KT_RTILE_BYTES(rows))   /* one operand block, all of K */
#define KT_A_BYTES (2 * KT_BLK_BYTES(128))
#define KT_B_BYTES KT_BLK_BYTES(KT_BN)
#define KT_SMEM_BYTES (1024 + KT_A_BYTES + KT_STAGES * KT_B_BYTES + KT_HS * KT_BN * 4 + 640) /* mirrored in knn_tc.py */


// bf16 x bf16 -> f32, A and B K-major, M = 128, N = KT_BN
#define KT_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | (((u32)KT_BN >> 3) << 17) | ((128u >> 4) << 24))

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
extern "C" __global__ void __launch_bounds__(320, 1)
kb200_knn_filter(const __grid_constant__ KtTensorMap map_x, const __grid_constant__ KtTensorMap map_xr,
                 const __grid_constant__ KtTensorMap map_y, const __grid_constant__ KtTensorMap map_yr,
                 const long long nrow, const long long ncol, const float* __restrict__ ypen,
                 const float* __restrict__ xpen, int* __restrict__ cand, int* __restrict__ cand_count,
                 const int tiles_per_split) {
  extern __shared__ unsigned char kt_raw[];
  const u32 raw = kt_smem(kt_raw);
  const u32 base = (raw + 1023u) & ~1023u;                       // swizzle atoms need (up to) 1024-byte alignment
  unsigned char* gbase = kt_raw + (base - raw);
  const u32 sA = base;                                           // 2 row blocks x [KBM main tiles | remainder tile]
  const u32 sB = sA + KT_A_BYTES;                                // STAGES x [KBM main tiles | remainder tile]
  const u32 sH = sB + KT_STAGES * KT_B_BYTES;                    // ring of column penalties (KT_HS x KT_BN floats)
  const float* hy = reinterpret_cast<const float*>(gbase + KT_A_BYTES + KT_STAGES * KT_B_BYTES);
  const u32 sBar = sH + KT_HS * KT_BN * 4;
  // barriers (8 bytes each): full[8] @0, empty[8] @64, tfull[buf][rb] (<= 8) @128, tempty[buf][rb] @192,
  //                          hfull[16] @256, hempty[16] @384, afull @512, tmem pointer @520
  const u32 bFull = sBar, bEmpty = sBar + 64, bTFull = sBar + 128, bTEmpty = sBar + 192, bAFull = sBar + 512;
  const u32 bHFull = sBar + 256, bHEmpty = sBar + 384;
  volatile u32* tmem_slot =
      reinterpret_cast<volatile u32*>(gbase + KT_A_BYTES + KT_STAGES * KT_B_BYTES + KT_HS * KT_BN * 4 + 520);

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
    for (int b = 0; b < 2 * KT_NBUF; b++) { kt_mbar_init(bTFull + 8 * b, 1); kt_mbar_init(bTEmpty + 8 * b, 4); }
    for (int h = 0; h < KT_HS; h++) { kt_mbar_init(bHFull + 8 * h, 1); kt_mbar_init(bHEmpty + 8 * h, 8); }
    kt_mbar_init(bAFull, 1);
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
    if (kt_elect()) {
      kt_mbar_expect_tx(bAFull, KT_A_BYTES);
      for (int rb = 0; rb < 2; rb++) {
        for (int kb = 0; kb < KT_KBM; kb++)
          kt_tma_2d(sA + rb * KT_BLK_BYTES(128) + kb * KT_TILE_BYTES(128), &map_x, kb * 64, (int)(row_base + rb * 128), bAFull);
        kt_tma_2d(sA + rb * KT_BLK_BYTES(128) + KT_KBM * KT_TILE_BYTES(128), &map_xr, 0, (int)(row_base + rb * 128), bAFull);
      }
    }
    __syncwarp();
    for (int t = 0; t < ntiles; t++) {
      const int s = t % KT_STAGES, h = t % KT_HS;
      kt_mbar_wait(bEmpty + 8 * s, ((t / KT_STAGES) & 1) ^ 1);
      kt_mbar_wait(bHEmpty + 8 * h, ((t / KT_HS) & 1) ^ 1);
      if (kt_elect()) {
        kt_mbar_expect_tx(bFull + 8 * s, KT_B_BYTES);
        for (int kb = 0; kb < KT_KBM; kb++)
          kt_tma_2d(sB + s * KT_B_BYTES + kb * KT_TILE_BYTES(KT_BN), &map_y, kb * 64, (t0 + t) * KT_BN, bFull + 8 * s);
        kt_tma_2d(sB + s * KT_B_BYTES + KT_KBM * KT_TILE_BYTES(KT_BN), &map_yr, 0, (t0 + t) * KT_BN, bFull + 8 * s);
        // column penalties of this tile (only the rare path of the epilogue reads them)
        kt_mbar_expect_tx(bHFull + 8 * h, KT_BN * 4);
        kt_bulk_1d(sH + h * KT_BN * 4, ypen + (long long)(t0 + t) * KT_BN, KT_BN * 4, bHFull + 8 * h);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    kt_mbar_wait(bAFull, 0);
    for (int t = 0; t < ntiles; t++) {
      const int s = t % KT_STAGES, buf = t % KT_NBUF;
      kt_mbar_wait(bFull + 8 * s, (t / KT_STAGES) & 1);
#pragma unroll
      for (int rb = 0; rb < 2; rb++) {
        // each (buffer, row block) accumulator has its own full / empty pair: the other row block's MMAs
        // cover the commit -> epilogue -> release round trip of this one
        kt_mbar_wait(bTEmpty + 8 * (buf * 2 + rb), ((t / KT_NBUF) & 1) ^ 1);
        kt_fence_after();
        if (kt_elect()) {
          const u32 d = tmem + (u32)((buf * 2 + rb) * KT_BN);
#pragma unroll
          for (int kb = 0; kb < KT_KBM; kb++) {
            const u64 ad = kt_smem_desc(sA + rb * KT_BLK_BYTES(128) + kb * KT_TILE_BYTES(128), 128);
            const u64 bd = kt_smem_desc(sB + s * KT_B_BYTES + kb * KT_TILE_BYTES(KT_BN), 128);
#pragma unroll
            for (int k = 0; k < 4; k++) {            // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle atom
              if (KT_DBG != 6) kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), KT_IDESC, (u32)((kb | k) != 0));
              if (KT_DBG == 7) kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), KT_IDESC, 1u);
            }
          }
          {
            const u64 ad = kt_smem_desc(sA + rb * KT_BLK_BYTES(128) + KT_KBM * KT_TILE_BYTES(128), KT_REM * 2);
            const u64 bd = kt_smem_desc(sB + s * KT_B_BYTES + KT_KBM * KT_TILE_BYTES(KT_BN), KT_REM * 2);
#pragma unroll
            for (int k = 0; k < KT_REM / 16; k++)
              kt_mma_bf16(d, ad + (u64)(2 * k), bd + (u64)(2 * k), KT_IDESC, (u32)((KT_KBM | k) != 0));
          }
          kt_mma_commit(bTFull + 8 * (buf * 2 + rb));  // this accumulator is complete
          if (rb == 1) kt_mma_commit(bEmpty + 8 * s);  // smem stage may be refilled once these MMAs have read it
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: one thread per row =====================
    const int q = warp & 3, rb = (warp - 2) >> 2;
    const long long i = row_base + rb * 128 + q * 32 + lane;
    const bool live = i < nrow;
    const float cx = live ? xpen[i] : 0.0f;
    float tau = live ? -KT_INF : KT_INF;
    float lst[KT_K];                                  // K largest lower bounds, descending
#pragma unroll
    for (int l = 0; l < KT_K; l++) lst[l] = -KT_INF;
    int cnt = 0;
    int* my_cand = cand + ((live ? i : 0) * (long long)gridDim.y + blockIdx.y) * (long long)KT_CAP;
    const float cy_scale = 2.0f * KT_C / (1.0f - KT_C);   // c|y|^2 = cy_scale * (1-c)|y|^2/2
    // One 64-column chunk of a tile: FMNMX max-reduction against tau_i, rare path for the survivors.
    auto process = [&](const float* u, const int t, const int c, const float* hyt, const int hslot, bool& hy_ready) {
      float mg[8];                                    // maxima of the eight 8-column sub-groups
#pragma unroll
      for (int g = 0; g < 8; g++) {
        float m = fmaxf(fmaxf(u[8 * g + 0], u[8 * g + 1]), u[8 * g + 2]);
        m = fmaxf(fmaxf(m, u[8 * g + 3]), u[8 * g + 4]);
        m = fmaxf(fmaxf(m, u[8 * g + 5]), u[8 * g + 6]);
        mg[g] = fmaxf(m, u[8 * g + 7]);
      }
      const float m01 = fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3]));
      const float m23 = fmaxf(fmaxf(mg[4], mg[5]), fmaxf(mg[6], mg[7]));
#if KT_DBG >= 2
      if (fmaxf(m01, m23) == 12345.678f) cnt++;
#endif
      if (KT_DBG == 0 && fmaxf(m01, m23) >= tau) {
        // rare path (a few hundred times per row over the whole stream): everything statically indexed
#pragma unroll
        for (int g = 0; g < 8; g++) {
          if (mg[g] >= tau) {
            u32 hits = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) hits |= (u[8 * g + e] >= tau) ? (1u << e) : 0u;
            const long long j0 = (long long)(t0 + t) * KT_BN + c * 64 + g * 8;
            while (hits) {
              const int e = __ffs(hits) - 1;
              hits &= hits - 1;
              // 8 -> 1 select tree instead of a dynamically indexed (local-memory) array
              const float a0 = (e & 1) ? u[8 * g + 1] : u[8 * g + 0], a1 = (e & 1) ? u[8 * g + 3] : u[8 * g + 2];
              const float a2 = (e & 1) ? u[8 * g + 5] : u[8 * g + 4], a3 = (e & 1) ? u[8 * g + 7] : u[8 * g + 6];
              const float b0 = (e & 2) ? a1 : a0, b1 = (e & 2) ? a3 : a2;
              const float ue = (e & 4) ? b1 : b0;
              if (ue >= tau && j0 + e < ncol) {
                if (cnt < KT_CAP) my_cand[cnt] = (int)(j0 + e);
                cnt++;
                if (!hy_ready) { kt_mbar_wait(bHFull + 8 * hslot, (t / KT_HS) & 1); hy_ready = true; }
                const float lo = ue - cy_scale * hyt[c * 64 + g * 8 + e];
                if (lo > lst[KT_K - 1]) {
#pragma unroll
                  for (int l = KT_K - 1; l > 0; --l) {
                    if (lo > lst[l - 1]) lst[l] = lst[l - 1];
                    else if (lo > lst[l]) lst[l] = lo;
                  }
                  if (lo > lst[0]) lst[0] = lo;
                  tau = lst[KT_K - 1] - cx;
                }
              }
            }
          }
        }
      }
    };
#if KT_PIPE && KT_BN == 128 && KT_DBG < 3
    // Software-pipelined TMEM reads (ncu of the un-pipelined loop: every epilogue warp idles ~2.4 cycles per issued
    // instruction on the tcgen05.ld round trip -- long_scoreboard -- with only two such warps per scheduler): the 64-column
    // half that is processed next is always in flight while the current one is reduced, across tile boundaries too.
    // uA / uB are two statically indexed register buffers; tcgen05.wait::ld retires every load issued so far.
    float uA[64], uB[64];
    if (ntiles > 0) {
      kt_mbar_wait(bTFull + 8 * (0 * 2 + rb), 0);
      kt_fence_after();
      kt_tmem_ld64_nowait(tmem + ((u32)(q * 32) << 16) + (u32)((0 * 2 + rb) * KT_BN), uA);
    }
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % KT_NBUF, hslot = t % KT_HS;
      const float* hyt = hy + hslot * KT_BN;
      bool hy_ready = false;                          // the penalty tile is waited for on first use only
      const u32 tbase = tmem + ((u32)(q * 32) << 16) + (u32)((buf * 2 + rb) * KT_BN);
      kt_tmem_wait_ld();                              // uA = columns 0..63 of tile t
      kt_tmem_ld64_nowait(tbase + 64u, uB);           // columns 64..127 fly while uA is reduced
      process(uA, t, 0, hyt, hslot, hy_ready);
      kt_tmem_wait_ld();                              // uB landed: the accumulator buffer goes back to the MMA warp
      kt_fence_before();
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bTEmpty + 8 * (buf * 2 + rb));
      if (t + 1 < ntiles) {
        const int nb = (t + 1) % KT_NBUF;
        kt_mbar_wait(bTFull + 8 * (nb * 2 + rb), ((t + 1) / KT_NBUF) & 1);
        kt_fence_after();
        kt_tmem_ld64_nowait(tmem + ((u32)(q * 32) << 16) + (u32)((nb * 2 + rb) * KT_BN), uA);
      }
      process(uB, t, 1, hyt, hslot, hy_ready);
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bHEmpty + 8 * hslot);
    }
#else
    for (int t = 0; t < ntiles; t++) {
      const int buf = t % KT_NBUF, hslot = t % KT_HS;
      kt_mbar_wait(bTFull + 8 * (buf * 2 + rb), (t / KT_NBUF) & 1);
      kt_fence_after();
      const float* hyt = hy + hslot * KT_BN;
      bool hy_ready = false;                          // the penalty tile is waited for on first use only
      const u32 tbase = tmem + ((u32)(q * 32) << 16) + (u32)((buf * 2 + rb) * KT_BN);
      // the whole row of this tile goes to registers first, so that the accumulator buffer is handed back to
      // the MMA warp before any (data-dependent) candidate processing
      float uu[KT_BN];
#if KT_DBG >= 3
#pragma unroll
      for (int e = 0; e < KT_BN; e++) uu[e] = -1.0f;
#else
#pragma unroll
      for (int c = 0; c < KT_BN / 64; c++) kt_tmem_ld64_nowait(tbase + (u32)(c * 64), uu + c * 64);
      kt_tmem_wait_ld();
#endif
      kt_fence_before();
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bTEmpty + 8 * (buf * 2 + rb));
#pragma unroll
      for (int c = 0; c < KT_BN / 64; c++) process(uu + c * 64, t, c, hyt, hslot, hy_ready);
      __syncwarp();
      if (lane == 0) kt_mbar_arrive(bHEmpty + 8 * hslot);
    }
#endif
    if (live) cand_count[i * (long long)gridDim.y + blockIdx.y] = cnt;
  }
  kt_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// select: exact fp32 distances of the candidates (the CUDA-core kernel's operation order: d0*d0, then
// fma(d_k, d_k, acc) for k = 1..D-1) and the K smallest in (distance, index) order
// ------------------------------------------------------------------------------------------------
// x row in registers (loaded once per warp row), y row streamed with 16-byte loads when D allows it
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

extern "C" __global__ void __launch_bounds__(256)
kb200_knn_select(const long long nrow, const long long ncol, const float* __restrict__ x, const float* __restrict__ y,
                 const int* __restrict__ cand, const int* __restrict__ cand_count, float* __restrict__ out,
                 int* __restrict__ overflow_count, int* __restrict__ overflow_rows, const int nsplit,
                 float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (row, split) pair
  const long long i = w / nsplit;
  if (i >= nrow) return;
  cand += w * (long long)KT_CAP - i * (long long)KT_CAP;      // below, cand[i * KT_CAP + e] addresses this pair's list
  const int n = cand_count[w];
  constexpr int OUTW = (KT_OUT == 2) ? 2 * KT_K : KT_K;
  if (n > KT_CAP) {                                   // candidate list overflowed: the caller recomputes this row
    if (lane == 0) { const int s = atomicAdd(overflow_count, 1); overflow_rows[s] = (int)i; }
    return;
  }
  constexpr int PER = KT_CAP / 32;
  float dv[PER];
  int dj[PER];
  float xr[KT_D];
#pragma unroll
  for (int k = 0; k < KT_D; k++) xr[k] = __ldg(x + i * KT_D + k);
#pragma unroll
  for (int p = 0; p < PER; p++) {
    const int e = p * 32 + lane;
    dv[p] = KT_INF;
    dj[p] = 0x7fffffff;
    if (e < n) {
      const int j = cand[i * (long long)KT_CAP + e];
      dj[p] = j;
      dv[p] = kt_sqdist(xr, y + (long long)j * KT_D);
    }
  }
  for (int l = 0; l < KT_K; l++) {
    float bv = dv[0];
    int bj = dj[0], bp = 0;
#pragma unroll
    for (int p = 1; p < PER; p++)
      if (dv[p] < bv || (dv[p] == bv && dj[p] < bj)) { bv = dv[p]; bj = dj[p]; bp = p; }
    float wv = bv;
    int wj = bj;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, wj, o);
      if (ov < wv || (ov == wv && oj < wj)) { wv = ov; wj = oj; }
    }
    if (wj == bj && bj != 0x7fffffff) {               // this lane owned the winner: retire it
#pragma unroll
      for (int p = 0; p < PER; p++)
        if (p == bp) { dv[p] = KT_INF; dj[p] = 0x7fffffff; }
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
