// kb200_tc_common.cuh -- PTX wrappers shared by the tensor-core kernels (kb200_knn_tc.cu, kb200_conv_tc.cu):
// mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05.mma / commit / ld, shared-memory matrix descriptors and bf16
// rounding helpers for sm_100a.  NVRTC has no include path here: the host code (codegen/knn_tc.py,
// codegen/conv_tc.py) prepends this text to the kernel source.
#ifndef KB200_TC_COMMON_CUH
#define KB200_TC_COMMON_CUH
typedef unsigned long long u64;
typedef unsigned int u32;

#ifndef KT_SPIN
#define KT_SPIN 0                             /* 1: poll barriers with test_wait instead of try_wait             */
#endif
#define KT_INF (__int_as_float(0x7f800000))

struct __align__(64) KtTensorMap { u64 opaque[16]; };


// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 kt_smem(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void kt_mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kt_mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void kt_mbar_wait(u32 bar, u32 parity) {
  // bounded spin: a protocol bug traps (the launch fails loudly) instead of hanging the GPU
  for (u32 spins = 0;; ++spins) {
    u32 done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
#if KT_SPIN
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#endif
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (spins > (1u << 28)) __trap();
  }
}
// non-blocking probe of a barrier phase
__device__ __forceinline__ bool kt_mbar_test(u32 bar, u32 parity) {
  u32 done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void kt_tma_2d(u32 dst, const KtTensorMap* map, int c0, int c1, u32 bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"((u64)map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// TMA tile load multicast to the CTAs of `mask` in the cluster: the tile lands at the same shared-memory offset in every
// destination CTA and completes bytes on the mbarrier at the same offset there
__device__ __forceinline__ void kt_tma_2d_mc(u32 dst, const KtTensorMap* map, int c0, int c1, u32 bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"((u64)map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void kt_mma_commit_mc(u32 bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ u32 kt_cluster_ctarank() {
  u32 r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void kt_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void kt_bulk_1d(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"((u64)src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void kt_mma_bf16(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (lane = row, two consecutive K elements per 32-bit column), B through a smem descriptor
__device__ __forceinline__ void kt_mma_bf16_ts(u32 tmem_d, u32 tmem_a, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: thread `lane` of the warp writes 32 consecutive 32-bit columns of its TMEM lane
__device__ __forceinline__ void kt_tmem_st32(u32 taddr, const u32* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void kt_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void kt_mma_commit(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a *converged* warp; keeping the warp converged lets the compiler hold descriptors and
// barrier addresses in uniform registers (a divergent `if (lane == 0)` wraps every UTCHMMA in an ELECT loop)
__device__ __forceinline__ bool kt_elect() {
  u32 pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void kt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void kt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 64 consecutive 32-bit columns: thread `lane` of the warp receives columns [col, col+64) of its TMEM lane
#define KT_R8(b) "%" #b
__device__ __forceinline__ void kt_tmem_ld64_nowait(u32 taddr, float* v) {
  u32 r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int e = 0; e < 64; e++) v[e] = __uint_as_float(r[e]);
}
// 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void kt_tmem_ld32_nowait(u32 taddr, float* v) {
  u32 r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int e = 0; e < 32; e++) v[e] = __uint_as_float(r[e]);
}
__device__ __forceinline__ void kt_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile whose rows are `row_bytes` (32 / 64 / 128) wide and swizzled at that width: rows are
// row_bytes apart, 8-row groups 8 * row_bytes apart (SBO); layout type 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B
__device__ __forceinline__ u64 kt_smem_desc(u32 saddr, u32 row_bytes) {
  const u64 layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  return (u64)((saddr & 0x3FFFF) >> 4) | ((u64)1 << 16) | ((u64)((8 * row_bytes) >> 4) << 32) | ((u64)1 << 46) | (layout << 61);
}

__device__ __forceinline__ unsigned short kt_bf16_rn(float v) {
  u32 b = __float_as_uint(v);
  b += 0x7FFFu + ((b >> 16) & 1u);            // round to nearest even (finite inputs)
  return (unsigned short)(b >> 16);
}
__device__ __forceinline__ float kt_bf16_to_f(unsigned short h) { return __uint_as_float(((u32)h) << 16); }

#endif  // KB200_TC_COMMON_CUH
