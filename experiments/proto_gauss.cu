// Prototype / micro-benchmarks used to choose the kernel skeleton emitted by
// keops_b200.codegen.  Not part of the product path.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -lineinfo -o proto_gauss proto_gauss.cu
// Gaussian conv  out_i = sum_j exp(-|x_i-y_j|^2) b_j   (D=3), coordinates pre-scaled by
// sqrt(log2 e) so the exponential is a bare ex2.approx.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---------------------------------------------------------------- V0: reference-like
__global__ void v0_ref(long nx, long ny, const float* __restrict__ x, const float* __restrict__ y,
                       const float* __restrict__ b, float* __restrict__ out) {
  extern __shared__ float yj[];
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  float xi[3] = {0, 0, 0};
  float acc = 0.f;
  if (i < nx) { xi[0] = x[i * 3]; xi[1] = x[i * 3 + 1]; xi[2] = x[i * 3 + 2]; }
  for (long jstart = 0, tile = 0; jstart < ny; jstart += blockDim.x, tile++) {
    long j = tile * blockDim.x + threadIdx.x;
    if (j < ny) {
      yj[threadIdx.x * 4 + 0] = y[j * 3]; yj[threadIdx.x * 4 + 1] = y[j * 3 + 1];
      yj[threadIdx.x * 4 + 2] = y[j * 3 + 2]; yj[threadIdx.x * 4 + 3] = b[j];
    }
    __syncthreads();
    if (i < nx) {
      float tmp = 0.f;
      float* yjrel = yj;
      for (long jrel = 0; (jrel < blockDim.x) && (jrel < ny - jstart); jrel++, yjrel += 4) {
        float d0 = xi[0] - yjrel[0], d1 = xi[1] - yjrel[1], d2 = xi[2] - yjrel[2];
        float s = d0 * d0; s += d1 * d1; s += d2 * d2;
        tmp += __expf(-s) * yjrel[3];
      }
      acc += tmp;
    }
    __syncthreads();
  }
  if (i < nx) out[i] = acc;
}

// ---------------------------------------------------------------- V1: scalar, R rows / thread
template <int R, int TJ, int UNR>
__global__ void __launch_bounds__(256) v1_scalar(long nx, long ny, const float* __restrict__ x,
                                                  const float* __restrict__ y, const float* __restrict__ b,
                                                  float* __restrict__ out, float scale) {
  __shared__ float4 sm[2][TJ];
  const int T = blockDim.x;
  long i0 = (long)blockIdx.x * T * R + threadIdx.x;
  float xi[R][3], acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    long i = min(i0 + (long)r * T, nx - 1);
#pragma unroll
    for (int k = 0; k < 3; k++) xi[r][k] = x[i * 3 + k] * scale;
    acc[r] = 0.f;
  }
  long ntiles = (ny + TJ - 1) / TJ;
  // prologue: stage tile 0
  for (int t = threadIdx.x; t < TJ; t += T) {
    long j = t; float4 v = make_float4(0, 0, 0, 0);
    if (j < ny) v = make_float4(y[j * 3] * scale, y[j * 3 + 1] * scale, y[j * 3 + 2] * scale, b[j]);
    sm[0][t] = v;
  }
  __syncthreads();
  for (long tile = 0; tile < ntiles; tile++) {
    int buf = tile & 1;
    // prefetch next tile into registers (TJ <= 2*T assumed)
    float4 pre[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int t = threadIdx.x + q * T;
      long j = (tile + 1) * TJ + t;
      pre[q] = make_float4(0, 0, 0, 0);
      if (t < TJ && j < ny) pre[q] = make_float4(y[j * 3] * scale, y[j * 3 + 1] * scale, y[j * 3 + 2] * scale, b[j]);
    }
    float tmp[R];
#pragma unroll
    for (int r = 0; r < R; r++) tmp[r] = 0.f;
    int jn = (int)min((long)TJ, ny - tile * TJ);
#pragma unroll UNR
    for (int jr = 0; jr < jn; jr++) {
      float4 v = sm[buf][jr];
#pragma unroll
      for (int r = 0; r < R; r++) {
        float d0 = xi[r][0] - v.x, d1 = xi[r][1] - v.y, d2 = xi[r][2] - v.z;
        float s = -d0 * d0; s = fmaf(-d1, d1, s); s = fmaf(-d2, d2, s);
        tmp[r] = fmaf(ex2(s), v.w, tmp[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += tmp[r];
#pragma unroll
    for (int q = 0; q < 2; q++) { int t = threadIdx.x + q * T; if (t < TJ) sm[buf ^ 1][t] = pre[q]; }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < R; r++) { long i = i0 + (long)r * T; if (i < nx) out[i] = acc[r]; }
}

// ---------------------------------------------------------------- V2: packed f32x2 over pairs of rows
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d; asm("sub.ftz.f32x2 %0, %1, %2;" : "=l"(*(unsigned long long*)&d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b)); return d;
}
__device__ __forceinline__ float2 nmul2(float2 a, float2 b) {  // -(a*b)
  float2 na = make_float2(-a.x, -a.y); return __fmul2_rn(na, b);
}
__device__ __forceinline__ float2 nfma2(float2 a, float2 b, float2 c) {  // -(a*b)+c
  float2 na = make_float2(-a.x, -a.y); return __ffma2_rn(na, b, c);
}

template <int P, int TJ, int UNR>   // P packs = 2P rows per thread
__global__ void __launch_bounds__(256) v2_packed(long nx, long ny, const float* __restrict__ x,
                                                  const float* __restrict__ y, const float* __restrict__ b,
                                                  float* __restrict__ out, float scale) {
  __shared__ float4 sm[2][TJ];
  const int T = blockDim.x;
  constexpr int R = 2 * P;
  long i0 = (long)blockIdx.x * T * R + threadIdx.x;
  float2 xi[P][3], acc[P];
  const float2 nl2e = make_float2(-1.4426950408889634f, -1.4426950408889634f);
  scale = 1.0f;
#pragma unroll
  for (int p = 0; p < P; p++) {
    long ia = min(i0 + (long)(2 * p) * T, nx - 1), ib = min(i0 + (long)(2 * p + 1) * T, nx - 1);
#pragma unroll
    for (int k = 0; k < 3; k++) xi[p][k] = make_float2(x[ia * 3 + k] * scale, x[ib * 3 + k] * scale);
    acc[p] = make_float2(0.f, 0.f);
  }
  long ntiles = (ny + TJ - 1) / TJ;
  for (int t = threadIdx.x; t < TJ; t += T) {
    long j = t; float4 v = make_float4(0, 0, 0, 0);
    if (j < ny) v = make_float4(y[j * 3] * scale, y[j * 3 + 1] * scale, y[j * 3 + 2] * scale, b[j]);
    sm[0][t] = v;
  }
  __syncthreads();
  for (long tile = 0; tile < ntiles; tile++) {
    int buf = tile & 1;
    float4 pre[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int t = threadIdx.x + q * T;
      long j = (tile + 1) * TJ + t;
      pre[q] = make_float4(0, 0, 0, 0);
      if (t < TJ && j < ny) pre[q] = make_float4(y[j * 3] * scale, y[j * 3 + 1] * scale, y[j * 3 + 2] * scale, b[j]);
    }
    float2 tmp[P];
#pragma unroll
    for (int p = 0; p < P; p++) tmp[p] = make_float2(0.f, 0.f);
    int jn = (int)min((long)TJ, ny - tile * TJ);
#pragma unroll UNR
    for (int jr = 0; jr < jn; jr++) {
      float4 v = sm[buf][jr];
      float2 y0 = make_float2(v.x, v.x), y1 = make_float2(v.y, v.y), y2 = make_float2(v.z, v.z), bb = make_float2(v.w, v.w);
#pragma unroll
      for (int p = 0; p < P; p++) {
        float2 d0 = sub2(xi[p][0], y0), d1 = sub2(xi[p][1], y1), d2 = sub2(xi[p][2], y2);
        float2 s = __fmul2_rn(d0, d0); s = __ffma2_rn(d1, d1, s); s = __ffma2_rn(d2, d2, s);
        s = __fmul2_rn(s, nl2e);
        float2 e = make_float2(ex2(s.x), ex2(s.y));
        tmp[p] = __ffma2_rn(e, bb, tmp[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < P; p++) acc[p] = __fadd2_rn(acc[p], tmp[p]);
#pragma unroll
    for (int q = 0; q < 2; q++) { int t = threadIdx.x + q * T; if (t < TJ) sm[buf ^ 1][t] = pre[q]; }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < P; p++) {
    long ia = i0 + (long)(2 * p) * T, ib = i0 + (long)(2 * p + 1) * T;
    if (ia < nx) out[ia] = acc[p].x;
    if (ib < nx) out[ib] = acc[p].y;
  }
}

// ---------------------------------------------------------------- pipe micro-benchmarks
template <int MODE>
__global__ void __launch_bounds__(256) pipe_bench(float* out, int iters, float seed) {
  // 8 independent chains per thread
  float a[8]; float2 p[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = seed + k + threadIdx.x * 1e-3f; p[k] = make_float2(a[k], a[k] + 0.5f); }
  float c1 = 0.999f, c2 = 1e-3f; float2 c1v = make_float2(c1, c1), c2v = make_float2(c2, c2);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) { a[k] = fmaf(a[k], c1, c2); }                                  // FFMA
      if (MODE == 1) { p[k] = __ffma2_rn(p[k], c1v, c2v); }                           // FFMA2
      if (MODE == 2) { a[k] = ex2(a[k]); }                                            // MUFU.EX2
      if (MODE == 3) { p[k] = __ffma2_rn(p[k], c1v, c2v); p[k] = __ffma2_rn(p[k], c1v, c2v);
                       p[k] = __ffma2_rn(p[k], c1v, c2v); p[k] = __ffma2_rn(p[k], c1v, c2v);
                       p[k].x = ex2(p[k].x); p[k].y = ex2(p[k].y); }                  // 4 FFMA2 + 2 MUFU (gauss mix 8:1)
      if (MODE == 4) { a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2);
                       a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2); a[k] = fmaf(a[k], c1, c2);
                       a[k] = ex2(a[k]); }                                            // 8 FFMA + 1 MUFU
      if (MODE == 5) { p[k] = __ffma2_rn(p[k], p[(k + 1) & 7], p[(k + 2) & 7]); }     // FFMA2 3 distinct regs
      if (MODE == 6) { a[k] = fmaf(a[k], a[(k + 1) & 7], a[(k + 2) & 7]); }           // FFMA 3 distinct regs
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k] + p[k].x + p[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float timeit(void (*launch)(void*), void* ctx, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(ctx); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(ctx); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

struct Ctx { long nx, ny; float *x, *y, *b, *out; float scale; int T; };

template <int R, int TJ, int UNR> void l_v1(void* c_) { Ctx* c = (Ctx*)c_; long rows = (long)c->T * R; v1_scalar<R, TJ, UNR><<<(c->nx + rows - 1) / rows, c->T>>>(c->nx, c->ny, c->x, c->y, c->b, c->out, c->scale); }
template <int P, int TJ, int UNR> void l_v2(void* c_) { Ctx* c = (Ctx*)c_; long rows = (long)c->T * 2 * P; v2_packed<P, TJ, UNR><<<(c->nx + rows - 1) / rows, c->T>>>(c->nx, c->ny, c->x, c->y, c->b, c->out, c->scale); }
void l_v0(void* c_) { Ctx* c = (Ctx*)c_; v0_ref<<<(c->nx + 191) / 192, 192, 192 * 16>>>(c->nx, c->ny, c->x, c->y, c->b, c->out); }

int main(int argc, char** argv) {
  long M = argc > 1 ? atol(argv[1]) : 1000000, N = argc > 2 ? atol(argv[2]) : 1000000;
  int reps = argc > 3 ? atoi(argv[3]) : 3;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs=%d clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  std::vector<float> hx(M * 3), hy(N * 3), hb(N);
  srand(0);
  auto rnd = []() { // approx normal
    float s = 0; for (int k = 0; k < 12; k++) s += rand() / (float)RAND_MAX; return s - 6.f; };
  for (auto& v : hx) v = rnd(); for (auto& v : hy) v = rnd(); for (auto& v : hb) v = rnd();
  Ctx c; c.nx = M; c.ny = N; c.scale = sqrtf(1.4426950408889634f);
  CK(cudaMalloc(&c.x, M * 12)); CK(cudaMalloc(&c.y, N * 12)); CK(cudaMalloc(&c.b, N * 4)); CK(cudaMalloc(&c.out, M * 4));
  CK(cudaMemcpy(c.x, hx.data(), M * 12, cudaMemcpyHostToDevice)); CK(cudaMemcpy(c.y, hy.data(), N * 12, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.b, hb.data(), N * 4, cudaMemcpyHostToDevice));
  std::vector<float> ref(M), got(M);
  double pairs = (double)M * N;

  // ---- pipe microbenchmarks
  {
    float* o; CK(cudaMalloc(&o, 148 * 8 * 256 * 4));
    int iters = 4096; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"FFMA(imm)", "FFMA2(imm)", "MUFU.EX2", "4FFMA2+2MUFU", "8FFMA+1MUFU", "FFMA2 3reg", "FFMA 3reg"};
    double ops_per_it[] = {8, 16, 8, 8 * (8 + 2), 8 * 9, 16, 8};
    for (int mode = 0; mode < 7; mode++) {
      for (int blocks_per_sm = 4; blocks_per_sm <= 8; blocks_per_sm += 4) {
        int grid = 148 * blocks_per_sm; float best = 1e30f;
        for (int r = 0; r < 3; r++) {
          cudaEventRecord(e0);
          switch (mode) {
            case 0: pipe_bench<0><<<grid, 256>>>(o, iters, 1.f); break; case 1: pipe_bench<1><<<grid, 256>>>(o, iters, 1.f); break;
            case 2: pipe_bench<2><<<grid, 256>>>(o, iters, 1.f); break; case 3: pipe_bench<3><<<grid, 256>>>(o, iters, 1.f); break;
            case 4: pipe_bench<4><<<grid, 256>>>(o, iters, 1.f); break; case 5: pipe_bench<5><<<grid, 256>>>(o, iters, 1.f); break;
            case 6: pipe_bench<6><<<grid, 256>>>(o, iters, 1.f); break;
          }
          cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double lane_ops = ops_per_it[mode] * iters * (double)grid * 256;
        printf("PIPE %-14s blocks/SM=%d  %.3f ms  %.1f lane-ops/clk/SM @1.965GHz (%.3e ops/s)\n", names[mode], blocks_per_sm, best,
               lane_ops / (best * 1e-3) / 148 / 1.965e9, lane_ops / (best * 1e-3));
      }
    }
    CK(cudaGetLastError());
  }

  float ms = timeit(l_v0, &c, reps);
  CK(cudaMemcpy(ref.data(), c.out, M * 4, cudaMemcpyDeviceToHost));
  printf("V0 ref-like 192thr            : %8.3f ms  %.3e pairs/s\n", ms, pairs / (ms * 1e-3));

  auto check = [&](const char* name, float ms) {
    CK(cudaGetLastError());
    CK(cudaMemcpy(got.data(), c.out, M * 4, cudaMemcpyDeviceToHost));
    double maxrel = 0; for (long i = 0; i < M; i++) { double e = fabs(got[i] - ref[i]) / (fabs(ref[i]) + 1e-3); if (e > maxrel) maxrel = e; }
    printf("%-30s: %8.3f ms  %.3e pairs/s  frac(4.65e12)=%.3f  maxrel=%.2e\n", name, ms, pairs / (ms * 1e-3), pairs / (ms * 1e-3) / 4.65e12, maxrel);
    CK(cudaMemset(c.out, 0, M * 4));
  };
  int Ts[] = {128, 256};
  for (int T : Ts) {
    c.T = T; char nm[64];
    snprintf(nm, 64, "V1 scalar R=1 T=%d U4", T); check(nm, timeit(l_v1<1, 256, 4>, &c, reps));
    snprintf(nm, 64, "V1 scalar R=2 T=%d U4", T); check(nm, timeit(l_v1<2, 256, 4>, &c, reps));
    snprintf(nm, 64, "V1 scalar R=4 T=%d U4", T); check(nm, timeit(l_v1<4, 256, 4>, &c, reps));
    snprintf(nm, 64, "V1 scalar R=4 T=%d U8", T); check(nm, timeit(l_v1<4, 256, 8>, &c, reps));
    snprintf(nm, 64, "V1 scalar R=8 T=%d U4", T); check(nm, timeit(l_v1<8, 256, 4>, &c, reps));
    snprintf(nm, 64, "V2 packed R=2 T=%d U4", T); check(nm, timeit(l_v2<1, 256, 4>, &c, reps));
    snprintf(nm, 64, "V2 packed R=2 T=%d U8", T); check(nm, timeit(l_v2<1, 256, 8>, &c, reps));
    snprintf(nm, 64, "V2 packed R=4 T=%d U4", T); check(nm, timeit(l_v2<2, 256, 4>, &c, reps));
    snprintf(nm, 64, "V2 packed R=4 T=%d U8", T); check(nm, timeit(l_v2<2, 256, 8>, &c, reps));
    snprintf(nm, 64, "V2 packed R=8 T=%d U4", T); check(nm, timeit(l_v2<4, 256, 4>, &c, reps));
    snprintf(nm, 64, "V2 packed R=8 T=%d U2", T); check(nm, timeit(l_v2<4, 256, 2>, &c, reps));
  }
  return 0;
}
