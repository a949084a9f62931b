"""Micro-benchmark: sustained MUFU.EX2 throughput of a B200 SM (ex2 / clk / SM), alone and in the
FFMA -> EX2 -> FFMA pattern of the Gaussian kernels.  The SURVEY 8(d) roofline assumes 16 / clk / SM.
Run on the GPU box:  python experiments/mufu_peak.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.runtime import capi, plan

SRC = r"""
__device__ __forceinline__ float ex2(float s) { float p; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(s)); return p; }
extern "C" __global__ void __launch_bounds__(512) k_pure(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + 0.001f * (threadIdx.x + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = ex2(a[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 123.456f) out[0] = s;
}
extern "C" __global__ void __launch_bounds__(512) k_mix(float* out, int iters, float c, float r, float b) {
  float u[8], acc[2] = {0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; k++) u[k] = -0.001f * (threadIdx.x + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float p = ex2(fmaf(u[k], c, r));
      acc[k & 1] = fmaf(p, b, acc[k & 1]);
      u[k] += 1e-7f;          // keeps the chain alive (one more FP32 op per pair)
    }
  }
  if (acc[0] + acc[1] == 123.456f) out[0] = acc[0];
}
"""


def main():
    dev = torch.device("cuda:0")
    cubin = capi.compile_to_cubin(SRC, plan.nvrtc_options(True), name="mufu_peak.cu")
    mod = capi.Module(cubin, 0)
    out = torch.zeros(4, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    iters = 20000
    for name in ("k_pure", "k_mix"):
        fn = mod.function(name)
        for ctas_per_sm, threads in ((1, 256), (1, 512), (2, 512), (4, 512)):
            if name == "k_pure":
                params = [C.c_void_p(out.data_ptr()), C.c_int(iters), C.c_float(-0.5)]
            else:
                params = [C.c_void_p(out.data_ptr()), C.c_int(iters), C.c_float(1.0), C.c_float(-1.0), C.c_float(0.5)]
            grid = 148 * ctas_per_sm
            capi.launch(fn, stream, (grid,), (threads,), 0, params)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.launch(fn, stream, (grid,), (threads,), 0, params)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            n = grid * threads * iters * 8.0
            print("%s  %d CTA/SM x %d thr: %.3f ms  %.2f ex2/clk/SM (at 1965 MHz)" % (
                name, ctas_per_sm, threads, ms, n / (ms * 1e-3) / 148 / 1.965e9), flush=True)


if __name__ == "__main__":
    main()
