"""Times every BASELINE.json config at full size on one B200 and prints one JSON line per kernel.

  python experiments/configs_bench.py [--only NAME] [--once] [--size M]

``--once`` launches each selected kernel exactly twice (warm-up + one) -- the form used under ncu.
Rooflines are the FP32-pipe / MUFU issue bounds of SURVEY.md 8(d) at the max SM clock.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
GX = "Grad_WithSavedForward(%s,Var(0,3,0),Var(3,1,0),Var(4,1,0))" % G
GY = "Grad_WithSavedForward(%s,Var(1,3,1),Var(3,1,0),Var(4,1,0))" % G
GB = "Grad_WithSavedForward(%s,Var(2,1,1),Var(3,1,0),Var(4,1,0))" % G
KNN3 = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"
KNN64 = "ArgKMin_Reduction(Sum(Square((Var(0,64,0) - Var(1,64,1)))),10,0)"
LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"
KS16 = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)"
REF_KEY = {"gauss_fwd": "gauss_fwd_f32", "gauss_dx": "gauss_dx_f32", "gauss_dy": "gauss_dy_f32", "gauss_db": "gauss_db_f32",
           "knn_d3": "argkmin10_d3_f32", "knn_d64": "argkmin10_d64_f32", "lse_f32": "lse_f32", "lse_f64": "lse_f64",
           "ksolve_d16": "gauss16_pm_f32"}
SM, HZ = 148, 1.965e9
# pairs / clk / SM bounds
BOUND = {"gauss_fwd": 16.0, "gauss_dx": 128 / 11.0, "gauss_dy": 128 / 11.0, "gauss_db": 16.0, "knn_d3": 128 / 7.0,
         "knn_d64": 128 / 129.0, "lse_f32": 128 / 12.0, "lse_f64": 64 / 36.0, "ksolve_d16": 128 / 34.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--size", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ref", action="store_true", help="also time the reference's own CUDA kernel (oracle/_ref) on a row sample")
    ap.add_argument("--cpu", action="store_true", help="also time the reference's own CPU code (oracle/_ref, OpenMP, all host cores) on a bounded row sample")
    ap.add_argument("--rows", type=int, default=None, help="output rows M (default: --size); the shard of one GPU out of 8 is 125000")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    N = args.size
    M = args.rows or N
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gauss_inputs():
        x, y, b = torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev), torch.randn(N, 1, device=dev)
        eta = torch.randn(M, 1, device=dev)
        fwd = get_binder(G, [], 3, "float32").genred_pytorch(dev, None, M, N, 0, None, x, y, b)
        return x, y, b, eta, fwd

    cases = {
        "gauss_fwd": lambda: (G, 3, "float32", gauss_inputs()[:3]),
        "gauss_dx": lambda: (GX, 5, "float32", gauss_inputs()),
        "gauss_dy": lambda: (GY, 5, "float32", gauss_inputs()),
        "gauss_db": lambda: (GB, 5, "float32", gauss_inputs()),
        "knn_d3": lambda: (KNN3, 2, "float32", (torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev))),
        "knn_d64": lambda: (KNN64, 2, "float32", (torch.randn(M, 64, device=dev), torch.randn(N, 64, device=dev))),
        "lse_f32": lambda: (LSE, 4, "float32", (torch.rand(N, 1, device=dev), torch.rand(M, 3, device=dev),
                                                torch.rand(N, 3, device=dev), torch.tensor([0.05], device=dev))),
        "lse_f64": lambda: (LSE, 4, "float64", (torch.rand(N, 1, device=dev, dtype=torch.float64),
                                                torch.rand(M, 3, device=dev, dtype=torch.float64),
                                                torch.rand(N, 3, device=dev, dtype=torch.float64),
                                                torch.tensor([0.05], device=dev, dtype=torch.float64))),
        "ksolve_d16": lambda: (KS16, 4, "float32", (torch.rand(M, 16, device=dev), torch.rand(N, 16, device=dev),
                                                    torch.randn(N, 1, device=dev), torch.tensor([2.0], device=dev))),
    }
    names = [args.only] if args.only else list(cases)
    for name in names:
        formula, nargs, dtype, targs = cases[name]()
        bd = get_binder(formula, [], nargs, dtype)
        out = bd.genred_pytorch(dev, None, M, N, 0, None, *targs)
        torch.cuda.synchronize()
        best = None
        for _ in range(1 if args.once else args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bd.genred_pytorch(dev, None, M, N, 0, out, *targs)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        rate = float(M) * N / (best * 1e-3)
        peak = BOUND[name] * SM * HZ
        engine = bd.last_launch.get("engine") or "cuda-core"
        extra = {}
        if engine == "tcgen05-conv":
            # the squared distance runs on the tensor cores: the CUDA-core bound becomes 1 MUFU.EX2 per pair (16 / clk / SM)
            extra = {"cuda_core_roofline_pairs_per_s": peak, "frac_of_cuda_core_roofline": rate / peak}
            peak = 16.0 * SM * HZ
        if args.ref:
            # the reference's generated GpuConv1DOnDevice for the same formula, same GPU, same (full) size
            from oracle import ref_runner
            key = REF_KEY.get(name)
            if key and ref_runner.available():
                tagI = ref_runner.manifest()[key]["tagI"]
                rows = M if tagI == 0 else N
                rc = ref_runner.RefCuda(key, 0)
                var_cat = {v.params[0]: v.params[2] for v in bd.red.formula.variables()} if hasattr(bd, "red") else {}
                rargs = [a[:rows].contiguous() if var_cat.get(k) == tagI else a for k, a in enumerate(targs)]
                nx, ny = (rows, N) if tagI == 0 else (M, rows)
                ro = rc(rargs, nx, ny)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.zero_()
                e0.record()
                rc(rargs, nx, ny, out=ro)
                e1.record()
                torch.cuda.synchronize()
                rms = e0.elapsed_time(e1)
                rrate = float(nx) * ny / (rms * 1e-3)
                extra.update({"ref_cuda_pair_evals_per_s": rrate, "ref_cuda_rows": rows, "ref_cuda_ms": rms,
                              "speedup_vs_ref_cuda": rate / rrate})
        if args.cpu:
            # the reference's generated CpuReduc code (g++ -O3 -fopenmp) on the host cores of this box, on a bounded sample
            # of output rows against all columns (~2 s of CPU work per case; 1e12 pairs would take minutes to hours)
            import time
            from oracle import ref_runner
            key = REF_KEY.get(name)
            if key and ref_runner.available():
                tagI = ref_runner.manifest()[key]["tagI"]
                var_cat = {v.params[0]: v.params[2] for v in bd.red.formula.variables()}
                host = [a.cpu().numpy() for a in targs]

                def cpu_run(rows):
                    cargs = [a[:rows] if var_cat.get(k) == tagI else a for k, a in enumerate(host)]
                    nx, ny = (rows, N) if tagI == 0 else (M, rows)
                    t0 = time.perf_counter()
                    ref_runner.run_cpu(key, cargs, nx, ny)
                    return time.perf_counter() - t0, float(nx) * ny
                cores = os.cpu_count() or 1
                cpu_run(cores)                                  # warm-up (thread pool)
                dt, pairs = cpu_run(8 * cores)
                rows = int(max(8 * cores, min(M if tagI == 0 else N, 2.0 * (8 * cores) / max(dt, 1e-6)) // cores * cores))
                dt, pairs = cpu_run(rows)
                extra.update({"ref_cpu_pair_evals_per_s": pairs / dt, "ref_cpu_rows": rows, "ref_cpu_cores": cores,
                              "speedup_vs_ref_cpu": rate / (pairs / dt)})
        print(json.dumps({"case": name, "M": M, "N": N, "dtype": dtype, "ms": best, "pair_evals_per_s": rate,
                          "engine": engine, **extra,
                          "roofline_pairs_per_s": peak, "frac": rate / peak, "cfg": bd.ks.cfg.key(),
                          "nsplit": bd.last_launch["nsplit"], "regs": bd.last_launch["regs"],
                          "blocks_per_sm": bd.last_launch["blocks_per_sm"]}), flush=True)
        del targs, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
