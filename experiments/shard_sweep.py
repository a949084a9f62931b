"""One-GPU prediction of the 8-GPU strong-scaling shard (rows [0, M/8) against the full streamed operand) for every
bench leg, and a sweep of the number of column splits for the CUDA-core legs (KB200_FORCE_NSPLIT) next to the planner's
own choice: shows how much of the lost efficiency at N = 8 is the launch shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
GX = "Grad_WithSavedForward(" + G + ",Var(0,3,0),Var(3,1,0),Var(4,1,0))"
KNN3 = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"
LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"


def timeit(fn, flush, reps=5):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    M = N = 1000000
    worlds = [int(w) for w in os.environ.get("WORLDS", "1,8").split(",")]
    x, y, b = torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev), torch.randn(N, 1, device=dev)
    eta, fwd = torch.randn(M, 1, device=dev), torch.randn(M, 1, device=dev)
    f = torch.rand(N, 1, device=dev)
    eps = torch.tensor([0.05], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cases = [
        ("gauss_fwd", G, 3, lambda Ms: (x[:Ms].contiguous(), y, b)),
        ("gauss_dx", GX, 5, lambda Ms: (x[:Ms].contiguous(), y, b, eta[:Ms].contiguous(), fwd[:Ms].contiguous())),
        ("knn_d3", KNN3, 2, lambda Ms: (x[:Ms].contiguous(), y)),
        ("lse", LSE, 4, lambda Ms: (f, x[:Ms].contiguous(), y, eps)),
    ]
    only = os.environ.get("ONLY")
    for name, formula, nargs, mk in cases:
        if only and name not in only.split(","):
            continue
        bd = get_binder(formula, [], nargs, "float32")
        base = None
        for world in worlds:
            Ms = M // world
            args = mk(Ms)
            os.environ.pop("KB200_FORCE_NSPLIT", None)
            ms = timeit(lambda: bd.genred_pytorch(dev, None, Ms, N, 0, None, *args), flush)
            base = ms if base is None else base
            ll = dict(bd.last_launch)
            rec = {"case": name, "world": world, "ms": round(ms, 3), "speedup_vs_1": round(base / ms, 3), "nsplit": ll.get("nsplit"),
                   "grid_x": ll.get("grid_x"), "blocks_per_sm": ll.get("blocks_per_sm"), "engine": ll.get("engine"), "segments": ll.get("segments")}
            print(json.dumps(rec, default=str), flush=True)
            if (world > 1 or os.environ.get("SWEEP_ALL") == "1") and os.environ.get("SWEEP", "1") == "1":
                sweep = {}
                for ns in [int(v) for v in os.environ.get("NS", "4,6,8,10,12,14,17,19,22,24,29,34,39,48").split(",")]:
                    os.environ["KB200_FORCE_NSPLIT"] = str(ns)
                    os.environ["KB200_TAILSPLIT"] = "0"
                    sweep[ns] = round(timeit(lambda: bd.genred_pytorch(dev, None, Ms, N, 0, None, *args), flush, reps=3), 3)
                os.environ.pop("KB200_FORCE_NSPLIT", None)
                os.environ.pop("KB200_TAILSPLIT", None)
                print(json.dumps({"case": name, "world": world, "forced_nsplit_ms": sweep}), flush=True)


if __name__ == "__main__":
    main()
