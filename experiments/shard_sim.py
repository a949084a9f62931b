"""Predicts the 1/2/4/8-GPU strong-scaling of bench.py on ONE GPU: times the per-rank shard
(rows [0, M/world) of the output index against the full streamed operand) for the forward Gaussian
conv and the K=10 KNN, and prints the per-rank kernel time, the launch shape and the implied speed-up.
The NCCL broadcast of y, b (16 MB) is not included (measured separately in bench.py at N>1)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
KNN3 = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    M = N = 1000000
    x, y, b = torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev), torch.randn(N, 1, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, formula, nargs, targs in (("gauss_fwd", G, 3, (y, b)), ("knn_d3", KNN3, 2, (y,))):
        bd = get_binder(formula, [], nargs, "float32")
        base = None
        for world in (1, 2, 4, 8):
            Ms = M // world
            xs = x[:Ms].contiguous()
            out = bd.genred_pytorch(dev, None, Ms, N, 0, None, xs, *targs)
            torch.cuda.synchronize()
            best = None
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                bd.genred_pytorch(dev, None, Ms, N, 0, out, xs, *targs)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            base = best if base is None else base
            print(json.dumps({"case": name, "world": world, "rows_per_rank": Ms, "ms": best, "speedup_vs_1": base / best,
                              "launch": {k: v for k, v in bd.last_launch.items()}}, default=str), flush=True)


if __name__ == "__main__":
    main()
