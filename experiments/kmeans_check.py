"""Nearest-centroid assignment (the reduction of a K-means step, ``pykeops/pykeops/tutorials/kmeans``: ``D_ij.argmin(dim=1)``)
through the filtered kernel vs the generic one: M points against N centroids in dimension D."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
x, c = torch.randn(M, D, device=dev), torch.randn(N, D, device=dev)
bd = get_binder("ArgMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),0)" % (D, D), [], 2, "float32")


def run(filtered):
    os.environ["KB200_KMIN_FILTER"] = "1" if filtered else "0"
    o = bd.genred_pytorch(dev, None, M, N, 0, None, x, c)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = bd.genred_pytorch(dev, None, M, N, 0, None, x, c); e1.record(); torch.cuda.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    return o, best, dict(bd.last_launch)


o_f, t_f, ll = run(True)
o_g, t_g, ll2 = run(False)
print("M=%d N=%d D=%d: filtered %.2f ms (%s, %.3e pair/s)   generic %.2f ms (%.3e pair/s)   x%.2f   identical: %s" % (
    M, N, D, t_f, ll.get("engine"), M * N / t_f * 1e3, t_g, M * N / t_g * 1e3, t_g / t_f, torch.equal(o_f, o_g)))
