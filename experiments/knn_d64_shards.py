"""ArgKMin K=10 D=64 on the row shards of an N-GPU run (rows = 1e6 / N against 1e6 columns), median of 7 timed calls with an
L2 flush in between: the effect of splitting the last partial wave of row blocks (KB200_KNN_TC_TAIL)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

dev = torch.device("cuda:0")
N, D = 1000000, 64
torch.manual_seed(0)
x, y = torch.randn(N, D, device=dev), torch.randn(N, D, device=dev)
bd = get_binder("ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),10,0)" % (D, D), [], 2, "float32")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
base = None
for world in (1, 2, 4, 8):
    M = N // world
    xs = x[:M].contiguous()
    res = {}
    for tail in ("1", "0"):
        os.environ["KB200_KNN_TC_TAIL"] = tail
        bd.genred_pytorch(dev, None, M, N, 0, None, xs, y)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bd.genred_pytorch(dev, None, M, N, 0, None, xs, y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[tail] = (ts[3], bd.last_launch.get("segments"))
    base = base or res["1"][0]
    print("world %d rows %d: tail split %.2f ms %s | unsplit %.2f ms | speed-up vs 1 GPU %.2fx" % (
        world, M, res["1"][0], res["1"][1], res["0"][0], base / res["1"][0]), flush=True)
