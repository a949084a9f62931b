"""ArgKMin K=10 in dimension D <= 15: filtered kernel vs generic kernel, M = N points (median of 3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
for D in [int(a) for a in sys.argv[2:]] or [3, 5, 8, 12, 15]:
    torch.manual_seed(D)
    x, y = torch.randn(n, D, device=dev), torch.randn(n, D, device=dev)
    bd = get_binder("ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),10,0)" % (D, D), [], 2, "float32")
    res = {}
    for filt in ("1", "0"):
        os.environ["KB200_KMIN_FILTER"] = filt
        o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[filt] = (sorted(ts)[1], o, bd.last_launch.get("engine"))
    print("D=%d n=%d: filtered %.1f ms (%s, %.2e pair/s)  generic %.1f ms (%.2e pair/s)  x%.2f  identical %s" % (
        D, n, res["1"][0], res["1"][2], n * n / res["1"][0] * 1e3, res["0"][0], n * n / res["0"][0] * 1e3,
        res["0"][0] / res["1"][0], torch.equal(res["1"][1], res["0"][1])), flush=True)
