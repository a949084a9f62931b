"""Attention-like sum with a wide streamed factor (dim b = 300, the reference's final-chunk test shape at scale):
chunked evaluation vs the single kernel whose accumulators spill.   python experiments/wide_b_check.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dev = torch.device("cuda:0")
for DV in (128, 300):
    x, y, b = torch.rand(n, 3, device=dev), torch.rand(n, 3, device=dev), torch.randn(n, DV, device=dev)
    bd = get_binder("Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,%d,1),0)" % DV, [], 3, "float32")
    for mode in ("1", "0"):
        os.environ["KB200_WIDE_CHUNKS"] = mode
        o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("dim b = %d n = %d chunks=%s: %.1f ms  %.3e pair-components/s (%s)" % (
            DV, n, mode, ms, float(n) * n * DV / ms * 1e3, bd.last_launch.get("engine")), flush=True)
os.environ.pop("KB200_WIDE_CHUNKS", None)
