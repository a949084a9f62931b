import os, sys
sys.path.insert(0, "/root/repo")
import torch
from keops_b200.binder import get_binder, LoadKeOpsB200
n = 50000
dev = torch.device("cuda:0")
LoadKeOpsB200.WIDE_MIN = 33
for C in (16, 32, 64):
    LoadKeOpsB200.WIDE_CHUNK = C
    for DV in (40, 64, 96, 300):
        x, y, b = torch.rand(n, 3, device=dev), torch.rand(n, 3, device=dev), torch.randn(n, DV, device=dev)
        bd = get_binder("Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,%d,1),0)" % DV, [], 3, "float32")
        bd._wide_binders.clear()
        for mode in ("1", "0"):
            if mode == "0" and C != 16: continue
            os.environ["KB200_WIDE_CHUNKS"] = mode
            o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y, b); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = bd.genred_pytorch(dev, None, n, n, 0, None, x, y, b); e1.record(); torch.cuda.synchronize()
            print("chunk %d dim b = %d chunks=%s: %.1f ms" % (C, DV, mode, e0.elapsed_time(e1)), flush=True)
