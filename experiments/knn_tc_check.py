"""GPU check of the tensor-core K-min path against the CUDA-core kernel (same exact arithmetic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder
from keops_b200.codegen import knn_tc

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
kind = sys.argv[5] if len(sys.argv) > 5 else "ArgKMin"
F = "%s_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,0)" % (kind, D, D, K)
torch.manual_seed(0)
x, y = torch.randn(M, D, device=dev), torch.randn(N, D, device=dev)
bd = get_binder(F, [], 2, "float32")

def run(tc):
    os.environ["KB200_KNN_TC"] = "1" if tc else "0"
    o = bd.genred_pytorch(dev, None, M, N, 0, None, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = bd.genred_pytorch(dev, None, M, N, 0, None, x, y); e1.record(); torch.cuda.synchronize()
    return o, e0.elapsed_time(e1), dict(bd.last_launch)

o_tc, ms_tc, ll = run(True)
print("tcgen05 path: %.3f ms  %s" % (ms_tc, ll))
if os.environ.get("KT_STATS"):
    eng = knn_tc.get(D, K, kind, True, 0)
    print("candidate stats:", eng.stats)
if M * N <= 4e10 or os.environ.get("KT_FULLCHECK"):
    o_cc, ms_cc, ll2 = run(False)
    print("cuda-core path: %.3f ms  %s" % (ms_cc, ll2))
    same = torch.equal(o_tc, o_cc)
    nbad = int((o_tc != o_cc).any(dim=1).sum())
    print("identical:", same, " rows differing:", nbad, "of", M)
    if nbad:
        r = int((o_tc != o_cc).any(dim=1).nonzero()[0])
        print("first bad row", r, "\n tc:", o_tc[r].tolist(), "\n cc:", o_cc[r].tolist())
    print("speedup %.2fx   queries/s %.3e" % (ms_cc / ms_tc, M / ms_tc * 1e3))
else:
    # sample check: exact top-K by torch on 2048 random rows
    idx = torch.randperm(M, device=dev)[:2048]
    d2 = torch.cdist(x[idx].double(), y.double()) ** 2
    ref = d2.topk(K, largest=False).indices
    got = o_tc[idx].long()
    agree = (ref.sort(1).values == got.sort(1).values).all(1).float().mean().item()
    print("sample rows agreeing with fp64 top-K sets: %.4f   queries/s %.3e" % (agree, M / ms_tc * 1e3))
