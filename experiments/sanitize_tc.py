"""Small runs of every hand-written tensor-core kernel, meant to be launched under compute-sanitizer:
  compute-sanitizer --tool memcheck python experiments/sanitize_tc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder

dev = torch.device("cuda:0")
torch.manual_seed(0)


def conv(D, DB, M, N, lse=False):
    x, y = torch.rand(M, D, device=dev), torch.rand(N, D, device=dev)
    g = torch.tensor([8.0 / D], device=dev)
    if lse:
        a = torch.rand(N, 1, device=dev)
        f = "Max_SumShiftExp_Reduction((Var(2,1,1) - Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1))),0)" % (D, D)
        args = (x, y, a, g)
    else:
        b = torch.randn(N, DB, device=dev)
        f = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,%d,1),0)" % (D, D, DB)
        args = (x, y, b, g)
    bd = get_binder(f, [], 4, "float32")
    o = bd.genred_pytorch(dev, None, M, N, 0, None, *args)
    torch.cuda.synchronize()
    print("conv D=%d DB=%d lse=%s %s nchunks=%s finite=%s" % (D, DB, lse, bd.last_launch.get("engine"),
                                                            bd.last_launch.get("nchunks"), bool(torch.isfinite(o).all())), flush=True)


def knn(D, K, M, N):
    x, y = torch.randn(M, D, device=dev), torch.randn(N, D, device=dev)
    bd = get_binder("ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,0)" % (D, D, K), [], 2, "float32")
    o = bd.genred_pytorch(dev, None, M, N, 0, None, x, y)
    torch.cuda.synchronize()
    print("knn D=%d %s nsplit=%s" % (D, bd.last_launch.get("engine"), bd.last_launch.get("nsplit")), flush=True)


conv(16, 1, 8192 + 37, 1024 + 5)          # 256-row kernel, TMEM operand, ragged rows and columns
conv(13, 2, 8192, 1100)                   # 256-row kernel, smem operand, remainder k-block
conv(21, 1, 8192, 1100)                   # penalty in the epilogue
conv(64, 2, 8192 + 5, 1100)               # high-dimensional kernel, TMEM operand
conv(95, 1, 8192, 1100)                   # high-dimensional kernel, smem operand
conv(128, 1, 8192, 1100)                  # both operands streamed
conv(16, 1, 1100, 66000)                  # few rows: column chunks + combine
conv(16, 1, 8192, 1100, lse=True)
conv(64, 1, 1100, 66000, lse=True)        # LSE chunks + combine
os.environ["KB200_CONV_TC_MC"] = "1"
conv(64, 1, 8192 + 200, 1100)             # cluster multicast variant (odd number of row blocks)
os.environ.pop("KB200_CONV_TC_MC")
knn(64, 10, 8192 + 9, 4100)
knn(32, 3, 1100, 66000)                   # column splits + merge
print("done")
