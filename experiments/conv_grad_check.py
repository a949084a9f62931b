"""Full-size timing of the Gaussian-sum position gradients (the reductions KernelSolve's backward launches): tensor-core
assembly (keops_b200/binder.py::_genred_conv_tc_grad) vs the CUDA-core gradient kernel, M = N = 1e6 by default."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from keops_b200.binder import get_binder

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
x, y = torch.rand(n, D, device=dev), torch.rand(n, D, device=dev)
b, eta = torch.randn(n, 1, device=dev), torch.randn(n, 1, device=dev)
g = torch.tensor([1.0 if D <= 21 else 16.0 / D], device=dev)
F = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,1,1),0)" % (D, D)


def timed(bd, nargs_args, reps=3):
    out = bd.genred_pytorch(dev, None, n, n, 0, None, *nargs_args)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = bd.genred_pytorch(dev, None, n, n, 0, None, *nargs_args)
        e1.record()
        torch.cuda.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    return out, best, dict(bd.last_launch)


fwd, t_fwd, ll = timed(get_binder(F, [], 4, "float32"), (x, y, b, g))
print("forward: %.1f ms (%s)" % (t_fwd, ll.get("engine")))
for var, v in (("x", "Var(0,%d,0)" % D), ("y", "Var(1,%d,1)" % D), ("b", "Var(2,1,1)")):
    f = "Grad_WithSavedForward(%s,%s,Var(4,1,0),Var(5,1,0))" % (F, v)
    bd = get_binder(f, [], 6, "float32")
    args = (x, y, b, g, eta, fwd)
    os.environ["KB200_CONV_TC_GRAD"] = "1"
    o_tc, t_tc, ll = timed(bd, args)
    line = "d/d%s: tensor cores %.1f ms (%s, %s weight columns, bound %s)" % (var, t_tc, ll.get("engine"), ll.get("weight_columns"), ll.get("error_bound"))
    if var != "b" and os.environ.get("COMPARE", "1") == "1":
        os.environ["KB200_CONV_TC_GRAD"] = "0"
        o_cc, t_cc, ll2 = timed(bd, args, reps=1)
        den = o_cc.abs().max()
        line += "   CUDA cores %.1f ms (%s)  x%.2f   max |tc-cc| / max|cc| %.2e" % (t_cc, ll2.get("engine"), t_cc / t_tc, float((o_tc - o_cc).abs().max() / den))
    print(line, flush=True)
