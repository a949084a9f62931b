"""Tensor-core Gaussian conv (codegen/conv_tc.py) vs the CUDA-core kernel: accuracy against an fp64 dense torch
evaluation (small size) and timings (full size).  Run on the GPU box:  python experiments/conv_tc_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder


def run(binder, x, y, b, g, tc, nrep=1):
    os.environ["KB200_CONV_TC"] = "1" if tc else "0"
    M, N = x.shape[0], y.shape[0]
    dev = x.device
    out = binder.genred_pytorch(dev, None, M, N, 0, None, x, y, b, g)
    torch.cuda.synchronize()
    ts = []
    for _ in range(nrep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = binder.genred_pytorch(dev, None, M, N, 0, None, x, y, b, g)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, min(ts), dict(binder.last_launch)


def main():
    dev = torch.device("cuda:0")
    sizes = [int(a) for a in sys.argv[1:]] or [20000, 1000000]
    cases = ((16, 1.0, 1.0), (16, 8.0, 1.0), (16, 1.0, 10.0), (3, 1.0, 1.0))
    if os.environ.get("CT_ONLY"):                     # profiling / timing experiments: the tensor-core launch only
        D = int(os.environ["CT_ONLY"])
        f = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,1,1),0)" % (D, D)
        os.environ["KB200_CONV_TC_DMIN"] = "3"
        binder = get_binder(f, [], 4, "float32")
        for n in sizes:
            x, y, b = torch.rand(n, D, device=dev), torch.rand(n, D, device=dev), torch.randn(n, 1, device=dev)
            gv = torch.tensor([1.0 if D <= 21 else 16.0 / D], device=dev)
            o, t, ll = run(binder, x, y, b, gv, True, nrep=3)
            line = "CT_ONLY D=%d n=%d DBG=%s STAGES=%s: %.3f ms  %.3e pair/s (%s, bound %s)" % (
                D, n, os.environ.get("CT_DBG", "0"), os.environ.get("CT_STAGES", "4"), t, n * n / t * 1e3,
                ll.get("engine"), ll.get("error_bound"))
            if os.environ.get("CT_COMPARE"):
                o2, t2, _ = run(binder, x, y, b, gv, False, nrep=1)
                line += "  cuda-core %.1f ms, max diff / max %.2e" % (t2, float((o - o2).abs().max() / o2.abs().max()))
            print(line, flush=True)
        return
    for D, gval, scale in cases:
        os.environ["KB200_CONV_TC_DMIN"] = "3"
        f = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,%d,0),Var(1,%d,1)))*Var(2,1,1),0)" % (D, D)
        binder = get_binder(f, [], 4, "float32")
        for n in sizes:
            torch.manual_seed(0)
            if D == 16:
                x = torch.rand(n, D, device=dev) * scale + (scale - 1.0)
                y = torch.rand(n, D, device=dev) * scale + (scale - 1.0)
            else:
                x, y = torch.randn(n, D, device=dev), torch.randn(n, D, device=dev)
            b = torch.randn(n, 1, device=dev)
            g = torch.tensor([gval / scale ** 2], device=dev)
            o_tc, t_tc, ll = run(binder, x, y, b, g, True, nrep=3)
            o_cc, t_cc, _ = run(binder, x, y, b, g, False, nrep=1 if n > 200000 else 3)
            line = "D=%d g=%.2f scale=%g n=%d  tc %.2f ms (%.3e pair/s, %s)  cuda-core %.2f ms (%.3e pair/s)" % (
                D, gval, scale, n, t_tc, n * n / t_tc * 1e3, ll.get("engine"), t_cc, n * n / t_cc * 1e3)
            den = o_cc.abs().max()
            line += "  |tc-cc|max/|cc|max %.2e" % float((o_tc - o_cc).abs().max() / den)
            if n <= 20000:
                xd, yd, bd = x.double(), y.double(), b.double()
                ref = torch.exp(-float(g.double()) * torch.cdist(xd, yd) ** 2) @ bd
                # honest scale for a sum with cancelling signs: sum_j |term_j|
                mag = torch.exp(-float(g.double()) * torch.cdist(xd, yd) ** 2) @ bd.abs()
                for nm, o in (("tc", o_tc), ("cc", o_cc)):
                    err = (o.double() - ref).abs()
                    line += "  %s: max/|ref|max %.2e  max(err/sum|terms|) %.2e" % (
                        nm, float(err.max() / ref.abs().max()), float((err / mag).max()))
            print(line, flush=True)


if __name__ == "__main__":
    main()
