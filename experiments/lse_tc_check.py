"""LogSumExp soft-min (Sinkhorn) in dimension D on the tensor-core path vs the CUDA-core kernel.
  python experiments/lse_tc_check.py N D [D ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder


def timed(bd, dev, n, args, tc, reps):
    os.environ["KB200_CONV_TC"] = "1" if tc else "0"
    out = bd.genred_pytorch(dev, None, n, n, 0, None, *args)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = bd.genred_pytorch(dev, None, n, n, 0, None, *args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return out, best, dict(bd.last_launch)


def main():
    n = int(sys.argv[1])
    dev = torch.device("cuda:0")
    for D in [int(a) for a in sys.argv[2:]]:
        torch.manual_seed(0)
        x, y, a = torch.rand(n, D, device=dev), torch.rand(n, D, device=dev), torch.rand(n, 1, device=dev)
        eps = torch.tensor([D / 24.0], device=dev)
        f = ("Max_SumShiftExp_Reduction((Var(2,1,1) - ((Sum(Square((Var(0,%d,0) - Var(1,%d,1)))) / IntCst(2)) / Var(3,1,2))),0)" % (D, D))
        bd = get_binder(f, [], 4, "float32")
        o_tc, t_tc, ll = timed(bd, dev, n, (x, y, a, eps), True, 3)
        o_cc, t_cc, _ = timed(bd, dev, n, (x, y, a, eps), False, 1)
        l_tc, l_cc = o_tc[:, 0] + o_tc[:, 1].log(), o_cc[:, 0] + o_cc[:, 1].log()
        print("LSE D=%d n=%d: %s %.2f ms (%.3e pair/s)  cuda-core %.2f ms  x%.2f  max |lse diff| / max |lse| %.2e  bound %s"
              % (D, n, ll.get("engine"), t_tc, n * n / t_tc * 1e3, t_cc, t_cc / t_tc,
                 float((l_tc - l_cc).abs().max() / l_cc.abs().max()), ll.get("error_bound")), flush=True)
    os.environ.pop("KB200_CONV_TC", None)


if __name__ == "__main__":
    main()
