"""BASELINE configs[4]: KernelSolve (conjugate gradient) with a Gaussian kernel, N = 10^6, D = 16, fp32, at full size.
Prints one JSON line per engine: iterations, time per iteration, true relative residual.

  python experiments/ksolve_full.py [N] [--cuda-core]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.torch import Genred, KernelSolve
from keops_b200.torch import kernel_solve as ks_mod


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000000
    engines = ["tcgen05"] + (["cuda-core"] if "--cuda-core" in sys.argv else [])
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    D, g, alpha, eps = 16, 2.0, 1000.0, 1e-6
    x = torch.rand(n, D, device=dev)
    b = torch.randn(n, 1, device=dev)
    gam = torch.tensor([g], device=dev)
    aliases = ["x=Vi(%d)" % D, "y=Vj(%d)" % D, "b=Vj(1)", "g=Pm(1)"]
    formula = "Exp(-g*SqDist(x,y))*b"
    Kinv = KernelSolve(formula, aliases, "b", axis=1)
    K = Genred(formula, aliases, reduction_op="Sum", axis=1)
    for eng in engines:
        os.environ["KB200_CONV_TC"] = "1" if eng == "tcgen05" else "0"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = Kinv(x, x, b, gam, alpha=alpha, eps=eps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        iters = ks_mod.conjugate_gradient.last_iterations
        os.environ["KB200_CONV_TC"] = "0"              # true residual with the CUDA-core (reference-order) matvec
        res = K(x, x, a, gam) + alpha * a - b
        rel = float(res.norm() / b.norm())
        print(json.dumps({"config": "KernelSolve Gaussian N=%d D=%d g=%g alpha=%g eps=%g fp32" % (n, D, g, alpha, eps),
                          "engine": eng, "iterations": iters, "total_s": dt, "ms_per_iteration": dt / max(iters, 1) * 1e3,
                          "pair_evals_per_s": iters * float(n) * n / dt, "rel_residual": rel}), flush=True)
    os.environ.pop("KB200_CONV_TC", None)


if __name__ == "__main__":
    main()
