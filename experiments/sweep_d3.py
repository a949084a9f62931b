"""D=3 Gaussian forward / grad_x / KNN tile-configuration sweep at full size (GPU box).  python experiments/sweep_d3.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.formulas.parser import parse_reduction
from keops_b200.codegen.emit import generate, KernelConfig
from keops_b200.runtime import plan, capi

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
GX = "Grad_WithSavedForward(%s,Var(0,3,0),Var(3,1,0),Var(4,1,0))" % G
KNN = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"
dev = torch.device("cuda:0")
torch.manual_seed(0)
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(formula, nargs, cfg, args, reps=3):
    red = parse_reduction(formula)
    ks = generate(red, "float32", nargs, cfg)
    lk = plan.load_kernel(ks, True, 0)
    rpb = cfg.T * cfg.R
    ns, chunk = plan.launch_shape(M, N, rpb, cfg.TJ, lk.sm_count * lk.blocks_per_sm, ks.dimpart, 4)
    out = torch.empty(M, ks.dim, device=dev)
    ws = torch.empty(max(1, ns * M * ks.dimpart), device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    run = lambda: capi.genred_launch(lk.fn_reduce, lk.fn_combine, stream, M, N, rpb, cfg.T, ns, chunk, out.data_ptr(), ws.data_ptr(), [a.data_ptr() for a in args])
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, ns, lk.regs, lk.blocks_per_sm, out


x, y, b, eta = torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev), torch.randn(N, 1, device=dev), torch.randn(M, 1, device=dev)
CFGS = [(256, 4, 512, 8, 4), (256, 8, 512, 4, 2), (256, 8, 512, 8, 2), (128, 8, 512, 4, 4), (128, 8, 512, 8, 4), (128, 8, 512, 4, 3),
        (256, 8, 1024, 4, 2), (128, 4, 512, 8, 8), (192, 4, 512, 8, 5), (256, 4, 512, 8, 3), (256, 6, 512, 4, 2)]
fwd = None
for name, f, nargs, extra in (("fwd", G, 3, ()), ("grad_x", GX, 5, ("eta", "fwd")), ("knn", KNN, 2, ())):
    for T, R, TJ, U, B in CFGS:
        cfg = KernelConfig(T=T, R=R, TJ=TJ, unroll=U, min_blocks=B, group=(4 if name == "knn" and R == 4 else (2 if name == "knn" else 1)))
        args = [x, y] if name == "knn" else ([x, y, b] + ([eta, fwd] if extra else []))
        try:
            ms, ns, regs, bps, out = timeit(f, nargs, cfg, args)
            print("%-6s T=%d R=%d TJ=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms" % (name, T, R, TJ, U, B, regs, bps, ns, ms), flush=True)
            if name == "fwd" and fwd is None:
                fwd = out
        except Exception as e:
            print("%-6s T=%d R=%d TJ=%d U=%d B=%d failed: %s" % (name, T, R, TJ, U, B, str(e)[:160]), flush=True)
