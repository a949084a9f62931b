"""Ad-hoc timing of generated kernels under different KernelConfig / split choices (GPU box)."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.formulas.parser import parse_reduction
from keops_b200.codegen.emit import generate, KernelConfig
from keops_b200.runtime import plan, capi

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
KNN = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"
KNN64 = "ArgKMin_Reduction(Sum(Square((Var(0,64,0) - Var(1,64,1)))),10,0)"
LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"
GX = "Grad_WithSavedForward(%s,Var(0,3,0),Var(3,1,0),Var(4,1,0))" % G
GY = "Grad_WithSavedForward(%s,Var(1,3,1),Var(3,1,0),Var(4,1,0))" % G
dev = torch.device("cuda:0")
torch.manual_seed(0)
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000

def timeit(formula, nargs, dtype, cfg, args, nrow, ncol, nsplit=None, reps=3):
    red = parse_reduction(formula)
    ks = generate(red, dtype, nargs, cfg)
    fast = dtype == "float32"
    lk = plan.load_kernel(ks, fast, 0)
    rpb = cfg.T * cfg.R
    esz = 4 if fast else 8
    ns, chunk = plan.launch_shape(nrow, ncol, rpb, cfg.TJ, lk.sm_count * lk.blocks_per_sm, ks.dimpart, esz)
    if nsplit is not None:
        ns = nsplit; ntiles = -(-ncol // cfg.TJ); chunk = -(-ntiles // ns) * cfg.TJ
    tdt = torch.float32 if fast else torch.float64
    out = torch.empty(nrow, ks.dim, dtype=tdt, device=dev)
    ws = torch.empty(max(1, ns * nrow * ks.dimpart), dtype=tdt, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    def run():
        capi.genred_launch(lk.fn_reduce, lk.fn_combine, stream, nrow, ncol, rpb, cfg.T, ns, chunk, out.data_ptr(), ws.data_ptr(), [a.data_ptr() for a in args])
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, ns, lk.regs, lk.blocks_per_sm, out

x = torch.randn(M, 3, device=dev); y = torch.randn(N, 3, device=dev); b = torch.randn(N, 1, device=dev)
eta = torch.randn(M, 1, device=dev)
print("== Gaussian conv M=N=%d" % M)
for T, R, U, B in [(256,4,8,4),(256,4,8,3),(256,4,4,4),(128,4,8,8),(128,4,8,6),(256,2,8,4),(128,2,8,8),(256,4,8,2),(192,4,8,5),(256,8,4,2)]:
    for nsplit in (1, None):
        cfg = KernelConfig(T=T, R=R, TJ=256, unroll=U, min_blocks=B)
        try:
            ms, ns, regs, bps, out = timeit(G, 3, "float32", cfg, [x, y, b], M, N, nsplit)
            print("T=%d R=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s (%.1f%% of 4.65e12)" % (T,R,U,B,regs,bps,ns,ms,M*N/ms*1e3,M*N/ms*1e3/4.65e12*100))
        except Exception as e:
            print("T=%d R=%d U=%d B=%d failed: %s" % (T,R,U,B,str(e)[:200]))
fwd = out
print("== grads")
for name, f in (("dx", GX), ("dy", GY)):
    for T,R,U,B in [(256,4,8,4),(256,4,4,3),(256,2,8,4),(128,4,4,4)]:
        cfg = KernelConfig(T=T, R=R, TJ=256, unroll=U, min_blocks=B)
        ms, ns, regs, bps, _ = timeit(f, 5, "float32", cfg, [x, y, b, eta, fwd], M, N)
        print("%s T=%d R=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s" % (name,T,R,U,B,regs,bps,ns,ms,M*N/ms*1e3))
print("== ArgKMin K=10 D=3")
for T,R,U,B in [(256,2,8,2),(256,2,8,3),(128,2,8,4),(128,4,4,2),(256,4,4,1),(128,2,4,6)]:
    for nsplit in (1, None):
        cfg = KernelConfig(T=T, R=R, TJ=256, unroll=U, min_blocks=B)
        ms, ns, regs, bps, _ = timeit(KNN, 2, "float32", cfg, [x, y], M, N, nsplit)
        print("T=%d R=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s  %.3e q/s (%.1f%% of 5.32e12)" % (T,R,U,B,regs,bps,ns,ms,M*N/ms*1e3,M/ms*1e3,M*N/ms*1e3/5.32e12*100))
print("== LSE fp32")
f = torch.rand(N, 1, device=dev); xu = torch.rand(M,3,device=dev); yu = torch.rand(N,3,device=dev); eps = torch.tensor([0.05], device=dev)
for T,R,U,B in [(256,4,8,4),(256,4,4,3),(256,2,8,4),(128,4,8,4)]:
    cfg = KernelConfig(T=T, R=R, TJ=256, unroll=U, min_blocks=B)
    ms, ns, regs, bps, _ = timeit(LSE, 4, "float32", cfg, [f, xu, yu, eps], M, N)
    print("T=%d R=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s" % (T,R,U,B,regs,bps,ns,ms,M*N/ms*1e3))
if M >= 200000:
    M2 = N2 = 200000
    x64 = torch.randn(M2, 64, device=dev); y64 = torch.randn(N2, 64, device=dev)
    print("== ArgKMin K=10 D=64 M=N=%d" % M2)
    for T,R,U,B,TJ in [(128,2,2,1,64),(128,2,1,1,64),(256,2,2,1,64),(128,2,4,1,64),(128,2,2,2,64),(128,4,1,1,32)]:
        cfg = KernelConfig(T=T, R=R, TJ=TJ, unroll=U, min_blocks=B)
        try:
            ms, ns, regs, bps, _ = timeit(KNN64, 2, "float32", cfg, [x64, y64], M2, N2)
            print("T=%d R=%d U=%d B=%d TJ=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s (%.1f%% of 2.89e11)" % (T,R,U,B,TJ,regs,bps,ns,ms,M2*N2/ms*1e3,M2*N2/ms*1e3/2.89e11*100))
        except Exception as e:
            print("failed", T,R,U,B,TJ, str(e)[:300])
    xd = torch.rand(M2,3,device=dev,dtype=torch.float64); yd = torch.rand(N2,3,device=dev,dtype=torch.float64); fd = torch.rand(N2,1,device=dev,dtype=torch.float64); epsd = torch.tensor([0.05], device=dev, dtype=torch.float64)
    print("== LSE fp64 M=N=%d" % M2)
    for T,R,U,B in [(128,2,8,1),(128,2,4,2),(256,2,4,1),(128,4,4,1),(128,1,8,2)]:
        cfg = KernelConfig(T=T, R=R, TJ=256, unroll=U, min_blocks=B)
        ms, ns, regs, bps, _ = timeit(LSE, 4, "float64", cfg, [fd, xd, yd, epsd], M2, N2)
        print("T=%d R=%d U=%d B=%d regs=%d blk/SM=%d nsplit=%d : %.2f ms  %.3e pairs/s" % (T,R,U,B,regs,bps,ns,ms,M2*N2/ms*1e3))
