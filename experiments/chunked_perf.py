"""GPU timing of the chunked-dimension kernel against the generic (register-resident, fully unrolled) kernel.
python experiments/chunked_perf.py [--prebuild]   (--prebuild: only cross-compile the kernels, no GPU needed)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.binder import get_binder
from keops_b200.runtime import plan
from keops_b200.codegen import emit_chunked

CASES = [
    ("matern_d256", "Sum_Reduction(Exp(-Sqrt(SqDist(Var(0,256,0),Var(1,256,1))))*Var(2,1,1),0)", 3, 100000, 100000, 256),
    ("laplace_d128", "Sum_Reduction(Exp(-Sum(Abs(Var(0,128,0)-Var(1,128,1))))*Var(2,1,1),0)", 3, 100000, 100000, 128),
    ("argkmin_d784", "ArgKMin_Reduction(Sum(Square((Var(0,784,0) - Var(1,784,1)))),10,0)", 2, 10000, 100000, 784),
]
prebuild = "--prebuild" in sys.argv
for name, f, nargs, M, N, D in CASES:
    bd = get_binder(f, [], nargs, "float32")
    if prebuild:
        t = time.time()
        ks = emit_chunked.generate_chunked(bd.red, bd._chunk_plan, "float32", nargs)
        plan.get_cubin(ks, True)
        t1 = time.time() - t
        if D <= 256:
            plan.get_cubin(bd.ks, True)
        print("%s: chunked compiled in %.1f s, generic in %.1f s" % (name, t1, time.time() - t - t1))
        continue
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x, y = torch.rand(M, D, device=dev) / D ** 0.5, torch.rand(N, D, device=dev) / D ** 0.5
    args = (x, y) if nargs == 2 else (x, y, torch.randn(N, 1, device=dev))
    res = {}
    for mode in (("chunked", "generic") if D <= 256 else ("chunked",)):
        os.environ["KB200_CHUNKS"] = "1" if mode == "chunked" else "0"
        os.environ["KB200_CONV_TC"] = "0"
        os.environ["KB200_KNN_TC"] = "0"
        o = bd.genred_pytorch(dev, None, M, N, 0, None, *args)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = bd.genred_pytorch(dev, None, M, N, 0, None, *args); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = (best, o, dict(bd.last_launch))
    ms, o, ll = res["chunked"]
    ops = 2.0 * D        # sub/abs + fma-class per component (lane-ops on the FP32 pipe)
    frac = M * N * ops / (ms * 1e-3) / (128 * 148 * 1.965e9)
    line = "%s M=%d N=%d: chunked %.2f ms (%.3e pairs/s, %.2f of the FP32-pipe roofline at %d lane-ops/pair) %s" % (name, M, N, ms, M * N / ms * 1e3, frac, int(ops), ll)
    if "generic" in res:
        g_ms, g_o, _ = res["generic"]
        err = float((o - g_o).abs().max() / g_o.abs().max())
        line += " | generic %.2f ms -> %.2fx, max rel diff %.2e" % (g_ms, g_ms / ms, err)
    print(line)
