#!/usr/bin/env python
"""Benchmark of the KeOps map-reduce hot path on B200 (driver contract: one JSON line on stdout).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): Gaussian-kernel convolution ``out_i = sum_j exp(-|x_i-y_j|^2) b_j``,
M = N = 10^6, D = 3, fp32, synthetic ``randn`` point clouds (seed 0).  One *step* = one pass of the
forward reduction over the whole problem = 10^12 pair evaluations.  ``value`` is pair-evals/s with the
inputs resident in HBM; ``e2e`` is the same metric through the public ``Genred`` call with pinned HOST
buffers (H2D of x, y, b and D2H of the result inside the timed region).

N > 1 (torchrun, one rank per GPU) goes through the product's multi-GPU API, ``keops_b200.distributed``: every array is
row-sharded by its own index (rank r owns slab r of x, y, b, ...), ``ShardedGenred`` shards the *output index* and
replicates the streamed operands with one NCCL ``all_gather`` per operand per step INSIDE the timed region; there is no
cross-GPU reduction.  Total work is fixed ("scaling": "strong").

Legs, each at every N and each through the sharded API (``legs`` key; median of >= 10 CUDA-event-timed repetitions, L2
flushed in between, max over ranks): ``fwd_bwd`` (forward + grad_x sharded by i + grad_y sharded by j,
``ShardedGenred.gradient``), ``knn`` / ``knn_d64`` (ArgKMin K=10, D = 3 / 64), ``lse_softmin`` (LogSumExp, D = 3) and
``ksolve`` (``ShardedKernelSolve``, Gaussian kernel D = 16, time per CG iteration).

Extra keys: ``roofline`` (FP32/SFU roofline of SURVEY.md 8(d): 16 pairs/clk/SM -- one MUFU.EX2 per pair),
``roofline_hbm`` (algorithmic unique bytes vs measured copy bandwidth: shows the kernel is not HBM-bound),
``cpu_baseline`` (the reference's own CpuReduc code on the host cores, bounded sample, thread count reported by OpenMP
itself), ``parity`` (achieved max errors of every BASELINE config against the reference's CPU code and CUDA kernel) and,
at N = 1, ``ref_cuda`` (the reference's own generated GpuConv1DOnDevice kernel on the same GPU).

``--impl reference`` times the reference's CPU implementation (oracle/_ref, built from the unmodified
reference code generator by oracle/build_ref.py) on all host cores, on a bounded sample of the same
workload, and prints the same line with ``"impl": "reference"``.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_FULL = N_FULL = 1_000_000
GAUSS = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
GRAD_X = "Grad_WithSavedForward(%s,Var(0,3,0),Var(3,1,0),Var(4,1,0))" % GAUSS
GRAD_Y = "Grad_WithSavedForward(%s,Var(1,3,1),Var(3,1,0),Var(4,1,0))" % GAUSS
KNN = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"
KSOLVE = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)"    # BASELINE configs[4] matvec
KSOLVE_G = "Sum_Reduction(Exp(-g*SqDist(x,y))*b,0)"
LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"
SM_COUNT, SM_MAX_HZ = 148, 1.965e9
# dram__bytes_read.sum + dram__bytes_write.sum of one kb200_reduce launch of the full-size forward conv, from the
# committed `ncu --set full` capture profiles/r01_ncu_gauss_fwd_v2.txt (28.148 MB + 42 KB; algorithmic bytes 32 MB)
NCU_DRAM_BYTES_FWD = 28147968 + 42496
# measured co-issue ceiling of the conv's instruction mix (8 packed-FP32 lane-ops per MUFU.EX2, independent chains, no
# memory): 118.6 lane-ops + 14.83 ex2 / clk / SM = 92.7 % of both pipes at once (experiments/ffma2_peak.py on B200)
MEASURED_MIX_EX2_PER_CLK_SM = 14.83
ROOFLINE_PAIRS_PER_CLK_SM = {"gauss_fwd": 16.0, "gauss_grad": 128.0 / 11.0, "knn_d3": 128.0 / 7.0}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, nm in enumerate(names):
            if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        # keep the samples taken under load (upper half)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        mx = float(self.rows[0][1]) if self.rows and len(self.rows[0]) >= 2 else None
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": med, "sm_max_mhz": mx, "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "reasons": reasons}


_JSON_FD = None


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------
def host_threads():
    """Host cores this process may run on (torchrun does not pin ranks; cgroup limits show up in the affinity mask)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def openmp_use_all_cores():
    """``torch.distributed.run`` exports OMP_NUM_THREADS=1 to every rank, which silently turned the reference's OpenMP code
    into a single-thread run.  Set the thread count explicitly -- in the environment (read by libgomp when
    oracle/_ref/libkeops_ref_cpu.so pulls it in) AND through omp_set_num_threads of that same libgomp.so.1 (in case it is
    already loaded) -- and return the count OpenMP itself reports."""
    import ctypes
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_dynamic(0)
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads()), n
    except OSError:
        return None, n


def reference_cpu_rate(budget_s):
    """pair-evals/s of the reference CpuReduc code on a bounded sample: rows of the M=N=1e6 problem."""
    import numpy as np
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    omp_threads, cores = openmp_use_all_cores()
    rng = np.random.default_rng(0)
    N = N_FULL
    y = rng.standard_normal((N, 3)).astype(np.float32)
    b = rng.standard_normal((N, 1)).astype(np.float32)
    rows = max(64, 8 * cores)
    x = rng.standard_normal((rows, 3)).astype(np.float32)
    ref_runner.run_cpu("gauss_fwd_f32", [x, y, b], rows, N)            # warm-up (thread pool start)
    t0 = time.perf_counter()
    ref_runner.run_cpu("gauss_fwd_f32", [x, y, b], rows, N)
    dt = time.perf_counter() - t0
    rate = rows * N / dt
    rows2 = int(min(M_FULL, max(rows, rate * budget_s / N)))
    rows2 = max(cores, rows2 // cores * cores)
    x = rng.standard_normal((rows2, 3)).astype(np.float32)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        ref_runner.run_cpu("gauss_fwd_f32", [x, y, b], rows2, N)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return dict(value=rows2 * N / best, unit="pair-evals/s", cores=omp_threads or cores, kind="reference",
                sample="%d of %d rows x all %d columns (%.2e pairs), best of 2; OpenMP threads %s (omp_get_max_threads) on %d usable host cores"
                       % (rows2, M_FULL, N, rows2 * N, omp_threads, cores), seconds=best, rows=rows2,
                omp_max_threads=omp_threads, host_cores=cores)


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    from oracle import ref_runner
    if not ref_runner.available():
        emit_json({"impl": "reference", "unavailable": "oracle/_ref was not built (run oracle/build_ref.py where /root/reference is mounted)"})
        return 0
    omp_threads, cores = openmp_use_all_cores()
    rng = np.random.default_rng(0)
    N = N_FULL
    y = rng.standard_normal((N, 3)).astype(np.float32)
    b = rng.standard_normal((N, 1)).astype(np.float32)
    probe = reference_cpu_rate(2.0)
    total = max(1, args.steps + args.warmup)
    per_step_s = min(20.0, 150.0 / total)
    rows = int(max(cores, min(M_FULL, probe["value"] * per_step_s / N) // cores * cores))
    x = rng.standard_normal((rows, 3)).astype(np.float32)
    for _ in range(args.warmup):
        ref_runner.run_cpu("gauss_fwd_f32", [x, y, b], rows, N)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_runner.run_cpu("gauss_fwd_f32", [x, y, b], rows, N)
    dt = time.perf_counter() - t0
    value = args.steps * rows * N / dt
    sample = ("%d of %d rows x all %d columns per step; OpenMP threads %s (omp_get_max_threads) on %d usable host cores, "
              "OMP_NUM_THREADS reset from the launcher's value" % (rows, M_FULL, N, omp_threads, cores))
    line = {
        "impl": "reference", "metric": "gaussian_conv_pair_evals_per_s", "value": value, "unit": "pair-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gaussian_conv_fwd M=N=1e6 D=3 fp32 (reference CpuReduc, bounded row sample)",
                   "M": M_FULL, "N": N_FULL, "D": 3, "sample_rows": rows},
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": omp_threads or cores, "kind": "reference",
                         "sample": sample, "omp_max_threads": omp_threads, "host_cores": cores},
        "e2e": {"value": value, "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """Per-rank state of a run: device, world, the L2-flush buffer, and the timing helpers."""

    def __init__(self, torch, dist, rank, world, dev):
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, dev
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def slab(self, n):
        from keops_b200.distributed import shard_bounds
        return shard_bounds(n, self.world, self.rank)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def time_leg(self, fn, reps=10, warm=1):
        """Median over ``reps`` of (max over ranks of) the CUDA-event time of ``fn()`` -- collectives included -- with the L2
        flushed and the ranks aligned before every repetition."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.sync_all()
        ms = []
        for _ in range(reps):
            self.flush.zero_()
            self.sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = sorted(self.max_over_ranks(ms))
        return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kb200", choices=["kb200", "reference"])
    ap.add_argument("--size", type=int, default=M_FULL, help="M = N (default: the BASELINE size 10^6)")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (no legs / comparators / parity)")
    ap.add_argument("--legs", default="fwd_bwd,knn,knn_d64,lse_softmin,ksolve", help="comma-separated legs to run")
    ap.add_argument("--reps", type=int, default=10, help="timed repetitions per leg (median reported)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "kb200" else args.warmup
    # the driver parses ONE JSON line from stdout: libraries that print to fd 1 (NCCL's version banner)
    # are sent to stderr, and the JSON line is written to the original stdout
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from keops_b200.binder import get_binder
    from keops_b200.distributed import ShardedGenred, ShardedKernelSolve

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cx = Ctx(torch, dist, rank, world, dev)
    M = N = args.size
    legs_wanted = [] if args.no_extras else [s for s in args.legs.split(",") if s]
    # ---- synthetic data (same bytes on every rank: seed 0); every array is row-sharded by its own index
    g = torch.Generator(device="cpu").manual_seed(0)
    hx = torch.randn(M, 3, generator=g)
    hy = torch.randn(N, 3, generator=g)
    hb = torch.randn(N, 1, generator=g)
    heta = torch.randn(M, 1, generator=g)
    i0, i1 = cx.slab(M)
    j0, j1 = cx.slab(N)
    hx_s, hy_s, hb_s = (t.contiguous().pin_memory() for t in (hx[i0:i1], hy[j0:j1], hb[j0:j1]))
    x, y, b = hx_s.to(dev), hy_s.to(dev), hb_s.to(dev)
    Ms = i1 - i0
    al = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"]
    conv = ShardedGenred("Exp(-SqDist(x,y))*b", al, "Sum", axis=1)
    out = torch.empty(Ms, 1, device=dev)

    def step():
        conv(x, y, b, n_streamed=N, mode="all_gather", out=out)

    for _ in range(args.warmup):
        cx.flush.zero_()
        step()
    cx.sync_all()
    from keops_b200.torch import genred as genred_mod
    fwd_binder = genred_mod.LAST_BINDER                 # the binder the public API call really used
    launches_per_step = int(fwd_binder.last_launch["launches"])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    cx.sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        cx.flush.zero_()                   # L2 flush between timed iterations (inputs are 28 MB < 126 MB L2)
        ev[k][0].record()
        step()
        ev[k][1].record()
    cx.sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = [a.elapsed_time(b_) for a, b_ in ev]
    tot_s = cx.max_over_ranks([sum(dev_ms)])[0] * 1e-3
    value = args.steps * float(M) * float(N) / tot_s
    kernel_launches = args.steps * launches_per_step

    # ---- e2e through the public API with host buffers: every rank uploads ITS slabs from pinned memory, the sharded call
    # replicates the streamed ones over NVLink, the result slab goes back to pinned host memory
    hout = torch.empty(Ms, 1).pin_memory()

    def e2e_step():
        xd = hx_s.to(dev, non_blocking=True)
        yd = hy_s.to(dev, non_blocking=True)
        bd = hb_s.to(dev, non_blocking=True)
        res = conv(xd, yd, bd, n_streamed=N, mode="all_gather")
        hout.copy_(res, non_blocking=True)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(3):                      # warm-up: pinned-memory registration, NCCL channel set-up for these sizes
        e2e_step()
    cx.sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    cx.sync_all()
    e2e_t = cx.max_over_ranks([time.perf_counter() - t0])[0]
    e2e_value = e2e_steps * float(M) * float(N) / e2e_t
    h2d = 4 * (M * 3 + N * 3 + N * 1)          # whole job: the slabs of all ranks
    d2h = 4 * M

    peaks, peaks_src = measured_peaks()
    sm_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    cc_peak = lambda ops: 128.0 / ops * SM_COUNT * sm_hz * world        # FP32-pipe roofline of an `ops`-instruction pair
    legs = {}

    # ---- forward + backward (BASELINE configs[1]): grad_x stays sharded by i, grad_y is sharded by j ------------------
    fwd_slab = None
    if "fwd_bwd" in legs_wanted:
        eta = heta[i0:i1].to(dev)
        fwd_slab = conv(x, y, b, n_streamed=N).clone()
        gx_op, gy_op = conv.gradient("x"), conv.gradient("y")
        f_ms, _ = cx.time_leg(step, args.reps)
        gx_ms, _ = cx.time_leg(lambda: gx_op(x, y, b, eta, fwd_slab, n_i=M, n_j=N), args.reps)
        gy_ms, _ = cx.time_leg(lambda: gy_op(x, y, b, eta, fwd_slab, n_i=M, n_j=N), args.reps)
        legs["fwd_bwd"] = {"value": 3.0 * M * N / ((f_ms + gx_ms + gy_ms) * 1e-3), "unit": "pair-evals/s", "n_gpus": world,
                           "fwd_ms": f_ms, "grad_x_ms": gx_ms, "grad_y_ms": gy_ms,
                           "sharding": "fwd and grad_x by i (all_gather y, b); grad_y by j (all_gather x, eta, saved forward)",
                           "grad_x_frac_of_roofline": M * N / (gx_ms * 1e-3) / (ROOFLINE_PAIRS_PER_CLK_SM["gauss_grad"] * SM_COUNT * sm_hz * world),
                           "grad_y_frac_of_roofline": M * N / (gy_ms * 1e-3) / (ROOFLINE_PAIRS_PER_CLK_SM["gauss_grad"] * SM_COUNT * sm_hz * world)}

    # ---- KNN K=10 (BASELINE configs[2]): every rank answers its slab of queries against the all-gathered database -------
    def knn_leg(D):
        gg = torch.Generator(device=dev).manual_seed(D)
        q_full = torch.randn(M, D, generator=gg, device=dev)
        d_full = torch.randn(N, D, generator=gg, device=dev)
        q, d = q_full[i0:i1].contiguous(), d_full[j0:j1].contiguous()
        del q_full, d_full
        op = ShardedGenred("SqDist(x,y)", ["x=Vi(%d)" % D, "y=Vj(%d)" % D], "ArgKMin", axis=1, opt_arg=10)
        ms, _ = cx.time_leg(lambda: op(q, d, n_streamed=N), args.reps)
        engine = (getattr(genred_mod.LAST_BINDER, "last_launch", None) or {}).get("engine") or "cuda-core"
        res = {"metric": "knn_k10_queries_per_s", "value": M / (ms * 1e-3), "unit": "queries/s", "D": D, "K": 10, "ms": ms,
               "pair_evals_per_s": float(M) * N / (ms * 1e-3), "n_gpus": world, "engine": engine}
        if engine == "cuda-core":
            res["frac_of_roofline"] = float(M) * N / (ms * 1e-3) / cc_peak(2 * D + 1)
        elif engine == "cuda-core/kmin-filter":
            # filtered kernel (codegen/emit_kmin_filter.py): D FMAs per pair on the FP32 pipe (+ 1/2 FMNMX3 on the ALU pipe)
            res["frac_of_roofline"] = float(M) * N / (ms * 1e-3) / cc_peak(D)
            res["roofline_note"] = ("FP32-pipe roofline of the %d-FMA filter (128 / %d pairs/clk/SM); against the reference's "
                                    "%d-operation pair (the round-1 roofline) the rate is %.2f" %
                                    (D, D, 2 * D + 1, float(M) * N / (ms * 1e-3) / cc_peak(2 * D + 1)))
        else:
            from keops_b200.codegen import knn_tc
            kdim = knn_tc.KnnTensorCore(D, 10, "ArgKMin", True, device=-1).DPG
            tflops = 2.0 * M * N * kdim / (ms * 1e-3) / 1e12
            res["roofline"] = {"bound": "tensor", "achieved": tflops, "peak": float(peaks.get("bf16_tflops", 1646.4)) * world,
                               "unit": "TFLOP/s", "frac": tflops / (float(peaks.get("bf16_tflops", 1646.4)) * world),
                               "note": "bf16 filter GEMM 2*M*N*%d flops (D + 3 penalty columns, padded) over the whole call "
                                       "(prep + filter + exact select) vs the measured cuBLAS bf16 burst peak (%s)" % (kdim, peaks_src)}
            res["x_cuda_core_fp32_roofline"] = float(M) * N / (ms * 1e-3) / cc_peak(2 * D + 1)
        return res

    for name, D in (("knn", 3), ("knn_d64", 64)):
        if name in legs_wanted:
            try:
                legs[name] = knn_leg(D)
            except Exception as e:
                legs[name] = {"unavailable": repr(e)[:300]}

    # ---- LogSumExp soft-min (BASELINE configs[3]) -------------------------------------------------------------------------
    if "lse_softmin" in legs_wanted:
        gg = torch.Generator(device=dev).manual_seed(7)
        lf, lx, ly = torch.rand(N, 1, generator=gg, device=dev), torch.rand(M, 3, generator=gg, device=dev), torch.rand(N, 3, generator=gg, device=dev)
        lf, lx, ly = lf[j0:j1].contiguous(), lx[i0:i1].contiguous(), ly[j0:j1].contiguous()
        eps = torch.tensor([0.05], device=dev)
        op = ShardedGenred("f - SqDist(x,y)/IntCst(2)/e", ["f=Vj(1)", "x=Vi(3)", "y=Vj(3)", "e=Pm(1)"], "LogSumExp", axis=1)
        ms, _ = cx.time_leg(lambda: op(lf, lx, ly, eps, n_streamed=N), args.reps)
        legs["lse_softmin"] = {"workload": "LogSumExp soft-min M=N=%d D=3 fp32 (BASELINE configs[3])" % M, "ms": ms,
                               "value": float(M) * N / (ms * 1e-3), "unit": "pair-evals/s", "n_gpus": world,
                               "frac_of_roofline": float(M) * N / (ms * 1e-3) / cc_peak(12)}
        del lf, lx, ly

    # ---- KernelSolve (BASELINE configs[4]): row-sharded CG, Gaussian kernel D = 16; one iteration = M*N pairs -------------
    if "ksolve" in legs_wanted:
        try:
            gg = torch.Generator(device=dev).manual_seed(16)
            kx = torch.rand(M, 16, generator=gg, device=dev)[i0:i1].contiguous()
            rhs = torch.randn(M, 1, generator=gg, device=dev)[i0:i1].contiguous()
            kg = torch.tensor([1.0], device=dev)
            ITER = 4
            ks = ShardedKernelSolve("Exp(-g*SqDist(x,y))*b", ["x=Vi(16)", "y=Vj(16)", "b=Vj(1)", "g=Pm(1)"], "b", axis=1)
            ms, _ = cx.time_leg(lambda: ks(kx, kx, rhs, kg, n_total=M, alpha=0.8, eps=0.0, max_iter=ITER), max(3, args.reps // 2))
            engine = (getattr(genred_mod.LAST_BINDER, "last_launch", None) or {}).get("engine") or "cuda-core"
            legs["ksolve"] = {"workload": "ShardedKernelSolve, Gaussian kernel D=16, N=%d fp32: %d CG iterations (all_gather of the "
                                          "point cloud once, of the search direction every iteration)" % (M, ITER),
                              "ms_per_iteration": ms / ITER, "iterations": ks.last_iterations, "value": ITER * float(M) * M / (ms * 1e-3),
                              "unit": "pair-evals/s", "n_gpus": world, "engine": engine,
                              "frac_of_mufu_roofline": ITER * float(M) * M / (ms * 1e-3) / (16.0 * SM_COUNT * sm_hz * world),
                              "x_cuda_core_fp32_roofline": ITER * float(M) * M / (ms * 1e-3) / cc_peak(34)}
            del kx, rhs
        except Exception as e:
            legs["ksolve"] = {"unavailable": repr(e)[:300]}

    if world > 1:
        cx.sync_all()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (kb200_reduce of the forward conv): measured live on this rank
    k_ms = sum(dev_ms) / len(dev_ms)
    pairs_per_launch = float(Ms) * float(N)
    achieved = pairs_per_launch / (k_ms * 1e-3)
    peak_pairs = ROOFLINE_PAIRS_PER_CLK_SM["gauss_fwd"] * SM_COUNT * sm_hz
    algo_bytes = 4.0 * (Ms * 3 + N * 4 + Ms * 1)
    line = {
        "metric": "gaussian_conv_pair_evals_per_s", "value": value, "unit": "pair-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gaussian_conv_fwd M=N=%d D=3 fp32 (BASELINE configs[1], forward reduction)" % M,
                   "M": M, "N": N, "D": 3,
                   "api": "keops_b200.distributed.ShardedGenred (output rows over %d GPU(s); y, b all-gathered over NCCL inside the timed step)" % world,
                   "l2": "256 MiB memset between timed steps (inputs are 28 MB < 126 MB L2)",
                   "kernel_config": fwd_binder.ks.cfg.key(), "nsplit": fwd_binder.last_launch.get("nsplit"),
                   "engine": fwd_binder.last_launch.get("engine") or "cuda-core",
                   "regs": fwd_binder.last_launch.get("regs"), "blocks_per_sm": fwd_binder.last_launch.get("blocks_per_sm")},
        "wall_ms_per_step": wall / args.steps * 1e3,
        "e2e": {"value": e2e_value, "unit": "pair-evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "ShardedGenred / Genred: pinned host slabs -> device -> (all_gather) -> reduction -> result slab back to pinned host"},
        "gpu_launches": int(kernel_launches),
        "clocks": clocks,
        "roofline": {"bound": "sfu", "achieved": achieved, "peak": peak_pairs, "unit": "pair-evals/s",
                     "frac": achieved / peak_pairs,
                     "traffic": NCU_DRAM_BYTES_FWD if (world == 1 and M == M_FULL) else None,
                     "note": "FP32/SFU roofline (SURVEY 8d): 1 MUFU.EX2 per pair at 16/clk/SM x 148 SM x %.0f MHz (max SM clock, %s); "
                             "per-GPU: rank 0's rows / rank 0's mean step time" % (sm_hz / 1e6, peaks_src),
                     "kernel": "kb200_reduce", "ms_per_launch": k_ms,
                     "frac_of_measured_mix_ceiling": achieved / (MEASURED_MIX_EX2_PER_CLK_SM * SM_COUNT * sm_hz),
                     "measured_mix_ceiling_note": "micro-benchmark of the same 8 FP32 : 1 EX2 mix reaches 14.83 pairs/clk/SM "
                                                  "(92.7 % of nominal) -- experiments/ffma2_peak.py"},
        "roofline_hbm": {"bound": "hbm", "achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                         "unit": "GB/s", "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
                         "traffic": None, "note": "algorithmic unique bytes 4*(M*3+N*4+M) per launch; the kernel is compute-bound (%s peak)" % peaks_src},
        "legs": sorted(legs),
    }
    if clocks and clocks.get("sm_mhz"):
        line["roofline"]["frac_at_sustained_clock"] = achieved / (ROOFLINE_PAIRS_PER_CLK_SM["gauss_fwd"] * SM_COUNT * clocks["sm_mhz"] * 1e6)
    line.update(legs)                      # one top-level key per leg: fwd_bwd, knn, knn_d64, lse_softmin, ksolve
    if world > 1:
        dist.destroy_process_group()       # the other ranks are done; rank 0 goes on alone with the host-side comparators

    if not args.no_extras:
        # ---- the reference CPU code beside it, at EVERY N (rank 0's host cores; OpenMP thread count set explicitly) ------
        try:
            cb = reference_cpu_rate(12.0)
            if cb is not None:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "omp_max_threads", "host_cores")}
        except Exception as e:
            line["cpu_baseline"] = {"unavailable": repr(e)[:200]}
    if not args.no_extras and world == 1:
        try:
            line["parity"] = parity_report(torch, dev, cx, M, N, (hx, hy, hb, heta), line)
        except Exception as e:
            line["parity"] = {"unavailable": repr(e)[:300]}
    emit_json(line)
    return 0


def parity_report(torch, dev, cx, M, N, host, line):
    """Achieved errors of every BASELINE config at full size against the reference's own code: its generated CPU code on a
    sample of output rows and its generated CUDA kernel on all rows (oracle/_ref).  Sum-type outputs: max over entries of
    |ours - ref| / sum_j |F_ij| (the condition scale, evaluated by an Abs-reduction on the GPU); LogSumExp: max relative
    error of m + log s; K-min: number of differing indices.  Also times the reference CUDA kernels (``ref_cuda``)."""
    import numpy as np
    from keops_b200.binder import get_binder
    from oracle import ref_runner
    if not ref_runner.available():
        return {"unavailable": "oracle/_ref not built"}
    from keops_b200.formulas.parser import parse_reduction
    hx, hy, hb, heta = host
    x, y, b, eta = (t.to(dev) for t in (hx, hy, hb, heta))
    rep = {"definition": "sum-type: max |ours-ref| / sum_j|F_ij|; lse: max rel. error of m+log(s); knn: differing indices / compared",
           "cpu_sample_rows": 256}
    rows = np.sort(np.random.default_rng(3).choice(M, size=256, replace=False))
    rsel = torch.from_numpy(rows).to(dev)

    def run(formula, nargs, *targs, dtype="float32"):
        o = get_binder(formula, [], nargs, dtype).genred_pytorch(dev, None, M, N, 0, None, *targs)
        torch.cuda.synchronize()
        return o

    def abs_scale(formula, nargs, *targs):
        red = parse_reduction(formula, [])
        return run("Sum_Reduction(Abs(%r),%d)" % (red.formula, red.tagI), nargs, *targs)

    def ref_cuda(key, targs, reps=1):
        rc = ref_runner.RefCuda(key, dev.index or 0)
        o = rc(targs, M, N)
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            cx.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc(targs, M, N, out=o)
            e1.record()
            torch.cuda.synchronize()
            best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        return o, best

    def sum_case(name, key, formula, nargs, targs, i_pos):
        ours = run(formula, nargs, *targs)
        sc = abs_scale(formula, nargs, *targs)
        tagI = parse_reduction(formula, []).tagI
        entry = {}
        refc, ms = ref_cuda(key, list(targs), reps=2 if name == "gauss_fwd" else 1)
        entry["vs_ref_cuda_all_rows"] = float(((ours - refc).abs() / sc).max())
        entry["ref_cuda_ms"] = ms
        if tagI == 0:                      # CPU sample: restrict the i-indexed arrays to the sampled rows
            sub = [t[rsel].cpu().numpy() if k in i_pos else t.cpu().numpy() for k, t in enumerate(targs)]
            cpu = ref_runner.run_cpu(key, sub, len(rows), N)
            entry["vs_ref_cpu_sampled_rows"] = float(((ours[rsel].cpu() - torch.from_numpy(cpu)).abs() / sc[rsel].cpu()).max())
        rep[name] = entry
        return ours, ms

    fwd, fwd_ref_ms = sum_case("gauss_fwd", "gauss_fwd_f32", GAUSS, 3, (x, y, b), {0})
    sum_case("gauss_grad_x", "gauss_dx_f32", GRAD_X, 5, (x, y, b, eta, fwd), {0, 3, 4})
    sum_case("gauss_grad_y", "gauss_dy_f32", GRAD_Y, 5, (x, y, b, eta, fwd), {0, 3, 4})
    ours_ms = line["roofline"]["ms_per_launch"]
    line["ref_cuda"] = {"value": M * N / (fwd_ref_ms * 1e-3), "unit": "pair-evals/s", "ms": fwd_ref_ms,
                        "kernel": "reference GpuConv1DOnDevice (GpuReduc1D, block 192, -use_fast_math), same GPU",
                        "max_err_vs_ours_over_condition_scale": rep["gauss_fwd"]["vs_ref_cuda_all_rows"],
                        "speedup_ours": fwd_ref_ms / ours_ms}
    # K-min: indices
    for D, key in ((3, "argkmin10_d3_f32"), (64, "argkmin10_d64_f32")):
        gg = torch.Generator(device=dev).manual_seed(D)
        q, d = torch.randn(M, D, generator=gg, device=dev), torch.randn(N, D, generator=gg, device=dev)
        f = "ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),10,0)" % (D, D)
        ours = run(f, 2, q, d)
        cpu = ref_runner.run_cpu(key, [q[rsel].cpu().numpy(), d.cpu().numpy()], len(rows), N)
        entry = {"vs_ref_cpu_sampled_rows": "%d / %d" % (int((ours[rsel].cpu() != torch.from_numpy(cpu)).sum()), cpu.size)}
        nq = M if D == 3 else 50000
        rc = ref_runner.RefCuda(key, dev.index or 0)
        refc = rc([q[:nq].contiguous(), d], nq, N)
        torch.cuda.synchronize()
        entry["vs_ref_cuda"] = "%d / %d" % (int((ours[:nq] != refc).sum()), refc.numel())
        rep["knn_d%d" % D] = entry
        del q, d
    # LogSumExp fp32 / fp64
    for dtype, key in (("float32", "lse_f32"), ("float64", "lse_f64")):
        tdt = torch.float32 if dtype == "float32" else torch.float64
        gg = torch.Generator(device=dev).manual_seed(7)
        lf = torch.rand(N, 1, generator=gg, device=dev, dtype=tdt)
        lx, ly = torch.rand(M, 3, generator=gg, device=dev, dtype=tdt), torch.rand(N, 3, generator=gg, device=dev, dtype=tdt)
        eps = torch.tensor([0.05], device=dev, dtype=tdt)
        ms_ = run(LSE, 4, lf, lx, ly, eps, dtype=dtype)
        ours = ms_[:, 0] + ms_[:, 1].log()
        cpu = ref_runner.run_cpu(key, [lf.cpu().numpy(), lx[rsel].cpu().numpy(), ly.cpu().numpy(), eps.cpu().numpy()], len(rows), N)
        refv = torch.from_numpy(cpu[:, 0] + np.log(cpu[:, 1]))
        entry = {"vs_ref_cpu_sampled_rows": float(((ours[rsel].cpu() - refv).abs() / refv.abs()).max())}
        if dtype == "float32":
            refc, _ = ref_cuda(key, [lf, lx, ly, eps])
            rv = refc[:, 0] + refc[:, 1].log()
            entry["vs_ref_cuda_all_rows"] = float(((ours - rv).abs() / rv.abs()).max())
        rep["lse_" + dtype] = entry
        del lf, lx, ly
    # KernelSolve matvec, D = 16 (tensor-core path, stated tolerance 5e-6)
    gg = torch.Generator(device=dev).manual_seed(16)
    kx, kb_ = torch.rand(M, 16, generator=gg, device=dev), torch.randn(M, 1, generator=gg, device=dev)
    kg = torch.tensor([1.0], device=dev)
    ours = run(KSOLVE, 4, kx, kx, kb_, kg)
    sc = abs_scale(KSOLVE, 4, kx, kx, kb_, kg)
    cpu = ref_runner.run_cpu("gauss16_pm_f32", [kx[rsel].cpu().numpy(), kx.cpu().numpy(), kb_.cpu().numpy(), kg.cpu().numpy()], len(rows), N)
    entry = {"vs_ref_cpu_sampled_rows": float(((ours[rsel].cpu() - torch.from_numpy(cpu)).abs() / sc[rsel].cpu()).max())}
    refc, ms = ref_cuda("gauss16_pm_f32", [kx, kx, kb_, kg])
    entry["vs_ref_cuda_all_rows"] = float(((ours - refc).abs() / sc).max())
    entry["ref_cuda_ms"] = ms
    rep["ksolve_matvec_d16"] = entry
    return rep


if __name__ == "__main__":
    sys.exit(main())
