#!/usr/bin/env python
"""Benchmark of the KeOps map-reduce hot path on B200 (driver contract: one JSON line on stdout).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): Gaussian-kernel convolution ``out_i = sum_j exp(-|x_i-y_j|^2) b_j``,
M = N = 10^6, D = 3, fp32, synthetic ``randn`` point clouds (seed 0).  One *step* = one pass of the
forward reduction over the whole problem = 10^12 pair evaluations.  ``value`` is pair-evals/s with the
inputs resident in HBM; ``e2e`` is the same metric through the public ``Genred`` call with pinned HOST
buffers (H2D of x, y, b and D2H of the result inside the timed region).

N > 1 (torchrun, one rank per GPU): the *output index* is sharded -- rank r owns rows
[r*M/N, (r+1)*M/N) of x and of the output -- and the streamed operands (y, b) are replicated with one
NCCL broadcast per step inside the timed region; there is no cross-GPU reduction.  Total work is fixed
("scaling": "strong").

Extra keys: ``roofline`` (FP32/SFU roofline of SURVEY.md 8(d): 16 pairs/clk/SM -- one MUFU.EX2 per
pair), ``roofline_hbm`` (algorithmic unique bytes vs measured copy bandwidth: shows the kernel is not
HBM-bound), ``cpu_baseline`` (the reference's own CpuReduc code on the host cores, bounded sample),
``ref_cuda`` (the reference's own generated GpuConv1DOnDevice kernel on the same GPU), ``knn``
(ArgKMin K=10 queries/s, configs[2]) and ``fwd_bwd`` (forward + grad_x + grad_y rate).

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
def reference_cpu_rate(budget_s, full=False):
    """pair-evals/s of the reference CpuReduc code on a bounded sample: rows of the M=N=1e6 problem."""
    import numpy as np
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    rng = np.random.default_rng(0)
    N = N_FULL
    y = rng.standard_normal((N, 3)).astype(np.float32)
    b = rng.standard_normal((N, 1)).astype(np.float32)
    cores = os.cpu_count() or 1
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
    return dict(value=rows2 * N / best, unit="pair-evals/s", cores=cores, kind="reference",
                sample="%d of %d rows x all %d columns (%.2e pairs), best of 2, OpenMP threads = all %d cores"
                       % (rows2, M_FULL, N, rows2 * N, cores), seconds=best, rows=rows2)


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    from oracle import ref_runner
    if not ref_runner.available():
        emit_json({"impl": "reference", "unavailable": "oracle/_ref was not built (run oracle/build_ref.py where /root/reference is mounted)"})
        return 0
    rng = np.random.default_rng(0)
    N = N_FULL
    cores = os.cpu_count() or 1
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
    sample = "%d of %d rows x all %d columns per step, OpenMP over %d host cores" % (rows, M_FULL, N, cores)
    line = {
        "impl": "reference", "metric": "gaussian_conv_pair_evals_per_s", "value": value, "unit": "pair-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gaussian_conv_fwd M=N=1e6 D=3 fp32 (reference CpuReduc, bounded row sample)",
                   "M": M_FULL, "N": N_FULL, "D": 3, "sample_rows": rows},
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kb200", choices=["kb200", "reference"])
    ap.add_argument("--size", type=int, default=M_FULL, help="M = N (default: the BASELINE size 10^6)")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / ref_cuda / knn / fwd_bwd legs")
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
    from keops_b200.torch import Genred

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    M = N = args.size
    # ---- synthetic data (same bytes on every rank: seed 0), row slab of this rank
    g = torch.Generator(device="cpu").manual_seed(0)
    hx = torch.randn(M, 3, generator=g)
    hy = torch.randn(N, 3, generator=g)
    hb = torch.randn(N, 1, generator=g)
    r0, r1 = rank * M // world, (rank + 1) * M // world
    hx_slab = hx[r0:r1].contiguous().pin_memory()
    hy_p, hb_p = hy.pin_memory(), hb.pin_memory()
    x = hx_slab.to(dev)
    if world > 1:
        # rank 0 owns the streamed operands; replicas are (re)filled by the broadcast of each step
        y = hy.to(dev) if rank == 0 else torch.empty(N, 3, device=dev)
        b = hb.to(dev) if rank == 0 else torch.empty(N, 1, device=dev)
    else:
        y, b = hy.to(dev), hb.to(dev)
    Ms = r1 - r0
    binder = get_binder(GAUSS, [], 3, "float32")
    out = torch.empty(Ms, 1, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        if world > 1:
            dist.broadcast(y, 0)
            dist.broadcast(b, 0)
        binder.genred_pytorch(dev, None, Ms, N, 0, out, x, y, b)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (inputs are 28 MB < L2)
        ev[k][0].record()
        step()
        ev[k][1].record()
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = [a.elapsed_time(b_) for a, b_ in ev]
    tot_ms = torch.tensor([sum(dev_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    tot_s = float(tot_ms.item()) * 1e-3
    value = args.steps * float(M) * float(N) / tot_s
    launches_per_step = binder.last_launch["launches"]
    kernel_launches = args.steps * launches_per_step

    # ---- e2e through the public API with host buffers ------------------------------------------
    conv = Genred("Exp(-SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"], reduction_op="Sum", axis=1)
    hout = torch.empty(Ms, 1).pin_memory()

    def e2e_step():
        xd = hx_slab.to(dev, non_blocking=True)
        if world > 1:
            yd = hy_p.to(dev, non_blocking=True) if rank == 0 else torch.empty(N, 3, device=dev)
            bd = hb_p.to(dev, non_blocking=True) if rank == 0 else torch.empty(N, 1, device=dev)
            dist.broadcast(yd, 0)
            dist.broadcast(bd, 0)
        else:
            yd, bd = hy_p.to(dev, non_blocking=True), hb_p.to(dev, non_blocking=True)
        res = conv(xd, yd, bd)
        hout.copy_(res, non_blocking=True)

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * float(M) * float(N) / float(e2e_t.item())
    h2d = hx_slab.numel() * 4 + (hy_p.numel() + hb_p.numel()) * 4 * (1 if (world == 1 or rank == 0) else 0)
    d2h = hout.numel() * 4

    # ---- KNN K=10 (BASELINE configs[2], D=3) on the same sharding: each rank answers its slab of
    # queries against the full (replicated) database; max over ranks; aggregate queries/s
    knn_line = None
    if not args.no_extras:
        knn_b = get_binder(KNN, [], 2, "float32")
        ko = knn_b.genred_pytorch(dev, None, Ms, N, 0, None, x, y)
        sync_all()
        knn_ms = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if world > 1:
                dist.broadcast(y, 0)
            knn_b.genred_pytorch(dev, None, Ms, N, 0, ko, x, y)
            e1.record()
            sync_all()
            knn_ms.append(e0.elapsed_time(e1))
        kt = torch.tensor([min(knn_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        knn_t = float(kt.item())
        knn_line = {"metric": "knn_k10_queries_per_s", "value": M / (knn_t * 1e-3), "unit": "queries/s", "D": 3, "K": 10,
                    "ms": knn_t, "pair_evals_per_s": float(M) * N / (knn_t * 1e-3), "n_gpus": world,
                    "frac_of_roofline": float(M) * N / (knn_t * 1e-3) / (ROOFLINE_PAIRS_PER_CLK_SM["knn_d3"] * SM_COUNT * SM_MAX_HZ * world)}
        del ko

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peaks_src = measured_peaks()
    sm_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
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
                   "M": M, "N": N, "D": 3, "sharding": "output rows over %d GPU(s); y,b replicated by NCCL broadcast" % world,
                   "l2": "256 MiB memset between timed steps (inputs are 28 MB < 126 MB L2)",
                   "kernel_config": binder.ks.cfg.key(), "nsplit": binder.last_launch["nsplit"],
                   "regs": binder.last_launch["regs"], "blocks_per_sm": binder.last_launch["blocks_per_sm"]},
        "wall_ms_per_step": wall / args.steps * 1e3,
        "e2e": {"value": e2e_value, "unit": "pair-evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "keops_b200.torch.Genred (pinned host buffers -> device -> result back to host)"},
        "gpu_launches": int(kernel_launches),
        "clocks": clocks,
        "roofline": {"bound": "sfu", "achieved": achieved, "peak": peak_pairs, "unit": "pair-evals/s",
                     "frac": achieved / peak_pairs,
                     "traffic": NCU_DRAM_BYTES_FWD if (world == 1 and M == M_FULL) else None,
                     "note": "FP32/SFU roofline (SURVEY 8d): 1 MUFU.EX2 per pair at 16/clk/SM x 148 SM x %.0f MHz (max SM clock, %s)"
                             % (sm_hz / 1e6, peaks_src),
                     "kernel": "kb200_reduce", "ms_per_launch": k_ms,
                     "frac_of_measured_mix_ceiling": achieved / (MEASURED_MIX_EX2_PER_CLK_SM * SM_COUNT * sm_hz),
                     "measured_mix_ceiling_note": "micro-benchmark of the same 8 FP32 : 1 EX2 mix reaches 14.83 pairs/clk/SM "
                                                  "(92.7 % of nominal) -- experiments/ffma2_peak.py"},
        "roofline_hbm": {"bound": "hbm", "achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                         "unit": "GB/s", "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
                         "traffic": None, "note": "algorithmic unique bytes 4*(M*3+N*4+M) per launch; the kernel is compute-bound (%s peak)" % peaks_src},
    }
    if clocks and clocks.get("sm_mhz"):
        line["roofline"]["frac_at_sustained_clock"] = achieved / (ROOFLINE_PAIRS_PER_CLK_SM["gauss_fwd"] * SM_COUNT * clocks["sm_mhz"] * 1e6)

    if knn_line is not None:
        line["knn"] = knn_line
    if not args.no_extras and world == 1:
        # ---- backward reductions + KNN (device-resident, CUDA events) -----------------------------
        def time_binder(bd, nrow, ncol, targs, reps=3):
            o = bd.genred_pytorch(dev, None, nrow, ncol, 0, None, *targs)
            torch.cuda.synchronize()
            best = None
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                bd.genred_pytorch(dev, None, nrow, ncol, 0, o, *targs)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            return best, o
        eta = torch.randn(M, 1, device=dev)
        fwd_ms, fwd = time_binder(binder, M, N, (x, y, b))
        gx_ms, _ = time_binder(get_binder(GRAD_X, [], 5, "float32"), M, N, (x, y, b, eta, fwd))
        gy_ms, _ = time_binder(get_binder(GRAD_Y, [], 5, "float32"), M, N, (x, y, b, eta, fwd))
        pk = ROOFLINE_PAIRS_PER_CLK_SM["gauss_grad"] * SM_COUNT * sm_hz
        line["fwd_bwd"] = {"value": 3.0 * M * N / ((fwd_ms + gx_ms + gy_ms) * 1e-3), "unit": "pair-evals/s",
                           "fwd_ms": fwd_ms, "grad_x_ms": gx_ms, "grad_y_ms": gy_ms,
                           "grad_x_frac_of_roofline": M * N / (gx_ms * 1e-3) / pk,
                           "grad_y_frac_of_roofline": M * N / (gy_ms * 1e-3) / pk}
        # ---- the other BASELINE configs at full size: KNN K=10 in D=64 (configs[2]) and the LogSumExp soft-min (configs[3])
        try:
            KNN64 = "ArgKMin_Reduction(Sum(Square((Var(0,64,0) - Var(1,64,1)))),10,0)"
            qx, qy = torch.randn(M, 64, device=dev), torch.randn(N, 64, device=dev)
            kb64 = get_binder(KNN64, [], 2, "float32")
            k64_ms, _ = time_binder(kb64, M, N, (qx, qy))
            line["knn_d64"] = {"metric": "knn_k10_queries_per_s", "value": M / (k64_ms * 1e-3), "unit": "queries/s", "D": 64,
                               "K": 10, "ms": k64_ms, "pair_evals_per_s": float(M) * N / (k64_ms * 1e-3),
                               "engine": kb64.last_launch.get("engine") or "cuda-core",
                               "x_cuda_core_fp32_roofline": float(M) * N / (k64_ms * 1e-3) / (128.0 / 129.0 * SM_COUNT * sm_hz)}
            del qx, qy
            LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"
            lf, lx, ly = torch.rand(N, 1, device=dev), torch.rand(M, 3, device=dev), torch.rand(N, 3, device=dev)
            lse_ms, _ = time_binder(get_binder(LSE, [], 4, "float32"), M, N, (lf, lx, ly, torch.tensor([0.05], device=dev)))
            line["lse_softmin"] = {"workload": "LogSumExp soft-min M=N=%d D=3 fp32 (BASELINE configs[3])" % M, "ms": lse_ms,
                                   "value": float(M) * N / (lse_ms * 1e-3), "unit": "pair-evals/s",
                                   "frac_of_roofline": float(M) * N / (lse_ms * 1e-3) / (128.0 / 12.0 * SM_COUNT * sm_hz)}
            del lf, lx, ly
        except Exception as e:
            line["knn_d64"] = {"unavailable": str(e)[:300]}
        # ---- KernelSolve matvec (BASELINE configs[4]): Gaussian kernel, D = 16, one CG iteration = M*N pairs.
        # Tensor-core path (tcgen05 GEMM for |x-y|^2, MUFU-bound) vs the CUDA-core kernel (FP32-pipe-bound, 34 ops/pair)
        try:
            kx, ky = torch.rand(M, 16, device=dev), torch.rand(N, 16, device=dev)
            kb_, kg = torch.randn(N, 1, device=dev), torch.tensor([1.0], device=dev)
            ks_b = get_binder(KSOLVE, [], 4, "float32")
            probe = torch.zeros(2, dtype=torch.int64, device=dev)
            ks_b.clock_probe = probe
            tc_ms, o_tc = time_binder(ks_b, M, N, (kx, ky, kb_, kg))
            engine, ebound = ks_b.last_launch.get("engine"), ks_b.last_launch.get("error_bound")
            cyc, ns = (int(v) for v in probe.tolist())
            real_hz = cyc / (ns * 1e-9) if ns > 0 else None
            os.environ["KB200_CONV_TC"] = "0"
            try:
                cc_ms, o_cc = time_binder(ks_b, M, N, (kx, ky, kb_, kg), reps=1)
            finally:
                os.environ.pop("KB200_CONV_TC", None)
            mufu_peak = 16.0 * SM_COUNT
            line["ksolve_matvec"] = {
                "workload": "Gaussian kernel matvec D=16 M=N=%d fp32 (one KernelSolve CG iteration)" % M,
                "value": M * N / (tc_ms * 1e-3), "unit": "pair-evals/s", "ms": tc_ms, "engine": engine,
                "cuda_core_ms": cc_ms, "speedup_vs_cuda_core": cc_ms / tc_ms,
                "max_rel_diff_vs_cuda_core": float((o_tc - o_cc).abs().max() / o_cc.abs().max()),
                "error_bound": ebound,
                "roofline": {"bound": "sfu", "note": "1 MUFU.EX2 per pair at 16/clk/SM (measured 15.9, experiments/mufu_peak.py)",
                             "frac_at_max_clock": M * N / (tc_ms * 1e-3) / (mufu_peak * sm_hz),
                             "sm_clock_hz_measured_in_kernel": real_hz,
                             "frac_at_measured_clock": (M * N / (tc_ms * 1e-3) / (mufu_peak * real_hz)) if real_hz else None},
                "cuda_core_frac_of_fp32_roofline": M * N / (cc_ms * 1e-3) / (128.0 / 34.0 * SM_COUNT * sm_hz)}
            del kx, ky, kb_, o_tc, o_cc
        except Exception as e:
            line["ksolve_matvec"] = {"unavailable": str(e)[:300]}
        # ---- the reference's own CUDA kernel on the same GPU ---------------------------------------
        try:
            from oracle import ref_runner
            if ref_runner.available():
                rc = ref_runner.RefCuda("gauss_fwd_f32", local_rank)
                o = rc([x, y, b], M, N)
                torch.cuda.synchronize()
                best = None
                for _ in range(2):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc([x, y, b], M, N, out=o)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    best = ms if best is None else min(best, ms)
                err = float((o - fwd).abs().max() / fwd.abs().max())
                line["ref_cuda"] = {"value": M * N / (best * 1e-3), "unit": "pair-evals/s", "ms": best,
                                    "kernel": "reference GpuConv1DOnDevice (GpuReduc1D, block 192, -use_fast_math), same GPU",
                                    "max_rel_diff_vs_ours": err, "speedup_ours": best / fwd_ms}
        except Exception as e:  # comparator only
            line["ref_cuda"] = {"unavailable": str(e)[:200]}
        try:
            cb = reference_cpu_rate(12.0)
            if cb is not None:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline"] = {"unavailable": str(e)[:200]}
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
