"""Compilation cache, kernel configuration heuristics and launch-shape tuner.

Replaces ``keopscore/keopscore/binders/LinkCompile.py`` (hash -> artefact on disk),
``keopscore/keopscore/utils/Cache.py`` (memoisation) and the block-size logic of
``KeOps_module::launch_kernel`` (``keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:441-458``).

Cache key = sha256(generator version + full generated source + NVRTC options): the source text
already contains the tile configuration, dtype, fast-math flag and the canonical reduction.
"""
from __future__ import annotations

import hashlib
import math
import os
import threading
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from ..codegen.emit import KernelConfig, KernelSource, GENERATOR_VERSION
from ..formulas.reductions import Reduction
from . import capi

_lock = threading.RLock()


def cache_dir() -> str:
    """Cubin cache: ``$KB200_CACHE_DIR``, else an in-tree folder (so that cubins cross-compiled by
    ``build()`` travel with the repository snapshot), else ``~/.cache/keops_b200``."""
    d = os.environ.get("KB200_CACHE_DIR")
    if not d:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_cache", GENERATOR_VERSION)
        try:
            os.makedirs(d, exist_ok=True)
            if not os.access(d, os.W_OK):
                raise OSError
        except OSError:
            d = os.path.join(os.path.expanduser("~"), ".cache", "keops_b200", GENERATOR_VERSION)
    os.makedirs(d, exist_ok=True)
    return d


def clean_cache():
    """Counterpart of ``pykeops.clean_pykeops()`` (``pykeops/pykeops/__init__.py:47-61``)."""
    with _lock:
        d = cache_dir()
        for f in os.listdir(d):
            if f.endswith(".cubin") or f.endswith(".cu"):
                try:
                    os.remove(os.path.join(d, f))
                except OSError:
                    pass
        for m in _modules.values():
            m.module.unload()
        _modules.clear()
        _cubins.clear()


_cubins: Dict[str, bytes] = {}


def nvrtc_options(fast: bool) -> List[str]:
    opts = ["-lineinfo"]
    if fast:
        opts.append("--use_fast_math")
    extra = os.environ.get("KB200_NVRTC_FLAGS")
    if extra:
        opts.extend(extra.split())
    return opts


def compile_cached(source: str, options: Sequence[str], prefix: str, version: str) -> bytes:
    """NVRTC -> sm_100a cubin for a hand-written kernel source, cached on disk by content hash (same in-tree cache
    as the generated kernels, so that cubins cross-compiled by ``build()`` travel with the repository snapshot)."""
    key = hashlib.sha256((version + source + " ".join(options)).encode()).hexdigest()[:20]
    path = os.path.join(cache_dir(), "%s_%s.cubin" % (prefix, key))
    with _lock:
        if os.path.exists(path):
            with open(path, "rb") as f:
                return f.read()
        cubin = capi.compile_to_cubin(source, list(options), name="%s_%s.cu" % (prefix, key))
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(cubin)
        os.replace(tmp, path)
        return cubin


def get_cubin(ks: KernelSource, fast: bool) -> bytes:
    opts = nvrtc_options(fast)
    key = hashlib.sha256((ks.hash + "|" + " ".join(opts)).encode()).hexdigest()[:20]
    with _lock:
        blob = _cubins.get(key)
        if blob is not None:
            return blob
        path = os.path.join(cache_dir(), "kb200_%s.cubin" % key)
        if os.path.exists(path):
            with open(path, "rb") as f:
                blob = f.read()
        else:
            blob = capi.compile_to_cubin(ks.source, opts, name="kb200_%s.cu" % key)
            tmp = path + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(blob)
            os.replace(tmp, path)
            if os.environ.get("KB200_KEEP_SOURCE"):
                with open(path[:-6] + ".cu", "w") as f:
                    f.write(ks.source)
        _cubins[key] = blob
        return blob


# ------------------------------------------------------------------------------------------
# configuration heuristics
# ------------------------------------------------------------------------------------------
def default_config(red: Reduction, dtype: str) -> KernelConfig:
    """Pick (T, R, TJ, unroll) from the formula's register / shared-memory footprint."""
    tagI = red.tagI
    dimx = sum(v.params[1] for v in red.variables() if v.params[2] == tagI)
    dimy = sum(v.params[1] for v in red.variables() if v.params[2] == 1 - tagI)
    f64 = dtype == "float64"
    env = os.environ.get("KB200_CONFIG")  # e.g. "T=128,R=4,TJ=256,U=8,B=2" (debug / tuning)
    acc_regs = red.dimred if red.kind not in ("Sum", "LSE") else 2 * red.dim
    work = dimx + acc_regs
    if f64:
        R = 2 if work <= 24 else 1
        T = 128
        min_blocks = 1
    else:
        kfam = red.kind in ("KMin", "ArgKMin", "KMin_ArgKMin")
        if kfam and dimx <= 8 and red.K * red.dimF <= 16:
            # measured on B200 (ArgKMin K=10, D=3, M=N=1e6, group=4): (T=256, R=4, 2 CTAs/SM, 128 regs) 232.9 ms;
            # (T=128, R=4, 4 CTAs/SM) 239.2 ms; without the register cap (148 regs, 3 CTAs/SM) 252.5 ms
            R, T, min_blocks = 4, 256, 2
        elif work <= 12:
            R, T, min_blocks = 4, 256, 4
        elif work <= 40:
            R, T, min_blocks = 2, 256, 2
        else:
            R, T, min_blocks = 2, 128, 1
    esz = 8 if f64 else 4
    vec = 2 if f64 else 4
    stride = max(vec, -(-dimy // vec) * vec)
    TJ = 16384 // (stride * esz)          # two buffers <= 32 KB
    # longer tiles amortise the per-tile barrier / bookkeeping; measured on B200 at M = N = 1e6 (experiments/tj_sweep.sh),
    # TJ = 256 / 512 / 1024: ArgKMin D=3 232.1 / 228.5 / 222.7 ms, LSE 350.1 / 349.3 / 342.8 ms, Gaussian conv 262.6 / 259.0 / 261.4 ms
    tj_cap = 256 if f64 else (512 if red.kind == "Sum" else 1024)     # fp64 LSE: 1604 ms at 256, 1615 ms at 512
    TJ = max(8, min(tj_cap, 1 << int(math.log2(max(TJ, 1)))))
    while TJ > 1 and 2 * TJ * stride * esz > 200 * 1024:      # very wide streamed rows: keep the double buffer inside one SM
        TJ //= 2
    body = max(1, dimx + dimy) * (R // (1 if f64 else 2))
    unroll = 8 if body <= 16 else (4 if body <= 64 else (2 if body <= 160 else 1))
    if red.kind in ("KMin", "ArgKMin", "KMin_ArgKMin"):
        unroll = min(unroll, 4)
    group = 1
    if red.kind not in ("Sum", "LSE") and not f64:
        # values held per group: group * dimF * R registers; keep the branch-free window small
        nval = max(1, red.dimF) * R
        if dimx + dimy <= 16:
            group = max(1, min(4, 16 // nval))
    cfg = KernelConfig(T=T, R=R, TJ=TJ, unroll=unroll, min_blocks=min_blocks, group=group)
    if env:
        kv = dict(item.split("=") for item in env.split(","))
        cfg = KernelConfig(T=int(kv.get("T", cfg.T)), R=int(kv.get("R", cfg.R)), TJ=int(kv.get("TJ", cfg.TJ)),
                           unroll=int(kv.get("U", cfg.unroll)), min_blocks=int(kv.get("B", cfg.min_blocks)),
                           group=int(kv.get("G", cfg.group)))
    return cfg


# ------------------------------------------------------------------------------------------
# loaded kernels + launch shape
# ------------------------------------------------------------------------------------------
@dataclass
class LoadedKernel:
    ks: KernelSource
    module: capi.Module
    fn_reduce: int
    fn_combine: int
    regs: int
    blocks_per_sm: int
    sm_count: int
    fn_seeded: int = 0        # K-min kernels: entry point whose column splits share per-row thresholds (kb200_reduce_seeded)


_modules: Dict[Tuple[str, int], LoadedKernel] = {}


def load_kernel(ks: KernelSource, fast: bool, device: int) -> LoadedKernel:
    key = (ks.hash + ("f" if fast else "p"), device)
    with _lock:
        lk = _modules.get(key)
        if lk is None:
            cubin = get_cubin(ks, fast)
            mod = capi.Module(cubin, device)
            fr = mod.function(ks.kernel_name)
            fc = mod.function(ks.combine_name) if ks.combine_name else 0
            regs, _, _ = capi.function_info(fr)
            if ks.dynamic_smem:
                capi.set_max_dynamic_smem(fr, ks.dynamic_smem)
            bps = max(1, capi.max_active_blocks(fr, ks.cfg.T, ks.dynamic_smem))
            sm = capi.device_info(device)["sm_count"]
            fs = 0
            if ks.kernel_name == "kb200_reduce" and "kb200_reduce_seeded" in ks.source:
                fs = mod.function("kb200_reduce_seeded")
                if ks.dynamic_smem:
                    capi.set_max_dynamic_smem(fs, ks.dynamic_smem)
            lk = LoadedKernel(ks, mod, fr, fc, regs, bps, sm, fs)
            _modules[key] = lk
        return lk


def _sm_rounds(nblocks: int, slots: int, sms: int) -> float:
    """Duration of a grid of ``nblocks`` equal CTAs in units of one CTA running alone at full occupancy.

    Measured on B200 (experiments/shard_sweep.py, 125 000-row shard of the Gaussian conv, 4 CTAs / SM): co-resident CTAs
    share their SM's pipes, so the time follows the work of the most loaded SM -- ``ceil(nblocks / sms)`` CTAs -- and not
    whole waves of ``slots`` CTAs in lockstep: 1476 CTAs (9.97 per SM) 29.3 ms, 1722 (11.6 -> 12) 30.0 ms,
    2091 (14.1 -> 15) 30.9 ms, 2952 (19.9 -> 20) 29.2 ms.  With ``sms = 0`` the older whole-wave model is used."""
    if sms <= 0:
        return float(-(-nblocks // slots))
    bps = max(1, slots // sms)
    per_sm = -(-nblocks // sms)
    # a CTA that is alone on its SM runs ~3.7x faster than with three neighbours (same sweep: 5 CTAs per SM at 4 resident
    # cost 1.27 rounds, not 2): the time is proportional to the CTAs of the most loaded SM, whatever the residency
    return per_sm / float(bps) * (1.0 if per_sm % bps == 0 else 1.02)


def launch_shape(nrow: int, ncol: int, rows_per_block: int, TJ: int, slots: int, dimpart: int, esz: int,
                 max_split: int = 64, ws_limit_bytes: int = 2 << 30, split_penalty_tiles: float = 0.0,
                 sms: int = 0) -> Tuple[int, int]:
    """Choose the number of column splits so that every SM gets the same number of CTAs.

    All CTAs do the same amount of work, so wall time ~ rounds * work-per-CTA (``_sm_rounds``).  Without splitting,
    ``gx`` CTAs over ``slots`` resident slots leave SMs idle at the end (M=1e6 rows -> 1.65 waves -> 2).  Splitting the
    streamed range into ``ns`` chunks multiplies the CTA count and evens the load out; partials are merged by
    ``kb200_combine``.  Returns (nsplit, col_chunk)."""
    gx = -(-nrow // rows_per_block)
    ntiles = max(1, -(-ncol // TJ))
    forced = os.environ.get("KB200_FORCE_NSPLIT")          # tuning experiments (experiments/shard_sweep.py)
    if forced:
        tiles_per = -(-ntiles // max(1, min(int(forced), ntiles)))
        return -(-ntiles // tiles_per), tiles_per * TJ
    best = None
    for ns in range(1, max_split + 1):
        if ns > ntiles:
            break
        tiles_per = -(-ntiles // ns)
        ns_eff = -(-ntiles // tiles_per)
        if ns_eff != ns:
            continue
        if ns > 1 and ns * nrow * dimpart * esz > ws_limit_bytes:
            break
        rounds = _sm_rounds(gx * ns, slots, sms)
        # cost in tile units: per-CTA prologue/epilogue overhead ~ 512 streamed rows, combine ~ 64 rows per split
        cost = rounds * (tiles_per + 512.0 / TJ + (split_penalty_tiles if ns > 1 else 0.0)) + (64.0 / TJ * ns if ns > 1 else 0.0)
        if sms > 0:
            cost += 0.03 * tiles_per       # SMs do not finish together: a few % of a CTA's duration (same sweep: ns = 6 vs 24)
        if best is None or cost < best[0] * (0.995 if sms > 0 else 0.98):
            best = (cost, ns, tiles_per * TJ)
    return best[1], best[2]


def launch_segments(nrow: int, ncol: int, rows_per_block: int, TJ: int, slots: int, dimpart: int, esz: int,
                    split_penalty_tiles: float = 0.0, ws_limit_bytes: int = 2 << 30, sms: int = 0,
                    prefer_uniform: bool = False):
    """Row segments ``[(row_begin, row_end, nsplit, col_chunk), ...]`` covering ``[0, nrow)``.

    Splitting the column range of *every* CTA fixes wave quantisation but makes each split restart its
    accumulator (for K-min lists every restart replays the ~K log(N/K) insertions of the transient).
    When the grid holds at least one full wave it is cheaper to launch the whole waves unsplit and to
    split only the rows of the last, partial wave so that they also fill the machine:
    time ~ full_waves * T + (T / ns) instead of (full_waves + 1) * T or full-split with restart costs.
    ``split_penalty_tiles``: extra cost of one restart, in tile units (0 for sums)."""
    gx = -(-nrow // rows_per_block)
    ntiles = max(1, -(-ncol // TJ))
    ns_all, chunk_all = launch_shape(nrow, ncol, rows_per_block, TJ, slots, dimpart, esz, ws_limit_bytes=ws_limit_bytes,
                                     split_penalty_tiles=split_penalty_tiles, sms=sms)
    tiles_all = -(-ntiles // ns_all)
    ovh, cmb = 512.0 / TJ, 64.0 / TJ
    cost_all = _sm_rounds(gx * ns_all, slots, sms) * (tiles_all + ovh + (split_penalty_tiles if ns_all > 1 else 0.0)) + (cmb * ns_all if ns_all > 1 else 0.0)
    full_waves, rem = divmod(gx, slots)
    if full_waves == 0 or rem == 0:
        return [(0, nrow, ns_all, chunk_all)]
    rows_a = full_waves * slots * rows_per_block
    nrow_b = nrow - rows_a
    ns_b, chunk_b = launch_shape(nrow_b, ncol, rows_per_block, TJ, slots, dimpart, esz, ws_limit_bytes=ws_limit_bytes,
                                 split_penalty_tiles=split_penalty_tiles, sms=sms)
    tiles_b = -(-ntiles // ns_b)
    cost_tail = full_waves * (ntiles + ovh) + _sm_rounds(rem * ns_b, slots, sms) * (tiles_b + ovh + (split_penalty_tiles if ns_b > 1 else 0.0)) \
        + (cmb * ns_b if ns_b > 1 else 0.0) + ovh / 2
    # ``prefer_uniform``: the uniform split wins ties and near-ties: its short CTAs backfill and even out per-SM speed differences.  Measured with the
    # filtered K-min kernel, whose splits share thresholds (experiments/shard_sweep.py, M x 1e6, D = 3, K = 10): M = 1e6 158.7 ms
    # (3 uniform splits) vs 159.1 (4 unsplit waves + split tail); 5e5: 80.6 vs 83.1; 2.5e5: 41.6 vs 44.0
    if cost_tail < (0.95 if prefer_uniform else 1.0) * cost_all:
        return [(0, rows_a, 1, ntiles * TJ), (rows_a, nrow, ns_b, chunk_b)]
    return [(0, nrow, ns_all, chunk_all)]
