"""Tensor-core path for K-min reductions of a squared Euclidean distance in large dimension.

``KMin / ArgKMin / KMin_ArgKMin_Reduction(Sum(Square(x_i - y_j)), K, tagI)`` with ``D >= 16`` is what the
reference runs through the generic ``GpuReduc1D`` kernel at 2 D + 1 FP32 instructions per pair
(``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:70-90``, reduction
``keopscore/keopscore/formulas/reductions/KMin_ArgKMin_Reduction.py:16-100``).  Here the contraction
``x_i . y_j`` is a tcgen05 GEMM used as a conservative *filter*; survivors are re-evaluated exactly, so
the output is bit-identical to the CUDA-core kernel (``keops_b200/csrc/kb200_knn_tc.cu`` has the bound).

This module recognises the pattern, JIT-compiles the three kernels for (D, K, output kind) and drives
them: ``prep`` (bf16 copies + penalties) -> ``filter`` (TMA + tcgen05 + TMEM) -> ``select`` (exact top-K);
rows whose candidate list overflowed are recomputed by the caller with the CUDA-core kernel.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

from ..formulas.reductions import Reduction
from ..runtime import capi, plan

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
_SRC_PATH = os.path.join(_CSRC, "kb200_knn_tc.cu")
_COMMON_PATH = os.path.join(_CSRC, "kb200_tc_common.cuh")      # PTX wrappers, prepended (NVRTC has no include path)
VERSION = "knn-tc-17"
BM, BN, STAGES, HS, KP, THREADS = 256, 128, 4, 8, 16, 576
MAX_SMEM = 227 * 1024
RESIDENT_MAX_COLS = 160      # D + 3 <= 160: the CTA's row operand stays in shared memory (>= 3 ring stages); beyond, k-blocks are streamed
MAX_D = 1021                 # operand rows of at most 1024 bf16 columns
ROW_BATCH = 1 << 20          # rows per filter launch: bounds the candidate workspace (CAP * 4 bytes per row)
TAIL_MAX_WAVES = int(os.environ.get("KB200_KNN_TC_TAIL_WAVES", 16))   # split the last partial wave only below this many waves of row blocks
_lock = threading.RLock()


def match(red: Reduction) -> Optional[Tuple[int, int, int]]:
    """``(row_var_index, col_var_index, D)`` if the reduction is a K-min of ``Sum(Square(Var - Var))`` between
    one output-side and one streamed variable of equal dimension; ``None`` otherwise."""
    if red.kind not in ("KMin", "ArgKMin", "KMin_ArgKMin") or red.dimF != 1:
        return None
    f = red.formula
    if f.op != "Sum" or len(f.args) != 1 or f.args[0].op != "Square":
        return None
    sub = f.args[0].args[0]
    if sub.op != "Subtract" or len(sub.args) != 2 or any(a.op != "Var" for a in sub.args):
        return None
    (ia, da, ca), (ib, db, cb) = sub.args[0].params, sub.args[1].params
    if da != db or {ca, cb} != {0, 1}:
        return None
    row, col = (ia, ib) if ca == red.tagI else (ib, ia)
    return row, col, da


def eligible(red: Reduction, dtype: str, nrow: int, ncol: int, nbatch: int) -> Optional[Tuple[int, int, int]]:
    """The tensor-core path pays off once the GEMM dominates: D >= 16 and a full machine of 256-row CTAs."""
    if os.environ.get("KB200_KNN_TC", "1") == "0" or dtype != "float32" or nbatch != 1:
        return None
    m = match(red)
    if m is None:
        return None
    D = m[2]
    # few rows against many columns (the reference's own KNN benchmark shape, 1e4 queries vs 1e6 points): the filter
    # splits the column range over gridDim.y, so the tensor-core path already pays off from ~1000 rows
    min_rows = 8192 if ncol < 65536 else 1024
    if D < 16 or D > MAX_D or red.K > 16 or nrow < min_rows or ncol < 4096 or ncol >= (1 << 27):
        return None                      # candidate tokens carry the 8-column group index in 24 bits
    return m


class KnnTensorCore:
    """Compiled kernels for one (D, K, output kind, fast-math) combination on one device."""

    def __init__(self, D: int, K: int, kind: str, fast: bool, device: int, cap: int = 320):
        self.D, self.K, self.kind, self.fast, self.cap = D, K, kind, fast, cap
        # operand rows: D data columns + 3 penalty columns, as KBM 64-column blocks + one remainder block
        total = D + 3
        self.stream = total > RESIDENT_MAX_COLS
        full, r = divmod(total, 64)
        if self.stream:                          # large D: whole 64-column k-blocks, streamed through the ring
            self.KBM, self.REM = full + (1 if r else 0), 0
        elif r == 0:
            self.KBM, self.REM = full - 1, 64
        else:
            self.KBM, self.REM = full, (16 if r <= 16 else (32 if r <= 32 else 64))
        self.DPG = 64 * self.KBM + self.REM
        self.outw = 2 * K if kind == "KMin_ArgKMin" else K
        out_code = {"ArgKMin": 0, "KMin": 1, "KMin_ArgKMin": 2}[kind]
        # filter margin c: two bf16 roundings per product (2 u + u^2, u = 2^-8 -> 0.0078278) + fp32 accumulation over the
        # K columns and the rounding of the exact path (each below D 2^-23 of sum |x_k y_k|)
        self.c = 0.008 if D <= 160 else 0.00783 + D * 2.0 ** -21
        self.BN = BN
        blk = lambda rows: self.KBM * rows * 128 + rows * self.REM * 2      # one operand block, all of K
        state = BM * (KP + 3) * 4                # per-row lists, threshold, token count, lock
        fixed = 1024 + HS * BN * 4 + state + 640
        if self.stream:
            a_bytes, b_bytes = 0, (BM + BN) * 128
        else:
            a_bytes, b_bytes = 2 * blk(128), blk(BN)
        self.STAGES = int(os.environ.get("KT_STAGES", min(STAGES, (MAX_SMEM - fixed - a_bytes) // b_bytes)))
        assert self.STAGES >= 2, (D, self.STAGES)
        with open(_COMMON_PATH) as f:
            body = f.read()
        with open(_SRC_PATH) as f:
            body += f.read()
        defs = ("#define KT_D %d\n#define KT_KBM %d\n#define KT_REM %d\n#define KT_K %d\n#define KT_CAP %d\n"
                "#define KT_OUT %d\n#define KT_FAST %d\n" % (D, self.KBM, self.REM, K, cap, out_code, 1 if fast else 0))
        defs += "#define KT_STREAM %d\n#define KT_STAGES %d\n#define KT_C %.9gf\n" % (int(self.stream), self.STAGES, self.c)
        defs += "#define KT_DBG %d\n" % int(os.environ.get("KT_DBG", 0))
        defs += "#define KT_SPIN %d\n" % int(os.environ.get("KT_SPIN", 0))
        src = defs + body
        self.cubin = cubin = plan.compile_cached(src, plan.nvrtc_options(fast), "kb200_knntc", VERSION)
        self.smem = fixed + a_bytes + self.STAGES * b_bytes
        self.smem = max(self.smem, 120 * 1024)     # one CTA per SM: each CTA allocates all 512 TMEM columns
        self.module = None
        if device >= 0:
            self.module = capi.Module(cubin, device)
            self.fn_prep = self.module.function("kb200_knn_prep")
            self.fn_filter = self.module.function("kb200_knn_filter")
            self.fn_select = self.module.function("kb200_knn_select")
            self.fn_merge = self.module.function("kb200_knn_merge")
            self.sm_count = capi.device_info(device)["sm_count"]
            capi.set_max_dynamic_smem(self.fn_filter, self.smem)

    def _maps(self, abf, n, box_rows):
        """TMA descriptors of a bf16 operand array (n, DPG): the 64-column main blocks (128-byte swizzle) and the
        remainder block (swizzled at its own width), with boxes of ``box_rows`` rows."""
        pitch = self.DPG * 2
        if self.REM == 0:              # streamed k-blocks: no remainder block
            main = capi.tensor_map_2d(abf.data_ptr(), 2, 64 * self.KBM, n, pitch, 64, box_rows, 128)
            return main, main
        rem = capi.tensor_map_2d(abf.data_ptr() + 128 * self.KBM, 2, self.REM, n, pitch, self.REM, box_rows, self.REM * 2)
        if self.KBM == 0:
            return rem, rem            # the main map is never dereferenced
        return capi.tensor_map_2d(abf.data_ptr(), 2, 64 * self.KBM, n, pitch, 64, box_rows, 128), rem

    def run(self, x, y, out, stream_obj):
        """x (M, D), y (N, D) fp32 contiguous CUDA tensors; out (M, outw) fp32.  Returns the tensor of row
        indices whose candidate list overflowed (usually empty) -- the caller recomputes those rows."""
        import torch
        M, N = x.shape[0], y.shape[0]
        dev = x.device
        stream = stream_obj.cuda_stream
        i64, ptr, i32 = C.c_longlong, C.c_void_p, C.c_int
        BN = self.BN
        ntiles = -(-N // BN)
        ybf = torch.empty((N, self.DPG), dtype=torch.int16, device=dev)
        ypen = torch.empty(ntiles * BN, dtype=torch.float32, device=dev)    # padded: the last tile is bulk-copied whole
        capi.launch(self.fn_prep, stream, (-(-N // 256),), (256,), 0,
                    [i64(N), ptr(y.data_ptr()), ptr(ybf.data_ptr()), ptr(ypen.data_ptr()), i32(0)])
        map_y, map_yr = self._maps(ybf, N, BN)
        # Row segments (first row, rows, column splits).  With fewer 256-row blocks than SMs the column range is divided over
        # gridDim.y; with many, the LAST partial wave of row blocks is split so that it also fills the machine: 489 row blocks
        # (one GPU's shard of an 8-GPU run at M = 1e6) are 3 full waves + 45 blocks -- unsplit that is 4 waves, with the tail
        # split 3-way 3.33 (the splits of a row keep their own candidate lists and are merged, so only the tail's rows pay the
        # restart of an empty list).
        def small_split(nb):
            return max(1, min(self.sm_count // nb, ntiles // 64, 32))
        forced = os.environ.get("KB200_KNN_TC_NSPLIT")      # experiments: forced column chunks (L2-sized y slabs)
        segments = []
        for r0 in range(0, M, ROW_BATCH):
            m = min(ROW_BATCH, M - r0)
            nb = -(-m // BM)
            if forced:
                segments.append((r0, m, max(1, min(int(forced), ntiles))))
            elif nb < self.sm_count:
                segments.append((r0, m, small_split(nb) if os.environ.get("KB200_KNN_TC_SPLIT", "1") != "0" else 1))
            else:
                full = (nb // self.sm_count) * self.sm_count
                rem = nb - full
                ns_tail = min(small_split(rem), 4) if rem else 1
                # (measured, experiments/knn_d64_shards.py, D = 64, N = 1e6, median of 7: 125 000 rows = 3.3 waves 22.97 -> 20.62 ms;
                # 500 000 rows = 13.2 waves 83.1 -> 80.9 ms; 1e6 rows = 26.4 waves 162.3 -> 161.7 ms: not worth a second launch)
                if (rem and ns_tail >= 2 and nb < TAIL_MAX_WAVES * self.sm_count
                        and os.environ.get("KB200_KNN_TC_TAIL", "1") != "0"):
                    segments += [(r0, full * BM, 1), (r0 + full * BM, m - full * BM, ns_tail)]
                else:
                    segments.append((r0, m, 1))
        max_split = max(sg[2] for sg in segments)
        overflow_count = torch.zeros(1, dtype=torch.int32, device=dev)
        overflow_rows = torch.empty(max(M, 1) * max_split, dtype=torch.int32, device=dev)
        ctas = 0
        for r0, m, nsplit in segments:
            tiles_per_split = -(-ntiles // nsplit)
            nsplit = -(-ntiles // tiles_per_split)
            ctas += -(-m // BM) * nsplit
            xs = x[r0:r0 + m]
            xbf = torch.empty((m, self.DPG), dtype=torch.int16, device=dev)
            xpen = torch.empty(m, dtype=torch.float32, device=dev)
            capi.launch(self.fn_prep, stream, (-(-m // 256),), (256,), 0,
                        [i64(m), ptr(xs.data_ptr()), ptr(xbf.data_ptr()), ptr(xpen.data_ptr()), i32(1)])
            map_x, map_xr = self._maps(xbf, m, 128)
            cand = torch.empty(m * nsplit * self.cap, dtype=torch.int32, device=dev)
            count = torch.empty(m * nsplit, dtype=torch.int32, device=dev)
            part = torch.empty(m * nsplit * self.K * 2 if nsplit > 1 else 1, dtype=torch.float32, device=dev)
            if os.environ.get("KT_TIME"):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream_obj)
            capi.launch(self.fn_filter, stream, (-(-m // BM), nsplit), (THREADS,), self.smem,
                        [map_x, map_xr, map_y, map_yr, i64(m), i64(N), ptr(ypen.data_ptr()), ptr(xpen.data_ptr()),
                         ptr(cand.data_ptr()), ptr(count.data_ptr()), i32(tiles_per_split)])
            if os.environ.get("KT_TIME"):
                ev1.record(stream_obj)
                torch.cuda.synchronize()
                print("KT_TIME filter %.3f ms (rows %d..%d STAGES=%d stream=%d nsplit=%d)" % (ev0.elapsed_time(ev1), r0, r0 + m, self.STAGES, int(self.stream), nsplit))
            # select writes rows r0.. of `out` (or per-split lists); overflowed rows are reported with their global index
            out_ptr = out.data_ptr() + r0 * self.outw * 4
            capi.launch(self.fn_select, stream, (-(-(m * nsplit) // 8),), (256,), 0,
                        [i64(m), i64(N), ptr(xs.data_ptr()), ptr(y.data_ptr()), ptr(cand.data_ptr()), ptr(count.data_ptr()),
                         ptr(out_ptr), ptr(overflow_count.data_ptr()), ptr(overflow_rows.data_ptr()), i32(nsplit),
                         ptr(part.data_ptr()), i64(r0)])
            if nsplit > 1:
                capi.launch(self.fn_merge, stream, (-(-m // 256),), (256,), 0,
                            [i64(m), i32(nsplit), ptr(part.data_ptr()), ptr(out_ptr)])
            if os.environ.get("KT_STATS"):
                cf = count.view(m, nsplit).sum(1).float()
                self.stats = dict(mean_candidates=float(cf.mean()), max_candidates=float(cf.max()),
                                  p99=float(cf.kthvalue(max(1, int(0.99 * m))).values))
            for t in (xbf, xpen, cand, count, part):
                t.record_stream(stream_obj)
        for t in (ybf, ypen):
            t.record_stream(stream_obj)
        n_over = int(overflow_count.item())          # one host sync per call (the reference syncs on every call)
        self.last = dict(overflow=n_over, ntiles=ntiles, ctas=ctas, nsplit=max_split,
                         segments=[(a_, a_ + m_, n_) for a_, m_, n_ in segments])
        rows = overflow_rows[:n_over].to(torch.int64)
        return torch.unique(rows) if max_split > 1 and n_over else rows      # a row may overflow in several splits


_cache: Dict[tuple, KnnTensorCore] = {}


def get(D: int, K: int, kind: str, fast: bool, device: int) -> KnnTensorCore:
    key = (D, K, kind, bool(fast), device)
    with _lock:
        k = _cache.get(key)
        if k is None:
            # candidate tokens per row: ~ K ln(N / K) list updates + the filter's false positives (measured at K = 10, N = 1e6:
            # mean 153, maximum 238 of 320 slots); longer lists get proportionally more slots.  Rows that overflow are
            # recomputed by the CUDA-core kernel -- correct, but slow.
            k = KnnTensorCore(D, K, kind, fast, device, cap=320 if K <= 10 else 512)
            _cache[key] = k
        return k
