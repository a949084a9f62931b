"""The drop-in seam: a binder with the call signature of the reference's ``LoadKeOps`` objects.

``pykeops.common.keops_io.keops_binder["nvrtc"]`` is looked up at call time by the reference's
``GenredAutograd_base._forward`` (``pykeops/pykeops/torch/generic/generic_red.py:89-129``) and
``KernelSolveAutograd.forward`` (``pykeops/pykeops/torch/operations.py:45-81``); an object with

* constructor ``B(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, formula, aliases, nargs,
  dtype, lang, optional_flags)`` (argument order of ``LoadKeOps.init``,
  ``pykeops/pykeops/common/keops_io/LoadKeOps.py:33-46``),
* ``.import_module() -> self`` and ``.genred_pytorch(device_args, ranges, nx, ny, nbatchdims, out,
  *args) -> Tensor`` (``LoadKeOps.py:177-186,245-252``),
* attributes ``.dimout`` and ``.tagIJ`` (``LoadKeOps.py:19-20``),

is all they need.  :class:`LoadKeOpsB200` is that object, backed by the sm_100a code generator and
``libkb200.so``; ``register_into_pykeops()`` installs it into an importable reference ``pykeops``.
There is deliberately no CPU path: ``tagCPUGPU == 0`` raises.
"""
from __future__ import annotations

import os
import threading
import types
from typing import Dict, List, Optional, Sequence

import torch

from .formulas.parser import parse_reduction, parse_alias
from .formulas.reductions import Reduction
from .codegen.emit import KernelConfig, generate, KernelSource
from .codegen import conv_tc, knn_tc, emit_chunked, emit_kmin_filter
from .runtime import plan, capi

_binder_cache: Dict[tuple, "LoadKeOpsB200"] = {}
_cache_lock = threading.RLock()


def _normalise_aliases(aliases: Sequence[str]) -> List[str]:
    """``"x = Vi(3)"`` -> ``"x=Var(k,3,0)"`` with k the list position (``LoadKeOps.py:47-64``)."""
    out = []
    for k, alias in enumerate(aliases):
        if alias is None:
            continue
        name, var = parse_alias(alias, k)
        ind, dim, cat = var.params
        out.append(("%s=" % name if name else "") + "Var(%d,%d,%d)" % (ind, dim, cat))
    return out


class LoadKeOpsB200:
    """Formula -> compiled sm_100a kernels -> asynchronous launch on the current torch stream."""

    def __new__(cls, tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id_request, formula, aliases,
                nargs, dtype, lang, optional_flags):
        key = (tagCPUGPU, tag1D2D, tagHostDevice, bool(use_ranges), device_id_request, formula,
               tuple(aliases), nargs, dtype, lang, tuple(sorted(optional_flags.items())))
        with _cache_lock:
            self = _binder_cache.get(key)
            if self is None:
                self = object.__new__(cls)
                self._init(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id_request, formula,
                           aliases, nargs, dtype, lang, optional_flags)
                _binder_cache[key] = self
            return self

    def _init(self, tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id_request, formula, aliases,
              nargs, dtype, lang, optional_flags):
        if tagCPUGPU == 0:
            raise ValueError("[KeOps-B200] this engine has no CPU backend (backend='CPU' is not available).")
        if lang not in ("torch", "pytorch", "numpy"):
            raise ValueError("[KeOps-B200] unknown binding %r (torch or numpy expected)." % (lang,))
        if dtype not in ("float32", "float64"):
            raise ValueError("[KeOps-B200] dtype %r is not implemented (float32 and float64 only; no half2 kernels)." % (dtype,))
        p = types.SimpleNamespace()
        p.aliases_old = list(aliases)
        p.aliases = _normalise_aliases(aliases)
        p.red_formula_string = formula
        p.dtype = dtype
        p.lang = lang
        p.nargs = nargs
        p.tag1D2D = tag1D2D
        p.tagCPUGPU, p.tagHostDevice = tagCPUGPU, tagHostDevice
        p.use_ranges = bool(use_ranges)
        p.device_id_request = device_id_request
        p.sum_scheme = optional_flags.get("sum_scheme", "auto")
        p.dtype_acc = optional_flags.get("dtype_acc", "auto")
        p.use_fast_math = bool(optional_flags.get("use_fast_math", 1))
        p.enable_chunks = optional_flags.get("enable_chunks", 1)
        p.mult_var_highdim = optional_flags.get("multVar_highdim", 0)
        if p.dtype_acc not in ("auto", "float32", "float64", "float16"):
            raise ValueError('[KeOps] invalid value for option dtype_acc : should be one of "auto", "float16", "float32" or "float64".')
        if dtype == "float64" and p.dtype_acc not in ("auto", "float64"):
            raise ValueError("[KeOps] invalid parameter dtype_acc : should be 'float64' when dtype is 'float64'")

        red: Reduction = parse_reduction(formula, p.aliases)
        self.red = red
        if p.sum_scheme == "auto":
            p.sum_scheme = "block_sum" if red.kind in ("Sum", "LSE") else "direct_sum"
        p.tagI = red.tagI
        p.tagZero = 1 if red.is_zero else 0
        p.dim, p.dimred = red.dim, red.dimred
        p.reduction_op = formula.split("(")[0]
        p.axis = 1 - red.tagI
        cfg = plan.default_config(red, dtype)
        self.ks: KernelSource = generate(red, dtype, nargs, cfg, use_fast_math=p.use_fast_math,
                                         sum_scheme=p.sum_scheme, dtype_acc=p.dtype_acc)
        if self.ks.is_zero:
            p.tagZero = 1
        # metadata in the reference's *post-swap* convention (``LoadKeOps.py:151-153``): indsi are
        # always the cat-0 variables, whatever tagI is.
        vs = sorted(red.variables(), key=lambda v: v.params[0])
        p.indsi = tuple(v.params[0] for v in vs if v.params[2] == 0)
        p.indsj = tuple(v.params[0] for v in vs if v.params[2] == 1)
        p.indsp = tuple(v.params[0] for v in vs if v.params[2] == 2)
        p.dimsx = tuple(v.params[1] for v in vs if v.params[2] == 0)
        p.dimsy = tuple(v.params[1] for v in vs if v.params[2] == 1)
        p.dimsp = tuple(v.params[1] for v in vs if v.params[2] == 2)
        p.dimy = sum(d for _, d in self.ks.col_vars)
        self.params = p
        self.dimout = p.dim
        self.tagIJ = p.tagI
        self._vars = {v.params[0]: (v.params[1], v.params[2]) for v in vs}
        self._loaded: Dict[object, plan.LoadedKernel] = {}
        self.ks_ranges: Optional[KernelSource] = None
        self._range_tables: Dict[tuple, tuple] = {}
        self._optional_flags = dict(optional_flags)
        self._init_args = (tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id_request, lang)
        self._wide = self._detect_wide(red)
        self._wide_binders: Dict[int, "LoadKeOpsB200"] = {}
        # high-dimensional Sum / Scalprod sub-formulas: the chunked kernel (the reference's chunk mode, switched by the same
        # ``enable_chunks`` flag: get_keops_dll.py:101-111)
        self._prescale_ks = None         # None: not tried yet; False: the formula does not have the Gaussian shape
        # K nearest neighbours in dimension <= 3: filtered kernel (codegen/emit_kmin_filter.py), generated on first use
        self._kf_match = emit_kmin_filter.match(red) if (dtype == "float32" and p.use_fast_math and not self.ks.is_zero) else None
        self.ks_kfilter: Optional[KernelSource] = None
        # gradient of a Gaussian-kernel sum w.r.t. a position variable in dimension 8..64: assembled from tensor-core sums
        self._conv_grad = conv_tc.match_grad(red) if (dtype == "float32" and not self.ks.is_zero) else None
        self._chunk_plan = emit_chunked.plan_chunks(red) if (p.enable_chunks and not self.ks.is_zero) else None
        self.ks_chunked: Optional[KernelSource] = None

    # ---- wide streamed factor (the reference's "final chunks" mode, GpuReduc1D_finalchunks.py) ----------------
    # measured on B200 (D = 3 Gaussian, M = N = 5e4; experiments/wide_b_sweep.py), single kernel vs chunks of 32:
    # dim b = 40: 4.7 / 5.0 ms, 64: 12.0 / 7.3 ms, 96: 18.8 / 11.0 ms, 300: 470 / 34 ms (chunks of 16: 51 ms, of 64: 53 ms)
    WIDE_MIN, WIDE_CHUNK = 64, 32

    @staticmethod
    def _detect_wide(red: Reduction):
        """``Sum_Reduction(E_ij * b, tagI)`` with ``E`` scalar and ``b`` a *streamed* variable of dimension >= 64 (attention-
        like sums): the accumulator (dim b values per output row) does not fit the register file.  Returns
        ``(b index, dim b, b category, text of E)`` when the reduction can be evaluated chunk by chunk of ``b``."""
        if red.kind != "Sum":
            return None
        f = red.formula
        if f.op != "Mult" or len(f.args) != 2:
            return None
        for e, b in ((f.args[0], f.args[1]), (f.args[1], f.args[0])):
            if b.op == "Var" and e.dim == 1 and b.params[1] >= LoadKeOpsB200.WIDE_MIN and b.params[2] == 1 - red.tagI \
                    and all(v.params[0] != b.params[0] for v in e.variables()):
                text = repr(e)
                try:                         # the chunk kernels are built from the printed sub-formula: it must round-trip
                    from .formulas.parser import parse_formula
                    if parse_formula(text) is not e:
                        return None
                except Exception:
                    return None
                return b.params[0], b.params[1], b.params[2], text
        return None

    def _genred_wide(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        """Chunks of ``WIDE_CHUNK`` components of the wide variable, one launch of the narrow kernel each (``E_ij`` is
        recomputed per chunk: a few operations against ``WIDE_CHUNK`` FMAs); the chunk results fill ``out[..., c0:c1]``."""
        ind, DV, cat, etext = self._wide
        p = self.params
        C = self.WIDE_CHUNK
        b = args[ind]
        if not isinstance(b, torch.Tensor) or b.shape[-1] != DV:
            raise ValueError("[KeOps] Wrong value of the 'vector size' dimension for arg number %d" % ind)
        res, launches = out, 0
        for c0 in range(0, DV, C):
            w = min(C, DV - c0)
            sub = self._wide_binders.get(w)
            if sub is None:
                formula = "Sum_Reduction(Mult(%s,Var(%d,%d,%d)),%d)" % (etext, ind, w, cat, p.tagI)
                t = self._init_args
                sub = LoadKeOpsB200(t[0], t[1], t[2], t[3], t[4], formula, [], p.nargs, p.dtype, t[5], self._optional_flags)
                sub._wide = None               # a chunk kernel is never chunked again
                self._wide_binders[w] = sub
            a2 = list(args)
            a2[ind] = b[..., c0:c0 + w].contiguous()
            so = sub._genred_device(device_args, ranges, nx, ny, nbatchdims, None, *a2)
            launches += sub.last_launch.get("launches", 1)
            if res is None:
                res = torch.empty(tuple(so.shape[:-1]) + (DV,), dtype=so.dtype, device=so.device)
            elif tuple(res.shape) != tuple(so.shape[:-1]) + (DV,):
                raise ValueError("[KeOps-B200] `out` must have shape %s" % (tuple(so.shape[:-1]) + (DV,),))
            res[..., c0:c0 + w] = so
        self.last_launch = dict(sub.last_launch, launches=launches, engine="cuda-core/wide-chunks", wide_chunks=-(-DV // C))
        return res

    # ---- reference-compatible surface ----------------------------------------------------------
    def import_module(self):
        return self

    class _Reset:
        """``reset`` is called both ways by the reference: on a binder object, and on the ``keops_binder[...]`` entry itself
        -- there a ``Cache_partial`` -- by ``pykeops.clean_pykeops()`` / ``pykeops.set_build_folder(path)``
        (``pykeops/pykeops/__init__.py:47-61,83-93``, ``keopscore/utils/Cache.py:99-109``), which for this engine is the
        class.  Instance: forget the loaded modules.  Class: forget every memoised binder; a new save folder becomes the
        cubin cache directory."""

        def __get__(self, obj, cls):
            if obj is not None:
                return obj._loaded.clear

            def reset_all(new_save_folder=None):
                with _cache_lock:
                    _binder_cache.clear()
                if new_save_folder is not None:
                    os.makedirs(new_save_folder, exist_ok=True)
                    os.environ["KB200_CACHE_DIR"] = str(new_save_folder)
            return reset_all

    reset = _Reset()

    def _kernel(self, device: int) -> plan.LoadedKernel:
        lk = self._loaded.get(device)
        if lk is None:
            lk = plan.load_kernel(self.ks, self.params.use_fast_math and self.params.dtype == "float32", device)
            self._loaded[device] = lk
        return lk

    def _genred_from_host(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        """``tagHostDevice == 0`` (the reference's ``GPU_1D_host`` / ``GPU_2D_host`` modes, what ``backend="auto"`` resolves
        to for host tensors on a GPU machine, ``pykeops/pykeops/common/get_options.py:84-120``): the reference's launcher
        copies the arrays to the device, runs the CUDA kernel and copies the result back
        (``keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:470-507,617-631``).  Same here -- the computation is still this
        engine's CUDA path; there is no CPU code."""
        p = self.params
        if not torch.cuda.is_available():
            raise ValueError("[KeOps-B200] no CUDA device: this engine has no CPU backend.")
        index = p.device_id_request if (p.device_id_request is not None and p.device_id_request >= 0) else torch.cuda.current_device()
        dev = torch.device("cuda", index)
        dargs = [a.to(dev, non_blocking=True) if isinstance(a, torch.Tensor) else a for a in args]
        host = next((a.device for a in args if isinstance(a, torch.Tensor)), torch.device("cpu"))
        res = self._genred_device(dev, ranges, nx, ny, nbatchdims, None, *dargs)
        if out is not None:
            out.copy_(res)
            return out
        return res.to(host)

    def genred(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        if self.params.tagHostDevice == 0 and not any(isinstance(a, torch.Tensor) and a.is_cuda for a in args):
            return self._genred_from_host(device_args, ranges, nx, ny, nbatchdims, out, *args)
        return self._genred_device(device_args, ranges, nx, ny, nbatchdims, out, *args)

    def _genred_device(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        p = self.params
        if self._wide is not None and len(args) == p.nargs and os.environ.get("KB200_WIDE_CHUNKS", "1") != "0":
            return self._genred_wide(device_args, ranges, nx, ny, nbatchdims, out, *args)
        if ranges is not None and len(ranges) == 0:
            ranges = None
        if ranges is not None and int(nbatchdims or 0) > 0:
            # reference: Ranges.h:88-96 ("not supported with batch processing")
            raise ValueError("[KeOps] The 'ranges' argument (block-sparse mode) is not supported with batch processing, "
                             "but we detected %d > 0 batch dimensions." % int(nbatchdims))
        if len(args) != p.nargs:
            raise ValueError("[KeOps-B200] expected %d arguments, got %d" % (p.nargs, len(args)))
        tdtype = torch.float32 if p.dtype == "float32" else torch.float64
        dev = None
        for k, a in enumerate(args):
            if not isinstance(a, torch.Tensor) or not a.is_cuda:
                raise ValueError("[KeOps-B200] argument %d must be a CUDA tensor (there is no CPU fallback)." % k)
            if a.dtype != tdtype:
                raise ValueError("[KeOps-B200] argument %d has dtype %s, expected %s" % (k, a.dtype, tdtype))
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise ValueError("[KeOps] Input arrays must be all located on the same device.")
        if dev is None:
            raise ValueError("[KeOps-B200] at least one argument is required")
        nx, ny = int(nx), int(ny)
        # the front-end computes max(arg.ndim) - 2 (generic_red.py:82): -1 when every argument is a parameter vector
        nb = max(0, int(nbatchdims or 0))
        # ---- shapes: [batch..., n, dim] for variables, [batch..., dim] for parameters
        #      (subset of Sizes<T>, keopscore/keopscore/include/Sizes.h:262-411)
        batch_shape = ()
        if nb:
            shapes = []
            for ind, (dim, cat) in self._vars.items():
                a = args[ind]
                lead = a.shape[:-1] if cat == 2 else a.shape[:-2]
                lead = (1,) * (nb - len(lead)) + tuple(lead)
                shapes.append(lead)
            batch_shape = tuple(torch.broadcast_shapes(*shapes)) if shapes else (1,) * nb
        nbatch = 1
        for b_ in batch_shape:
            nbatch *= int(b_)
        args = list(args)
        strides = [0] * p.nargs
        for ind, (dim, cat) in self._vars.items():
            a = args[ind]
            tail = (dim,) if cat == 2 else ((nx if cat == 0 else ny), dim)
            if cat == 2 and nb == 0:
                if a.numel() != dim:
                    raise ValueError("[KeOps] Wrong value of the 'vector size' dimension for arg number %d : is %d but should be %d"
                                     % (ind, a.numel(), dim))
                args[ind] = a.reshape(dim).contiguous()
                continue
            if a.dim() < len(tail) or tuple(a.shape[-len(tail):]) != tail:
                raise ValueError("[KeOps] Wrong shape for arg number %d : is %s but should end with %s"
                                 % (ind, tuple(a.shape), tail))
            lead = tuple(a.shape[:-len(tail)])
            lead = (1,) * (nb - len(lead)) + lead
            if len(lead) != nb:
                raise ValueError("[KeOps] Wrong number of batch dimensions for arg number %d" % ind)
            per = 1
            for t_ in tail:
                per *= t_
            if all(l == 1 for l in lead):
                args[ind] = a.reshape(tail).contiguous()           # shared by every batch
            else:
                if lead != batch_shape:                           # partial broadcast: materialise it
                    a = a.reshape(lead + tail).expand(batch_shape + tail)
                args[ind] = a.contiguous()
                strides[ind] = per
        for k in range(p.nargs):
            if k not in self._vars:
                args[k] = args[k].contiguous()

        nrow, ncol = (nx, ny) if p.tagI == 0 else (ny, nx)
        shapeout = batch_shape + (nrow, p.dim)
        if out is None:
            # block-sparse mode: rows covered by no range are never written (garbage in the reference) -> zeros
            out = (torch.zeros if ranges is not None else torch.empty)(shapeout, dtype=tdtype, device=dev)
        else:
            if tuple(out.shape) != shapeout or out.dtype != tdtype or out.device != dev or not out.is_contiguous():
                raise ValueError("[KeOps-B200] `out` must be a contiguous %s tensor of shape %s on %s"
                                 % (tdtype, shapeout, dev))
        if p.tagZero:
            out.zero_()
            return out
        if nrow == 0 or nbatch == 0:
            return out
        if ranges is not None:
            return self._genred_ranges(dev, ranges, nrow, ncol, out, args, tdtype)
        tc = knn_tc.eligible(self.red, p.dtype, nrow, ncol, nbatch)
        if tc is not None:
            return self._genred_knn_tc(dev, tc, nrow, ncol, out, args, strides, tdtype)
        if (self._conv_grad is not None and nbatch == 1 and p.use_fast_math and p.sum_scheme in ("block_sum", "auto")
                and p.dtype_acc in ("float32", "auto") and os.environ.get("KB200_CONV_TC", "1") != "0"
                and os.environ.get("KB200_CONV_TC_GRAD", "1") != "0"
                and nrow >= (8192 if ncol < 65536 else 1024) and ncol >= 1024 and ncol < (1 << 31) - 128):
            if self._genred_conv_tc_grad(dev, self._conv_grad, nrow, ncol, out, args):
                return out
        ct = conv_tc.eligible(self.red, p.dtype, nrow, ncol, nbatch, p.use_fast_math, p.sum_scheme, p.dtype_acc)
        if ct is not None:
            if self._genred_conv_tc(dev, ct, nrow, ncol, out, args):
                return out
        if self.red.kind == "LSE":
            cl = conv_tc.eligible_lse(self.red, p.dtype, nrow, ncol, nbatch, p.use_fast_math, p.sum_scheme, p.dtype_acc)
            if cl is not None and self._genred_lse_tc(dev, cl, nrow, ncol, out, args):
                return out
        if self._chunk_plan is not None and os.environ.get("KB200_CHUNKS", "1") != "0":
            if self._launch_chunked(dev, nrow, ncol, nbatch, out, args, strides, tdtype):
                return out
        return self._launch_dense(dev, nrow, ncol, nbatch, out, args, strides, tdtype)

    def _launch_chunked(self, dev, nrow, ncol, nbatch, out, args, strides, tdtype) -> bool:
        """High-dimensional formulas on the chunked kernel (``codegen/emit_chunked.py``); False when the formula is outside
        that scheme (the generic kernel then runs)."""
        import ctypes as C
        p = self.params
        if self.ks_chunked is None:
            ks = emit_chunked.generate_chunked(self.red, self._chunk_plan, p.dtype, p.nargs, use_fast_math=p.use_fast_math,
                                               sum_scheme=p.sum_scheme, dtype_acc=p.dtype_acc)
            if ks is None:
                self._chunk_plan = None
                return False
            self.ks_chunked = ks
        ks = self.ks_chunked
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        lk = self._loaded.get(("chunked", index))
        if lk is None:
            lk = plan.load_kernel(ks, p.use_fast_math and p.dtype == "float32", index)
            self._loaded[("chunked", index)] = lk
        ch = ks.chunk
        esz = 4 if p.dtype == "float32" else 8
        args = list(args)
        if ch["align16"]:
            for ind in ch["hd_j"]:
                if args[ind].data_ptr() % 16:
                    args[ind] = args[ind].clone()         # 16-byte cp.async needs an aligned base (offset views)
        G, rows_pb, TJ = ch["G"], ch["ROWS"], ch["TJ"]
        slots = max(1, (lk.sm_count * lk.blocks_per_sm) // min(nbatch, lk.sm_count * lk.blocks_per_sm))
        nsplit, col_chunk = plan.launch_shape(nrow, ncol, rows_pb, TJ, slots, ks.dimpart * G, esz,
                                              ws_limit_bytes=max(1, (2 << 30) // nbatch))
        ntot = nsplit * G
        i64, ptr, i32 = C.c_longlong, C.c_void_p, C.c_int
        gx = -(-nrow // rows_pb)
        launches = 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            out_flat = out.view(nbatch, nrow * p.dim)
            for b0 in range(0, nbatch, 65535):
                bn = min(65535, nbatch - b0)
                ws = torch.empty(bn * ntot * nrow * ks.dimpart, dtype=tdtype, device=dev) if ntot > 1 else None
                out_ptr = out_flat[b0].data_ptr()
                params = [i64(nrow), i64(ncol), i64(max(col_chunk, 1)), ptr(out_ptr), ptr(ws.data_ptr() if ws is not None else 0)]
                params += [ptr(a.data_ptr() + b0 * strides[k] * esz) for k, a in enumerate(args)]
                params += [i64(int(st)) for st in strides]
                capi.launch(lk.fn_reduce, stream, (gx, nsplit, bn), (ks.cfg.T,), ks.dynamic_smem, params)
                launches += 1
                if ws is not None:
                    capi.launch(lk.fn_combine, stream, (-(-nrow // 256), 1, bn), (256,), 0,
                                [i64(nrow), i32(ntot), ptr(ws.data_ptr()), ptr(out_ptr)])
                    launches += 1
                    ws.record_stream(torch.cuda.current_stream(dev))
        self.last_launch = dict(nsplit=nsplit, col_chunk=col_chunk, grid_x=gx, nbatch=nbatch, regs=lk.regs,
                                blocks_per_sm=lk.blocks_per_sm, launches=launches, engine="cuda-core/chunked",
                                chunk=dict(D=ch["D"], rows=rows_pb, groups=G, tile=TJ))
        return True

    def _genred_knn_tc(self, dev, tc, nrow, ncol, out, args, strides, tdtype):
        """K-min of a squared distance in large dimension: tcgen05 filter + exact re-evaluation
        (``codegen/knn_tc.py``); rows whose candidate list overflowed go through the CUDA-core kernel."""
        p = self.params
        row_ind, col_ind, D = tc
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        eng = knn_tc.get(D, self.red.K, self.red.kind, p.use_fast_math, index)
        with torch.cuda.device(dev):
            over = eng.run(args[row_ind], args[col_ind], out.view(nrow, p.dim), torch.cuda.current_stream(dev))
        self.last_launch = dict(nsplit=eng.last["nsplit"], col_chunk=ncol, grid_x=eng.last["ctas"], nbatch=1, regs=0,
                                blocks_per_sm=1, launches=4 + (eng.last["nsplit"] > 1), engine="tcgen05",
                                overflow_rows=eng.last["overflow"], segments=eng.last["segments"])
        if over.numel():
            sub = list(args)
            sub[row_ind] = args[row_ind][over].contiguous()
            res = torch.empty((over.numel(), p.dim), dtype=tdtype, device=dev)
            self._launch_dense(dev, over.numel(), ncol, 1, res, sub, strides, tdtype)
            out.view(nrow, p.dim)[over] = res
            self.last_launch["engine"] = "tcgen05"
            self.last_launch["overflow_rows"] = int(over.numel())
        return out

    def _genred_conv_tc(self, dev, ct, nrow, ncol, out, args):
        """Gaussian-kernel Sum reduction with the squared distance on the tensor cores (``codegen/conv_tc.py``):
        tcgen05 GEMM on three-term bf16 splits + ex2 / weighted accumulation epilogue, under the stated tolerance."""
        p = self.params
        row_ind, col_ind, D, b_ind, DB, terms = ct
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        eng = conv_tc.get(D, DB, index)
        with torch.cuda.device(dev):
            coef = conv_tc.coefficient(terms, args, dev)
            # `clock_probe` (bench.py sets it): int64[2] tensor receiving (SM cycles, ns) of one CTA of the kernel
            ok = eng.run(args[row_ind], args[col_ind], args[b_ind], coef, out.view(nrow, p.dim),
                         torch.cuda.current_stream(dev), clock_probe=getattr(self, "clock_probe", None))
        if not ok:
            return False            # accuracy guard: the caller runs the CUDA-core kernel
        self.last_launch = dict(nsplit=1, col_chunk=ncol, grid_x=eng.last["ctas"], nbatch=1, regs=0, blocks_per_sm=1,
                                launches=3 + (eng.last["nchunks"] > 1), engine="tcgen05-conv",
                                error_bound=eng.last_bound, nchunks=eng.last["nchunks"])
        return True

    def _genred_conv_tc_grad(self, dev, gp, nrow, ncol, out, args):
        """Gradient of ``sum_j exp(c |x_i - y_j|^2) b_j`` w.r.t. ``x`` or ``y`` from forward-type sums on the tensor cores
        (``conv_tc.match_grad`` has the algebra and the error estimate).  False: the accuracy guard refused -- the caller
        runs the CUDA-core gradient kernel."""
        kind, row_ind, col_ind, D, b_ind, eta_ind, fwd_ind, terms = gp
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        chunk = conv_tc.grad_chunk(D)
        if chunk is None:
            return False
        MAXDB, stages = chunk
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            coef = conv_tc.coefficient(terms, args, dev)
            xr, yc, b, eta = args[row_ind], args[col_ind], args[b_ind], args[eta_ind]
            if kind == "x":
                # rows = the forward's rows; streamed side = (y, b):  2 c e_i [ (x_i - m) S0_i - sum_j k_ij b_j (y_j - m) ]
                rows, cols, lead_w, scale_rows = xr, yc, b, eta
                s0 = args[fwd_ind] if fwd_ind is not None else None
            else:
                # rows = the forward's streamed points y_j; streamed side = (x, e): 2 c b_j [ (y_j - m') T0_j - sum_i k_ij e_i (x_i - m') ]
                rows, cols, lead_w, scale_rows = yc, xr, eta, b
                s0 = None
            m = cols.mean(dim=0, dtype=torch.float32)
            w = lead_w * (cols - m)                                    # (ncol, D) weights of the D-wide sum
            if s0 is None:                                             # the scalar sum rides in the first chunk
                w = torch.cat([lead_w, w], dim=1)
            nw = w.shape[1]
            sums = torch.empty((nrow, nw), dtype=torch.float32, device=dev)
            launches, bound, nchunks = 0, 0.0, 0
            for c0 in range(0, nw, MAXDB):
                c1 = min(nw, c0 + MAXDB)
                eng = conv_tc.get(D, c1 - c0, index, stages=stages)
                part = sums if (c0 == 0 and c1 == nw) else torch.empty((nrow, c1 - c0), dtype=torch.float32, device=dev)
                if not eng.run(rows, cols, w[:, c0:c1].contiguous(), coef, part, stream):
                    return False
                if part is not sums:
                    sums[:, c0:c1] = part
                launches += 3 + (eng.last["nchunks"] > 1)
                nchunks = max(nchunks, eng.last["nchunks"])
                bound = max(bound, eng.last_bound or 0.0)
            if s0 is None:
                s0, s1 = sums[:, :1], sums[:, 1:]
            else:
                s1 = sums
            res = (2.0 * coef) * scale_rows * ((rows - m) * s0 - s1)
            out.view(nrow, D).copy_(res)
        self.last_launch = dict(nsplit=1, col_chunk=ncol, grid_x=-(-nrow // 256), nbatch=1, regs=0, blocks_per_sm=1,
                                launches=launches, engine="tcgen05-conv-grad", error_bound=bound, nchunks=nchunks,
                                weight_columns=nw)
        return True

    def _genred_lse_tc(self, dev, cl, nrow, ncol, out, args):
        """``Max_SumShiftExp[Weight]`` of ``a_j + c |x_i - y_j|^2`` (soft-min of a Sinkhorn iteration) in dimension
        D >= 5: the same tensor-core kernels in LSE mode (``codegen/conv_tc.py``), same tolerance guard."""
        p = self.params
        row_ind, col_ind, D, a_ind, w_ind, DB, terms = cl
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        eng = conv_tc.get(D, DB, index, lse=True)
        with torch.cuda.device(dev):
            coef = conv_tc.coefficient(terms, args, dev)
            w = args[w_ind] if w_ind is not None else torch.ones((ncol, 1), dtype=torch.float32, device=dev)
            acol = args[a_ind].reshape(-1) if a_ind is not None else None
            ok = eng.run(args[row_ind], args[col_ind], w, coef, out.view(nrow, p.dim), torch.cuda.current_stream(dev),
                         acol=acol)
        if not ok:
            return False
        self.last_launch = dict(nsplit=1, col_chunk=ncol, grid_x=eng.last["ctas"], nbatch=1, regs=0, blocks_per_sm=1,
                                launches=3 + (eng.last["nchunks"] > 1), engine="tcgen05-lse", error_bound=eng.last_bound,
                                nchunks=eng.last["nchunks"])
        return True

    # ---- pre-scaled Gaussian-type sums in low dimension (codegen/emit.py::_prescale_rewrite) ------------------------
    # Taken when the problem is large enough to amortise the pack kernel, the statistics and one host read-back, and when
    # the run-time bound on the rounding of the scaled coordinates is small: per-term relative error <= PRESCALE_TOL in the
    # worst case (all six coordinate roundings aligned); measured on the BASELINE data (randn clouds, M = N = 1e6) the
    # reduced sums differ from the reference's CUDA / CPU code by ~2e-7 of sum_j |F_ij| (bench.py "parity").
    PRESCALE_MIN_PAIRS = 2e10
    PRESCALE_TOL = 3e-5

    def _launch_prescaled(self, dev, nrow, ncol, out, args, tdtype) -> bool:
        import ctypes as C
        import math
        p = self.params
        if self._prescale_ks is None:
            ks = generate(self.red, p.dtype, p.nargs, self.ks.cfg, use_fast_math=True, sum_scheme=p.sum_scheme,
                          dtype_acc=p.dtype_acc, prescale=True)
            pairs = ks.prescale["pairs"] if ks is not None else []
            if ks is None or len({q[0] for q in pairs}) != 1 or len({q[2] for q in pairs}) != 1:
                self._prescale_ks = False
                return False
            self._prescale_ks = ks
        ks = self._prescale_ks
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        lk = self._loaded.get(("prescaled", index))
        if lk is None:
            lk = plan.load_kernel(ks, True, index)
            lk.fn_pack = lk.module.function("kb200_pack")
            self._loaded[("prescaled", index)] = lk
        pairs, stride = ks.prescale["pairs"], ks.prescale["stride"]
        ix, iy = pairs[0][0], pairs[0][2]
        cx, cy = [q[1] for q in pairs], [q[3] for q in pairs]
        xa, ya = args[ix], args[iy]
        xv = xa if cx == list(range(xa.shape[1])) else xa[:, cx]
        yv = ya if cy == list(range(ya.shape[1])) else ya[:, cy]
        i64, ptr = C.c_longlong, C.c_void_p
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            shf = yv.mean(0).contiguous()                                   # common shift: the streamed cloud's mean
            ypk = torch.empty(ncol * stride, dtype=tdtype, device=dev)
            info = torch.empty(2, dtype=tdtype, device=dev)
            capi.launch(lk.fn_pack, stream, (-(-ncol // 256), 1, 1), (256,), 0,
                        [i64(ncol), ptr(ypk.data_ptr()), ptr(shf.data_ptr()), ptr(info.data_ptr())] + [ptr(a.data_ptr()) for a in args])
            # bound on the exponent error of a term: 4 * 2^-24 |d'| (|x'| + |y'|) with |d'| <= min(|x'| + |y'|, 11.3) for every
            # term that does not underflow (2^-126); times ln 2 for the relative error of the term
            t = ((xv - shf).square().sum(1).amax().sqrt() + (yv - shf).square().sum(1).amax().sqrt()) * info[1]
            bound = (4.0 * 2.0 ** -24 * math.log(2.0)) * t.clamp(max=11.3) * t
            k_val, b_val = torch.stack([info[0].reshape(()), bound.reshape(())]).tolist()     # one host read-back per call
            if not (k_val > 0.0 and b_val <= self.PRESCALE_TOL):            # also catches NaN / inf
                self._prescale_rejected = dict(K=k_val, bound=b_val)
                return False
            rows_per_block = ks.cfg.T * ks.cfg.R
            slots = lk.sm_count * lk.blocks_per_sm
            nsplit, col_chunk = plan.launch_shape(nrow, ncol, rows_per_block, ks.cfg.TJ, slots, ks.dimpart, 4, sms=lk.sm_count)
            ws = torch.empty(nsplit * nrow * ks.dimpart, dtype=tdtype, device=dev) if nsplit > 1 else None
            params = [i64(nrow), i64(ncol), i64(max(col_chunk, 1)), ptr(out.data_ptr()), ptr(ws.data_ptr() if ws is not None else 0)]
            params += [ptr(a.data_ptr()) for a in args] + [i64(0)] * p.nargs + [ptr(ypk.data_ptr()), ptr(shf.data_ptr())]
            gx = -(-nrow // rows_per_block)
            capi.launch(lk.fn_reduce, stream, (gx, nsplit, 1), (ks.cfg.T,), ks.dynamic_smem, params)
            if ws is not None:
                capi.launch(lk.fn_combine, stream, (-(-nrow // 256), 1, 1), (256,), 0,
                            [i64(nrow), C.c_int(nsplit), ptr(ws.data_ptr()), ptr(out.data_ptr())])
                ws.record_stream(torch.cuda.current_stream(dev))
            for t_ in (ypk, shf, info):
                t_.record_stream(torch.cuda.current_stream(dev))
        self.last_launch = dict(nsplit=nsplit, col_chunk=col_chunk, grid_x=gx, nbatch=1, regs=lk.regs,
                                blocks_per_sm=lk.blocks_per_sm, launches=2 + (nsplit > 1), engine="cuda-core/prescaled",
                                error_bound=b_val)
        return True

    def _launch_dense(self, dev, nrow, ncol, nbatch, out, args, strides, tdtype):
        p = self.params
        if (self._prescale_ks is not False and self.red.kind == "Sum" and p.dtype == "float32" and p.use_fast_math
                and nbatch == 1 and p.tag1D2D == 0 and float(nrow) * float(ncol) >= self.PRESCALE_MIN_PAIRS
                and os.environ.get("KB200_PRESCALE", "1") != "0"):
            if self._launch_prescaled(dev, nrow, ncol, out, args, tdtype):
                return out
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        lk = self._kernel(index)
        ks = self.ks
        engine = None
        kf_env = os.environ.get("KB200_KMIN_FILTER", "1")
        if (self._kf_match is not None and nbatch == 1 and p.tag1D2D == 0 and kf_env != "0"
                and (kf_env == "force" or (float(nrow) * float(ncol) >= self.KFILTER_MIN_PAIRS
                                           and ncol >= self.KFILTER_MIN_COLS * max(1, self.red.K // 2)))):
            kf = self._kmin_filter_setup(dev, index, ncol, args, tdtype)
            if kf is not None:
                lk, ks, extra = kf
                args, strides = list(args) + extra, list(strides) + [0, 0]
                engine = "cuda-core/kmin-filter"
        rows_per_block = ks.cfg.T * ks.cfg.R
        esz = 4 if p.dtype == "float32" else 8
        slots = max(1, (lk.sm_count * lk.blocks_per_sm) // min(nbatch, lk.sm_count * lk.blocks_per_sm))
        ws_limit = max(1, (2 << 30) // nbatch)
        kfam = self.red.kind in ("KMin", "ArgKMin", "KMin_ArgKMin")
        # Tail-split pays only when a restart is expensive (K-min lists).  For sums, measured on B200
        # (Gaussian conv, M=N=1e6): one long unsplit wave + a split tail = 272 ms vs 262 ms for the uniform
        # 3-way split, whose short CTAs backfill and average out per-SM speed differences.
        tail_split = kfam if os.environ.get("KB200_TAILSPLIT") is None else os.environ["KB200_TAILSPLIT"] == "1"
        # restart cost of a K-min list, in units of the whole column range: independent splits each replay the
        # transient of an empty list; splits that share thresholds (kb200_reduce_seeded) pay much less
        shared = kfam and nbatch == 1 and os.environ.get("KB200_KMIN_SEED", "1") != "0"
        kpen = float(os.environ.get("KB200_SPLIT_PENALTY", 0.002 if shared else 0.02))   # measured sweep: experiments/knn_seed_sweep.sh
        if nbatch == 1 and p.tag1D2D == 0 and tail_split:
            # restart cost of a K-min list, in tile units: ~ K * log(N/K) insertions per row, each stalling its warp
            penalty = (kpen * (-(-ncol // ks.cfg.TJ))) if kfam else 0.0
            segments = plan.launch_segments(nrow, ncol, rows_per_block, ks.cfg.TJ, slots, ks.dimpart, esz,
                                            split_penalty_tiles=penalty, ws_limit_bytes=ws_limit,
                                            sms=lk.sm_count if nbatch == 1 else 0, prefer_uniform=engine is not None)
        else:
            nsplit, col_chunk = plan.launch_shape(nrow, ncol, rows_per_block, ks.cfg.TJ, slots, ks.dimpart, esz,
                                                  ws_limit_bytes=ws_limit, sms=lk.sm_count if nbatch == 1 else 0,
                                                  split_penalty_tiles=(kpen * (-(-ncol // ks.cfg.TJ))) if kfam else 0.0)
            if p.tag1D2D == 1 and nsplit == 1 and ncol > 4 * ks.cfg.TJ:
                # backend="GPU_2D": the caller asked for the split-column scheme explicitly
                ntiles = -(-ncol // ks.cfg.TJ)
                nsplit = min(8, ntiles)
                col_chunk = -(-ntiles // nsplit) * ks.cfg.TJ
                nsplit = -(-ncol // col_chunk)
            segments = [(0, nrow, nsplit, col_chunk)]
        row_dims = {ind: dim for ind, (dim, cat) in self._vars.items() if cat == p.tagI}
        BMAX = 65535
        launches = 0
        # column splits of a K-min reduction share per-row thresholds through global memory (kb200_reduce_seeded)
        n_seed = 0
        if (kfam and nbatch == 1 and lk.fn_seeded and any(sg[2] > 1 for sg in segments)
                and os.environ.get("KB200_KMIN_SEED", "1") != "0"):
            n_seed = 1
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            out_flat = out.view(nbatch, nrow * p.dim)
            if n_seed:
                launches += self._launch_kmin_seeded(lk, ks, stream, dev, tdtype, esz, nrow, ncol, n_seed, segments, out_flat,
                                                     args, strides, row_dims, rows_per_block)
                segments_to_run = []
            else:
                segments_to_run = segments
            for (r0, r1, nsplit, col_chunk) in segments_to_run:
                nseg = r1 - r0
                for b0 in range(0, nbatch, BMAX):
                    bn = min(BMAX, nbatch - b0)
                    ws = None
                    if nsplit > 1:
                        ws = torch.empty(bn * nsplit * nseg * ks.dimpart, dtype=tdtype, device=dev)
                    ptrs = [a.data_ptr() + (b0 * strides[k] + r0 * row_dims.get(k, 0)) * esz for k, a in enumerate(args)]
                    capi.genred_launch(lk.fn_reduce, lk.fn_combine, stream, nseg, ncol, rows_per_block, ks.cfg.T,
                                       nsplit, max(col_chunk, 1), out_flat[b0].data_ptr() + r0 * p.dim * esz,
                                       ws.data_ptr() if ws is not None else 0, ptrs, bn, strides,
                                       dynamic_smem=ks.dynamic_smem)
                    launches += 2 if nsplit > 1 else 1
                    if ws is not None:
                        ws.record_stream(torch.cuda.current_stream(dev))
        nsplit, col_chunk = segments[-1][2], segments[-1][3]
        self.last_launch = dict(nsplit=nsplit, col_chunk=col_chunk, grid_x=-(-nrow // rows_per_block),
                                nbatch=nbatch, regs=lk.regs, blocks_per_sm=lk.blocks_per_sm, launches=launches,
                                segments=[(a, b, c) for a, b, c, _ in segments], kmin_seed_columns=n_seed)
        if engine is not None:
            self.last_launch["engine"] = engine
            self.last_launch["launches"] = launches + 1            # + kb200_kfpack
            for t_ in args[-2:]:
                t_.record_stream(torch.cuda.current_stream(dev))
        return out

    # ---- K nearest neighbours in dimension <= 3: filtered kernel (codegen/emit_kmin_filter.py) ----------------------------
    KFILTER_MIN_PAIRS = 1e8        # below, the pack launch and the mean are not worth it
    # The filter pays once survivors are rare per WARP: a warp's 128 rows together see ~128 K ln(N / K) list updates whatever
    # N is, against N / 8 column groups.  Measured (experiments/kmeans_check.py, ArgMin, D = 3): 1e6 x 1e6 204.8 -> 134.9 ms,
    # but 1e7 points against 1000 centroids 2.59 -> 3.54 ms (every group of the short stream holds an update for some lane).
    KFILTER_MIN_COLS = 50000       # streamed points, times K / 2 for lists (K = 10: 250 000)

    def _kfilter_source(self) -> Optional[KernelSource]:
        """The filtered kernel's source for this reduction (generated once), or None when it does not apply."""
        if self.ks_kfilter is None:
            # 128-thread CTAs, three per SM: 170 registers per thread keep the K-lists of four rows without spills
            # (256 threads x 2 CTAs cap at 128 registers: 20 values spill inside the pair loop)
            c0 = self.ks.cfg
            lists = emit_kmin_filter.KINDS[self.red.kind][1]
            D = self._kf_match[2]
            s4 = -(-(D + 1) // 4) + -(-D // 4)               # float4 per staged column: filter row + coordinates
            # D <= 3: four rows per thread, 1024-column tiles (64 KB), three 128-thread CTAs per SM (170 registers keep the
            # K-lists of four rows without spills; Min / ArgMin: 89 registers, five CTAs, 512-column tiles).  4 <= D <= 15: two
            # rows per thread (the -2 x' values alone take 2 D registers), tiles sized to ~48-64 KB
            if D <= 3:
                R, TJ, B = c0.R, (1024 if lists else 512), (3 if lists else 5)
            else:
                R, TJ, B = 2, (512 if s4 <= 4 else 256), 3
            cfg = KernelConfig(T=int(os.environ.get("KB200_KF_T", 128)), R=int(os.environ.get("KB200_KF_R", R)),
                               TJ=int(os.environ.get("KB200_KF_TJ", TJ)), unroll=c0.unroll,
                               min_blocks=int(os.environ.get("KB200_KF_B", B)), group=c0.group)
            ks = emit_kmin_filter.generate_filtered(self.red, self.params.nargs, cfg, self.ks)
            self.ks_kfilter = ks if ks is not None else False
        return self.ks_kfilter or None

    def _kmin_filter_setup(self, dev, index, ncol, args, tdtype):
        """``(loaded kernel, kernel source, [packed rows, info])`` of the filtered K-min kernel for this call, after
        ``kb200_kfpack`` has written the shifted streamed rows and the largest squared norm; ``None`` if the kernel does
        not apply."""
        import ctypes as C
        p = self.params
        ks = self._kfilter_source()
        if ks is None:
            return None
        lk = self._loaded.get(("kfilter", index))
        if lk is None:
            lk = plan.load_kernel(ks, True, index)
            lk.fn_pack = lk.module.function("kb200_kfpack")
            self._loaded[("kfilter", index)] = lk
        D, iy = ks.kfilter["D"], ks.kfilter["col"]
        y = args[iy]
        i64, ptr = C.c_longlong, C.c_void_p
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            info = torch.zeros(16 + -(-ncol // 256), dtype=tdtype, device=dev)    # [shift (16), max |y'|^2 per 256 columns]
            # common shift (x - y is unchanged): the mean of the streamed cloud, non-finite entries counted as 0 -- any finite
            # vector near the cloud will do, the margin is computed from what the shift leaves
            info[:D] = torch.nan_to_num(torch.nan_to_num(y, nan=0.0, posinf=0.0, neginf=0.0).mean(dim=0, dtype=torch.float32),
                                        nan=0.0, posinf=0.0, neginf=0.0)
            ypk = torch.empty(ncol * ks.kfilter["stride"], dtype=tdtype, device=dev)   # per column: the filter's row + the coordinates
            capi.launch(lk.fn_pack, stream, (-(-ncol // 256), 1, 1), (256,), 0,
                        [i64(ncol), ptr(y.data_ptr()), ptr(ypk.data_ptr()), ptr(info.data_ptr())])
        return lk, ks, [ypk, info]

    def _launch_kmin_seeded(self, lk, ks, stream, dev, tdtype, esz, nrow, ncol, n_seed, segments, out_flat, args, strides,
                            row_dims, rows_per_block):
        """Row segments through ``kb200_reduce_seeded`` / ``kb200_combine``: the column splits of a row publish their
        K-th smallest value to ``shthr[row][component]`` (atomic min, starts at +inf) and skip candidates above the
        smallest value published so far -- an upper bound of the final K-th value, so the result is exactly the
        unsplit kernel's while each split is spared the transient of a list that starts empty."""
        import ctypes as C
        p = self.params
        i64, ptr, i32 = C.c_longlong, C.c_void_p, C.c_int
        d = self.red.dimF
        T = ks.cfg.T
        shthr = torch.full((nrow * d,), float("inf"), dtype=tdtype, device=dev)
        count = 0
        for (r0, r1, nsplit, col_chunk) in segments:
            nseg = r1 - r0
            ws = torch.empty(nsplit * nseg * ks.dimpart, dtype=tdtype, device=dev) if nsplit > 1 else None
            out_ptr = out_flat[0].data_ptr() + r0 * p.dim * esz
            ptrs = [a.data_ptr() + r0 * row_dims.get(k, 0) * esz for k, a in enumerate(args)]
            if nsplit == 1:          # nothing to share: the plain entry point (fewer registers)
                capi.genred_launch(lk.fn_reduce, lk.fn_combine, stream, nseg, ncol, rows_per_block, T, 1,
                                   max(col_chunk, 1), out_ptr, 0, ptrs, 1, strides, dynamic_smem=ks.dynamic_smem)
                count += 1
                continue
            params = [i64(nseg), i64(ncol), i64(max(col_chunk, 1)), ptr(out_ptr), ptr(ws.data_ptr() if ws is not None else 0)]
            params += [ptr(q) for q in ptrs] + [i64(int(st)) for st in strides]
            params += [ptr(shthr.data_ptr() + r0 * d * esz), i32(0)]
            capi.launch(lk.fn_seeded, stream, (-(-nseg // rows_per_block), nsplit, 1), (T,), ks.dynamic_smem, params)
            count += 1
            if nsplit > 1:
                capi.launch(lk.fn_combine, stream, (-(-nseg // 256), 1, 1), (256,), 0,
                            [i64(nseg), i32(nsplit), ptr(ws.data_ptr()), ptr(out_ptr)])
                count += 1
                ws.record_stream(torch.cuda.current_stream(dev))
        shthr.record_stream(torch.cuda.current_stream(dev))
        return count

    # ---- block-sparse mode -----------------------------------------------------------------------
    def _kernel_ranges(self, device: int) -> plan.LoadedKernel:
        lk = self._loaded.get(("ranges", device))
        if lk is None:
            p = self.params
            if self.ks_ranges is None:
                self.ks_ranges = generate(self.red, p.dtype, p.nargs, self.ks.cfg, use_fast_math=p.use_fast_math,
                                          sum_scheme=p.sum_scheme, dtype_acc=p.dtype_acc, ranges=True)
            lk = plan.load_kernel(self.ks_ranges, p.use_fast_math and p.dtype == "float32", device)
            self._loaded[("ranges", device)] = lk
        return lk

    def _genred_ranges(self, dev, ranges, nrow, ncol, out, args, tdtype):
        """``ranges = (ranges_i, slices_i, redranges_j, ranges_j, slices_j, redranges_i)``: the first three
        drive a reduction over j (tagI == 0), the last three a reduction over i
        (``pykeops/pykeops/torch/generic/generic_red.py:581-639``, ``keops_nvrtc.cpp:166-170``)."""
        p = self.params
        if len(ranges) != 6:
            raise ValueError("[KeOps] the 'ranges' argument should be a tuple of 6 integer arrays.")
        rx, sx, ry = ranges[0:3] if p.tagI == 0 else ranges[3:6]
        rx_user = rx
        rx, sx, ry = (r.to(device=dev, dtype=torch.int64).contiguous() for r in (rx, sx, ry))
        nranges = int(rx.shape[0])
        if rx.dim() != 2 or rx.shape[1] != 2 or ry.dim() != 2 or ry.shape[1] != 2 or sx.numel() != nranges:
            raise ValueError("[KeOps] wrong shapes for the 'ranges' arrays: expected (Nc,2), (Nc,), (Ncc,2).")
        if nranges == 0:
            return out
        lk = self._kernel_ranges(dev.index if dev.index is not None else torch.cuda.current_device())
        if self.ks_ranges.dynamic_smem:
            raise ValueError("[KeOps-B200] block-sparse reductions are limited to streamed rows that fit 48 KB of shared memory "
                             "(%d components here)." % self.ks_ranges.stride)
        cfg = self.ks_ranges.cfg
        rpb = cfg.T * cfg.R
        key = (rx_user.data_ptr(), rx_user._version, nranges, rpb)
        hit = self._range_tables.get(key)
        if hit is None:
            # chunks per row-range, prefix-summed on the device; the grid size is the one value the host needs
            block_ends = torch.cumsum((rx[:, 1] - rx[:, 0] + (rpb - 1)).clamp_(min=0) // rpb, 0)
            grid_x = int(block_ends[-1].item())
            if len(self._range_tables) > 64:
                self._range_tables.clear()
            hit = (block_ends, grid_x, rx_user)     # keep the tensor alive so that its data_ptr stays unique
            self._range_tables[key] = hit
        block_ends, grid_x, _ = hit
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.genred_launch_ranges(lk.fn_reduce, stream, nrow, ncol, cfg.T, grid_x, out.data_ptr(),
                                      [a.data_ptr() for a in args], rx.data_ptr(), sx.data_ptr(), ry.data_ptr(),
                                      block_ends.data_ptr(), nranges)
            for t in (rx, sx, ry, block_ends):
                t.record_stream(torch.cuda.current_stream(dev))
        self.last_launch = dict(nsplit=1, col_chunk=ncol, grid_x=grid_x, nbatch=1, regs=lk.regs,
                                blocks_per_sm=lk.blocks_per_sm, launches=1, ranges=nranges)
        return out

    genred_pytorch = genred

    def genred_numpy(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        """The same call as made by the reference's numpy front-end (``LoadKeOps.py:246``,
        ``pykeops/pykeops/numpy/generic/generic_red.py:357``): host numpy arrays (and numpy ``ranges``) in, numpy array out.
        The arrays are staged on the CUDA device for the call, as the reference's ``*_host`` launch modes do."""
        import numpy as np
        targs = [torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a for a in args]
        if ranges is not None and len(ranges):
            ranges = tuple(torch.from_numpy(np.ascontiguousarray(r)) if isinstance(r, np.ndarray) else r for r in ranges)
        res = self._genred_from_host(device_args, ranges, nx, ny, nbatchdims, None, *targs).numpy()
        if out is not None:
            out[...] = res
            return out
        return res


def get_binder(formula: str, aliases: Sequence[str], nargs: int, dtype: str, optional_flags: Optional[dict] = None,
               device_id: int = -1) -> LoadKeOpsB200:
    flags = dict(dtype_acc="auto", sum_scheme="auto", use_fast_math=1, enable_chunks=1, multVar_highdim=0)
    if optional_flags:
        flags.update(optional_flags)
    return LoadKeOpsB200(1, 0, 1, False, device_id, formula, list(aliases), nargs, dtype, "torch", flags)


def register_into_pykeops():
    """Install this engine as the GPU binder of an importable reference ``pykeops``:
    ``keops_binder['nvrtc'] = LoadKeOpsB200`` (``pykeops/pykeops/common/keops_io/__init__.py:6-13``)."""
    import pykeops  # type: ignore
    from pykeops.common import keops_io  # type: ignore

    keops_io.keops_binder["nvrtc"] = LoadKeOpsB200
    # The reference resolves backend="auto"/"GPU" through ``pykeops.config.gpu_available``
    # (``pykeops/pykeops/common/get_options.py:84-86``), which it sets from its OWN CUDA probing (libcuda + libnvrtc found by
    # ``ctypes.util.find_library`` and its nvrtc binder built, ``keopscore/keopscore/config/cuda.py:117-150``).  This engine is
    # the GPU backend now, so availability follows the CUDA driver this engine uses.
    if torch.cuda.is_available():
        pykeops.config.gpu_available = True
    return keops_io.keops_binder
