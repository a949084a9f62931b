"""Loaders for the prebuilt reference comparators in ``oracle/_ref`` (see ``oracle/build_ref.py``).
TEST / BENCHMARK INFRASTRUCTURE ONLY -- never imported by the product.

* :func:`run_cpu`   -- the reference's generated ``CpuReduc`` code (OpenMP over all host cores);
* :class:`RefCuda`  -- the reference's generated ``GpuConv1DOnDevice`` kernel, launched exactly like
  ``KeOps_module::launch_kernel`` does (grid ``ceil(nx/192)``, block 192, dynamic shared memory
  ``192 * dimy * sizeof(T)``, device array of argument pointers;
  ``keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:441-449,596-614``).
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "manifest.json")) and os.path.exists(os.path.join(REF, "libkeops_ref_cpu.so"))


def manifest():
    with open(os.path.join(REF, "manifest.json")) as f:
        return json.load(f)


_lib = None


def _cpu_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(REF, "libkeops_ref_cpu.so"))
    return _lib


def run_cpu(key, args, nx, ny):
    """Run the reference CPU code for ``key``; ``args`` are contiguous numpy arrays. Returns (nrow, dim)."""
    m = manifest()[key]
    ct = ctypes.c_float if m["ctype"] == "float" else ctypes.c_double
    npdt = np.float32 if m["ctype"] == "float" else np.float64
    args = [np.ascontiguousarray(a, dtype=npdt) for a in args]
    nrow = nx if m["tagI"] == 0 else ny
    out = np.empty((nrow, m["dim"]), npdt)
    fn = getattr(_cpu_lib(), "kref_" + key)
    fn.restype = ctypes.c_int
    P = ctypes.POINTER(ct)
    argv = (P * len(args))(*[a.ctypes.data_as(P) for a in args])
    rc = fn(ctypes.c_long(nx), ctypes.c_long(ny), ctypes.c_int(m["tagI"]), out.ctypes.data_as(P), argv)
    if rc != 0:
        raise RuntimeError("reference CPU kernel %s returned %d" % (key, rc))
    return out


class RefCuda:
    """The reference's own CUDA kernel for ``key`` on the current device (comparator only)."""

    def __init__(self, key, device=0):
        import torch
        from keops_b200.runtime import capi
        self.m = manifest()[key]
        self.key = key
        with open(os.path.join(REF, "ref_cuda_%s.cubin" % key), "rb") as f:
            self.mod = capi.Module(f.read(), device)
        self.fn = self.mod.function(self.m["cuda_kernel"])
        self.capi = capi
        self.torch = torch
        self.device = device

    def __call__(self, args, nx, ny, out=None):
        torch, m = self.torch, self.m
        dt = torch.float32 if m["ctype"] == "float" else torch.float64
        esz = 4 if m["ctype"] == "float" else 8
        nrow, ncol = (nx, ny) if m["tagI"] == 0 else (ny, nx)
        if out is None:
            out = torch.empty(nrow, m["dim"], dtype=dt, device=args[0].device)
        table = torch.tensor([a.data_ptr() for a in args], dtype=torch.int64, device=args[0].device)
        block = m["cuda_block"]
        self._keep = (table, args)
        c = ctypes
        params = [c.c_long(nrow), c.c_long(ncol), c.c_void_p(out.data_ptr()), c.c_void_p(table.data_ptr())]
        stream = torch.cuda.current_stream().cuda_stream
        self.capi.launch(self.fn, stream, (-(-nrow // block),), (block,), block * m["dimy"] * esz, params)
        return out
