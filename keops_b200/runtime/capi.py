"""ctypes binding of ``libkb200.so`` (the C ABI declared in ``include/kb200.h``).

The product path has no fallback: if the shared library is missing, importing this module raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C keops_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_HERE, "libkb200.so")

KB200_ENOCUDA = 20001
KB200_ENONVRTC = 20002
KB200_NVRTC_BASE = 10000


class KB200Error(RuntimeError):
    """Driver / launch failure (the reference raises ``RuntimeError('[KeOps] Cuda error.')``,
    ``keopscore/keopscore/include/utils_pe.h:27-37``)."""

    def __init__(self, status: int, message: str):
        super().__init__("[KeOps-B200] status %d: %s" % (status, message))
        self.status = status


class KB200CompileError(ValueError):
    """NVRTC failure; carries the compiler log (the reference maps compile failures to
    ``KeOps_Error`` -> ``ValueError``, ``binders/nvrtc/Gpu_link_compile.py:88-91``)."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "keops_b200: native runtime %s is missing -- run __graft_entry__.build() first; "
            "there is no CPU or pure-Python fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    vp, i, u, sz, i64 = C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.c_int64
    cp = C.c_char_p
    sig = {
        "kb200_abi_version": (i, []),
        "kb200_last_error": (cp, []),
        "kb200_compile": (i, [cp, cp, C.POINTER(cp), i, cp, cp, sz]),
        "kb200_compile_to_memory": (i, [cp, cp, C.POINTER(cp), i, C.POINTER(vp), C.POINTER(sz), cp, sz]),
        "kb200_free": (None, [vp]),
        "kb200_module_load": (i, [vp, sz, i, C.POINTER(vp)]),
        "kb200_module_load_file": (i, [cp, i, C.POINTER(vp)]),
        "kb200_module_unload": (i, [vp]),
        "kb200_module_get_function": (i, [vp, cp, C.POINTER(vp)]),
        "kb200_function_info": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "kb200_max_active_blocks": (i, [vp, i, sz, C.POINTER(i)]),
        "kb200_device_info": (i, [i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i64), C.POINTER(i64)]),
        "kb200_launch": (i, [vp, vp, u, u, u, u, u, u, u, C.POINTER(vp)]),
        "kb200_function_set_max_dynamic_smem": (i, [vp, i]),
        "kb200_tensor_map_2d": (i, [vp, i, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, i]),
        "kb200_genred_launch": (i, [vp, vp, vp, i64, i64, i, i, i, i64, vp, vp, i, C.POINTER(vp), i, C.POINTER(i64)]),
        "kb200_genred_launch_smem": (i, [vp, vp, vp, i64, i64, i, i, i, i64, vp, vp, i, C.POINTER(vp), i, C.POINTER(i64), u]),
        "kb200_genred_launch_ranges": (i, [vp, vp, i64, i64, i, i64, vp, i, C.POINTER(vp), vp, vp, vp, vp, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED_SYMBOLS = (
    "kb200_abi_version", "kb200_last_error", "kb200_compile", "kb200_compile_to_memory", "kb200_free",
    "kb200_module_load", "kb200_module_load_file", "kb200_module_unload", "kb200_module_get_function",
    "kb200_function_info", "kb200_max_active_blocks", "kb200_device_info", "kb200_launch",
    "kb200_genred_launch", "kb200_genred_launch_smem", "kb200_genred_launch_ranges", "kb200_function_set_max_dynamic_smem", "kb200_tensor_map_2d",
)


def last_error() -> str:
    return (lib.kb200_last_error() or b"").decode("utf-8", "replace")


def check(status: int):
    if status != 0:
        raise KB200Error(status, last_error())


def compile_to_cubin(source: str, options: Sequence[str] = (), name: str = "kb200_kernel.cu") -> bytes:
    """NVRTC -> sm_100a cubin.  Works without a GPU."""
    opts = (C.c_char_p * len(options))(*[o.encode() for o in options])
    out, size = C.c_void_p(), C.c_size_t()
    log = C.create_string_buffer(1 << 16)
    st = lib.kb200_compile_to_memory(source.encode(), name.encode(), opts, len(options),
                                     C.byref(out), C.byref(size), log, len(log))
    if st != 0:
        msg = last_error()
        if st >= KB200_NVRTC_BASE and st < 20000:
            raise KB200CompileError("[KeOps-B200] NVRTC error %d\n%s\n%s"
                                    % (st - KB200_NVRTC_BASE, msg, log.value.decode("utf-8", "replace")))
        raise KB200Error(st, msg)
    try:
        return C.string_at(out, size.value)
    finally:
        lib.kb200_free(out)


class Module:
    """A cubin loaded on one device."""

    def __init__(self, cubin: bytes, device: int):
        self._h = C.c_void_p()
        self._cubin = cubin  # keep alive
        check(lib.kb200_module_load(cubin, len(cubin), int(device), C.byref(self._h)))
        self.device = int(device)

    def function(self, name: str) -> int:
        fn = C.c_void_p()
        check(lib.kb200_module_get_function(self._h, name.encode(), C.byref(fn)))
        return fn.value

    def unload(self):
        if self._h:
            lib.kb200_module_unload(self._h)
            self._h = C.c_void_p()


def function_info(fn: int) -> Tuple[int, int, int]:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib.kb200_function_info(fn, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def max_active_blocks(fn: int, block: int, dyn_smem: int = 0) -> int:
    v = C.c_int()
    check(lib.kb200_max_active_blocks(fn, block, dyn_smem, C.byref(v)))
    return v.value


def device_info(device: int):
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    l2, mem = C.c_int64(), C.c_int64()
    check(lib.kb200_device_info(device, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2), C.byref(mem)))
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), l2_bytes=l2.value, total_mem=mem.value)


def genred_launch(reduce_fn: int, combine_fn: Optional[int], stream: int, nrow: int, ncol: int,
                  rows_per_block: int, block: int, nsplit: int, col_chunk: int, out_ptr: int,
                  workspace_ptr: int, arg_ptrs: Sequence[int], nbatch: int = 1,
                  batch_strides: Optional[Sequence[int]] = None, dynamic_smem: int = 0):
    n = len(arg_ptrs)
    arr = (C.c_void_p * max(n, 1))(*arg_ptrs)
    bs = None if batch_strides is None else (C.c_int64 * max(n, 1))(*batch_strides)
    check(lib.kb200_genred_launch_smem(reduce_fn, combine_fn or None, stream or None, nrow, ncol, rows_per_block,
                                       block, nsplit, col_chunk, out_ptr, workspace_ptr or None, n, arr, nbatch, bs,
                                       int(dynamic_smem)))


def genred_launch_ranges(reduce_fn: int, stream: int, nrow: int, ncol: int, block: int, grid_x: int, out_ptr: int,
                         arg_ptrs: Sequence[int], ranges_x_ptr: int, slices_x_ptr: int, ranges_y_ptr: int,
                         block_ends_ptr: int, nranges: int):
    n = len(arg_ptrs)
    arr = (C.c_void_p * max(n, 1))(*arg_ptrs)
    check(lib.kb200_genred_launch_ranges(reduce_fn, stream or None, nrow, ncol, block, grid_x, out_ptr, n, arr,
                                         ranges_x_ptr, slices_x_ptr, ranges_y_ptr, block_ends_ptr, nranges))


def set_max_dynamic_smem(fn: int, nbytes: int):
    check(lib.kb200_function_set_max_dynamic_smem(fn, int(nbytes)))


class TensorMap(C.Structure):
    """Opaque 128-byte CUtensorMap, 64-byte aligned (pass by value as a __grid_constant__ kernel parameter)."""
    _fields_ = [("opaque", C.c_uint64 * 16)]
    _align_ = 64


def tensor_map_2d(base_ptr: int, elem_bytes: int, inner: int, outer: int, row_stride_bytes: int, box_inner: int,
                  box_outer: int, swizzle_bytes: int) -> "TensorMap":
    buf = C.create_string_buffer(128 + 64)
    addr = (C.addressof(buf) + 63) & ~63
    check(lib.kb200_tensor_map_2d(addr, elem_bytes, base_ptr, inner, outer, row_stride_bytes, box_inner, box_outer,
                                  swizzle_bytes))
    tm = TensorMap()
    C.memmove(C.addressof(tm), addr, 128)
    return tm


def launch(fn: int, stream: int, grid, block, smem: int, params: Sequence):
    """Generic launch; ``params`` is a sequence of ctypes objects (kept alive for the call)."""
    arr = (C.c_void_p * len(params))(*[C.cast(C.byref(p), C.c_void_p) for p in params])
    gx, gy, gz = (list(grid) + [1, 1])[:3]
    bx, by, bz = (list(block) + [1, 1])[:3]
    check(lib.kb200_launch(fn, stream or None, gx, gy, gz, bx, by, bz, smem, arr))
