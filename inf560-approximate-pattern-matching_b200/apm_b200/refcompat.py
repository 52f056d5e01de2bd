"""ctypes view of libapm_refcompat.so -- the reference's GPU entry points under their original names
(include/apm_refcompat.h).  Used by the tests; a C caller simply links the library."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_HERE, "libapm_refcompat.so")
EXPORTS = ["invoke_kernel", "write_kernel_result", "initializeGPU", "getGPUResult", "getDeviceCount", "setDevice",
           "apm_patterns_over_ranks_hybrid", "apm_database_over_ranks"]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing -- run `make -C {_HERE}`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.invoke_kernel.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp]
        L.invoke_kernel.restype = vp
        L.write_kernel_result.argtypes = [vp, vp]
        L.write_kernel_result.restype = None
        L.initializeGPU.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.getGPUResult.argtypes = [C.c_int]
        L.getGPUResult.restype = C.POINTER(C.c_int)
        L.getDeviceCount.argtypes = [vp]
        L.getDeviceCount.restype = None
        L.setDevice.argtypes = [C.c_int, C.c_int]
        L.setDevice.restype = None
        for name in ("apm_patterns_over_ranks_hybrid", "apm_database_over_ranks"):
            getattr(L, name).argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def invoke_and_fetch(buf: bytes, pattern: bytes, approx_factor: int, initial: int = 0) -> int:
    """invoke_kernel + write_kernel_result exactly as patterns_over_ranks.c:323,380 call them."""
    L = lib()
    tb = C.create_string_buffer(buf, len(buf))
    pb = C.create_string_buffer(pattern, len(pattern))
    local = C.c_int(initial)
    handle = L.invoke_kernel(C.addressof(tb), len(buf), C.addressof(pb), len(pattern), approx_factor, C.byref(local))
    L.write_kernel_result(C.byref(local), handle)
    return local.value


def initialize_and_fetch(buf: bytes, patterns: Sequence[bytes], last_gpu_pattern: int, finish_without_extra: int,
                         my_rank: int, nb_ranks: int, start: int, approx_factor: int,
                         initial: Sequence[int] | None = None) -> list[int]:
    """initializeGPU + getGPUResult exactly as database_over_ranks.c:273,561 call them."""
    L = lib()
    n = len(patterns)
    tb = C.create_string_buffer(buf, len(buf))
    pbufs = [C.create_string_buffer(p, len(p)) for p in patterns]
    pptr = (C.c_void_p * max(n, 1))(*[C.addressof(b) for b in pbufs])
    sizes = (C.c_int * max(n, 1))(*[len(p) for p in patterns])
    init = (C.c_int * max(n, 1))(*(list(initial) if initial is not None else [0] * n))
    rc = L.initializeGPU(C.addressof(tb), len(buf), pptr, n, last_gpu_pattern, sizes, finish_without_extra, my_rank,
                         nb_ranks, start, approx_factor, init)
    assert rc == 1
    res = L.getGPUResult(n)
    out = [int(res[i]) for i in range(n)]
    C.CDLL(None).free(res)
    return out


def device_count() -> int:
    n = C.c_int(-1)
    lib().getDeviceCount(C.byref(n))
    return n.value


def run_approach(which: str, argv: Sequence[str], rank: int = 0, world: int = 2) -> int:
    """apm_patterns_over_ranks_hybrid / apm_database_over_ranks with a C argv (prints on stdout like the reference)."""
    L = lib()
    enc = [a.encode() for a in argv]
    arr = (C.c_char_p * (len(enc) + 1))(*enc, None)
    fn = L.apm_patterns_over_ranks_hybrid if which == "patterns" else L.apm_database_over_ranks
    return int(fn(len(enc), arr, rank, world, 1))
