"""apm_b200 -- Python (ctypes) binding of libapm_b200.so, the B200-native approximate-pattern-matching
hot path.  The binding is a thin mirror of include/apm_b200.h; every computation happens in the CUDA
kernels of the shared library.  There is no CPU fallback: importing works anywhere (so the C-ABI can be
inspected), but every compute call raises ApmError when the library or a CUDA device is missing.

Reference interfaces mirrored here (linomp/INF560-approximate-pattern-matching):
  count_matches(text, patterns, k)      <- the search loop of src/sequential.c:105-144
  count_matches_file(path, patterns, k) <- src/utils.c:12-68 read_input_file + the search loop
  Plan / Plan.count_device              <- invoke_kernel / write_kernel_result (patterns_over_ranks.c:33-36)
                                           and initializeGPU / getGPUResult (database_over_ranks.c:18-22)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libapm_b200.so")
CLI_PATH = os.path.join(PKG_DIR, "apm")

APM_OK, APM_EINVAL, APM_ENODEVICE, APM_ECUDA, APM_EIO, APM_ENOMEM = range(6)

# every symbol include/apm_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "apm_count_matches", "apm_count_matches_file", "apm_set_option", "apm_get_option", "apm_last_error",
    "apm_device_count", "apm_set_device", "apm_plan_create", "apm_plan_destroy", "apm_plan_count_device",
    "apm_plan_set_pattern_shard", "apm_plan_zero_counts", "apm_plan_counts_device_ptr",
    "apm_plan_read_counts", "apm_plan_max_pattern_len", "apm_synth_text_device", "apm_int_peak",
    "apm_launch_count", "apm_version", "apm_release_cache", "apm_find_matches", "apm_plan_set_hit_buffer",
    "apm_text_pack_bytes", "apm_text_pack_device", "apm_plan_count_device_packed",
]


class ApmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"apm_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libapm_b200.so (built in-tree by `make -C inf560-approximate-pattern-matching_b200`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ApmError(APM_ENODEVICE, f"{LIB_PATH} is not built; run __graft_entry__.build() "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ull, ll = C.c_void_p, C.c_ulonglong, C.c_longlong
    L.apm_count_matches.argtypes = [vp, C.c_size_t, vp, vp, C.c_int, C.c_int, vp]
    L.apm_count_matches_file.argtypes = [C.c_char_p, vp, vp, C.c_int, C.c_int, vp, vp]
    L.apm_set_option.argtypes = [C.c_char_p, C.c_char_p]
    L.apm_get_option.argtypes = [C.c_char_p]
    L.apm_get_option.restype = C.c_char_p
    L.apm_last_error.restype = C.c_char_p
    L.apm_device_count.argtypes = [vp]
    L.apm_set_device.argtypes = [C.c_int]
    L.apm_plan_create.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.apm_plan_destroy.argtypes = [vp]
    L.apm_plan_count_device.argtypes = [vp, vp, ull, ull, ull, ull, ull, vp]
    L.apm_plan_count_device_packed.argtypes = [vp, vp, vp, ull, ull, ull, ull, ull, vp]
    L.apm_text_pack_bytes.argtypes = [ull]
    L.apm_text_pack_bytes.restype = ull
    L.apm_text_pack_device.argtypes = [vp, ull, vp, vp]
    L.apm_plan_set_pattern_shard.argtypes = [vp, C.c_int, C.c_int]
    L.apm_plan_zero_counts.argtypes = [vp, vp]
    L.apm_find_matches.argtypes = [vp, C.c_size_t, vp, vp, C.c_int, C.c_int, vp, ull, vp, vp, vp]
    L.apm_plan_set_hit_buffer.argtypes = [vp, vp, ull, vp]
    L.apm_plan_counts_device_ptr.argtypes = [vp, vp]
    L.apm_plan_read_counts.argtypes = [vp, vp, vp]
    L.apm_plan_max_pattern_len.argtypes = [vp, vp]
    L.apm_synth_text_device.argtypes = [vp, ull, ull, ull, vp]
    L.apm_int_peak.argtypes = [C.c_int, vp, vp]
    L.apm_launch_count.restype = ull
    L.apm_version.restype = C.c_char_p
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != APM_OK:
        raise ApmError(rc, lib().apm_last_error().decode("utf-8", "replace"))


def _pattern_arrays(patterns: Sequence[bytes]):
    """char*[] / int[] views of the patterns: ONE flat host buffer + pointers into it (4096 patterns: ~1 ms)."""
    pats = [bytes(p) for p in patterns]
    n = len(pats)
    flat = C.create_string_buffer(b"".join(pats) + b"\0")
    base = C.addressof(flat)
    addrs, pos = [], 0
    for p in pats:
        addrs.append(base + pos)
        pos += len(p)
    ptrs = (C.c_void_p * max(n, 1))(*addrs)
    lens = (C.c_int * max(n, 1))(*[len(p) for p in pats])
    return flat, ptrs, lens, n


def set_option(key: str, value) -> None:
    _check(lib().apm_set_option(key.encode(), str(value).encode()))


def release_cache() -> None:
    """Hand the device memory the library caches between calls back to the driver."""
    _check(lib().apm_release_cache())


def get_option(key: str) -> str | None:
    v = lib().apm_get_option(key.encode())
    return None if v is None else v.decode()


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().apm_device_count(C.byref(n)))
    return n.value


def set_device(device: int) -> None:
    _check(lib().apm_set_device(device))


def launch_count() -> int:
    return int(lib().apm_launch_count())


def version() -> str:
    return lib().apm_version().decode()


def count_matches(text, patterns: Sequence[bytes], approx_factor: int) -> list[int]:
    """n_matches per pattern, HOST buffers in / out (the drop-in for sequential.c:105-144)."""
    tb = bytes(text) if not isinstance(text, (bytes, bytearray)) else text
    tbuf = (C.c_ubyte * max(len(tb), 1)).from_buffer_copy(tb if len(tb) else b"\0")
    bufs, ptrs, lens, n = _pattern_arrays(patterns)
    out = (C.c_longlong * max(n, 1))()
    _check(lib().apm_count_matches(C.addressof(tbuf) if len(tb) else None, len(tb), ptrs, lens, n,
                                   approx_factor, out))
    return [int(out[i]) for i in range(n)]


def find_matches(text, patterns: Sequence[bytes], approx_factor: int, max_hits: int = 1 << 20):
    """-> (n_matches per pattern, [(pattern index, window start), ...] sorted, total number of hits)."""
    tb = bytes(text) if not isinstance(text, (bytes, bytearray)) else text
    tbuf = (C.c_ubyte * max(len(tb), 1)).from_buffer_copy(tb if len(tb) else b"\0")
    bufs, ptrs, lens, n = _pattern_arrays(patterns)
    out = (C.c_longlong * max(n, 1))()
    hp = (C.c_int * max(max_hits, 1))()
    hs = (C.c_ulonglong * max(max_hits, 1))()
    nh = C.c_ulonglong(0)
    _check(lib().apm_find_matches(C.addressof(tbuf) if len(tb) else None, len(tb), ptrs, lens, n, approx_factor, out,
                                  max_hits, hp, hs, C.byref(nh)))
    k = min(int(nh.value), max_hits)
    return [int(out[i]) for i in range(n)], [(int(hp[i]), int(hs[i])) for i in range(k)], int(nh.value)


def count_matches_ptr(text_ptr: int, n_bytes: int, patterns: Sequence[bytes], approx_factor: int) -> list[int]:
    """Same, for a host buffer given by address (numpy / pinned torch tensor) -- no extra copy."""
    bufs, ptrs, lens, n = _pattern_arrays(patterns)
    out = (C.c_longlong * max(n, 1))()
    _check(lib().apm_count_matches(C.c_void_p(text_ptr), n_bytes, ptrs, lens, n, approx_factor, out))
    return [int(out[i]) for i in range(n)]


def count_matches_file(path: str, patterns: Sequence[bytes], approx_factor: int) -> list[int]:
    bufs, ptrs, lens, n = _pattern_arrays(patterns)
    out = (C.c_longlong * max(n, 1))()
    nbytes = C.c_ulonglong(0)
    _check(lib().apm_count_matches_file(os.fsencode(path), ptrs, lens, n, approx_factor, out, C.byref(nbytes)))
    return [int(out[i]) for i in range(n)]


class Plan:
    """Device-resident plan: Peq tables + pattern groups + 64-bit counters on the current CUDA device."""

    def __init__(self, patterns: Sequence[bytes], approx_factor: int):
        self._h = C.c_void_p(None)
        self._bufs, ptrs, lens, n = _pattern_arrays(patterns)
        self.nb_patterns = n
        self.approx_factor = approx_factor
        _check(lib().apm_plan_create(ptrs, lens, n, approx_factor, C.byref(self._h)))
        mm = C.c_int(0)
        _check(lib().apm_plan_max_pattern_len(self._h, C.byref(mm)))
        self.m_max = mm.value

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().apm_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_pattern_shard(self, rank: int, world: int) -> None:
        _check(lib().apm_plan_set_pattern_shard(self._h, rank, world))

    def zero_counts(self, stream: int = 0) -> None:
        _check(lib().apm_plan_zero_counts(self._h, C.c_void_p(stream)))

    def count_device(self, d_buf: int, buf_offset: int, buf_len: int, n_total: int, j_begin: int,
                     j_end: int, stream: int = 0) -> None:
        """Accumulate matches of window starts [j_begin, j_end); d_buf is a raw device address."""
        _check(lib().apm_plan_count_device(self._h, C.c_void_p(d_buf), buf_offset, buf_len, n_total,
                                           j_begin, j_end, C.c_void_p(stream)))

    def count_device_packed(self, d_buf: int, d_packed: int, buf_offset: int, buf_len: int, n_total: int, j_begin: int,
                            j_end: int, stream: int = 0) -> None:
        """count_device with the resident 2-bit copy of the same buffer (text_pack_device) at hand."""
        _check(lib().apm_plan_count_device_packed(self._h, C.c_void_p(d_buf), C.c_void_p(d_packed), buf_offset, buf_len,
                                                  n_total, j_begin, j_end, C.c_void_p(stream)))

    def counts_device_ptr(self) -> int:
        p = C.c_void_p(None)
        _check(lib().apm_plan_counts_device_ptr(self._h, C.byref(p)))
        return int(p.value or 0)

    def read_counts(self, stream: int = 0) -> list[int]:
        out = (C.c_longlong * max(self.nb_patterns, 1))()
        _check(lib().apm_plan_read_counts(self._h, out, C.c_void_p(stream)))
        return [int(out[i]) for i in range(self.nb_patterns)]


def synth_text_device(d_out: int, seed: int, offset: int, count: int, stream: int = 0) -> None:
    _check(lib().apm_synth_text_device(C.c_void_p(d_out), seed, offset, count, C.c_void_p(stream)))


def text_pack_bytes(buf_len: int) -> int:
    """bytes of the resident 2-bit copy of a text buffer of buf_len bytes"""
    return int(lib().apm_text_pack_bytes(buf_len))


def text_pack_device(d_buf: int, buf_len: int, d_packed: int, stream: int = 0) -> None:
    """d_packed <- 2-bit copy of the 16-byte aligned device buffer d_buf (for Plan.count_device_packed)"""
    _check(lib().apm_text_pack_device(C.c_void_p(d_buf), buf_len, C.c_void_p(d_packed), C.c_void_p(stream)))


def int_peak(kind: int = 0) -> tuple[float, float]:
    """(int32 lane-ops per second, seconds) of the dependency-free ALU microbenchmark."""
    ops, sec = C.c_double(0), C.c_double(0)
    _check(lib().apm_int_peak(kind, C.byref(ops), C.byref(sec)))
    return ops.value, sec.value
