"""ctypes front-end to the CPU checker (oracle/apm_oracle.c and, when built, oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; the product package never imports this module.

  count_matches(text, patterns, k)       our C restatement of sequential.c:105-144 + utils.c:76-99
  ref_count_matches(text, patterns, k)   the reference's own levenshtein() (oracle/_ref/libapm_ref.so)
  levenshtein_py / count_matches_py      pure-Python restatement for tiny cases
  synth_text(seed, offset, count)        SURVEY.md section 8d counter-based ACGT generator
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libapm_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libapm_ref.so")
REF_BINARY = os.path.join(_HERE, "_ref", "apm_sequential")

_lib = None
_ref = None


def build(ref: bool = True) -> None:
    """Compile the checker (gcc).  `ref` also rebuilds oracle/_ref when /root/reference exists."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def _flatten(patterns: Sequence[bytes]):
    lens = np.array([len(p) for p in patterns], dtype=np.int32)
    offs = np.zeros(len(patterns), dtype=np.int64)
    if len(patterns):
        offs[1:] = np.cumsum(lens[:-1], dtype=np.int64)
    flat = np.frombuffer(b"".join(bytes(p) for p in patterns) or b"\0", dtype=np.uint8)
    return flat, offs, lens


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        L = C.CDLL(_ORACLE_SO)
        L.apm_oracle_levenshtein.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.apm_oracle_levenshtein.restype = C.c_int
        L.apm_oracle_count.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.apm_oracle_count.restype = C.c_int
        L.apm_oracle_count_range.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                             C.c_longlong, C.c_longlong]
        L.apm_oracle_count_range.restype = C.c_longlong
        L.apm_oracle_distances.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p]
        L.apm_oracle_distances.restype = C.c_int
        L.apm_oracle_synth_text.argtypes = [C.c_ulonglong, C.c_longlong, C.c_longlong, C.c_void_p]
        L.apm_oracle_synth_text.restype = None
        L.apm_oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(_REF_SO)
        L.ref_count_matches.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_count_matches.restype = C.c_int
        L.ref_levenshtein.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_levenshtein.restype = C.c_int
        L.ref_max_threads.restype = C.c_int
        _ref = L
    return _ref


def max_threads() -> int:
    return int(lib().apm_oracle_max_threads())


def levenshtein(a: bytes, b: bytes, length: int | None = None) -> int:
    n = min(len(a), len(b)) if length is None else length
    ua, ub = _as_u8(a), _as_u8(b)
    return int(lib().apm_oracle_levenshtein(ua.ctypes.data, ub.ctypes.data, n))


def count_matches(text, patterns: Sequence[bytes], k: int, threads: int = 0) -> list[int]:
    """Our restatement of the reference search loop; threads=0 -> all host cores."""
    t = _as_u8(text)
    flat, offs, lens = _flatten(patterns)
    out = np.zeros(max(len(patterns), 1), dtype=np.int64)
    if threads <= 0:
        threads = max_threads()
    rc = lib().apm_oracle_count(t.ctypes.data if t.size else None, t.size, flat.ctypes.data,
                                offs.ctypes.data, lens.ctypes.data, len(patterns), k, threads,
                                out.ctypes.data)
    if rc != 0:
        raise ValueError(f"apm_oracle_count rc={rc}")
    return [int(x) for x in out[: len(patterns)]]


def count_range(text, pattern: bytes, k: int, j0: int, j1: int) -> int:
    """Matches among window starts [j0, j1) (clamped to n-k) of ONE pattern, global tail semantics."""
    t = _as_u8(text)
    p = _as_u8(pattern)
    return int(lib().apm_oracle_count_range(t.ctypes.data, t.size, p.ctypes.data, p.size, k, j0, j1))


def distances(text, pattern: bytes, k: int) -> np.ndarray:
    t = _as_u8(text)
    p = _as_u8(pattern)
    n = max(t.size - k, 0)
    out = np.zeros(max(n, 1), dtype=np.int32)
    lib().apm_oracle_distances(t.ctypes.data, t.size, p.ctypes.data, p.size, k, out.ctypes.data)
    return out[:n]


def ref_count_matches(text, patterns: Sequence[bytes], k: int, mode: int = 1,
                      threads: int = 0) -> list[int]:
    """The reference's own levenshtein() in the reference's loop (oracle/_ref)."""
    t = _as_u8(text)
    flat, offs, lens = _flatten(patterns)
    out = np.zeros(max(len(patterns), 1), dtype=np.int64)
    if threads <= 0:
        threads = int(ref().ref_max_threads())
    rc = ref().ref_count_matches(t.ctypes.data if t.size else None, t.size, flat.ctypes.data,
                                 offs.ctypes.data, lens.ctypes.data, len(patterns), k, mode, threads,
                                 out.ctypes.data)
    if rc != 0:
        raise ValueError(f"ref_count_matches rc={rc}")
    return [int(x) for x in out[: len(patterns)]]


def ref_levenshtein(a: bytes, b: bytes, length: int) -> int:
    ua, ub = _as_u8(a), _as_u8(b)
    return int(ref().ref_levenshtein(ua.ctypes.data, ub.ctypes.data, length))


# ---------------------------------------------------------------------------------------------
# Pure-Python restatement (small cases only) -- utils.c:76-99 and sequential.c:105-144.
# ---------------------------------------------------------------------------------------------
def levenshtein_py(a: bytes, b: bytes, length: int) -> int:
    col = list(range(length + 1))
    for c in range(1, length + 1):
        diag = col[0]
        col[0] = c
        for r in range(1, length + 1):
            up_left = diag
            diag = col[r]
            col[r] = min(up_left + (0 if a[r - 1] == b[c - 1] else 1), col[r] + 1, col[r - 1] + 1)
    return col[length]


def count_matches_py(text: bytes, patterns: Sequence[bytes], k: int) -> list[int]:
    n = len(text)
    out = []
    for p in patterns:
        hits = 0
        for j in range(0, max(n - k, 0)):
            size = min(len(p), n - j)
            if levenshtein_py(p, text[j:j + size], size) <= k:
                hits += 1
        out.append(hits)
    return out


# ---------------------------------------------------------------------------------------------
# Synthetic inputs
# ---------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def synth_text(seed: int, offset: int, count: int) -> np.ndarray:
    out = np.empty(max(count, 1), dtype=np.uint8)
    lib().apm_oracle_synth_text(seed, offset, count, out.ctypes.data)
    return out[:count]
