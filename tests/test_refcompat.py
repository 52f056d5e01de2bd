"""libapm_refcompat.so: the reference's GPU entry points (invoke_kernel / write_kernel_result,
initializeGPU / getGPUResult, getDeviceCount / setDevice, the two approaches) under their original names.
CPU part: the library loads and exports what include/apm_refcompat.h declares.  GPU part: each function
against the oracle's statement of what the reference function computes (patterns_over_ranks.cu:19-73,
database_over_ranks.cu:20-134), bit-exact."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from apm_b200 import refcompat
from oracle import oracle
from tests.golden_util import cases, fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = fixtures()
CASES = cases()


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "apm_refcompat.h")).read()
    body = hdr[hdr.index('extern "C" {'):]
    declared = set(re.findall(r"\b\*?(\w+)\s*\(", re.sub(r"/\*.*?\*/", "", body, flags=re.S)))
    declared = {d for d in declared if d not in ("defined",)}
    assert declared == set(refcompat.EXPORTS), declared ^ set(refcompat.EXPORTS)
    L = refcompat.lib()
    for sym in refcompat.EXPORTS:
        getattr(L, sym)


def test_compat_library_does_not_touch_the_oracle():
    out = subprocess.run(["ldd", refcompat.LIB_PATH], capture_output=True, text=True).stdout
    assert "libapm_b200.so" in out and "oracle" not in out and "apm_ref.so" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["config1_readme", "x100_k2", "small_k25", "easy_k2", "small_m200_k10"])
def test_invoke_kernel_matches_reference_loop(name):
    """invoke_kernel = the search loop of ONE pattern over buf[0, n_bytes), plus the caller's initial value."""
    case = next(c for c in CASES if c["name"] == name)
    text = FX[case["text"]]
    for p, want in list(zip(case["patterns"], case["expected"]))[:4]:
        assert refcompat.invoke_and_fetch(text, p, case["k"], initial=0) == want
        assert refcompat.invoke_and_fetch(text, p, case["k"], initial=7) == want + 7


@pytest.mark.gpu
def test_invoke_kernel_on_a_prefix_truncates_at_the_prefix_end():
    """patterns_over_ranks.c:316-327 hands the GPU the first 75 % of the text (+ m-1 ghost bytes): the windows are
    truncated at THAT end, exactly like ComputeMatches (patterns_over_ranks.cu:40-43)."""
    text = oracle.synth_text(0x5EED0001, 5, 60_000).tobytes()
    pat = text[44_970:45_000] + b"ACGTACGTAC"  # its 30-byte prefix ends exactly at the cut below
    cut = 45_000
    got = refcompat.invoke_and_fetch(text[:cut], pat, 2)
    assert got == oracle.count_matches(text[:cut], [pat], 2)[0]
    assert got >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("nb_ranks", [2, 4])
def test_initialize_gpu_matches_search_pattern_kernel(nb_ranks):
    """Every worker rank's piece as database_over_ranks.c:119-166 cuts it; per pattern the kernel's own range
    [start, end_i - k) with truncation at end_i = finish (+ m_i - 1 unless last rank)."""
    text = FX["small_chrY_x100"][:40_000]
    n = len(text)
    rng = np.random.default_rng(11)
    pats = []
    for m in (20, 50, 50, 64, 33, 7):
        off = int(rng.integers(0, n - m))
        pats.append(text[off:off + m])
    k = 2
    workers = nb_ranks - 1
    size_piece = n // workers
    for rank in range(1, nb_ranks):
        start = (rank - 1) * size_piece
        finish = n if rank == nb_ranks - 1 else rank * size_piece
        last_gpu = len(pats) // 2 + 1  # "first half of the patterns go to the GPU" (database_over_ranks.c:256-258)
        init = [3 * i for i in range(len(pats))]
        got = refcompat.initialize_and_fetch(text, pats, last_gpu, finish, rank, nb_ranks, start, k, init)
        want = list(init)
        for i in range(last_gpu):
            end = finish + (len(pats[i]) - 1 if rank != nb_ranks - 1 else 0)
            end = min(end, n)
            want[i] += oracle.count_range(text[:end], pats[i], k, start, end)
        assert got == want, rank


@pytest.mark.gpu
def test_device_utils():
    import torch
    assert refcompat.device_count() == torch.cuda.device_count()
    refcompat.lib().setDevice(1, refcompat.device_count())  # worker rank 1 -> device 0


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["patterns", "db"])
def test_approach_entry_points_print_the_reference_lines(which, tmp_path):
    case = next(c for c in CASES if c["name"] == "config1_readme")
    path = tmp_path / "text.fa"
    path.write_bytes(FX[case["text"]])
    argv = ["apm", str(case["k"]), str(path)] + [p.decode() for p in case["patterns"]]
    code = ("import sys; sys.path[:0] = %r; from apm_b200 import refcompat; "
            "rc0 = refcompat.run_approach(%r, %r, 0, 3); rc1 = refcompat.run_approach(%r, %r, 1, 3); "
            "sys.exit(rc0 * 10 + rc1)") % ([ROOT, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200")],
                                           which, argv, which, argv)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    counts = [int(m.group(2)) for m in (re.match(r"Number of matches for pattern <(.*)>: (\d+)$", l) for l in lines) if m]
    assert counts == case["expected"]
    assert any(l.startswith("(Rank 0) - TOTAL TIME using 3 mpi_ranks and 4 omp_thread(s) per rank:") for l in lines)
    banner = "Approximate Pattern Matching:" if which == "patterns" else "Approximate Pattern Mathing:"
    assert lines[0].startswith(banner)
    assert len(counts) == len(case["patterns"])  # rank 1 printed nothing


@pytest.mark.gpu
def test_invoke_kernel_text_cache_reuse_and_invalidation():
    """patterns_over_ranks.c:323 calls invoke_kernel once per pattern with the same broadcast text: the device copy is
    reused; new content at the same address (or another length) is noticed; several jobs may be in flight."""
    import ctypes as C
    L = refcompat.lib()
    n, k = 80_000, 2
    text = oracle.synth_text(0x5EED0001, 9, n).tobytes()
    pats = [text[1000:1032], text[40_000:40_050], text[79_000:79_064], b"ACGTTGCAACGTTGCAACGT"]
    want = oracle.count_matches(text, pats, k)
    tb = C.create_string_buffer(text, n)

    def invoke(buf, nbytes, p, local):
        pb = C.create_string_buffer(p, len(p))
        return L.invoke_kernel(C.addressof(buf), nbytes, C.addressof(pb), len(p), k, C.byref(local))

    # all four jobs in flight on the same text, results fetched afterwards (as an OpenMP team of callers would)
    locals_ = [C.c_int(3) for _ in pats]
    handles = [invoke(tb, n, p, loc) for p, loc in zip(pats, locals_)]
    for h, loc in zip(handles, locals_):
        L.write_kernel_result(C.byref(loc), h)
    assert [loc.value - 3 for loc in locals_] == want
    # same address, same length, different content in the middle and at the end
    text2 = bytearray(text)
    text2[40_000:40_050] = b"T" * 50
    text2[-64:] = b"A" * 64
    C.memmove(tb, bytes(text2), n)
    want2 = oracle.count_matches(bytes(text2), pats, k)
    got2 = []
    for p in pats:
        loc = C.c_int(0)
        L.write_kernel_result(C.byref(loc), invoke(tb, n, p, loc))
        got2.append(loc.value)
    assert got2 == want2 and want2 != want
    # a prefix of the same buffer
    loc = C.c_int(0)
    L.write_kernel_result(C.byref(loc), invoke(tb, 50_000, pats[0], loc))
    assert loc.value == oracle.count_matches(bytes(text2[:50_000]), [pats[0]], k)[0]
