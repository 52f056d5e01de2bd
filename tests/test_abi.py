"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/apm_b200.h, validates arguments, and fails loudly (no CPU fallback) without a CUDA device."""
import os
import re
import subprocess

import pytest

import apm_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "apm_b200.h")).read()
    declared = set(re.findall(r"\b(apm_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(apm_b200.EXPORTS), declared ^ set(apm_b200.EXPORTS)
    L = apm_b200.lib()
    for sym in declared:
        assert getattr(L, sym) is not None
    nm = subprocess.run(["nm", "-D", "--defined-only", apm_b200.LIB_PATH], capture_output=True, text=True).stdout
    for sym in declared:
        assert re.search(rf"\bT {sym}\b", nm), f"{sym} not exported as a C symbol"


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "inf560-approximate-pattern-matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower() or f.endswith((".cuh", ".cu")) and "import" not in src, f
                assert "libapm_oracle" not in src and "libapm_ref.so" not in src and "_ref/" not in src, f
    ldd = subprocess.run(["ldd", apm_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "apm_ref" not in ldd


def test_option_validation():
    apm_b200.set_option("kernel", "dp")
    assert apm_b200.get_option("kernel") == "dp"
    apm_b200.set_option("kernel", "sliced")
    assert apm_b200.get_option("kernel") == "sliced"
    apm_b200.set_option("kernel", "auto")
    apm_b200.set_option("shard", "DB_OVER_RANKS")
    assert apm_b200.get_option("shard") == "db"
    apm_b200.set_option("shard", "auto")
    for key, val in (("kernel", "cpu"), ("cell", "bogus"), ("reduce", "mpi"), ("text_chunk_mb", "0"), ("mode", "fast"), ("gpus", "0"), ("rblock", "3"), ("tile", "100"), ("nope", "1")):
        with pytest.raises(apm_b200.ApmError) as ei:
            apm_b200.set_option(key, val)
        assert ei.value.code == apm_b200.APM_EINVAL


def test_argument_errors_match_reference_contract():
    # approx_factor < 0 is UB in the reference (sequential.c:121); here it is an error
    with pytest.raises(apm_b200.ApmError) as ei:
        apm_b200.count_matches(b"ACGT", [b"AC"], -1)
    assert ei.value.code == apm_b200.APM_EINVAL
    # empty pattern: the reference exits with "Error while parsing argument" (sequential.c:64-67)
    with pytest.raises(apm_b200.ApmError) as ei:
        apm_b200.count_matches(b"ACGT", [b"AC", b""], 0)
    assert ei.value.code == apm_b200.APM_EINVAL


@pytest.mark.skipif(_has_cuda(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    with pytest.raises(apm_b200.ApmError) as ei:
        apm_b200.count_matches(b"ACGTACGT", [b"ACG"], 0)
    assert ei.value.code == apm_b200.APM_ENODEVICE
    with pytest.raises(apm_b200.ApmError):
        apm_b200.Plan([b"ACG"], 0)
    for mode in ("band", "filter"):  # the exact shortcut modes and the position output have no CPU path either
        apm_b200.set_option("mode", mode)
        try:
            with pytest.raises(apm_b200.ApmError) as ei:
                apm_b200.find_matches(b"ACGTACGTACGTACGT", [b"ACGTACGTAC"], 0)
            assert ei.value.code == apm_b200.APM_ENODEVICE
        finally:
            apm_b200.set_option("mode", "direct")
    # the reference's entry points by name: no device -> the error is reported and the caller's counter is untouched
    from apm_b200 import refcompat
    assert refcompat.device_count() == 0
    assert refcompat.invoke_and_fetch(b"ACGTACGT", b"ACG", 0, initial=5) == 5


def test_cli_usage_and_errors():
    # sequential.c:35-41: argc < 4 -> usage on stdout, exit status 1
    r = subprocess.run([apm_b200.CLI_PATH, "0", "x"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage: ")
    assert "approximation_factor dna_database pattern1 pattern2 ..." in r.stdout
    if not _has_cuda():
        # banner is printed before the text is read (sequential.c:79-84); then the failure is loud
        r = subprocess.run([apm_b200.CLI_PATH, "0", "/nonexistent.fa", "ACGT"], capture_output=True, text=True)
        assert r.returncode == 1
        assert r.stdout.startswith("Approximate Pattern Mathing: looking for 1 pattern(s) in file /nonexistent.fa w/ distance of 0")
