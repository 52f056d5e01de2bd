"""Pins the CPU oracle (oracle/apm_oracle.c) to the reference: README golden, golden.json produced by
the reference's apm_sequential, levenshtein() known answers, and -- when oracle/_ref is present --
live random comparison with the reference's own levenshtein()."""
import numpy as np
import pytest

from oracle import oracle
from tests.golden_util import cases, fixtures, lev_cases

FX = fixtures()
CASES = cases()


def test_readme_golden_config1():
    # /root/reference/README.md:58-63
    c = next(c for c in CASES if c["name"] == "config1_readme")
    assert c["expected"] == [0, 4, 4, 4, 4, 4]
    assert oracle.count_matches(FX[c["text"]], c["patterns"], c["k"]) == [0, 4, 4, 4, 4, 4]


@pytest.mark.parametrize("case", [c for c in CASES if c["name"] != "config1_readme"],
                         ids=lambda c: c["name"])
def test_oracle_matches_reference_binary(case):
    got = oracle.count_matches(FX[case["text"]], case["patterns"], case["k"])
    assert got == case["expected"]


def test_levenshtein_known_answers():
    for a, b, ln, d in lev_cases():
        assert oracle.levenshtein(a, b, ln) == d


def test_pure_python_restatement_agrees():
    for a, b, ln, d in lev_cases()[:150]:
        if ln <= 60:
            assert oracle.levenshtein_py(a, b, ln) == d
    c = next(c for c in CASES if c["name"] == "easy_k2")
    assert oracle.count_matches_py(FX["easy"], c["patterns"], 2) == c["expected"]


def test_serial_equals_threaded():
    text = FX["small_chrY"]
    pats = [b"ACGT", b"TTTGTCCCGG", text[-7:], text[-30:] + b"GG"]
    for k in (0, 1, 3):
        assert oracle.count_matches(text, pats, k, threads=1) == oracle.count_matches(text, pats, k, threads=4)


def test_edge_semantics():
    t = b"ACGTACGTAC"
    # S6: m <= k -> every window start in [0, n-k) matches
    assert oracle.count_matches(t, [b"AC"], 2) == [len(t) - 2]
    # k >= n -> no windows at all
    assert oracle.count_matches(t, [b"AC"], 10) == [0]
    assert oracle.count_matches(t, [b"AC"], 11) == [0]
    # S5: tail truncation -- a text suffix equal to a pattern prefix counts
    assert oracle.count_matches(b"GGGGGGAC", [b"ACGT"], 0) == [1 + 0]  # "AC" vs prefix "AC" at j=6; j=7 'C' vs 'A' no
    # m > n: every window truncated
    assert oracle.count_matches(b"ACG", [b"ACGTTTTT"], 0) == [1]
    # empty text
    assert oracle.count_matches(b"", [b"ACGT"], 0) == [0]


def test_synth_text_generator():
    a = oracle.synth_text(0x5EED0001, 0, 1000)
    b = oracle.synth_text(0x5EED0001, 500, 500)
    assert bytes(a[500:]) == bytes(b)
    assert set(bytes(a)) <= set(b"ACGT")
    ref = bytes(b"ACGT"[oracle.splitmix64(0x5EED0001 + i) >> 62] for i in range(64))
    assert bytes(a[:64]) == ref


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (reference tree absent)")
def test_live_against_reference_levenshtein():
    rng = np.random.default_rng(1)
    alph = np.frombuffer(b"ACGT\n", dtype=np.uint8)
    for _ in range(300):
        n = int(rng.integers(1, 400))
        m = int(rng.integers(1, 70))
        k = int(rng.integers(0, 6))
        text = alph[rng.integers(0, 5, n)].tobytes()
        pats = [alph[rng.integers(0, 5, m)].tobytes(), text[n // 2:n // 2 + m] or b"A"]
        assert oracle.count_matches(text, pats, k, threads=1) == oracle.ref_count_matches(text, pats, k, mode=0)
    # the OpenMP work splits of the harness agree with the serial loop
    text = FX["small_chrY"]
    pats = [b"ACGT", b"TTTGTCCCGGTCCC", b"GGGGACC"]
    base = oracle.ref_count_matches(text, pats, 2, mode=0)
    assert oracle.ref_count_matches(text, pats, 2, mode=1, threads=4) == base
    assert oracle.ref_count_matches(text, pats, 2, mode=2, threads=4) == base


def test_numpy_text_generator_equals_oracle_generator():
    from apm_b200.synth import make_patterns, text_slice
    for off, cnt in ((0, 257), (2**34 - 100, 100), (123456789, 4096)):
        assert text_slice(0x5EED0001, off, cnt).tobytes() == oracle.synth_text(0x5EED0001, off, cnt).tobytes()
    pats, offs, nsub = make_patterns(0x5EED0001, 1 << 20, 16, 64, 7)
    assert len(pats) == 16 and all(len(p) == 64 for p in pats)
    t = oracle.synth_text(0x5EED0001, offs[0], 64).tobytes()
    assert pats[0] == t and nsub[0] == 0
    t3 = oracle.synth_text(0x5EED0001, offs[3], 64).tobytes()
    assert sum(a != b for a, b in zip(pats[3], t3)) == 3
