"""Truncated tail windows (sequential.c:131-136: size = n_bytes - j < m, the pattern PREFIX is compared) through the
bit-parallel tail kernels of apm_tail.cuh, against the CPU oracle and against the explicit-DP tail kernel."""
import numpy as np
import pytest

import apm_b200
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _default_options():
    for k, v in (("kernel", "auto"), ("gpus", "1"), ("shard", "auto"), ("mode", "direct"), ("cell", "auto"),
                 ("tail", "auto")):
        apm_b200.set_option(k, v)
    yield
    apm_b200.set_option("tail", "auto")
    apm_b200.set_option("mode", "direct")


def _tail_case(rng, m, n, alphabet=b"ACGT", edits=2):
    """text whose END resembles the pattern's beginning: the pattern starts with (an edited copy of) the last
    `cut` bytes of the text, so several truncated windows are within the threshold."""
    text = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
    cut = int(rng.integers(1, min(m, n)))
    p = bytearray(text[n - cut:] + bytes(rng.choice(list(alphabet), size=m - cut).astype(np.uint8)))
    for _ in range(edits):
        p[int(rng.integers(0, cut))] = alphabet[int(rng.integers(0, len(alphabet)))]
    return text, bytes(p)


@pytest.mark.parametrize("m", [2, 5, 31, 32, 33, 50, 64, 65, 100, 200, 255, 256, 257, 300, 511, 512, 1000, 1024])
def test_tail_windows_vs_oracle(m):
    rng = np.random.default_rng(9100 + m)
    n = m + int(rng.integers(0, 40))
    pats, text = [], None
    text, p0 = _tail_case(rng, m, n)
    pats.append(p0)
    pats.append(text[-(m - 1):] + b"A")               # the text suffix itself: every shorter suffix matches too
    pats.append(bytes(rng.choice(list(b"ACGT"), size=m).astype(np.uint8)))
    for k in (0, 1, 3, 8):
        want = oracle.count_matches(text, pats, k)
        for mode in ("direct", "band", "filter"):
            apm_b200.set_option("mode", mode)
            for tail in ("bitpar", "dp"):
                apm_b200.set_option("tail", tail)
                assert apm_b200.count_matches(text, pats, k) == want, (m, k, mode, tail)


@pytest.mark.parametrize("seed", range(8))
def test_text_shorter_than_the_pattern(seed):
    """n_bytes < m: EVERY window is truncated (sizes n_bytes .. k+1)."""
    rng = np.random.default_rng(9200 + seed)
    m = int(rng.choice([20, 64, 130, 300, 700]))
    n = int(rng.integers(1, m))
    text, p = _tail_case(rng, m, n) if n > 1 else (b"A", b"A" * m)
    pats = [p, text + b"C" * (m - len(text))]
    for k in (0, 2, 5):
        want = oracle.count_matches(text, pats, k)
        for tail in ("bitpar", "dp"):
            apm_b200.set_option("tail", tail)
            assert apm_b200.count_matches(text, pats, k) == want, (seed, k, tail)


def test_tail_mixed_lengths_large_alphabet_and_positions():
    """mixed pattern lengths in one plan (both tail kernels in one call), bytes >= 0x80 and '\\n' in the alphabet,
    and the reported positions of the tail matches."""
    rng = np.random.default_rng(9300)
    alphabet = bytes([10, 65, 67, 71, 84, 78, 97, 200, 255])
    n = 5000
    text = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
    pats = []
    for m in (12, 40, 64, 100, 256, 300, 640, 1024):
        cut = int(rng.integers(m // 2, m))
        pats.append(text[n - cut:] + bytes(rng.choice(list(alphabet), size=m - cut).astype(np.uint8)))
    k = 3
    want = oracle.count_matches(text, pats, k)
    assert sum(want) >= len(pats)
    for tail in ("bitpar", "dp"):
        apm_b200.set_option("tail", tail)
        got, hits, nh = apm_b200.find_matches(text, pats, k)
        assert got == want, tail
        assert nh == sum(want) and len(set(hits)) == nh
        for p, j in hits:
            size = min(len(pats[p]), n - j)
            assert oracle.levenshtein(pats[p][:size], text[j:j + size]) <= k


def test_tail_db_shards_only_the_last_one_truncates():
    """window-start ranges ending before / inside / at the tail region: the tail kernels honour [j_begin, j_end)."""
    import torch
    rng = np.random.default_rng(9400)
    n, k = 20000, 2
    text, p = _tail_case(rng, 300, n)
    pats = [p, text[-63:] + b"G", text[-30:] + b"ACGTACGTAC"]
    want = oracle.count_matches(text, pats, k)
    d = torch.tensor(list(text), dtype=torch.uint8, device="cuda")
    with apm_b200.Plan(pats, k) as plan:
        for cuts in ([0, n], [0, n - 400, n - 299, n - 150, n - 10, n], [0, 7, n - 1, n]):
            plan.zero_counts()
            for a, b in zip(cuts[:-1], cuts[1:]):
                plan.count_device(d.data_ptr(), 0, n, n, a, b)
            assert plan.read_counts() == want, cuts
