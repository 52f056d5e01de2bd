"""GPU parity tests: the CUDA path (through the C-ABI of libapm_b200.so) against the CPU oracle and the
committed golden vectors produced by the reference's own apm_sequential.  Bit-exact (integer counts)."""
import os
import re
import subprocess

import numpy as np
import pytest

import apm_b200
from oracle import oracle
from tests.golden_util import cases, fixtures

pytestmark = pytest.mark.gpu

FX = fixtures()
CASES = cases()


@pytest.fixture(autouse=True)
def _default_options():
    for k, v in (("kernel", "auto"), ("rblock", "auto"), ("tile", "auto"), ("gpus", "1"), ("shard", "auto"),
                 ("variant", "0"), ("mode", "direct"), ("cell", "auto"), ("reduce", "auto")):
        apm_b200.set_option(k, v)
    yield


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------
# golden vectors (reference apm_sequential outputs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["auto", "myers", "sliced"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_golden_host_api(case, kernel):
    """auto = window-sliced kernel for m <= 64 / small alphabets, row-parallel Myers kernel otherwise."""
    apm_b200.set_option("kernel", kernel)
    before = apm_b200.launch_count()
    got = apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"])
    assert got == case["expected"]
    assert apm_b200.launch_count() > before, "no CUDA kernel launched"


@pytest.mark.parametrize("mode", ["direct", "band"])
@pytest.mark.parametrize("cell", ["auto", "lop3", "fma3", "fma", "fma3r"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_golden_both_cell_codes(case, cell, mode):
    """The DP cell as 5 LOP3 (lop3), 4 LOP3 + 3 FMA-pipe subtractions (fma3), 4 LOP3 + 2 subtractions on the
    (minus, nonzero) / (plus, nonzero) delta encoding (fma), the 4 + 3 cell ordered for the operand reuse cache with
    v- / h- taken from x (fma3r); auto picks per pattern-length class."""
    apm_b200.set_option("kernel", "sliced")
    apm_b200.set_option("cell", cell)
    apm_b200.set_option("mode", mode)
    assert apm_b200.get_option("cell") == cell
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


@pytest.mark.parametrize("seed", range(6))
def test_cell_codes_agree_on_random_slabs(seed):
    """all DP-cell codes on 2 MB of random text, mixed lengths (1..300), direct and band mode."""
    rng = np.random.default_rng(4000 + seed)
    n = 2_000_000
    text = oracle.synth_text(0x5EED0001, 1000 * seed, n).tobytes()
    pats = []
    for m in (1, 7, 31, 32, 33, 50, 64, 65, 100, 128, 200, 300):
        off = int(rng.integers(0, n - m))
        p = bytearray(text[off:off + m])
        for s in range(int(rng.integers(0, 6))):
            pos = int(rng.integers(0, m))
            p[pos] = ord("ACGT"[int(rng.integers(0, 4))])
        pats.append(bytes(p))
    k = int(rng.integers(0, 7))
    out = {}
    for mode in ("direct", "band"):
        for cell in ("lop3", "fma3", "fma", "fma3r"):
            apm_b200.set_option("mode", mode)
            apm_b200.set_option("cell", cell)
            out[(mode, cell)] = apm_b200.count_matches(text, pats, k)
    ref = out[("direct", "lop3")]
    assert sum(ref) > 0
    for key, v in out.items():
        assert v == ref, key


@pytest.mark.parametrize("name", ["config1_readme", "easy_k2", "small_k25", "small_m200_k10", "config2"])
def test_golden_dp_kernel(name):
    """The explicit-DP fallback kernel alone reproduces the reference too (north_star item 2)."""
    case = next(c for c in CASES if c["name"] == name)
    apm_b200.set_option("kernel", "dp")
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


@pytest.mark.parametrize("rblock", ["1", "2", "4"])
@pytest.mark.parametrize("name", ["config1_readme", "x100_m64_k4", "small_m200_k10", "small_k2"])
def test_golden_every_register_blocking(name, rblock):
    case = next(c for c in CASES if c["name"] == name)
    apm_b200.set_option("kernel", "myers")
    apm_b200.set_option("rblock", rblock)
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("rblock", ["1", "2", "4"])
@pytest.mark.parametrize("name", ["config1_readme", "x100_m64_k4", "small_m200_k10", "small_k2", "x100_k10"])
def test_golden_every_step_variant(name, rblock, variant):
    """The FMA-pipe formulations of the column step (myers_step_fma) are bit-identical."""
    case = next(c for c in CASES if c["name"] == name)
    apm_b200.set_option("kernel", "myers")
    apm_b200.set_option("rblock", rblock)
    apm_b200.set_option("variant", variant)
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("seed", range(12))
def test_random_vs_oracle_step_variants(seed, variant):
    rng = np.random.default_rng(7000 + seed)
    text, pats, k = _random_case(rng)
    apm_b200.set_option("kernel", "myers")
    apm_b200.set_option("variant", variant)
    assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


# ---------------------------------------------------------------------------------------------
# exact band mode (Ukkonen band): bit-identical to the full evaluation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_golden_band_mode(case):
    apm_b200.set_option("mode", "band")
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


@pytest.mark.parametrize("seed", range(40))
def test_random_vs_oracle_band_mode(seed):
    rng = np.random.default_rng(5000 + seed)
    text, pats, k = _random_case(rng)
    apm_b200.set_option("mode", "band")
    assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 17])
def test_band_mode_every_band_width(k):
    """Every instantiated band half-width (and k = 17: falls back to the full kernel), distances around k."""
    n = 60_000
    base = bytearray(oracle.synth_text(0x5EED0001, 1234, n).tobytes())
    rng = np.random.default_rng(k)
    pats = []
    for idx, m in enumerate((40, 64, 100, 200)):
        p = bytes(base[3000 * (idx + 1):3000 * (idx + 1) + m])
        pats.append(p)
        for e in range(0, k + 3):  # plant copies with e edits (substitutions, an insertion+deletion pair)
            q = bytearray(p)
            for _ in range(e):
                q[int(rng.integers(0, m))] = int(rng.choice(list(b"ACGT")))
            if e >= 2:
                pos = int(rng.integers(1, m - 2))
                del q[pos]
                q.insert(int(rng.integers(1, m - 2)), ord("A"))
            off = 20000 + 2000 * idx + 300 * e
            base[off:off + m] = q
    text = bytes(base)
    apm_b200.set_option("kernel", "dp")
    want = apm_b200.count_matches(text, pats, k)
    apm_b200.set_option("kernel", "auto")
    assert apm_b200.count_matches(text, pats, k) == want
    apm_b200.set_option("mode", "band")
    assert apm_b200.count_matches(text, pats, k) == want
    apm_b200.set_option("mode", "filter")
    assert apm_b200.count_matches(text, pats, k) == want
    assert sum(want) > 0


# ---------------------------------------------------------------------------------------------
# exact filter mode (pigeonhole seeds + banded verification): bit-identical counts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_golden_filter_mode(case):
    apm_b200.set_option("mode", "filter")
    assert apm_b200.get_option("mode") == "filter"
    before = apm_b200.launch_count()
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]
    assert apm_b200.launch_count() > before


@pytest.mark.parametrize("seed", range(40))
def test_random_vs_oracle_filter_mode(seed):
    rng = np.random.default_rng(5000 + seed)
    text, pats, k = _random_case(rng)
    apm_b200.set_option("mode", "filter")
    assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


def _edited_copies_case(rng, n, lengths, k):
    """text with copies of every pattern carrying 0..k+2 random edits (substitutions, insertions, deletions)."""
    base = bytearray(oracle.synth_text(0x5EED0001, int(rng.integers(0, 1 << 30)), n).tobytes())
    pats = []
    pos = 1000
    for m in lengths:
        off = int(rng.integers(0, n - m))
        p = bytes(base[off:off + m])
        pats.append(p)
        for e in range(0, k + 3):
            q = bytearray(p)
            for _ in range(e):
                op = int(rng.integers(0, 3))
                x = int(rng.integers(0, len(q)))
                if op == 0:
                    q[x] = int(rng.choice(list(b"ACGT")))
                elif op == 1:
                    q.insert(x, int(rng.choice(list(b"ACGT"))))
                elif len(q) > 1:
                    del q[x]
            if pos + len(q) + 50 < n:
                base[pos:pos + len(q)] = q
                pos += len(q) + int(rng.integers(1, 40))
    return bytes(base), pats


@pytest.mark.parametrize("seed", range(12))
def test_filter_mode_indels_and_shifts(seed):
    """Copies with insertions / deletions make the verbatim piece sit at a SHIFTED offset inside the window (|shift|
    <= k); windows near a planted copy are witnessed by several (piece, shift) pairs and must be counted once."""
    rng = np.random.default_rng(9100 + seed)
    k = int(rng.integers(0, 7))
    lengths = [int(x) for x in rng.integers(8 * (k + 1), 140, size=5)] + [12, 30]
    text, pats = _edited_copies_case(rng, 120_000, lengths, k)
    want = oracle.count_matches(text, pats, k)
    assert sum(want) >= len(lengths)
    for mode in ("filter", "band", "direct"):
        apm_b200.set_option("mode", mode)
        assert apm_b200.count_matches(text, pats, k) == want, mode


@pytest.mark.parametrize("cand_mb", ["1", "128"])
def test_filter_mode_low_complexity_text_overflows_to_the_band_kernel(cand_mb):
    """Poly-A / short-period text: every position is a seed hit of every repeat pattern.  With a 1 MiB candidate
    buffer the scan overflows and the band kernel takes the round; the counts stay exact either way."""
    n = 300_000
    text = (b"A" * 100_000) + (b"ACAC" * 25_000) + oracle.synth_text(0x5EED0001, 3, 100_000).tobytes()
    pats = [b"A" * 64, b"ACAC" * 16, b"A" * 31 + b"C" + b"A" * 32, text[250_000:250_064], b"CACA" * 10]
    k = 3
    apm_b200.set_option("mode", "band")
    want = apm_b200.count_matches(text, pats, k)
    assert want[0] > 90_000 and want[1] > 40_000
    apm_b200.set_option("mode", "filter")
    apm_b200.set_option("filter_cand_mb", cand_mb)
    try:
        assert apm_b200.count_matches(text, pats, k) == want
    finally:
        apm_b200.set_option("filter_cand_mb", "128")
    assert want[:2] == oracle.count_matches(text[:n], pats[:2], k)


def test_filter_mode_shards_and_binary_alphabet():
    """Window-range calls (database shards) add up in filter mode; byte patterns outside ACGT use the same path."""
    rng = np.random.default_rng(77)
    n = 400_000
    text = bytes(rng.integers(0, 256, size=n, dtype=np.uint8))
    pats = [text[1000:1064], text[200_000:200_100], text[399_900:399_990], bytes(rng.integers(0, 256, size=80, dtype=np.uint8))]
    k = 4
    apm_b200.set_option("mode", "filter")
    whole = apm_b200.count_matches(text, pats, k)
    assert whole == oracle.count_matches(text, pats, k)
    assert whole[:3] == [5, 5, 5]  # a shift by 1 or 2 positions costs 2 or 4 edits: 5 windows per planted copy at k = 4
    torch = _torch()
    dev = torch.tensor(np.frombuffer(text, dtype=np.uint8), device="cuda")
    with apm_b200.Plan(pats, k) as plan:
        cuts = [0, 1001, 1064, 200_050, 399_950, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            plan.count_device(dev.data_ptr(), 0, n, n, a, b)
        assert plan.read_counts() == whole


def test_cli_drop_in(tmp_path):
    """`apm` prints what sequential.c:79-82,151,157-160 prints (timing value aside)."""
    case = next(c for c in CASES if c["name"] == "config1_readme")
    f = tmp_path / "small_chrY_x100.fa"
    f.write_bytes(FX[case["text"]])
    argv = [apm_b200.CLI_PATH, str(case["k"]), str(f)] + [p.decode() for p in case["patterns"]]
    r = subprocess.run(argv, capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert lines[0] == (f"Approximate Pattern Mathing: looking for {len(case['patterns'])} pattern(s) "
                        f"in file {f} w/ distance of {case['k']}")
    assert re.fullmatch(r"APM done in \d+\.\d{6} s", lines[1])
    want = [f"Number of matches for pattern <{p.decode()}>: {n}" for p, n in zip(case["patterns"], case["expected"])]
    assert lines[2:] == want
    # trailing approach flag of the parallel binary (main.c:66-85) is accepted and not searched for
    r2 = subprocess.run(argv + ["DB_OVER_RANKS"], capture_output=True, text=True, check=True)
    assert r2.stdout.splitlines()[2:] == want
    r3 = subprocess.run(argv + ["PATTERNS_OVER_RANKS"], capture_output=True, text=True, check=True)
    assert r3.stdout.splitlines()[2:] == want


def test_file_api(tmp_path):
    case = next(c for c in CASES if c["name"] == "x100_k2")
    f = tmp_path / "t.fa"
    f.write_bytes(FX[case["text"]])
    assert apm_b200.count_matches_file(str(f), case["patterns"], case["k"]) == case["expected"]
    with pytest.raises(apm_b200.ApmError) as ei:
        apm_b200.count_matches_file(str(tmp_path / "missing.fa"), [b"A"], 0)
    assert ei.value.code == apm_b200.APM_EIO


# ---------------------------------------------------------------------------------------------
# edge semantics S4-S8 of SURVEY.md
# ---------------------------------------------------------------------------------------------
EDGE = [
    (b"ACGTACGTAC", [b"AC"], 2),             # m <= k: every start matches (n - k)
    (b"ACGTACGTAC", [b"AC"], 10),            # k >= n: no window
    (b"ACGTACGTAC", [b"AC"], 11),
    (b"GGGGGGAC", [b"ACGT"], 0),             # suffix == pattern prefix counts (tail truncation)
    (b"ACG", [b"ACGTTTTT"], 0),              # m > n
    (b"ACG", [b"ACGTTTTT", b"A" * 300], 1),  # long pattern on a tiny text
    (b"", [b"ACGT"], 0),                     # empty text
    (b"A", [b"A", b"C", b"AA"], 0),
    (bytes(range(256)) * 3, [bytes(range(100, 140)), bytes([255, 0, 1]), b"\n\n"], 1),  # all byte values
    (b"AC\nGT\nAC\nGT\n" * 20, [b"C\nG", b"\nAC\nGT\nAC\nGT\nAC\nGT\nAC\nGT\nAC\nG", b"GT\n"], 1),
]


@pytest.mark.parametrize("idx", range(len(EDGE)))
@pytest.mark.parametrize("kernel", ["auto", "myers", "dp"])
def test_edge_cases(idx, kernel):
    text, pats, k = EDGE[idx]
    apm_b200.set_option("kernel", kernel)
    assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


def test_alphabet_with_all_256_byte_values():
    rng = np.random.default_rng(3)
    text = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    pats = [bytes(range(256)), text[100:140], text[4000:4033], bytes(rng.integers(0, 256, 70, dtype=np.uint8))]
    for k in (0, 3):
        assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


# ---------------------------------------------------------------------------------------------
# randomized property test: GPU == oracle
# ---------------------------------------------------------------------------------------------
def _random_case(rng, max_n=3000):
    alph = [b"ACGT", b"ACGT\n", b"AC", b"ACGTNacgt\n", bytes(range(256))][int(rng.integers(0, 5))]
    al = np.frombuffer(alph, dtype=np.uint8)
    n = int(rng.integers(0, max_n))
    text = al[rng.integers(0, len(al), n)].tobytes()
    pats = []
    for _ in range(int(rng.integers(1, 9))):
        m = int(rng.choice([1, 2, 5, 16, 31, 32, 33, 50, 63, 64, 65, 96, 100, 128, 129, 200, 256, 257, 300]))
        if n > m and rng.random() < 0.7:
            off = int(rng.integers(0, n - m))
            p = bytearray(text[off:off + m])
            for _ in range(int(rng.integers(0, 6))):
                p[int(rng.integers(0, m))] = int(al[int(rng.integers(0, len(al)))])
        else:
            p = bytearray(al[rng.integers(0, len(al), m)].tobytes())
        pats.append(bytes(p))
    if n > 10 and rng.random() < 0.5:  # a pattern whose prefix is the text's suffix
        cut = int(rng.integers(1, min(n, 60)))
        pats.append(text[-cut:] + al[rng.integers(0, len(al), int(rng.integers(1, 40)))].tobytes())
    k = int(rng.choice([0, 0, 1, 2, 4, 10, 40]))
    return text, pats, k


@pytest.mark.parametrize("kernel", ["auto", "myers"])
@pytest.mark.parametrize("seed", range(40))
def test_random_vs_oracle(seed, kernel):
    rng = np.random.default_rng(1000 + seed)
    text, pats, k = _random_case(rng)
    apm_b200.set_option("kernel", kernel)
    assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k)


def test_myers_vs_dp_kernel_large_slice():
    """Kernel-vs-kernel on a slice far larger than the oracle can do in seconds."""
    text = oracle.synth_text(0x5EED0001, 12345, 400_000).tobytes()
    pats = [text[1000:1064], text[5000:5200], text[70000:70032], b"ACGT" * 16, text[-50:] + b"TTTT"]
    k = 6
    apm_b200.set_option("kernel", "dp")
    dp = apm_b200.count_matches(text, pats, k)
    for kernel in ("myers", "sliced", "auto"):
        apm_b200.set_option("kernel", kernel)
        assert apm_b200.count_matches(text, pats, k) == dp, kernel


@pytest.mark.parametrize("m", [65, 100, 128, 129, 200, 256, 257, 500, 1000, 1024])
def test_long_patterns_sliced_column_blocks_vs_dp(m):
    """m > 64: the sliced kernel chains ceil(m/64) column blocks through the boundary scratch."""
    text = bytearray(oracle.synth_text(0x5EED0001, 31337, 30_000).tobytes())
    rng = np.random.default_rng(m)
    p0 = bytearray(text[7000:7000 + m])
    for _ in range(5):  # a copy with 5 substitutions elsewhere in the text
        p0[int(rng.integers(0, m))] = ord("A")
    text[20000:20000 + m] = p0
    text = bytes(text)
    pats = [bytes(text[7000:7000 + m]), bytes(p0), oracle.synth_text(0x5EED0003, m, m).tobytes(),
            text[-(m // 2):] + b"ACGT" * (m // 8 + 1)]
    k = 6
    apm_b200.set_option("kernel", "dp")
    want = apm_b200.count_matches(text, pats, k)
    assert want[0] >= 2 or m > 1000
    for kernel in ("sliced", "auto"):
        apm_b200.set_option("kernel", kernel)
        assert apm_b200.count_matches(text, pats, k) == want, kernel
    if m <= 300:
        assert oracle.count_matches(text, pats[:2], k) == want[:2]


# ---------------------------------------------------------------------------------------------
# device-resident API: shards with halo, unaligned buffers, tiles
# ---------------------------------------------------------------------------------------------
def test_synth_text_device_matches_oracle_generator():
    torch = _torch()
    for off, cnt in ((0, 1000), (7, 4099), (123456789012, 65536 + 5)):
        buf = torch.empty(cnt + 32, dtype=torch.uint8, device="cuda")
        apm_b200.synth_text_device(buf.data_ptr() + 3, 0x5EED0001, off, cnt)  # unaligned destination
        torch.cuda.synchronize()
        got = buf[3:3 + cnt].cpu().numpy().tobytes()
        assert got == oracle.synth_text(0x5EED0001, off, cnt).tobytes()


@pytest.mark.parametrize("kernel", ["auto", "myers"])
@pytest.mark.parametrize("nshards,misalign", [(1, 0), (2, 1), (3, 5), (7, 15), (8, 0)])
def test_db_shards_with_halo_are_exact(nshards, misalign, kernel):
    """Sum over shards == unsharded; each shard sees only its own bytes + (m_max-1)-byte halo, at an
    arbitrary (unaligned) device address; seams fall on / next to planted matches."""
    torch = _torch()
    n = 150_000
    text = bytearray(oracle.synth_text(0x5EED0001, 999, n).tobytes())
    pat64 = bytes(text[20000:20064])
    pat200 = bytes(text[90000:90200])
    W0 = n - 3
    for g in range(1, nshards):  # plant an exact copy straddling every seam
        seam = W0 * g // nshards
        text[seam - 20:seam - 20 + 64] = pat64
    text = bytes(text)
    pats = [pat64, pat200, text[50:82], text[-30:] + b"ACGTAC"]
    k = 3
    want = oracle.count_matches(text, pats, k)
    W = n - k
    apm_b200.set_option("kernel", kernel)
    with apm_b200.Plan(pats, k) as plan:
        for g in range(nshards):
            j0, j1 = W * g // nshards, W * (g + 1) // nshards
            b0, b1 = j0, min(n, j1 + plan.m_max - 1)
            host = torch.frombuffer(bytearray(text[b0:b1]), dtype=torch.uint8)
            dev = torch.empty(b1 - b0 + 64, dtype=torch.uint8, device="cuda")
            dev[misalign:misalign + (b1 - b0)].copy_(host)
            plan.count_device(dev.data_ptr() + misalign, b0, b1 - b0, n, j0, j1)
        assert plan.read_counts() == want
        # too small a buffer (halo missing) is rejected, not silently truncated
        if nshards > 1:
            j0, j1 = 0, W // nshards
            dev = torch.zeros(j1, dtype=torch.uint8, device="cuda")
            with pytest.raises(apm_b200.ApmError):
                plan.count_device(dev.data_ptr(), 0, j1, n, j0, j1)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_pattern_shards_sum_to_whole(world):
    torch = _torch()
    text = oracle.synth_text(0x5EED0001, 5, 60_000).tobytes()
    pats = [text[i * 997:i * 997 + m] for i, m in enumerate([32, 64, 64, 50, 200, 20, 64, 33, 100, 64, 300])]
    k = 2
    want = oracle.count_matches(text, pats, k)
    dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    total = [0] * len(pats)
    for r in range(world):
        with apm_b200.Plan(pats, k) as plan:
            plan.set_pattern_shard(r, world)
            plan.count_device(dev.data_ptr(), 0, len(text), len(text), 0, len(text))
            part = plan.read_counts()
        assert all(part[i] == 0 for i in range(len(pats)) if i % world != r)
        total = [a + b for a, b in zip(total, part)]
    assert total == want


@pytest.mark.parametrize("tile", ["256", "512", "2048", "4096"])
def test_tile_sizes(tile):
    case = next(c for c in CASES if c["name"] == "x100_k5")
    apm_b200.set_option("kernel", "myers")
    apm_b200.set_option("tile", tile)
    assert apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"]) == case["expected"]


def test_counts_accumulate_and_zero():
    torch = _torch()
    text = FX["small_chrY"]
    dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    pats = [b"ACGT", b"TTTT"]
    one = oracle.count_matches(text, pats, 1)
    with apm_b200.Plan(pats, 1) as plan:
        for _ in range(3):
            plan.count_device(dev.data_ptr(), 0, len(text), len(text), 0, len(text))
        assert plan.read_counts() == [3 * x for x in one]
        plan.zero_counts()
        plan.count_device(dev.data_ptr(), 0, len(text), len(text), 0, len(text))
        assert plan.read_counts() == one


# ---------------------------------------------------------------------------------------------
# BASELINE-sized inputs: size-independent properties + sampled-slice oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["auto", "myers"])
def test_large_synthetic_sampled_slices(kernel):
    """64 MiB of the config-3 text, 32 patterns of length 64, k = 4: per-slice counts equal the oracle's
    on random slices + the tail slice; planted patterns are found; sharding does not change anything."""
    torch = _torch()
    from tests.synth import make_patterns
    n = 64 << 20
    seed = 0x5EED0001
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    apm_b200.synth_text_device(dev.data_ptr(), seed, 0, n)
    pats, offs, nsub = make_patterns(seed, n, 32, 64, 7)
    k = 4
    W = n - k
    apm_b200.set_option("kernel", kernel)
    with apm_b200.Plan(pats, k) as plan:
        plan.count_device(dev.data_ptr(), 0, n, n, 0, W)
        whole = plan.read_counts()
        for p in range(32):  # patterns cut from the text with <= k substitutions must be found
            if offs[p] is not None and nsub[p] <= k:
                assert whole[p] >= 1, p
        plan.zero_counts()
        cuts = [0, W // 3 + 1, (2 * W) // 3 + 7, W]
        for a, b in zip(cuts[:-1], cuts[1:]):
            plan.count_device(dev.data_ptr(), 0, n, n, a, b)
        assert plan.read_counts() == whole
        # sampled slices against the oracle: windows [a, a+L) of the global text
        rng = np.random.default_rng(5)
        L = 20000
        starts = [0, W - L] + [int(x) for x in rng.integers(0, W - L, 3)]
        starts += [int(o) - 50 for o in offs[:3] if o is not None and o > 50]
        for a in starts:
            plan.zero_counts()
            plan.count_device(dev.data_ptr(), 0, n, n, a, a + L)
            got = plan.read_counts()
            seg_end = min(n, a + L + 63)
            seg = oracle.synth_text(seed, a, seg_end - a).tobytes()
            for p in (0, 1, 5, 13, 31):
                if seg_end == n:  # slice touches the global end: tail truncation applies
                    want = oracle.count_range(seg, pats[p], k, 0, L)
                else:  # interior slice: pad so the oracle does not truncate at the slice end
                    want = oracle.count_range(seg + b"\0" * 64, pats[p], k, 0, L)
                assert got[p] == want, (a, p)


# ---------------------------------------------------------------------------------------------
# one-process-per-GPU host layer (apm_b200/dist.py) with the CUDA counter
# ---------------------------------------------------------------------------------------------
def test_dist_layer_cuda_counter_single_rank():
    from apm_b200 import dist as adist
    text = oracle.synth_text(0x5EED0001, 4242, 50_000)
    pats = [text[100:164].tobytes(), text[30000:30050].tobytes(), text[-25:].tobytes() + b"ACGT", b"ACGTACGTAC"]
    k = 2
    want = oracle.count_matches(text.tobytes(), pats, k)
    got = adist.count_matches_distributed(lambda off, cnt: text[off:off + cnt], len(text), pats, k)
    assert got == want
    # emulate 3 DB shards / 3 pattern shards in one process: partial sums add up to the whole
    m_max = max(len(p) for p in pats)
    total = [0] * len(pats)
    for r in range(3):
        j0, j1, b0, b1 = adist.db_shard(len(text), k, m_max, r, 3)
        part = adist.cuda_counter(text[b0:b1], b0, len(text), j0, j1, pats, k)
        total = [a + b for a, b in zip(total, part)]
    assert total == want
    total = [0] * len(pats)
    for r in range(3):
        part = adist.cuda_counter(text, 0, len(text), 0, len(text) - k, pats, k, pattern_shard=(r, 3))
        total = [a + b for a, b in zip(total, part)]
    assert total == want


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from apm_b200 import dist as adist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    text = oracle.synth_text(0x5EED0001, 99, 300_000)
    pats = [text[1000:1064].tobytes(), text[150000:150064].tobytes(), text[-30:].tobytes() + b"ACGTAC"]
    out = {}
    for shard in ("db", "patterns"):
        out[shard] = adist.count_matches_distributed(lambda off, cnt: text[off:off + cnt], len(text), pats, 3,
                                                     shard=shard, device=torch.device("cuda", rank))
    q.put((rank, out))
    dist.destroy_process_group()


def test_dist_layer_nccl_two_gpus():
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    text = oracle.synth_text(0x5EED0001, 99, 300_000)
    pats = [text[1000:1064].tobytes(), text[150000:150064].tobytes(), text[-30:].tobytes() + b"ACGTAC"]
    apm_b200.set_option("kernel", "dp")
    want = apm_b200.count_matches(text.tobytes(), pats, 3)
    for rank, out in res:
        assert out["db"] == want and out["patterns"] == want


def test_one_shot_api_two_gpus_single_process():
    """gpus=2 inside ONE process (the C launcher drives both devices): DB shards and pattern shards."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    text = oracle.synth_text(0x5EED0001, 2024, 3_000_000).tobytes()
    pats = [text[i * 40009:i * 40009 + m] for i, m in enumerate([64, 64, 50, 32, 200, 64, 100])] + [text[-33:] + b"ACGT"]
    k = 3
    apm_b200.set_option("gpus", "1")
    want = apm_b200.count_matches(text, pats, k)
    _, want_hits, want_n = apm_b200.find_matches(text, pats, k)
    apm_b200.set_option("gpus", "2")
    for reduce in ("auto", "p2p", "nccl", "host"):  # peer-memory kernel / NCCL all-reduce over NVLink / host-side sum
        apm_b200.set_option("reduce", reduce)
        for shard in ("db", "patterns", "auto"):
            apm_b200.set_option("shard", shard)
            assert apm_b200.count_matches(text, pats, k) == want, (reduce, shard)
    apm_b200.set_option("reduce", "auto")
    for mode in ("filter", "band"):  # the exact shortcut modes and the match positions shard the same way
        apm_b200.set_option("mode", mode)
        for shard in ("db", "patterns"):
            apm_b200.set_option("shard", shard)
            counts, hits, n_hits = apm_b200.find_matches(text, pats, k)
            assert counts == want and hits == want_hits and n_hits == want_n, (mode, shard)


def test_device_memory_cache_release_and_limits():
    """The library reuses device blocks between calls; releasing the cache or disabling it changes nothing."""
    text = oracle.synth_text(0x5EED0001, 77, 400_000).tobytes()
    pats = [text[1000:1064], text[5000:5200], text[-20:] + b"ACGTTT"]
    want = apm_b200.count_matches(text, pats, 3)
    assert apm_b200.count_matches(text, pats, 3) == want
    apm_b200.release_cache()
    assert apm_b200.count_matches(text, pats, 3) == want
    apm_b200.set_option("cache_mb", "0")
    try:
        assert apm_b200.count_matches(text, pats, 3) == want
        assert apm_b200.count_matches(text, pats, 3) == want
    finally:
        apm_b200.set_option("cache_mb", "4096")
    apm_b200.set_option("kernel", "dp")
    assert apm_b200.count_matches(text[:50_000], pats, 3) == oracle.count_matches(text[:50_000], pats, 3)


@pytest.mark.parametrize("mode", ["direct", "band"])
def test_file_ingest_pipeline_multi_chunk(tmp_path, mode):
    """apm_count_matches_file streams the file through 32 MiB pinned chunks and counts the windows that are
    complete while the next chunk is read: same counts as the host-buffer call, with patterns planted ACROSS the
    chunk seams (32 MiB, 64 MiB) and at the very end of the file (truncated tail windows)."""
    n = (72 << 20) + 12345
    text = oracle.synth_text(0x5EED0001, 4242, n).tobytes()
    chunk = 32 << 20
    pats = [text[chunk - 30:chunk + 34], text[2 * chunk - 1:2 * chunk + 63], text[chunk - 199:chunk + 1],
            text[5_000_000:5_000_050], text[-40:] + b"ACGTACGTACGT", text[chunk + 7:chunk + 39]]
    f = tmp_path / "big.fa"
    f.write_bytes(text)
    apm_b200.set_option("mode", mode)
    k = 3
    want = apm_b200.count_matches(text, pats, k)
    assert all(w >= 1 for w in want)
    assert apm_b200.count_matches_file(str(f), pats, k) == want


# ---------------------------------------------------------------------------------------------
# match positions: every kernel reports WHERE it counted
# ---------------------------------------------------------------------------------------------
def _oracle_positions(text, pats, k):
    out = []
    for p, pat in enumerate(pats):
        d = oracle.distances(text, pat, k)
        out += [(p, int(j)) for j in np.nonzero(d <= k)[0]]
    return out


@pytest.mark.parametrize("mode,kernel", [("direct", "auto"), ("band", "auto"), ("filter", "auto"), ("direct", "myers"),
                                         ("direct", "dp")])
def test_find_matches_positions(mode, kernel):
    """(pattern, window start) of every match, incl. shifted neighbours of planted copies and truncated tail windows."""
    rng = np.random.default_rng(31)
    text, pats = _edited_copies_case(rng, 90_000, [64, 40, 100, 33, 200], 3)
    pats.append(text[-30:] + b"ACGTACGTAC")  # prefix == text suffix: truncated tail windows match
    k = 3
    apm_b200.set_option("mode", mode)
    apm_b200.set_option("kernel", kernel)
    counts, hits, n_hits = apm_b200.find_matches(text, pats, k)
    want = _oracle_positions(text, pats, k)
    assert counts == oracle.count_matches(text, pats, k)
    assert n_hits == len(want) == sum(counts)
    assert hits == want
    assert any(j > len(text) - 40 for p, j in hits if p == len(pats) - 1)


def test_find_matches_capacity_and_golden():
    case = next(c for c in CASES if c["name"] == "x100_k2")
    counts, hits, n_hits = apm_b200.find_matches(FX[case["text"]], case["patterns"], case["k"], max_hits=50)
    assert counts == case["expected"] and n_hits == sum(case["expected"]) and len(hits) == 50
    assert hits == sorted(hits)
    counts, hits, n_hits = apm_b200.find_matches(FX[case["text"]], case["patterns"], case["k"], max_hits=0)
    assert counts == case["expected"] and hits == [] and n_hits == 0


@pytest.mark.parametrize("mode", ["direct", "filter"])
def test_one_shot_api_streams_large_shards_in_segments(tmp_path, mode):
    """text_chunk_mb = 1: a 5.3 MB text goes through two 1 MiB (+halo) device buffers, six segments; patterns planted
    across every segment seam, match positions, host-buffer and file entry points."""
    n = 5_300_000
    text = oracle.synth_text(0x5EED0001, 808, n).tobytes()
    seg = 1 << 20
    pats = [text[s * seg - 20:s * seg + 44] for s in range(1, 6)] + [text[seg - 1:seg + 199], text[-50:] + b"ACGTACGT"]
    k = 2
    apm_b200.set_option("mode", mode)
    want, want_hits, want_n = apm_b200.find_matches(text, pats, k)
    assert all(w >= 1 for w in want)
    f = tmp_path / "t.fa"
    f.write_bytes(text)
    apm_b200.set_option("text_chunk_mb", "1")
    try:
        assert apm_b200.count_matches(text, pats, k) == want
        assert apm_b200.find_matches(text, pats, k) == (want, want_hits, want_n)
        assert apm_b200.count_matches_file(str(f), pats, k) == want
    finally:
        apm_b200.set_option("text_chunk_mb", "32768")


def test_cli_pattern_file_positions_and_info(tmp_path):
    """CLI extras that stay out of argv: APM_PATTERN_FILE (one pattern per line), APM_POSITIONS, APM_INFO, APM_MODE."""
    case = next(c for c in CASES if c["name"] == "x100_k2")
    f = tmp_path / "t.fa"
    f.write_bytes(FX[case["text"]])
    pf = tmp_path / "patterns.txt"
    pf.write_bytes(b"\n".join(case["patterns"][1:]) + b"\n\n")
    argv = [apm_b200.CLI_PATH, str(case["k"]), str(f), case["patterns"][0].decode(), "DB_OVER_RANKS"]
    for mode in ("filter", "direct"):
        env = dict(os.environ, APM_PATTERN_FILE=str(pf), APM_POSITIONS="5", APM_INFO="1", APM_MODE=mode, OMP_NUM_THREADS="3")
        r = subprocess.run(argv, capture_output=True, text=True, env=env, check=True)
        lines = r.stdout.splitlines()
        counts = [int(l.rsplit(": ", 1)[1]) for l in lines if l.startswith("Number of matches for pattern <")]
        assert counts == case["expected"], mode
        assert any(l.startswith("(Rank 0) - TOTAL TIME using 1 mpi_ranks and 3 omp_thread(s) per rank:") for l in lines)
        assert sum(l.startswith("Match of pattern <") for l in lines) == 5
    r = subprocess.run([apm_b200.CLI_PATH, "1", str(f)], capture_output=True, text=True,
                       env=dict(os.environ, APM_PATTERN_FILE=str(pf)), check=True)
    assert sum(l.startswith("Number of matches") for l in r.stdout.splitlines()) == len(case["patterns"]) - 1


def test_adversarial_inputs_all_modes_agree_with_the_dp_kernel():
    """tools/stress_modes.py: tiny alphabets, periodic / run-length / self-copying texts, indels at piece borders,
    k <= 12: filter (also with an overflowing candidate buffer), band and direct mode == explicit-DP kernel, counts
    and positions."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.sys.executable, os.path.join(root, "tools", "stress_modes.py"), "60", "20261018"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "60 cases, 0 mismatches" in r.stdout


@pytest.mark.parametrize("mode", ["band", "filter"])
def test_band_kernel_too_wide_for_shared_memory_falls_back(mode):
    """8 pattern symbols x m = 300 x k = 16: the band kernel's U table (rows of 32 + 2K windows) exceeds 227 KB, the
    launcher must take the direct kernel instead (found by tools/stress_modes.py)."""
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGTNacg", dtype=np.uint8)
    text = alpha[rng.integers(0, 8, size=30_000)].tobytes()
    pats = [text[1000:1300], text[5000:5290], text[9000:9100]]
    for k in (12, 16):
        apm_b200.set_option("mode", "direct"); apm_b200.set_option("kernel", "dp")
        want = apm_b200.count_matches(text, pats, k)
        apm_b200.set_option("kernel", "auto"); apm_b200.set_option("mode", mode)
        assert apm_b200.count_matches(text, pats, k) == want


# ---------------------------------------------------------------------------------------------
# threaded host: the one-shot C-ABI call is re-entrant
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["direct", "filter"])
def test_concurrent_host_threads(mode):
    """Four host threads call apm_count_matches at the same time (ctypes releases the GIL), each with its own text
    and pattern set; every result must equal the serial oracle.  Exercises the device-block cache, the pooled
    pinned buffers, the per-kernel shared-memory opt-in table and the option snapshot under contention."""
    import threading

    apm_b200.set_option("mode", mode)
    jobs = []
    for t in range(4):
        rng = np.random.default_rng(9100 + t)
        n = 300_000 + 50_000 * t
        text = oracle.synth_text(0x5EED0001, 7777 * t, n).tobytes()
        pats = []
        for m in (20, 32, 50, 64, 64, 100, 200)[: 4 + t % 4]:
            off = int(rng.integers(0, n - m))
            p = bytearray(text[off:off + m])
            for s in range(int(rng.integers(0, 4))):
                p[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
            pats.append(bytes(p))
        pats.append(text[-30:] + b"ACGTACGTAC")  # truncated tail windows
        jobs.append((text, pats, 3))
    want = [oracle.count_matches(text, pats, k) for text, pats, k in jobs]
    got = [None] * len(jobs)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = apm_b200.count_matches(*jobs[i])
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    assert got == want


# ---------------------------------------------------------------------------------------------
# automatic kernel routing of the direct mode: every list class against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cell", ["auto", "fma3r"])
def test_auto_routing_all_length_classes(cell):
    """auto: m = 32 -> two-row sweep (RG 3), ragged m < 32 / 33..63 -> compile-time widths (RG 1), m = 64 and 65..224 ->
    CELL 3, longer -> CELL 0 -- one pattern set that populates every list, planted copies with substitutions and
    indels so that hits exist, k = 0, 3, 5; fma3r: the same set through the generic kernel with CELL 3."""
    apm_b200.set_option("kernel", "sliced")
    apm_b200.set_option("cell", cell)
    rng = np.random.default_rng(31337)
    n = 150_000
    text = bytearray(oracle.synth_text(0x5EED0001, 4242, n).tobytes())
    pats = []
    for m in (32, 64, 64, 32, 50, 64, 100, 64, 32, 64, 20, 31, 33, 63, 65, 128, 200, 224, 225, 300):
        off = int(rng.integers(0, n - m - 8))
        p = bytearray(text[off:off + m + 4])
        for s_ in range(int(rng.integers(0, 5))):
            pos = int(rng.integers(0, m))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                p[pos] = ord("ACGT"[int(rng.integers(0, 4))])
            elif kind == 1:
                del p[pos]
            else:
                p.insert(pos, ord("ACGT"[int(rng.integers(0, 4))]))
        pats.append(bytes(p[:m]))
    pats.append(bytes(text[-64:]))
    pats.append(bytes(text[-32:]))
    for k in (0, 3, 5):
        got = apm_b200.count_matches(bytes(text), pats, k)
        assert got == oracle.count_matches(bytes(text), pats, k), (cell, k)


def test_all_m32_and_all_m64_sets():
    """pattern sets of a single length (the lists the two-row sweep and the m = 64 kernel take alone), odd pattern
    counts, a text shorter than one tile and one that ends inside a tile"""
    rng = np.random.default_rng(99)
    for n in (3000, 70_001):
        text = oracle.synth_text(0x5EED0001, 11 * n, n).tobytes()
        for m in (32, 64):
            pats = []
            for i in range(7):
                off = int(rng.integers(0, n - m))
                p = bytearray(text[off:off + m])
                for s_ in range(i % 4):
                    p[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
                pats.append(bytes(p))
            for k in (0, 2, 4):
                assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k), (n, m, k)


@pytest.mark.parametrize("alphabet", [b"ACGTN", b"ACGTNMRY"])
def test_auto_routing_larger_alphabets(alphabet):
    """5 and 8 symbol planes: the U table leaves two / one CTA per SM, where the automatic cell choice falls back to
    plain LOP3 for multi-block (two CTAs) or all 64-column lists (one CTA); every list class against the oracle"""
    apm_b200.set_option("kernel", "sliced")
    rng = np.random.default_rng(len(alphabet))
    n = 120_000
    arr = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)]
    text = arr.tobytes()
    pats = []
    for m in (32, 32, 64, 64, 20, 50, 63, 65, 128, 200, 224, 225, 300):
        off = int(rng.integers(0, n - m))
        p = bytearray(text[off:off + m])
        for s_ in range(int(rng.integers(0, 4))):
            p[int(rng.integers(0, m))] = alphabet[int(rng.integers(0, len(alphabet)))]
        pats.append(bytes(p))
    pats.append(text[-50:] + alphabet[:4] * 3)
    for k in (0, 3):
        assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k), k
