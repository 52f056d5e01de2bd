"""Exact filter mode through the 2-bit q-gram scan + seed-hit verification (apm_dna.cuh): bit-identical to the
oracle, to the reference goldens, to the hashed scan and to the band kernel -- every probe stride H = 1..4, q-gram
lengths 8..10, indels (shifted witnesses), text bytes outside ACGT (they alias in the 2-bit code), unaligned
buffers, window-range calls, candidate overflow."""
import numpy as np
import pytest

import apm_b200
from oracle import oracle
from tests.golden_util import cases, fixtures
from tests.test_gpu_parity import _edited_copies_case

pytestmark = pytest.mark.gpu

FX = fixtures()
CASES = cases()


@pytest.fixture(autouse=True)
def _options():
    for k, v in (("kernel", "auto"), ("gpus", "1"), ("shard", "auto"), ("cell", "auto"), ("tail", "auto"),
                 ("filter_cand_mb", "128")):
        apm_b200.set_option(k, v)
    apm_b200.set_option("mode", "filter")
    apm_b200.set_option("filter_scan", "dna")
    yield
    apm_b200.set_option("filter_scan", "auto")
    apm_b200.set_option("mode", "direct")
    apm_b200.set_option("filter_cand_mb", "128")


def _is_dna(pats):
    return all(set(p) <= set(b"ACGT") for p in pats)


@pytest.mark.parametrize("case", [c for c in CASES if _is_dna(c["patterns"])], ids=lambda c: c["name"])
def test_golden_dna_scan(case):
    """reference apm_sequential outputs; the fixture texts hold '\\n', lower case and N: those bytes alias to one
    of the four 2-bit codes, which may only add false seed hits"""
    got = apm_b200.count_matches(FX[case["text"]], case["patterns"], case["k"])
    assert got == case["expected"]


def test_dna_scan_is_refused_for_other_alphabets():
    with pytest.raises(apm_b200.ApmError) as ei:
        apm_b200.count_matches(b"ACGTNNNNACGTACGTACGT" * 10, [b"ACGTNNNNACGT"], 0)
    assert ei.value.code == apm_b200.APM_EINVAL
    apm_b200.set_option("filter_scan", "auto")  # auto: the hashed scan takes such patterns
    text = b"ACGTNNNNACGTACGTACGT" * 10
    assert apm_b200.count_matches(text, [b"ACGTNNNNACGT"], 0) == oracle.count_matches(text, [b"ACGTNNNNACGT"], 0)


# (m, k) -> piece length l = m // (k+1), q = min(l, 10), H = min(4, l - q + 1)
@pytest.mark.parametrize("m,k", [(8, 0), (9, 0), (10, 0), (11, 0), (12, 0), (13, 0), (40, 0), (16, 1), (21, 1), (22, 1), (27, 1),
                                 (64, 4), (65, 4), (64, 3), (64, 5), (64, 7), (50, 2), (32, 2), (200, 10), (200, 16),
                                 (136, 16), (300, 5), (1000, 8), (1024, 16)])
def test_every_stride_and_qgram_length(m, k):
    rng = np.random.default_rng(1000 * m + k)
    if k > 15:  # the DNA scan maps the 2k+1 shifts to the lanes of a warp: k = 16 stays on the hashed scan
        apm_b200.set_option("filter_scan", "auto")
    n = 60_000 + int(rng.integers(0, 97))
    text, pats = _edited_copies_case(rng, n, [m, m, m], k)
    pats.append(text[-m:])                      # last full window
    pats.append(text[:m])                       # first window
    if m <= 200:
        want = oracle.count_matches(text, pats, k)
    else:  # the CPU oracle would need minutes: the band kernel (checked against it elsewhere) is the reference
        apm_b200.set_option("mode", "band")
        want = apm_b200.count_matches(text, pats, k)
        apm_b200.set_option("mode", "filter")
    assert sum(want) >= 5
    assert apm_b200.count_matches(text, pats, k) == want
    apm_b200.set_option("filter_scan", "hash")
    assert apm_b200.count_matches(text, pats, k) == want


@pytest.mark.parametrize("seed", range(16))
def test_indels_shifts_and_mixed_lengths(seed):
    rng = np.random.default_rng(4400 + seed)
    k = int(rng.integers(0, 9))
    lengths = [int(x) for x in rng.integers(8 * (k + 1), 8 * (k + 1) + 90, size=6)]
    text, pats = _edited_copies_case(rng, 150_000, lengths, k)
    want = oracle.count_matches(text, pats, k)
    assert sum(want) >= len(lengths)
    assert apm_b200.count_matches(text, pats, k) == want
    got, hits, nh = apm_b200.find_matches(text, pats, k)
    assert got == want and nh == sum(want) and len(set(hits)) == nh
    for p, j in hits[:200]:
        size = min(len(pats[p]), len(text) - j)
        assert oracle.levenshtein(pats[p][:size], text[j:j + size]) <= k


@pytest.mark.parametrize("seed", range(6))
def test_text_with_newlines_lowercase_and_n(seed):
    """FASTA-like text: '\\n' every 61st byte, runs of N, lower-case stretches; patterns stay pure ACGT"""
    rng = np.random.default_rng(4500 + seed)
    n, k = 200_000, int(rng.integers(0, 5))
    t = bytearray(oracle.synth_text(0x5EED0001, 17 * seed, n).tobytes())
    for i in range(60, n, 61):
        t[i] = 10
    t[5000:5400] = b"N" * 400
    t[9000:9900] = bytes(t[9000:9900]).lower()
    pats = []
    for m in (8 * (k + 1), 64, 100, 150):
        for off in (int(rng.integers(0, n - 200)), 4990 - m // 2, 8990 - m // 2, 59):
            p = bytearray(t[off:off + m])
            pats.append(bytes(x if x in b"ACGT" else ord("A") for x in p))
    text = bytes(t)
    want = oracle.count_matches(text, pats, k)
    assert apm_b200.count_matches(text, pats, k) == want


def test_many_patterns_dense_bitmap_and_second_level():
    """4096 patterns of length 64 at k = 4 (the config-5 pattern set): 6 % of the q-gram bitmap is set, so the second
    level (queue, CSR look-up, packed compare) carries real load; compared with the band kernel on 48 MiB."""
    import torch
    from tests.synth import TEXT_SEED, make_patterns
    n = 48 << 20
    pats, offs, nsub = make_patterns(TEXT_SEED, n, 4096, 64, 7)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    apm_b200.synth_text_device(dev.data_ptr(), TEXT_SEED, 0, n)
    with apm_b200.Plan(pats, 4) as plan:
        plan.count_device(dev.data_ptr(), 0, n, n, 0, n)
        got = plan.read_counts()
    apm_b200.set_option("mode", "band")
    with apm_b200.Plan(pats, 4) as plan:
        plan.count_device(dev.data_ptr(), 0, n, n, 0, n)
        want = plan.read_counts()
    assert got == want
    assert sum(want) >= sum(1 for p in range(4096) if offs[p] is not None and nsub[p] <= 4)


@pytest.mark.parametrize("misalign", [0, 1, 7, 15])
def test_window_ranges_and_unaligned_buffers(misalign):
    import torch
    rng = np.random.default_rng(4600 + misalign)
    n, k = 300_000, 3
    text, pats = _edited_copies_case(rng, n, [40, 64, 64, 90, 130], k)
    want = oracle.count_matches(text, pats, k)
    dev = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    dev[misalign:misalign + n] = torch.tensor(np.frombuffer(text, dtype=np.uint8), device="cuda")
    ptr = dev.data_ptr() + misalign
    with apm_b200.Plan(pats, k) as plan:
        for cuts in ([0, n], [0, 1, 1003, 1064, 150_001, 299_000, n]):
            plan.zero_counts()
            for a, b in zip(cuts[:-1], cuts[1:]):
                plan.count_device(ptr, 0, n, n, a, b)
            assert plan.read_counts() == want, cuts
        # a shard: the buffer starts at the first window start and ends with the halo
        plan.zero_counts()
        cuts = [0, 100_000 + misalign, 200_003, n]
        mmax = max(len(p) for p in pats)
        for a, b in zip(cuts[:-1], cuts[1:]):
            e = min(n, b + mmax - 1)
            plan.count_device(ptr + a, a, e - a, n, a, b)
        assert plan.read_counts() == want


@pytest.mark.parametrize("cand_mb", ["1", "128"])
def test_low_complexity_text_overflows_to_the_band_kernel(cand_mb):
    text = (b"A" * 100_000) + (b"ACAC" * 25_000) + oracle.synth_text(0x5EED0001, 3, 100_000).tobytes()
    pats = [b"A" * 64, b"ACAC" * 16, b"A" * 31 + b"C" + b"A" * 32, text[250_000:250_064], b"CACA" * 10]
    k = 3
    apm_b200.set_option("mode", "band")
    want = apm_b200.count_matches(text, pats, k)
    assert want[0] > 90_000 and want[1] > 40_000
    apm_b200.set_option("mode", "filter")
    apm_b200.set_option("filter_cand_mb", cand_mb)
    assert apm_b200.count_matches(text, pats, k) == want


def test_tiny_texts():
    for text, pats, k in ((b"ACGTACGT", [b"ACGTACGT"], 0), (b"ACGTACGTA", [b"ACGTACGT"], 0), (b"ACGTACG", [b"ACGTACGT"], 0),
                          (b"A" * 40, [b"A" * 16, b"A" * 17], 1), (b"ACGTACGTACGTACGTACGT", [b"CGTACGTACGTACGTA"], 1)):
        assert apm_b200.count_matches(text, pats, k) == oracle.count_matches(text, pats, k), (text, pats, k)


# ---------------------------------------------------------------------------------------------------------------
# resident 2-bit copy of the text (apm_text_pack_device / apm_plan_count_device_packed)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,P,m,k", [(3_000_001, 64, 64, 4), (777_777, 16, 32, 2), (1_234_567, 8, 200, 10), (70_000, 5, 50, 0),
                                     (4099, 3, 40, 1)])
def test_packed_text_counts_equal_plain(n, P, m, k):
    """every probe-stride class; text lengths that are not multiples of 16 or of the scan's regions; sub-ranges of window
    starts; a text with bytes outside ACGT (they alias in the 2-bit code, the byte-exact stage decides)"""
    import torch
    rng = np.random.default_rng(n + m)
    base = bytearray(oracle.synth_text(0x5EED0001, 3 * n, n).tobytes())
    for pos in rng.integers(0, n, size=n // 5000):  # sprinkle newlines / N / lower case
        base[int(pos)] = int(rng.choice(list(b"\nNacgt")))
    pats = []
    for i in range(P):
        off = int(rng.integers(0, n - m))
        p = bytearray(bytes(base[off:off + m]).upper().replace(b"\n", b"A").replace(b"N", b"C"))
        for s_ in range(i % (k + 2)):
            p[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
        pats.append(bytes(p))
    pats.append(bytes(base[-m:]).upper().replace(b"\n", b"A").replace(b"N", b"C"))
    text = torch.frombuffer(base, dtype=torch.uint8).cuda()
    pk = torch.empty(apm_b200.text_pack_bytes(n), dtype=torch.uint8, device="cuda")
    apm_b200.text_pack_device(text.data_ptr(), n, pk.data_ptr())
    torch.cuda.synchronize()
    # the packed words themselves: word w = codes of bytes [16 w, 16 w + 16)
    words = pk.cpu().numpy().view(np.uint32)
    arr = np.frombuffer(bytes(base), dtype=np.uint8)
    for w in (0, 1, (n // 16) // 2, n // 16 - 1, (n + 15) // 16 - 1):
        chunk = arr[16 * w: 16 * w + 16]
        want = sum(((int(c) >> 1) & 3) << (2 * i) for i, c in enumerate(chunk))
        assert int(words[w]) == want, w
    with apm_b200.Plan(pats, k) as plan:
        before = apm_b200.launch_count()
        plan.count_device(text.data_ptr(), 0, n, n, 0, n)
        plain = plan.read_counts()
        assert apm_b200.launch_count() > before
        plan.zero_counts()
        plan.count_device_packed(text.data_ptr(), pk.data_ptr(), 0, n, n, 0, n)
        assert plan.read_counts() == plain
        # window sub-ranges through the packed path add up to the whole
        plan.zero_counts()
        cuts = [0, n // 3 + 5, n // 2, n - m, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            plan.count_device_packed(text.data_ptr(), pk.data_ptr(), 0, n, n, a, b)
        assert plan.read_counts() == plain
        # d_packed = NULL is the plain call
        plan.zero_counts()
        plan.count_device_packed(text.data_ptr(), 0, 0, n, n, 0, n)
        assert plan.read_counts() == plain
    if n <= 100_000:
        assert plain == oracle.count_matches(bytes(base), pats, k)


def test_packed_text_shard_view_and_errors():
    """a database shard: the packed copy belongs to the shard's own buffer (global offset b0 > 0); misaligned buffers
    are refused"""
    import torch
    n, m, k = 2_000_000, 64, 4
    raw = oracle.synth_text(0x5EED0001, 99, n).tobytes()
    rng = np.random.default_rng(5)
    pats = [raw[o:o + m] for o in (int(x) for x in rng.integers(0, n - m, size=12))]
    text = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    with apm_b200.Plan(pats, k) as plan:
        plan.count_device(text.data_ptr(), 0, n, n, 0, n)
        whole = plan.read_counts()
        plan.zero_counts()
        W = n - k
        cuts = [(W * g // 3) & ~15 for g in range(3)] + [W]
        for a, b in zip(cuts[:-1], cuts[1:]):
            e = min(n, b + m - 1)
            view = text[a:e]                       # a is a multiple of 16: the view is 16-byte aligned
            pk = torch.empty(apm_b200.text_pack_bytes(e - a), dtype=torch.uint8, device="cuda")
            apm_b200.text_pack_device(view.data_ptr(), e - a, pk.data_ptr())
            plan.count_device_packed(view.data_ptr(), pk.data_ptr(), a, e - a, n, a, b)
        assert plan.read_counts() == whole
        pk = torch.empty(apm_b200.text_pack_bytes(n), dtype=torch.uint8, device="cuda")
        with pytest.raises(apm_b200.ApmError) as ei:
            apm_b200.text_pack_device(text.data_ptr() + 4, n - 4, pk.data_ptr())
        assert ei.value.code == apm_b200.APM_EINVAL
        with pytest.raises(apm_b200.ApmError) as ei:
            plan.count_device_packed(text.data_ptr() + 4, pk.data_ptr(), 4, n - 4, n, 4, n)
        assert ei.value.code == apm_b200.APM_EINVAL
