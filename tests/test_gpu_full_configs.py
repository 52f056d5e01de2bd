"""BASELINE.json configs 3 and 4 at FULL size (1 GiB synthetic ACGT text) on the GPU, through the exact band
mode (bit-identical to the direct evaluation, see test_gpu_parity.py), checked with size-independent
properties: planted patterns are found, database shards add up to the unsharded run, pattern shards add up,
a 32 MiB sub-range agrees with the direct (every-cell) evaluation, and sampled slices agree with the oracle."""
import numpy as np
import pytest

import apm_b200
from oracle import oracle
from tests.synth import TEXT_SEED, make_patterns

pytestmark = pytest.mark.gpu

N = 1 << 30


@pytest.fixture(scope="module")
def text_1g():
    import torch
    dev = torch.empty(N, dtype=torch.uint8, device="cuda")
    apm_b200.synth_text_device(dev.data_ptr(), TEXT_SEED, 0, N)
    torch.cuda.synchronize()
    yield dev
    del dev


@pytest.fixture(autouse=True)
def _opts():
    apm_b200.set_option("kernel", "auto")
    apm_b200.set_option("mode", "band")
    yield
    apm_b200.set_option("mode", "direct")


@pytest.mark.parametrize("name,P,m,k,submod", [("config3", 1024, 64, 4, 7), ("config4", 256, 200, 10, 14)])
def test_full_config(text_1g, name, P, m, k, submod):
    pats, offs, nsub = make_patterns(TEXT_SEED, N, P, m, submod)
    W = N - k
    ptr = text_1g.data_ptr()
    with apm_b200.Plan(pats, k) as plan:
        plan.count_device(ptr, 0, N, N, 0, W)
        whole = plan.read_counts()
        # (1) every pattern cut from the text with <= k substitutions is found; the window it was cut from
        #     and its neighbours within k - nsub shifts are matches, so the count is >= 1
        for p in range(P):
            if offs[p] is not None and nsub[p] <= k:
                assert whole[p] >= 1, (name, p)
        assert sum(whole) >= sum(1 for p in range(P) if offs[p] is not None and nsub[p] <= k)
        # (1b) exact filter mode over the whole text gives the same vector
        apm_b200.set_option("mode", "filter")
        with apm_b200.Plan(pats, k) as fplan:
            fplan.count_device(ptr, 0, N, N, 0, W)
            assert fplan.read_counts() == whole, name
            fplan.zero_counts()
            for a, b in ((0, W // 3 + 5), (W // 3 + 5, W)):
                fplan.count_device(ptr, 0, N, N, a, b)
            assert fplan.read_counts() == whole, name
        apm_b200.set_option("mode", "band")
        # (2) 4 database shards (unaligned cuts) add up to the whole
        plan.zero_counts()
        cuts = [0, W // 4 + 3, W // 2 + 17, (3 * W) // 4 + 1, W]
        for a, b in zip(cuts[:-1], cuts[1:]):
            plan.count_device(ptr, 0, N, N, a, b)
        assert plan.read_counts() == whole
        # (3) a 32 MiB sub-range: band == direct (every DP cell)
        a, b = 300 << 20, 332 << 20
        plan.zero_counts()
        plan.count_device(ptr, 0, N, N, a, b)
        band_part = plan.read_counts()
    apm_b200.set_option("mode", "direct")
    sel = list(range(0, P, max(1, P // 48)))
    with apm_b200.Plan([pats[i] for i in sel], k) as plan:
        plan.count_device(ptr, 0, N, N, a, b)
        assert plan.read_counts() == [band_part[i] for i in sel]
    apm_b200.set_option("mode", "band")
    # (4) pattern shards add up
    total = [0] * P
    for r in range(2):
        with apm_b200.Plan(pats, k) as plan:
            plan.set_pattern_shard(r, 2)
            plan.count_device(ptr, 0, N, N, 0, W // 8)
            part = plan.read_counts()
        total = [x + y for x, y in zip(total, part)]
    with apm_b200.Plan(pats, k) as plan:
        plan.count_device(ptr, 0, N, N, 0, W // 8)
        assert plan.read_counts() == total
    # (5) sampled slices against the oracle, incl. the global tail and the neighbourhood of planted patterns
    L = 6000
    starts = [0, W - L] + [int(o) - 100 for o in offs[:4] if o is not None and o > 100]
    chk = [0, 1, 2, 3, P - 1]
    with apm_b200.Plan([pats[i] for i in chk], k) as plan:
        for s in starts:
            plan.zero_counts()
            plan.count_device(ptr, 0, N, N, s, s + L)
            got = plan.read_counts()
            end = min(N, s + L + m - 1)
            seg = oracle.synth_text(TEXT_SEED, s, end - s).tobytes()
            pad = b"" if end == N else b"\0" * m
            want = [oracle.count_range(seg + pad, pats[i], k, 0, L) for i in chk]
            assert got == want, (name, s)
