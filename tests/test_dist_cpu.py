"""world_size-2/3 gloo tests (CPU) of the one-process-per-GPU host logic in apm_b200/dist.py: shard
arithmetic with halo, pattern sharding, the count all-reduce.  The counting is done by the checker here
(no GPU in this container); the CUDA counter is covered by tests/test_gpu_parity.py and the GPU dist test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apm_b200 import dist as adist
from oracle import oracle


def _oracle_counter(shard, b0, n_total, j0, j1, patterns, k, pattern_shard=None):
    """Counts of window starts [j0, j1) given only the shard bytes; truncation only at the GLOBAL end."""
    seg = bytes(shard)
    touches_end = b0 + len(seg) == n_total
    out = []
    for idx, p in enumerate(patterns):
        if pattern_shard is not None and idx % pattern_shard[1] != pattern_shard[0]:
            out.append(0)
            continue
        pad = b"" if touches_end else b"\0" * len(p)
        # local coordinates; interior shards are padded so the checker does not truncate at the shard end
        hi = min(j1, n_total - k) - b0
        out.append(oracle.count_range(seg + pad, p, k, j0 - b0, hi))
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, text, pats, k, shard, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arr = np.frombuffer(text, dtype=np.uint8)
    seen = []

    def read_bytes(off, cnt):
        seen.append((off, cnt))
        return arr[off:off + cnt]

    got = adist.count_matches_distributed(read_bytes, len(text), pats, k, shard=shard, counter=_oracle_counter)
    q.put((rank, got, seen))
    dist.destroy_process_group()


def _run(world, text, pats, k, shard):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, text, pats, k, shard, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def _case():
    n = 40_000
    text = bytearray(oracle.synth_text(0x5EED0001, 77, n).tobytes())
    p64 = bytes(text[1000:1064])
    for world in (2, 3):  # plant exact copies straddling every seam
        for g in range(1, world):
            seam = ((n - 2) * g // world) & ~15
            text[seam - 30:seam + 34] = p64
    text = bytes(text)
    pats = [p64, text[5000:5050], text[-20:] + b"ACGTACGT", b"ACGTAC", text[20000:20200]]
    return text, pats, 2


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shard", ["db", "patterns", None])
def test_sharded_counts_equal_unsharded(world, shard):
    text, pats, k = _case()
    want = oracle.count_matches(text, pats, k)
    res = _run(world, text, pats, k, shard)
    for rank, got, seen in res:
        assert got == want, (rank, shard)
    if shard == "db":  # every rank read only its shard + halo, never the whole text
        m_max = max(len(p) for p in pats)
        for rank, got, seen in res:
            j0, j1, b0, b1 = adist.db_shard(len(text), k, m_max, rank, world)
            assert seen == [(b0, b1 - b0)]
            assert b1 - b0 <= (len(text) // world) + m_max + 32


def test_db_shard_arithmetic():
    for n, k, m, world in ((1 << 34, 4, 64, 8), (132803, 2, 50, 3), (100, 0, 7, 8), (10, 20, 5, 2), (0, 0, 3, 2)):
        cover = []
        for r in range(world):
            j0, j1, b0, b1 = adist.db_shard(n, k, m, r, world)
            assert j0 % 16 == 0 and j0 <= j1 and b0 == j0 and b1 <= max(n, b0)
            assert b1 >= min(n, j1 + m - 1) or j1 == j0
            cover.append((j0, j1))
        assert cover[0][0] == 0 and cover[-1][1] == max(0, n - k)
        assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    assert adist.choose_shard(1 << 34, 4, 4096, 8) == "db"
    assert adist.choose_shard(132803, 0, 64, 8) == "patterns"
    assert adist.choose_shard(132803, 0, 2, 8) == "db"
    assert adist.choose_shard(132803, 0, 64, 8, "DB_OVER_RANKS") == "db"
