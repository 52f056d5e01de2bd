"""One-process-per-GPU sharding of the hot path on top of torch.distributed.

The path shards with no data-path collective: a rank evaluates its own window-start range (database
shard, with an (m_max-1)-byte halo) or its own patterns (pattern shard, p % world == rank, as
src/patterns_over_ranks.c:161) and the per-pattern count vectors are summed with ONE all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests).  This replaces the MPI master/worker exchange of the
reference (src/patterns_over_ranks.c:139-218,389, src/database_over_ranks.c:119-195,573).

The counting itself is delegated to `counter`; the default is the CUDA path (apm_b200.Plan).  The CPU test
suite passes a checker-backed counter to exercise the shard arithmetic and the reduction without a GPU.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

SHARD_DB = "db"
SHARD_PATTERNS = "patterns"


def db_shard(n_total: int, k: int, m_max: int, rank: int, world: int):
    """-> (j0, j1, b0, b1): window starts [j0, j1) owned by `rank`, bytes [b0, b1) it must hold.
    Cuts are 16-byte aligned (TMA-friendly); only the last shard sees the truncated tail windows."""
    W = max(0, n_total - k)  # sequential.c:121
    j0 = (W * rank // world) & ~15
    j1 = W if rank == world - 1 else (W * (rank + 1) // world) & ~15
    b0, b1 = j0, min(n_total, j1 + max(m_max, 1) - 1)
    return j0, j1, b0, max(b1, b0)


def choose_shard(n_total: int, k: int, nb_patterns: int, world: int, forced: str | None = None) -> str:
    """Replaces the CPU-thread cost model of src/main.c:88-123: text shards when every GPU still gets a few
    tiles per SM, pattern shards when the text is small and there are enough patterns."""
    if forced in ("db", "DB_OVER_RANKS"):
        return SHARD_DB
    if forced in ("patterns", "PATTERNS_OVER_RANKS"):
        return SHARD_PATTERNS
    W = max(0, n_total - k)
    if world <= 1 or W // world >= 148 * 4 * 1024 or nb_patterns < world:
        return SHARD_DB
    return SHARD_PATTERNS


def cuda_counter(shard: np.ndarray, b0: int, n_total: int, j0: int, j1: int, patterns: Sequence[bytes], k: int,
                 pattern_shard=None) -> list[int]:
    """Default counter: the CUDA path through the C-ABI (fails loudly without a GPU)."""
    import torch

    import apm_b200
    if not torch.cuda.is_available():
        raise apm_b200.ApmError(apm_b200.APM_ENODEVICE, "no CUDA device; apm_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    apm_b200.set_device(dev.index)
    d = torch.from_numpy(np.ascontiguousarray(shard)).to(dev) if len(shard) else torch.zeros(16, dtype=torch.uint8, device=dev)
    with apm_b200.Plan(patterns, k) as plan:
        if pattern_shard is not None:
            plan.set_pattern_shard(*pattern_shard)
        plan.count_device(d.data_ptr(), b0, len(shard), n_total, j0, j1, torch.cuda.current_stream().cuda_stream)
        return plan.read_counts(torch.cuda.current_stream().cuda_stream)


def count_matches_distributed(read_bytes: Callable[[int, int], np.ndarray], n_total: int, patterns: Sequence[bytes],
                              k: int, *, shard: str | None = None, counter: Callable = cuda_counter,
                              group=None, device=None) -> list[int]:
    """Every rank calls this; every rank gets the global n_matches (all-reduced).

    read_bytes(offset, count) returns the text bytes [offset, offset+count) as a uint8 array -- each rank
    loads only its own shard (no text ever crosses ranks)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    P = len(patterns)
    m_max = max((len(p) for p in patterns), default=1)
    mode = choose_shard(n_total, k, P, world, shard)
    if mode == SHARD_DB:
        j0, j1, b0, b1 = db_shard(n_total, k, m_max, rank, world)
        part = counter(read_bytes(b0, b1 - b0), b0, n_total, j0, j1, patterns, k) if j1 > j0 else [0] * P
    else:
        part = counter(read_bytes(0, n_total), 0, n_total, 0, max(0, n_total - k), patterns, k,
                       pattern_shard=(rank, world)) if rank < P else [0] * P
    counts = torch.tensor(part, dtype=torch.int64, device=device if device is not None else "cpu")
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return [int(x) for x in counts.tolist()]
