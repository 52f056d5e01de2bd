#!/usr/bin/env python
"""BASELINE config 5 at FULL size in DIRECT mode (every DP cell of every window evaluated): 2^34 B synthetic ACGT text x
4096 patterns (m = 64, k = 4), database-sharded over the GPUs of one box, one process per GPU (torchrun), counts summed by
one NCCL all-reduce.  2.9e17 cells: about 5 minutes on 8 B200 (40 minutes on one).  Rank 0 compares the complete count
vector with the sha256 pinned in tests/test_gpu_full_configs.py (the 412 s exact-band-mode run of round 1, reproduced by
the filter mode in every `pytest -m gpu` run) and with the exact filter mode run here on the same shards.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/config5_direct_full.py [--windows-per-rank W]     (W: only the first W window starts of every shard -- dry run)
Prints one JSON line (rank 0)."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import apm_b200  # noqa: E402
from apm_b200.dist import db_shard  # noqa: E402
from apm_b200.synth import TEXT_SEED, make_patterns  # noqa: E402

CONFIG5_SHA256 = "a3f964173cd6c44b4c638930ee412798843f9e777dcc45a39345a2332c6bc3a7"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows-per-rank", type=int, default=0)
    ap.add_argument("--slab", type=int, default=1 << 27)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    apm_b200.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, P, M, K = 1 << 34, 4096, 64, 4
    j0, j1, b0, b1 = db_shard(N, K, M, rank, world)
    if args.windows_per_rank:
        j1 = min(j1, j0 + args.windows_per_rank)
        b1 = min(N, j1 + M - 1)
    text = torch.empty(b1 - b0, dtype=torch.uint8, device="cuda")
    apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, b0, b1 - b0)
    torch.cuda.synchronize()
    pats, offs, nsub = make_patterns(TEXT_SEED, N, P, M, 7)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    times = {}
    for mode in ("filter", "direct"):
        apm_b200.set_option("mode", mode)
        with apm_b200.Plan(pats, K) as plan:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            e0.record()
            a = j0
            while a < j1:
                b = min(j1, a + args.slab)
                plan.count_device(text.data_ptr(), b0, b1 - b0, N, a, b, st)
                a = b
                if mode == "direct" and rank == 0 and ((a - j0) // args.slab) % 4 == 0:
                    torch.cuda.synchronize()
                    print(f"[rank 0] {a - j0} / {j1 - j0} windows, {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            c = torch.tensor(plan.read_counts(st), dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(c)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[mode] = [int(x) for x in c.tolist()]
            times[mode] = float(ms.item())
    if rank == 0:
        windows = (j1 - j0) * world if args.windows_per_rank else N - K
        cells = float(windows) * P * M * M
        out = {
            "config": "config5 FULL, direct mode: 2^34 B text x 4096 patterns m=64, k=4, every DP cell evaluated"
                      if not args.windows_per_rank else f"config5 DRY RUN: first {args.windows_per_rank} windows of every shard",
            "n_gpus": world, "shard": "db (16-byte aligned cuts, 63 B halo), NCCL all-reduce of the int64[4096] count vector",
            "direct_job_s": times["direct"] / 1e3, "direct_TCUPS": cells / (times["direct"] * 1e-3) / 1e12,
            "filter_job_ms": times["filter"],
            "total_matches": sum(res["direct"]),
            "direct_equals_filter": res["direct"] == res["filter"],
            "counts_sha256": hashlib.sha256(json.dumps(res["direct"]).encode()).hexdigest(),
        }
        if not args.windows_per_rank:
            out["sha256_equals_pinned_band_mode_vector"] = out["counts_sha256"] == CONFIG5_SHA256
            out["planted_found"] = all(res["direct"][p] >= 1 for p in range(P) if offs[p] is not None and nsub[p] <= K)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
