#!/usr/bin/env python
"""bench.py -- the contract benchmark of the approximate-pattern-matching hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the configuration the metric "GCUPS at 1/2/4/8 B200" is quoted on):
synthetic ACGT text of 2^34 bytes (counter-based generator, materialised in HBM, database-sharded over
the N ranks with a 63-byte halo), 4096 patterns of length 64, approx_factor 4.  Evaluating the whole
16 GiB x 4096 job directly is 2.9e17 DP cells (about an hour on one B200), so a STEP is one pass of the
hot path over one SLAB of the rank's shard: `slab` consecutive window starts x all 4096 patterns, every
DP cell evaluated (mode "direct", no filter).  Successive steps take successive slabs, so the text
touched is never L2-resident from a previous step.  GCUPS is size independent for this path; scaling is
weak (every rank processes its own slab per step).

JSON line (rank 0): metric GCUPS (+ text GB/s in roofline_hbm), value = cells of all ranks / max-over-
ranks device time, e2e = the same through the C-ABI one-shot call with HOST buffers (pinned text in,
counts out, H2D/D2H and plan construction inside the timed region), roofline against the measured
integer-pipe peak, cpu_baseline = the reference's own levenshtein() (oracle/_ref) on the host cores.

--impl reference times the reference's CPU implementation (oracle/_ref/libapm_ref.so: unmodified
src/utils.c levenshtein() in the OpenMP work split of src/patterns_over_ranks.c:353-357) on a bounded
sample of the same workload with all host threads.  The oracle is used here only as the thing measured
for that arm / the cpu_baseline, and as the parity checker -- never inside the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "inf560-approximate-pattern-matching_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

N_TOTAL = 1 << 34
NB_PATTERNS = 4096
M = 64
K_ERR = 4
SUBMOD = 7
METRIC = "GCUPS"
WORKLOAD = ("config5: synthetic 16 GiB ACGT text (2^34 B, DB-sharded with 63 B halo), 4096 patterns m=64, k=4; "
            "step = one slab of the rank's shard x all patterns, every DP cell evaluated")


def cells_full(windows: int, nb_patterns: int, m: int) -> float:
    return float(windows) * nb_patterns * m * m


def cells_text(n: int, nb_patterns: int, m: int, k: int) -> float:
    """cells of a complete text of n bytes (full windows + truncated tail)."""
    full = max(0, n - m + 1) * m * m
    tail = sum(s * s for s in range(k + 1, m)) if n >= m else sum(s * s for s in range(k + 1, n + 1))
    return float(nb_patterns) * (full + tail)


def algorithmic_ops(windows: int, nb_patterns: int, m: int) -> float:
    """10 * s * ceil(s/32) int32 ops per (pattern, window of size s)."""
    return float(windows) * nb_patterns * 10 * m * ((m + 31) // 32)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.ok = index, [], set(), None, False
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            self.ok = True
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report it, do not fail the bench
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": getattr(self, "error", "no samples")}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own levenshtein() on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_bytes: int, nb_patterns: int, repeats: int = 1):
    """-> (GCUPS, seconds, cores, kind, sample description).  Bounded sample of the same workload."""
    from apm_b200.synth import TEXT_SEED, make_patterns, text_slice
    from oracle import oracle
    text = text_slice(TEXT_SEED, 0, sample_bytes).tobytes()
    pats, _, _ = make_patterns(TEXT_SEED, N_TOTAL, NB_PATTERNS, M, SUBMOD)
    pats = pats[:nb_patterns]
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if oracle.have_ref():  # all host threads, whatever OMP_NUM_THREADS torchrun exported
        kind, cores = "reference", ncpu
        fn = lambda: oracle.ref_count_matches(text, pats, K_ERR, mode=1, threads=cores)  # noqa: E731
    else:
        kind, cores = "port", ncpu
        fn = lambda: oracle.count_matches(text, pats, K_ERR, threads=cores)  # noqa: E731
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cells = cells_text(sample_bytes, nb_patterns, M, K_ERR)
    sample = (f"first {sample_bytes} B of the config-5 text x first {nb_patterns} patterns (m=64, k=4), "
              f"{cells:.3g} cells, OpenMP position split of patterns_over_ranks.c:353-357")
    return cells / best / 1e9, best, cores, kind, sample


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_bytes, npat = 96 * 1024, 16
    times = []
    gc = cores = kind = sample = None
    for i in range(args.warmup + args.steps):
        gc, dt, cores, kind, sample = cpu_reference_rate(sample_bytes, npat)
        if i >= args.warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    value = cells_text(sample_bytes, npat, M, K_ERR) / mean_t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": sample, "threads": cores},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)



# ---------------------------------------------------------------------------------------------------
# BASELINE configs 3 and 4 (1 GiB synthetic text): complete job in exact filter mode + direct / band slabs
# ---------------------------------------------------------------------------------------------------
def run_config_1gib(torch, apm_b200, dev, stream, name: str, P: int, m: int, k: int, submod: int, hbm_peak: float) -> dict:
    from apm_b200.synth import TEXT_SEED, make_patterns
    n = 1 << 30
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
    torch.cuda.synchronize()
    pats, offs, nsub = make_patterns(TEXT_SEED, n, P, m, submod)
    W = n - k
    out = {"workload": f"{name}: 2^30 B synthetic ACGT text, {P} patterns m={m}, k={k}"}

    def timed(plan, a, b, reps):
        ms = []
        for _ in range(reps):
            plan.zero_counts(stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.count_device(text.data_ptr(), 0, n, n, a, b, stream)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms, plan.read_counts(stream)

    apm_b200.set_option("mode", "filter")
    with apm_b200.Plan(pats, k) as plan:
        ms, whole = timed(plan, 0, W, 5)
        fms = sum(ms[2:]) / len(ms[2:])  # the first two passes warm the tables
        _, fsub = timed(plan, 0, 1 << 24, 1)
    planted = [p for p in range(P) if offs[p] is not None and nsub[p] <= k]
    out["filter_full_job"] = {"ms": fms, "text_gbs": n / (fms * 1e-3) / 1e9, "hbm_frac": n / (fms * 1e-3) / 1e9 / hbm_peak,
                              "effective_gcups": cells_text(n, P, m, k) / (fms * 1e-3) / 1e9, "total_matches": int(sum(whole)),
                              "planted_found": all(whole[p] >= 1 for p in planted)}
    slab = 1 << 20
    apm_b200.set_option("mode", "band")
    with apm_b200.Plan(pats, k) as plan:
        _, bsub = timed(plan, 0, 1 << 24, 1)
        ms, bslab = timed(plan, 4 << 20, (4 << 20) + slab, 4)
        bms = sum(ms[1:]) / len(ms[1:])
    apm_b200.set_option("mode", "direct")
    with apm_b200.Plan(pats, k) as plan:
        ms, dslab = timed(plan, 4 << 20, (4 << 20) + slab, 4)
        dms = sum(ms[1:]) / len(ms[1:])
    out["direct_slab"] = {"gcups": cells_full(slab, P, m) / (dms * 1e-3) / 1e9, "ms": dms, "slab_windows": slab}
    out["band_slab"] = {"effective_gcups": cells_full(slab, P, m) / (bms * 1e-3) / 1e9, "ms": bms, "slab_windows": slab}
    out["parity"] = {"filter_eq_band_first_2p24_windows": fsub == bsub, "band_eq_direct_slab": bslab == dslab}
    del text
    torch.cuda.empty_cache()
    return out




# ---------------------------------------------------------------------------------------------------
# BASELINE configs 1 and 2: the reference's own CPU-runnable cases on dna/small_chrY_x100.fa (committed as
# tests/golden/fixtures.npz with the outputs of the reference's apm_sequential, tests/golden/golden.json).
# Launch-bound jobs: per-call latency through the one-shot C-ABI call with host buffers, counts against the goldens.
# ---------------------------------------------------------------------------------------------------
def run_configs_small(apm_b200) -> dict:
    sys.path.insert(0, ROOT)
    from tests.golden_util import cases, fixtures
    fx = fixtures()
    out = {}
    for key, name in (("1", "config1_readme"), ("2", "config2")):
        c = next(x for x in cases() if x["name"] == name)
        text = fx[c["text"]]
        res = {"workload": f"{name}: dna/{c['text']}.fa ({len(text)} B), {len(c['patterns'])} patterns "
                           f"m={sorted(set(len(p) for p in c['patterns']))}, k={c['k']}",
               "cells": float(sum(cells_text(len(text), 1, len(p), c["k"]) for p in c["patterns"]))}
        for mode in ("direct", "filter"):
            apm_b200.set_option("mode", mode)
            ts = []
            got = None
            for i in range(12):
                t0 = time.perf_counter()
                got = apm_b200.count_matches(text, c["patterns"], c["k"])
                ts.append(time.perf_counter() - t0)
            ms = sorted(ts[2:])[len(ts[2:]) // 2] * 1e3
            res[mode] = {"ms_per_call": ms, "gcups": res["cells"] / (ms * 1e-3) / 1e9,
                         "equals_reference_golden": got == c["expected"]}
        apm_b200.set_option("mode", "direct")
        out[key] = res
    return out

# ---------------------------------------------------------------------------------------------------
# End to end with the INGEST inside the timed region: config 3 (1 GiB) in exact filter mode, where
# the search itself is 0.4 ms, through the one-shot C-ABI calls a user makes -- host buffer (pinned / pageable) and file
# ---------------------------------------------------------------------------------------------------
def run_ingest_e2e(torch, apm_b200, dev) -> dict:
    import numpy as np
    from apm_b200.synth import TEXT_SEED, make_patterns
    n, P, m, k = 1 << 30, 1024, 64, 4
    pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(d.data_ptr(), TEXT_SEED, 0, n)
    torch.cuda.synchronize()
    pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
    pinned.copy_(d)
    torch.cuda.synchronize()
    del d
    pageable = pinned.numpy().copy()
    out = {"workload": "config3 (2^30 B, 1024 patterns m=64, k=4), exact filter mode, text NOT resident: "
                       "H2D / file read inside the timed region", "h2d_bytes_per_call": n + P * m, "d2h_bytes_per_call": 8 * P}
    apm_b200.set_option("mode", "filter")
    ref = None

    def timed(fn):
        best, got = None, None
        for rep in range(3):  # the first call allocates the pooled device / pinned buffers
            t0 = time.perf_counter()
            got = fn()
            dt = time.perf_counter() - t0
            if rep:
                best = dt if best is None else min(best, dt)
        return best, got

    try:
        for name, ptr in (("host_pinned", pinned.data_ptr()), ("host_pageable", pageable.ctypes.data)):
            dt, got = timed(lambda: apm_b200.count_matches_ptr(ptr, n, pats, k))
            ref = ref or got
            out[name] = {"ms": dt * 1e3, "text_gbs": n / dt / 1e9, "counts_equal": got == ref, "total_matches": int(sum(got))}
        path = "/dev/shm/apm_bench_config3.bin"
        try:
            pageable.tofile(path)
            dt, got = timed(lambda: apm_b200.count_matches_file(path, pats, k))
            out["file_tmpfs"] = {"ms": dt * 1e3, "text_gbs": n / dt / 1e9, "counts_equal": got == ref,
                                 "reader_threads": apm_b200.get_option("ingest_threads"), "host_cpus": os.cpu_count()}
        except OSError as e:
            out["file_tmpfs"] = {"skipped": repr(e)}
        finally:
            if os.path.exists(path):
                os.remove(path)
    finally:
        apm_b200.set_option("mode", "direct")
    return out

# ---------------------------------------------------------------------------------------------------
# N > 1: the multi-GPU paths that a one-GPU test box can never run
# ---------------------------------------------------------------------------------------------------
def multi_gpu_parity(torch, dist, apm_b200, dev, rank: int, world: int) -> dict | None:
    """(a) every rank: one PATTERN-shard step (p % world == rank, patterns_over_ranks.c:161) + all-reduce;
    (b) rank 0 alone: the single-process multi-GPU C path apm_count_matches(gpus=N) with database shards and
    pattern shards, count reduction by our peer kernel (p2p) and by ncclAllReduce -- all compared with rank 0's
    single-GPU result on a config-2-sized input (64 patterns of length 32, k = 2)."""
    from apm_b200.synth import TEXT_SEED, make_patterns
    import numpy as np
    n, P, m, k = 1 << 20, 64, 32, 2
    pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 4)
    pats = pats[:48] + [pats[0][:20], pats[1] + pats[2], pats[3] * 4, b"ACGT" * 16]  # mixed lengths: 20 .. 128
    pats += pats[48:] * 3
    pats = pats[:P]
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    with apm_b200.Plan(pats, k) as plan:  # single GPU, everything
        plan.count_device(text.data_ptr(), 0, n, n, 0, n, stream)
        want = torch.tensor(plan.read_counts(stream), dtype=torch.int64, device=dev)
    dist.broadcast(want, src=0)
    with apm_b200.Plan(pats, k) as plan:  # (a)
        plan.set_pattern_shard(rank, world)
        plan.count_device(text.data_ptr(), 0, n, n, 0, n, stream)
        part = torch.tensor(plan.read_counts(stream), dtype=torch.int64, device=dev)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    ok = torch.tensor([int(bool((part == want).all().item()))], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["pattern_shard_allreduce"] = "ok" if int(ok.item()) == 1 else "MISMATCH"
    dist.barrier()
    if rank == 0:  # (b) the other ranks idle at the barrier below
        host = text.cpu().numpy().tobytes()
        w = want.cpu().tolist()
        apm_b200.set_option("gpus", str(world))
        try:
            for name, shard, reduce in (("single_process_db_p2p", "db", "p2p"), ("single_process_db_nccl", "db", "nccl"),
                                        ("single_process_patterns_p2p", "patterns", "p2p"),
                                        ("single_process_patterns_host", "patterns", "host")):
                apm_b200.set_option("shard", shard)
                apm_b200.set_option("reduce", reduce)
                try:
                    got = apm_b200.count_matches(host, pats, k)
                    res[name] = "ok" if got == w else "MISMATCH"
                except apm_b200.ApmError as e:
                    res[name] = f"error: {e}"
        finally:
            apm_b200.set_option("gpus", "1")
            apm_b200.set_option("shard", "auto")
            apm_b200.set_option("reduce", "auto")
            apm_b200.set_device(dev.index)
    dist.barrier()
    del text
    return res if rank == 0 else None

# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import apm_b200
    from apm_b200.synth import TEXT_SEED, make_patterns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    apm_b200.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    apm_b200.set_option("kernel", args.kernel)
    apm_b200.set_option("gpus", "1")

    # ---- this rank's database shard: window starts [j0, j1), bytes [j0, j1 + M - 1) --------------
    W = N_TOTAL - K_ERR
    j0 = (W * rank // world) & ~15
    j1 = W if rank == world - 1 else (W * (rank + 1) // world) & ~15
    b0, b1 = j0, min(N_TOTAL, j1 + M - 1)
    shard = torch.empty(b1 - b0, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(shard.data_ptr(), TEXT_SEED, b0, b1 - b0)
    torch.cuda.synchronize()

    pats, offs, nsub = make_patterns(TEXT_SEED, N_TOTAL, NB_PATTERNS, M, SUBMOD)
    plan = apm_b200.Plan(pats, K_ERR)
    slab = args.slab_windows
    nslabs = max(1, (j1 - j0 - 2 * M) // slab)
    stream = torch.cuda.current_stream().cuda_stream
    counts_ptr = plan.counts_device_ptr()
    reduced = torch.zeros(NB_PATTERNS, dtype=torch.int64, device=dev) if world > 1 else None

    def step(i: int) -> None:
        a = j0 + (i % nslabs) * slab
        plan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, a, a + slab, stream)
        if world > 1:
            # the only exchange of the path: the per-pattern count vector (32 KB) summed over NCCL/NVLink.
            # A copy is reduced so that the plan's accumulating counters stay per-rank partial sums.
            reduced.copy_(_tensor_from_ptr(torch, counts_ptr, NB_PATTERNS, dev))
            dist.all_reduce(reduced, op=dist.ReduceOp.SUM)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = apm_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = apm_b200.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    cells_step = cells_full(slab, NB_PATTERNS, M) * world
    value = cells_step / (ms_step * 1e-3) / 1e9

    # ---- exact band mode: same slabs, only the 2k+1 diagonals that can matter for D <= k.
    #      Reported separately as EFFECTIVE GCUPS (cells of the reference DP per second, most never touched).
    band = None
    filt = None
    if not args.no_band:
        apm_b200.set_option("mode", "band")
        bplan = apm_b200.Plan(pats, K_ERR)
        for i in range(3):
            bplan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, j0 + (i % nslabs) * slab, j0 + (i % nslabs) * slab + slab, stream)
        sync_all()
        b_e0, b_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b_e0.record()
        for i in range(args.steps):
            a = j0 + ((3 + i) % nslabs) * slab
            bplan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, a, a + slab, stream)
        b_e1.record()
        sync_all()
        bms = b_e0.elapsed_time(b_e1) / args.steps
        if world > 1:
            t = torch.tensor([bms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            bms = float(t.item())
        # parity of the two modes on identical slabs
        plan.zero_counts(stream)
        bplan.zero_counts(stream)
        plan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, j0, j0 + slab, stream)
        bplan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, j0, j0 + slab, stream)
        same = plan.read_counts(stream) == bplan.read_counts(stream)
        band = {"value": cells_step / (bms * 1e-3) / 1e9, "unit": "effective GCUPS", "ms_per_step": bms,
                "counts_equal_direct": bool(same),
                "note": "exact Ukkonen band (9 of 64 diagonals at k=4): (2k+1)*5+(k+1) = 50 LOP3 per row and 32 windows"}
        bplan.close()
        # ---- exact filter mode (pigeonhole seeds + banded verification): the WHOLE shard of this rank, i.e. the
        #      complete config-5 job at N ranks, not a slab.  Counts of its first slab must equal the direct kernel's.
        apm_b200.set_option("mode", "filter")
        fplan = apm_b200.Plan(pats, K_ERR)
        fplan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, j0, j0 + slab, stream)
        slab_same = fplan.read_counts(stream) == plan.read_counts(stream)
        fplan.zero_counts(stream)
        sync_all()
        f_e0, f_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f_e0.record()
        fplan.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, j0, j1, stream)
        if world > 1:
            reduced.copy_(_tensor_from_ptr(torch, fplan.counts_device_ptr(), NB_PATTERNS, dev))
            dist.all_reduce(reduced, op=dist.ReduceOp.SUM)
        f_e1.record()
        sync_all()
        fms = f_e0.elapsed_time(f_e1)
        total_matches = sum(fplan.read_counts(stream))
        if world > 1:
            t = torch.tensor([fms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fms = float(t.item())
            total_matches = int(reduced.sum().item())
        filt = {"value": cells_text(N_TOTAL, NB_PATTERNS, M, K_ERR) / (fms * 1e-3) / 1e9, "unit": "effective GCUPS",
                "full_job_ms": fms, "text_gbs": N_TOTAL / (fms * 1e-3) / 1e9, "total_matches": int(total_matches),
                "counts_equal_direct_on_slab": bool(slab_same),
                "note": "complete config-5 job (2^34 B x 4096 patterns, every window start of every rank's shard + "
                        "count all-reduce) in exact filter mode; the truncated tail windows go through the DP kernel"}
        # the same job with a resident 2-bit copy of the text (packed once, outside the timed region: the repeated-search
        # case -- database in HBM, query batches change)
        try:
            pk = torch.empty(apm_b200.text_pack_bytes(b1 - b0), dtype=torch.uint8, device=dev)
            p_e0, p_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p_e0.record()
            apm_b200.text_pack_device(shard.data_ptr(), b1 - b0, pk.data_ptr(), stream)
            p_e1.record()
            whole = fplan.read_counts(stream)
            fplan.zero_counts(stream)
            sync_all()
            q_e0, q_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q_e0.record()
            fplan.count_device_packed(shard.data_ptr(), pk.data_ptr(), b0, b1 - b0, N_TOTAL, j0, j1, stream)
            if world > 1:
                reduced.copy_(_tensor_from_ptr(torch, fplan.counts_device_ptr(), NB_PATTERNS, dev))
                dist.all_reduce(reduced, op=dist.ReduceOp.SUM)
            q_e1.record()
            sync_all()
            pms = q_e0.elapsed_time(q_e1)
            same = fplan.read_counts(stream) == whole
            if world > 1:
                t = torch.tensor([pms, 0.0 if same else 1.0], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                pms, same = float(t[0].item()), float(t[1].item()) == 0.0
            filt["packed_text"] = {"full_job_ms": pms, "text_symbols_per_s": N_TOTAL / (pms * 1e-3),
                                   "pack_ms": p_e0.elapsed_time(p_e1), "packed_bytes_per_rank": int(pk.numel()),
                                   "counts_equal_unpacked": bool(same),
                                   "note": "apm_plan_count_device_packed: the scan streams the resident 2-bit copy "
                                           "(0.25 B per symbol); verification still reads the raw bytes"}
            del pk
        except Exception as e:  # noqa: BLE001
            filt["packed_text"] = {"error": repr(e)}
        fplan.close()
        apm_b200.set_option("mode", "direct")

    # ---- end-to-end through the one-shot C-ABI call with HOST buffers ------------------------------
    e2e_windows = args.e2e_windows
    host_text = torch.empty(e2e_windows + M - 1, dtype=torch.uint8).pin_memory()
    host_text.copy_(shard[: e2e_windows + M - 1])
    torch.cuda.synchronize()
    e2e_times = []
    for i in range(2 + args.e2e_steps):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        got = apm_b200.count_matches_ptr(host_text.data_ptr(), host_text.numel(), pats, K_ERR)
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
    e2e_dt = sum(e2e_times) / len(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_cells = cells_text(host_text.numel(), NB_PATTERNS, M, K_ERR) * world
    e2e = {"value": e2e_cells / e2e_dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(host_text.numel() + NB_PATTERNS * M),
           "d2h_bytes_per_step": 8 * NB_PATTERNS, "ms_per_step": e2e_dt * 1e3,
           "call": "apm_count_matches(host text slab + halo, 4096 patterns, k=4): plan build + H2D + kernels + D2H"}

    mgp = multi_gpu_parity(torch, dist, apm_b200, dev, rank, world) if world > 1 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: measured integer-pipe peak, algorithmic ops of SURVEY.md section 8d --------------
    peak0, _ = apm_b200.int_peak(0)  # LOP3 + IADD3 (ptxas turns the add into IMAD.IADD: both pipes)
    peak3, _ = apm_b200.int_peak(3)  # LOP3 + IMAD
    peak1, _ = apm_b200.int_peak(1)  # LOP3 only: the ALU pipe every boolean instruction has to use
    int_peak = max(peak0, peak3)
    alg_ops = algorithmic_ops(slab, NB_PATTERNS, M)
    achieved = alg_ops / (ms_step * 1e-3)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    text_gbs = slab * world / (ms_step * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this very
    # configuration (profiles/traffic.json); null when the bench runs a configuration that was not captured
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            captures = json.load(f)
        for name, tr in captures.items():
            if (name.startswith("sliced_count_kernel<64, 3>") and args.kernel in ("auto", "sliced")
                    and tr.get("slab_windows") == slab and tr.get("patterns") == NB_PATTERNS):
                traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    algorithmic_bytes = slab + (M - 1) + NB_PATTERNS * (M + 16) + 8 * NB_PATTERNS  # text + halo + patterns + counts
    # what the window-sliced kernel EXECUTES per unit (pattern, window): cell = 4 LOP3 (ALU pipe) + 3 IMAD (FMA pipe)
    # per 32 windows for m = 64; the Myers kernel executes the algorithmic 10 ops per word step
    units_per_s = float(slab) * NB_PATTERNS / (ms_step * 1e-3)
    sliced = args.kernel in ("auto", "sliced")
    lop3_per_unit = 4.0 * M * M / 32 if sliced else 7.0 * M * ((M + 31) // 32)
    exec_per_unit = 7.0 * M * M / 32 if sliced else 10.0 * M * ((M + 31) // 32)
    roofline = {
        "bound": "int_alu", "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tiop/s",
        "frac": achieved / int_peak, "frac_executed": exec_per_unit * units_per_s / int_peak,
        "alu_pipe_frac": lop3_per_unit * units_per_s / peak1,
        "executed_ops_per_unit": exec_per_unit, "alu_pipe_ops_per_unit": lop3_per_unit,
        "traffic": traffic, "traffic_source": "static: profiles/traffic.json, ncu --set full capture of this configuration "
        "(dram__bytes_read.sum + dram__bytes_write.sum per launch); null when this slab size was not captured",
        "algorithmic_bytes": algorithmic_bytes,
        "kernel": "sliced_count_kernel<64, 3>" if args.kernel in ("auto", "sliced") else "myers_count_kernel<2,4,0>",
        "algorithmic_ops_per_unit": "10*m*ceil(m/32) = 1280 int32 ops per (pattern, window), SURVEY.md 8d",
        "peak_source": "measured in this run: max(LOP3+IADD3, LOP3+IMAD) dependency-free microbenchmark",
        "lop3_only_peak": peak1 / 1e12,
        "note": "frac is on SURVEY's ALGORITHMIC op count (10 ops per 32-cell word step of the row-parallel "
                "formulation). The window-sliced kernel EXECUTES fewer: 4 LOP3 (ALU pipe) + 3 IMAD (FMA pipe) per "
                "cell and 32 windows = 512 + 384 instructions per unit, so frac can exceed 1; frac_executed is the "
                "same on the ops it executes (vs the LOP3+IMAD peak) and alu_pipe_frac its LOP3 rate vs lop3_only_peak, "
                "the pipe that bounds it (DESIGN.md 4.1)",
    }
    if filt is not None:  # the filter scan reads the text exactly once: its roofline is HBM
        filt["roofline"] = {"bound": "hbm", "achieved": filt["text_gbs"], "peak": hbm_peak * world, "unit": "GB/s",
                            "frac": filt["text_gbs"] / (hbm_peak * world),
                            "note": "whole job incl. seed-hit verification, tail windows and the count all-reduce; the "
                                    "2-bit q-gram scan kernel alone runs at 2.6 TB/s per GPU at 4096 patterns (DESIGN.md 4.1c)"}
    roofline_hbm = {"bound": "hbm", "achieved": text_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": text_gbs / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "note": "text bytes per second; the path is compute bound by construction (>1e5 int ops per text byte)"}

    # ---- parity spot check of the timed path against the oracle -------------------------------------
    parity = "skipped"
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle
        L = 12288
        a = j0 + slab // 2
        sel = [0, 1, 2, 3, 4, 5, 6, 3500]
        sub = apm_b200.Plan([pats[i] for i in sel], K_ERR)
        sub.count_device(shard.data_ptr(), b0, b1 - b0, N_TOTAL, a, a + L, stream)
        got = sub.read_counts(stream)
        seg = oracle.synth_text(TEXT_SEED, a, L + M - 1).tobytes() + b"\0" * M
        want = [oracle.count_range(seg, pats[i], K_ERR, 0, L) for i in sel]
        sub.close()
        parity = "ok" if got == want else f"MISMATCH got={got} want={want}"
        gc, dt, cores, kind, sample = cpu_reference_rate(160 * 1024, 16)
        cpu = {"value": gc, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample, "seconds": dt}

    configs = ingest = None
    if world == 1 and not args.no_configs:
        del shard
        torch.cuda.empty_cache()
        apm_b200.release_cache()
        configs = run_configs_small(apm_b200)
        configs["3"] = run_config_1gib(torch, apm_b200, dev, stream, "config3", 1024, 64, 4, 7, hbm_peak)
        configs["4"] = run_config_1gib(torch, apm_b200, dev, stream, "config4", 256, 200, 10, 14, hbm_peak)
        ingest = run_ingest_e2e(torch, apm_b200, dev)

    line = {
        "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "slab_windows": slab, "patterns": NB_PATTERNS, "m": M, "k": K_ERR,
                   "kernel": args.kernel, "mode": "direct", "shard": "db", "l2": "successive steps read successive "
                   "slabs of a 16 GiB text (inputs larger than L2)", "text_bytes_per_rank": int(b1 - b0)},
        "text_gbs": text_gbs, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "parity": parity, "band_mode": band, "filter_mode": filt,
        "configs": configs, "e2e_ingest": ingest, "multi_gpu_parity": mgp,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _tensor_from_ptr(torch, ptr: int, n: int, dev):
    """int64 tensor aliasing n device words at `ptr` (CUDA array interface)."""

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Holder(), device=dev)


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """Library chatter (e.g. "NCCL version ..." printed by NCCL on fd 1) must not precede the ONE JSON line: from
    here on fd 1 points at stderr, and _emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--kernel", choices=["auto", "sliced", "myers"], default="auto")
    ap.add_argument("--slab-windows", type=int, default=1 << 23,
                    help="window starts per step and rank (2^23: 1.2 s per step, so --steps 20 times >= 20 s of windows)")
    ap.add_argument("--no-configs", action="store_true", help="skip the config-3 / config-4 (1 GiB) measurements")
    ap.add_argument("--e2e-windows", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip cpu_baseline + parity spot check")
    ap.add_argument("--no-band", action="store_true", help="skip the extra exact-band-mode measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
