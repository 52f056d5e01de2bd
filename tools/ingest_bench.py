#!/usr/bin/env python
"""File -> counts end to end: the chunked ingest of apm_count_matches_file / the `apm` CLI in exact
filter mode on BASELINE-sized texts held in a tmpfs file (page cache), 1 .. N GPUs, reader-thread sweep; plus the
host-buffer entry point with pageable and with pinned memory.  One JSON line per measurement on stdout.

    python tools/ingest_bench.py [--sizes-gib 1,16] [--gpus 1,8] [--dir /dev/shm]

The reference reads the file outside its timer (sequential.c:84 vs :102); here the read IS the job: in filter mode the
16 GiB config-5 search is 7 ms of GPU time, so file -> counts is bounded by pread + H2D."""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import apm_b200  # noqa: E402
from apm_b200.synth import TEXT_SEED, make_patterns  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def make_file(path: str, n: int) -> float:
    """synthetic ACGT text (the BASELINE generator) written through the GPU generator in 256 MiB pieces"""
    t0 = time.perf_counter()
    piece = 256 << 20
    dev = torch.empty(min(piece, n), dtype=torch.uint8, device="cuda")
    with open(path, "wb") as f:
        for off in range(0, n, piece):
            cnt = min(piece, n - off)
            apm_b200.synth_text_device(dev.data_ptr(), TEXT_SEED, off, cnt)
            torch.cuda.synchronize()
            f.write(dev[:cnt].cpu().numpy().tobytes())
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-gib", default="1,16")
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--threads", default="1,2,4,8")
    args = ap.parse_args()
    ngpu = torch.cuda.device_count()
    gpus = [g for g in (int(x) for x in args.gpus.split(",")) if g <= ngpu]
    emit(what="host", cpus=os.cpu_count(), gpus_visible=ngpu)
    for gib in (int(x) for x in args.sizes_gib.split(",")):
        n = gib << 30
        P, m, k = (1024, 64, 4) if gib < 16 else (4096, 64, 4)
        path = os.path.join(args.dir, f"apm_text_{gib}g.bin")
        pfile = os.path.join(args.dir, f"apm_patterns_{gib}g.txt")
        try:
            st = os.statvfs(args.dir)
            if st.f_bavail * st.f_frsize < n + (64 << 20):
                emit(what="skip", size_gib=gib, reason="not enough space in " + args.dir)
                continue
            secs = make_file(path, n)
            pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
            with open(pfile, "wb") as f:
                f.write(b"\n".join(pats) + b"\n")
            emit(what="file written", size_gib=gib, seconds=secs, path=path)
            apm_b200.set_option("mode", "filter")
            want = None
            for g in gpus:
                apm_b200.set_option("gpus", str(g))
                for thr in ["auto"] + [t for t in args.threads.split(",")]:
                    apm_b200.set_option("ingest_threads", thr)
                    best = None
                    for rep in range(3):  # rep 0 allocates pinned buffers and device memory (kept in the pool)
                        t0 = time.perf_counter()
                        got = apm_b200.count_matches_file(path, pats, k)
                        dt = time.perf_counter() - t0
                        if rep:
                            best = dt if best is None else min(best, dt)
                    if want is None:
                        want = got
                    emit(what="apm_count_matches_file", size_gib=gib, patterns=P, m=m, k=k, gpus=g, ingest_threads=thr,
                         seconds=best, text_gbs=n / best / 1e9, total_matches=sum(got), same_counts=got == want)
                # the CLI as a user runs it: process start, CUDA context creation, ingest, search, output
                env = dict(os.environ, APM_GPUS=str(g), APM_PATTERN_FILE=pfile, APM_MODE="filter")
                t0 = time.perf_counter()
                r = subprocess.run([apm_b200.CLI_PATH, str(k), path], env=env, capture_output=True, text=True)
                wall = time.perf_counter() - t0
                mt = re.search(r"APM done in ([0-9.]+) s", r.stdout)
                total = sum(int(x) for x in re.findall(r">: (\d+)", r.stdout))
                emit(what="apm CLI", size_gib=gib, gpus=g, rc=r.returncode, apm_done_s=float(mt.group(1)) if mt else None,
                     process_wall_s=wall, text_gbs_in_timer=n / float(mt.group(1)) / 1e9 if mt else None,
                     total_matches=total, same_total=total == sum(want))
            apm_b200.set_option("gpus", "1")
            apm_b200.set_option("ingest_threads", "auto")
            # host-buffer entry point: pageable memory (staged through the pinned pipeline) vs page-locked memory
            if gib <= 4:
                host = np.fromfile(path, dtype=np.uint8)
                for kind in ("pageable", "pinned"):
                    if kind == "pinned":
                        t = torch.from_numpy(host).pin_memory()
                        ptr = t.data_ptr()
                    else:
                        ptr = host.ctypes.data
                    best = None
                    for rep in range(3):
                        t0 = time.perf_counter()
                        got = apm_b200.count_matches_ptr(ptr, n, pats, k)
                        dt = time.perf_counter() - t0
                        if rep:
                            best = dt if best is None else min(best, dt)
                    emit(what="apm_count_matches host buffer", memory=kind, size_gib=gib, seconds=best,
                         text_gbs=n / best / 1e9, same_counts=got == want)
                del host
        finally:
            for f in (path, pfile):
                if os.path.exists(f):
                    os.remove(f)
    apm_b200.set_option("mode", "direct")


if __name__ == "__main__":
    main()
