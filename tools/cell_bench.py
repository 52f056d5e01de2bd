#!/usr/bin/env python
"""Exploration bench: direct-mode slab timings of the sliced kernel for the DP-cell codes given on the command line
(default: all) at a few pattern lengths.  One JSON line per (m, cell).  (The experiments recorded under profiles/r02_cell_*
used earlier revisions of this tool with the since-removed options rows2 / ragcell and cells fma3x / lop3x.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))

import torch  # noqa: E402

import apm_b200  # noqa: E402
from apm_b200.synth import TEXT_SEED, make_patterns  # noqa: E402


def main():
    cells = sys.argv[1].split(",") if len(sys.argv) > 1 else ["auto", "lop3", "fma3", "fma", "fma3r"]
    lens = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 32, 128, 200]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n = 256 << 20
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    apm_b200.set_option("mode", sys.argv[3] if len(sys.argv) > 3 else "direct")
    for m in lens:
        P = 256 if m <= 64 else 64
        slab = max((1 << 40) // (P * m * m), 1 << 19) // 4096 * 4096
        if len(sys.argv) > 3 and sys.argv[3] == "band":
            slab *= 4
        pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
        ref = None
        for cell in cells:
            apm_b200.set_option("cell", cell)
            with apm_b200.Plan(pats, 4) as plan:
                def step(i):
                    a = (i * slab) % (n - slab - m)
                    plan.count_device(text.data_ptr(), 0, n, n, a, a + slab, st)
                for i in range(2):
                    step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for i in range(reps):
                    step(2 + i)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                counts = plan.read_counts()
                if ref is None:
                    ref = counts
                print(json.dumps({"m": m, "P": P, "cell": cell, "slab": slab, "ms": round(ms, 3),
                                  "TCUPS": round(slab * P * m * m / ms / 1e9, 2), "counts_equal_first": counts == ref,
                                  "total": int(sum(counts))}), flush=True)


if __name__ == "__main__":
    main()
