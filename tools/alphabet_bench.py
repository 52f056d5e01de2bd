#!/usr/bin/env python
"""Exploration bench: direct-mode slab rate of every DP-cell code for texts over 4, 5 and 8 symbols (the U table of the
sliced kernel grows with the number of symbol planes: three, two, one CTA per SM) -- the data behind the automatic cell
choice by occupancy (profiles/r02_alphabet_bench.txt)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, numpy as np, apm_b200
n = 64 << 20
for alpha in (b"ACGT", b"ACGTN", b"ACGTNMRY"):
    rng = np.random.default_rng(1)
    arr = np.frombuffer(alpha, dtype=np.uint8)[rng.integers(0, len(alpha), size=n)]
    text = torch.from_numpy(arr.copy()).cuda()
    for m in (64, 128, 32):
        P = 128
        pats = [arr[o:o + m].tobytes() for o in rng.integers(0, n - m, size=P)]
        slab = max((1 << 39) // (P * m * m), 1 << 18) // 4096 * 4096
        apm_b200.set_option("mode", "direct")
        for cell in ("auto", "lop3", "fma3", "fma", "fma3r"):
            apm_b200.set_option("cell", cell)
            with apm_b200.Plan(pats, 3) as plan:
                def step(i):
                    a = (i * slab) % (n - slab - m)
                    plan.count_device(text.data_ptr(), 0, n, n, a, a + slab)
                step(0); step(1); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); step(2); step(3); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                print(json.dumps({"alphabet": len(alpha), "m": m, "cell": cell, "TCUPS": round(slab * P * m * m / ms / 1e9, 1)}), flush=True)
