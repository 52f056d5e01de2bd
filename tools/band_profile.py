#!/usr/bin/env python
"""Small driver for ncu: band-mode launches on a config-3-shaped slab (m=64, k=4, 256 patterns)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
n = 64 << 20
text = torch.empty(n, dtype=torch.uint8, device="cuda")
apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
mode = sys.argv[1] if len(sys.argv) > 1 else "band"
m, k, P = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (64, 4, 256)
apm_b200.set_option("mode", mode)
pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
slab = 8 << 20
with apm_b200.Plan(pats, k) as plan:
    for i in range(4):
        plan.count_device(text.data_ptr(), 0, n, n, i * slab, (i + 1) * slab)
    print(sum(plan.read_counts()))
