#!/usr/bin/env python
"""Driver for ncu / timing of the exact filter mode with the resident 2-bit text copy (apm_plan_count_device_packed) on a
config-3-shaped job (1 GiB text, P x m=64, k=4); counterpart of tools/filter_profile.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
P, m, k = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 64, 4)
text = torch.empty(n, dtype=torch.uint8, device="cuda")
apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
pk = torch.empty(apm_b200.text_pack_bytes(n), dtype=torch.uint8, device="cuda")
apm_b200.text_pack_device(text.data_ptr(), n, pk.data_ptr())
apm_b200.set_option("mode", "filter")
pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7 if m < 100 else 14)
with apm_b200.Plan(pats, k) as plan:
    for it in range(3):
        plan.zero_counts()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.count_device_packed(text.data_ptr(), pk.data_ptr(), 0, n, n, 0, n)
        e1.record(); torch.cuda.synchronize()
        print(f"packed filter pass {it}: {e0.elapsed_time(e1):.3f} ms  ({n / e0.elapsed_time(e1) / 1e6:.1f} G symbols/s)  matches {sum(plan.read_counts())}")
