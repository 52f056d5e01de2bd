#!/usr/bin/env python
"""Phase timing of the one-shot C-ABI call on the bench's e2e workload (APM_TRACE=1 -> stderr)."""
import os, sys, time
os.environ["APM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
n = (1 << 20) + 63
dev = torch.device("cuda:0")
t = torch.empty(n, dtype=torch.uint8, device=dev)
apm_b200.synth_text_device(t.data_ptr(), TEXT_SEED, 0, n)
host = torch.empty(n, dtype=torch.uint8).pin_memory(); host.copy_(t); torch.cuda.synchronize()
pats, _, _ = make_patterns(TEXT_SEED, 1 << 34, 4096, 64, 7)
if len(sys.argv) > 1:
    apm_b200.set_option("mode", sys.argv[1])
for i in range(3):
    t0 = time.perf_counter()
    apm_b200.count_matches_ptr(host.data_ptr(), n, pats, 4)
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr)
