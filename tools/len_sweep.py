#!/usr/bin/env python
"""Direct-mode throughput of the window-sliced kernel over pattern lengths (JSON lines), default cell choice vs
APM_CELL given in argv[1] (e.g. "lop3": forces the generic kernel, i.e. the round-1 path for ragged lengths)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200")]
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
cell = sys.argv[1] if len(sys.argv) > 1 else "auto"
apm_b200.set_option("cell", cell)
n = 64 << 20
text = torch.empty(n, dtype=torch.uint8, device="cuda")
apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n); torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
LENS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
for m in LENS or (12, 20, 24, 28, 31, 32, 36, 40, 48, 50, 52, 56, 60, 63, 64, 65, 72, 96, 100, 128, 136, 192, 200, 256, 500, 1000):
    P = 256 if m <= 64 else 64; slab = max(1 << 18, (8 << 20) * 64 * 64 // (m * m) * 64 // P)
    slab = min(slab, n - 2000 - m)
    pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
    with apm_b200.Plan(pats, 4) as plan:
        for i in range(2): plan.count_device(text.data_ptr(), 0, n, n, 0, slab, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3): plan.count_device(text.data_ptr(), 0, n, n, i * 1000, i * 1000 + slab, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"m": m, "cell": cell, "P": P, "slab": slab, "ms": ms, "TCUPS": slab * P * m * m / ms / 1e9}), flush=True)
