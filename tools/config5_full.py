#!/usr/bin/env python
"""BASELINE config 5 at FULL size on one B200: 2^34 B synthetic ACGT text x 4096 patterns (m=64, k=4).
Counts of the exact filter mode over the whole text vs the exact band mode over the whole text (about 7 minutes),
plus oracle slices.  Writes one JSON line (profiles/r01_config5_full.json is a copy of its output)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
from oracle import oracle
N, P, M, K = 1 << 34, 4096, 64, 4
dev = torch.device("cuda:0")
text = torch.empty(N, dtype=torch.uint8, device=dev)
apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, N)
torch.cuda.synchronize()
pats, offs, nsub = make_patterns(TEXT_SEED, N, P, M, 7)
out = {"config": "config5 full: 2^34 B text, 4096 patterns m=64, k=4, one B200"}
res = {}
for mode in ("filter", "band"):
    apm_b200.set_option("mode", mode)
    with apm_b200.Plan(pats, K) as plan:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.count_device(text.data_ptr(), 0, N, N, 0, N)
        e1.record(); torch.cuda.synchronize()
        res[mode] = plan.read_counts()
        out[mode + "_ms"] = e0.elapsed_time(e1)
        out[mode + "_effective_TCUPS"] = N * P * M * M / out[mode + "_ms"] / 1e9
    print(mode, out[mode + "_ms"], "ms", file=sys.stderr, flush=True)
out["counts_equal"] = res["filter"] == res["band"]
out["total_matches"] = sum(res["band"])
out["planted_found"] = all(res["band"][p] >= 1 for p in range(P) if offs[p] is not None and nsub[p] <= K)
import hashlib
out["counts_sha256"] = hashlib.sha256(json.dumps(res["band"]).encode()).hexdigest()
# oracle slices (CPU restatement of sequential.c) around planted patterns and at the global tail
apm_b200.set_option("mode", "filter")
L, chk = 5000, [0, 1, 2, 3, 4095]
ok = True
with apm_b200.Plan([pats[i] for i in chk], K) as plan:
    for s in [0, N - K - L] + [int(offs[i]) - 100 for i in (0, 1, 2) if offs[i] is not None and offs[i] > 100]:
        plan.zero_counts()
        plan.count_device(text.data_ptr(), 0, N, N, s, s + L)
        got = plan.read_counts()
        end = min(N, s + L + M - 1)
        seg = oracle.synth_text(TEXT_SEED, s, end - s).tobytes() + (b"" if end == N else b"\0" * M)
        want = [oracle.count_range(seg, pats[i], K, 0, L) for i in chk]
        ok = ok and got == want
out["oracle_slices_ok"] = ok
print(json.dumps(out))
