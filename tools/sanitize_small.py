#!/usr/bin/env python
"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck): golden
fixtures through auto / myers / dp kernels, band mode, long patterns, unaligned shards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch
import apm_b200
from tests.golden_util import cases, fixtures
from oracle import oracle

FX = fixtures()
ok = True
for name in ("easy_k2", "small_k2", "small_k25", "small_m200_k10"):
    c = next(c for c in cases() if c["name"] == name)
    for kernel, mode in (("auto", "direct"), ("auto", "band"), ("myers", "direct"), ("dp", "direct")):
        apm_b200.set_option("kernel", kernel); apm_b200.set_option("mode", mode)
        got = apm_b200.count_matches(FX[c["text"]], c["patterns"], c["k"])
        ok &= got == c["expected"]
        print(name, kernel, mode, "ok" if got == c["expected"] else f"MISMATCH {got}")
apm_b200.set_option("kernel", "auto")
text = oracle.synth_text(0x5EED0001, 5, 20000).tobytes()
pats = [text[100:164], text[3000:3300], text[9000:9033], text[-20:] + b"ACGTAC"]
want = oracle.count_matches(text, pats, 3)
for mode in ("direct", "band"):
    apm_b200.set_option("mode", mode)
    dev = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
    dev[7:7 + len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
    with apm_b200.Plan(pats, 3) as plan:
        plan.count_device(dev.data_ptr() + 7, 0, len(text), len(text), 0, 9999)
        plan.count_device(dev.data_ptr() + 7, 0, len(text), len(text), 9999, len(text))
        got = plan.read_counts()
    ok &= got == want
    print("shards", mode, "ok" if got == want else f"MISMATCH {got} {want}")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
