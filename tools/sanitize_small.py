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
    for kernel, mode in (("auto", "direct"), ("auto", "band"), ("auto", "filter"), ("myers", "direct"), ("myers", "filter"),
                         ("dp", "direct")):
        apm_b200.set_option("kernel", kernel); apm_b200.set_option("mode", mode)
        got = apm_b200.count_matches(FX[c["text"]], c["patterns"], c["k"])
        ok &= got == c["expected"]
        print(name, kernel, mode, "ok" if got == c["expected"] else f"MISMATCH {got}")
apm_b200.set_option("kernel", "auto")
text = oracle.synth_text(0x5EED0001, 5, 20000).tobytes()
pats = [text[100:164], text[3000:3300], text[9000:9033], text[-20:] + b"ACGTAC"]
want = oracle.count_matches(text, pats, 3)
for cell in ("lop3", "fma3", "fma", "fma3r"):
    apm_b200.set_option("cell", cell)
    got = apm_b200.count_matches(text, pats, 3)
    ok &= got == want
    print("cell", cell, "ok" if got == want else f"MISMATCH {got} {want}")
apm_b200.set_option("cell", "auto")
# round 2: both filter scans, both tail kernels, ragged single-block lengths, k up to 15
pats2 = [text[200:212], text[500:520], text[700:750], text[1000:1063], text[5000:5200], text[-31:] + b"ACGTACGTACGTA"]
for k2 in (0, 1, 5):
    want2 = oracle.count_matches(text, pats2, k2)
    for mode, scan, tail in (("direct", "auto", "bitpar"), ("direct", "auto", "dp"), ("filter", "dna", "bitpar"),
                             ("filter", "hash", "dp"), ("band", "auto", "bitpar")):
        apm_b200.set_option("mode", mode); apm_b200.set_option("filter_scan", "auto" if k2 == 0 and scan == "dna" else scan)
        apm_b200.set_option("tail", tail)
        got = apm_b200.count_matches(text, pats2, k2)
        ok &= got == want2
        print("round2", k2, mode, scan, tail, "ok" if got == want2 else f"MISMATCH {got} {want2}")
apm_b200.set_option("filter_scan", "auto"); apm_b200.set_option("tail", "auto"); apm_b200.set_option("mode", "direct")
for mode in ("direct", "band", "filter"):
    apm_b200.set_option("mode", mode)
    counts, hits, n_hits = apm_b200.find_matches(text, pats, 3, max_hits=64)
    ok &= counts == want and n_hits == sum(want)
    print("positions", mode, "ok" if counts == want and n_hits == sum(want) else f"MISMATCH {counts} {n_hits}")
    dev = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
    dev[7:7 + len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
    with apm_b200.Plan(pats, 3) as plan:
        plan.count_device(dev.data_ptr() + 7, 0, len(text), len(text), 0, 9999)
        plan.count_device(dev.data_ptr() + 7, 0, len(text), len(text), 9999, len(text))
        got = plan.read_counts()
    ok &= got == want
    print("shards", mode, "ok" if got == want else f"MISMATCH {got} {want}")
# round 2 (late): automatic routing -- m = 32 only (two-row sweep), m = 64 only, 65..224 and longer lists -- and the
# resident 2-bit text copy
apm_b200.set_option("mode", "direct"); apm_b200.set_option("cell", "auto")
for pset in ([text[i:i + 32] for i in (10, 500, 7000)], [text[i:i + 64] for i in (10, 500, 7000)],
             [text[100:200], text[300:524], text[600:825], text[-40:] + b"ACGT" * 10]):
    got = apm_b200.count_matches(text, pset, 2)
    w3 = oracle.count_matches(text, pset, 2)
    ok &= got == w3
    print("auto routing", [len(p) for p in pset], "ok" if got == w3 else f"MISMATCH {got} {w3}")
apm_b200.set_option("mode", "filter")
dtext = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
pk = torch.empty(apm_b200.text_pack_bytes(len(text)), dtype=torch.uint8, device="cuda")
apm_b200.text_pack_device(dtext.data_ptr(), len(text), pk.data_ptr())
with apm_b200.Plan(pats2, 1) as plan:
    plan.count_device_packed(dtext.data_ptr(), pk.data_ptr(), 0, len(text), len(text), 0, len(text))
    got = plan.read_counts()
w4 = oracle.count_matches(text, pats2, 1)
ok &= got == w4
print("packed text", "ok" if got == w4 else f"MISMATCH {got} {w4}")
apm_b200.set_option("mode", "direct")
# filter mode with an overflowing candidate buffer (fallback to the band kernel) on low-complexity text
apm_b200.set_option("mode", "filter"); apm_b200.set_option("filter_cand_mb", "1")
lc = b"A" * 200000
got = apm_b200.count_matches(lc, [b"A" * 40, b"A" * 20 + b"C" + b"A" * 19], 2)
apm_b200.set_option("mode", "band")
want_lc = apm_b200.count_matches(lc, [b"A" * 40, b"A" * 20 + b"C" + b"A" * 19], 2)
ok &= got == want_lc
print("filter overflow", "ok" if got == want_lc else f"MISMATCH {got} {want_lc}")
apm_b200.set_option("filter_cand_mb", "128"); apm_b200.set_option("mode", "direct")
# the reference's entry points by name
from apm_b200 import refcompat
r = refcompat.invoke_and_fetch(text, pats[0], 3, initial=2)
ok &= r == want[0] + 2
print("refcompat invoke_kernel", "ok" if r == want[0] + 2 else f"MISMATCH {r}")
r = refcompat.initialize_and_fetch(text, pats, 3, 10000, 1, 3, 0, 3)
w = [oracle.count_range(text[:10000 + len(p) - 1], p, 3, 0, 10000 + len(p) - 1) for p in pats[:3]] + [0]
ok &= r == w
print("refcompat initializeGPU", "ok" if r == w else f"MISMATCH {r} {w}")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
