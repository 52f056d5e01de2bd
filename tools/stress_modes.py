#!/usr/bin/env python
"""Randomised cross-check of the exact shortcut modes against the explicit-DP kernel on adversarial inputs:
tiny alphabets, periodic texts and patterns (identical seeds in several pieces, overlapping witnesses), edits near the
piece boundaries, k up to 17, pattern lengths 2..1100, every DP-cell code.  usage: stress_modes.py [cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import numpy as np
import apm_b200

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    sigma = int(rng.choice([1, 2, 2, 3, 4, 4, 6, 20]))
    alpha = rng.choice(256, size=sigma, replace=False).astype(np.uint8)
    n = int(rng.integers(2_000, 60_000))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        text = alpha[rng.integers(0, sigma, size=n)]
    elif kind == 1:  # periodic
        per = alpha[rng.integers(0, sigma, size=int(rng.integers(1, 40)))]
        text = np.tile(per, n // len(per) + 1)[:n].copy()
        flips = rng.integers(0, n, size=n // 200)
        text[flips] = alpha[rng.integers(0, sigma, size=len(flips))]
    elif kind == 2:  # runs
        text = np.repeat(alpha[rng.integers(0, sigma, size=n // 7 + 1)], rng.integers(1, 14, size=n // 7 + 1))[:n].copy()
    else:  # mostly random with long copies of itself
        text = alpha[rng.integers(0, sigma, size=n)]
        for _ in range(5):
            a, b, L = int(rng.integers(0, n - 500)), int(rng.integers(0, n - 500)), int(rng.integers(50, 400))
            text[b:b + L] = text[a:a + L]
    text = text.tobytes()
    k = int(rng.integers(0, 18)) if rng.integers(0, 4) == 0 else int(rng.integers(0, 9))
    pats = []
    for _ in range(int(rng.integers(1, 8))):
        m = int(rng.integers(max(2, k), 1100 if rng.integers(0, 10) == 0 else 300))
        m = min(m, len(text) - 1)
        off = int(rng.integers(0, max(1, len(text) - m)))
        p = bytearray(text[off:off + m])
        for _ in range(int(rng.integers(0, k + 3))):
            op, x = int(rng.integers(0, 3)), int(rng.integers(0, len(p)))
            if op == 0: p[x] = int(alpha[rng.integers(0, sigma)])
            elif op == 1: p.insert(x, int(alpha[rng.integers(0, sigma)]))
            elif len(p) > 1: del p[x]
        if p: pats.append(bytes(p))
    apm_b200.set_option("kernel", "dp"); apm_b200.set_option("mode", "direct")
    want, want_hits, _ = apm_b200.find_matches(text, pats, k, max_hits=1 << 22)
    apm_b200.set_option("kernel", "auto")
    for mode, extra in (("filter", {}), ("filter", {"filter_cand_mb": "1"}), ("band", {}), ("direct", {})):
        apm_b200.set_option("mode", mode)
        apm_b200.set_option("cell", str(rng.choice(["auto", "lop3", "fma3", "fma", "fma3r"])))
        for kk, vv in extra.items(): apm_b200.set_option(kk, vv)
        got, hits, _ = apm_b200.find_matches(text, pats, k, max_hits=1 << 22)
        for kk in extra: apm_b200.set_option(kk, "128")
        if got != want or hits != want_hits:
            bad += 1
            print(f"MISMATCH case {c} mode {mode} {extra} sigma {sigma} kind {kind} n {n} k {k} m {[len(p) for p in pats]} got {got} want {want}", flush=True)
print(f"{cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
