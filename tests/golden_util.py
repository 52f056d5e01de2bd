"""Loaders for the committed golden vectors (tests/golden/, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures() -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, "fixtures.npz"))
    return {k: z[k].tobytes() for k in z.files}


def cases() -> list:
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        cs = json.load(f)["cases"]
    for c in cs:
        c["patterns"] = [p.encode("latin1") for p in c["patterns_latin1"]]
    return cs


def lev_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "lev_cases.npz"))
    lens, dist, a, b = z["lens"], z["dist"], z["a"], z["b"]
    off = 0
    out = []
    for ln, d in zip(lens, dist):
        out.append((a[off:off + ln].tobytes(), b[off:off + ln].tobytes(), int(ln), int(d)))
        off += int(ln)
    return out
