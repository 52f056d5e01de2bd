"""Synthetic workloads of the BASELINE configs, host side (numpy).

text[i] = "ACGT"[splitmix64(seed + i) >> 62] -- counter based, so any slice can be produced on the CPU
(here) and on the GPU (apm_synth_text_device) without materialising a file.  Patterns: the first three
quarters are cut from the text at pseudo-random offsets and given `p mod submod` substitutions at fixed
positions (so some lie within the threshold and some beyond); the last quarter is uniformly random.
"""
from __future__ import annotations

import numpy as np

TEXT_SEED = 0x5EED0001
OFFSET_SEED = 0x5EED0002
RANDOM_SEED = 0x5EED0003
_SYM = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def text_slice(seed: int, offset: int, count: int) -> np.ndarray:
    idx = np.arange(offset, offset + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(idx + np.uint64(seed))
    return _SYM[(h >> np.uint64(62)).astype(np.int64)]


def make_patterns(seed: int, n_total: int, nb_patterns: int, m: int, submod: int):
    """-> (patterns, offsets, n_substitutions); offsets[p] is None for the random quarter."""
    pats, offs, nsubs = [], [], []
    cut = (3 * nb_patterns) // 4
    for p in range(nb_patterns):
        if p < cut:
            h = int(splitmix64(np.array([OFFSET_SEED + p], dtype=np.uint64))[0])
            off = h % (n_total - m)
            b = text_slice(seed, off, m).copy()
            nsub = p % submod
            for s in range(nsub):
                pos = (3 + 9 * s) % m
                b[pos] = _SYM[(int(np.where(_SYM == b[pos])[0][0]) + 1) % 4]
            pats.append(b.tobytes()); offs.append(off); nsubs.append(nsub)
        else:
            idx = np.arange(m, dtype=np.uint64) + np.uint64(RANDOM_SEED + 1000003 * p)
            b = _SYM[(splitmix64(idx) >> np.uint64(62)).astype(np.int64)]
            pats.append(b.tobytes()); offs.append(None); nsubs.append(None)
    return pats, offs, nsubs
