"""CPU check of the bit-sliced DP-cell codes of csrc/apm_sliced.cuh (CELL 0..3): the LOP3 truth tables are READ FROM THE
HEADER, the cell bodies are transliterated below (LOP3 = 8-bit truth table on three words, fma_sub = c - a mod 2^32), and a
complete window-sliced DP over 32 consecutive windows (bit b of a word <-> window j0 + b, exactly the kernel's layout) must
reproduce utils.c:76-99 as restated by the oracle.  No GPU: this pins the algebra (and the constants) the kernels execute."""
import os
import re

import numpy as np
import pytest

from oracle import oracle

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "inf560-approximate-pattern-matching_b200",
                   "csrc", "apm_sliced.cuh")
M32 = 0xFFFFFFFF


def _luts():
    src = open(HDR).read()
    return {m.group(1): int(m.group(2), 16) for m in re.finditer(r"constexpr int (kLut\w+) = (0x[0-9A-Fa-f]+);", src)}


LUT = _luts()


def lop3(name, a, b, c):
    """d = LUT[(a << 2) | (b << 1) | c] per bit position (PTX lop3.b32: the table is evaluated on 0xF0, 0xCC, 0xAA)"""
    t = LUT[name]
    d = 0
    for idx in range(8):
        if (t >> idx) & 1:
            d |= (a if idx & 4 else ~a) & (b if idx & 2 else ~b) & (c if idx & 1 else ~c)
    return d & M32


def fma_sub(c, a):
    return (c - a) & M32


def cell(code, q, ap, am, bp, bm):
    """-> (ap', am', bp', bm') as sliced_cell<CELL> of apm_sliced.cuh"""
    if code == 0:
        d0 = lop3("kLutOr3", q, am, bm)
        vm = lop3("kLutAndOr", bp, q, am)
        vp = lop3("kLutOrNor", bm, d0, bp)
        hp2 = lop3("kLutOrNor", am, d0, ap)
        hm2 = lop3("kLutAndOr", ap, q, bm)
        return vp, vm, hp2, hm2
    if code == 1:
        x = lop3("kLutNor3", q, am, bm)
        vm = lop3("kLutAndOr", bp, q, am)
        vp = lop3("kLutOrAndN", bm, x, bp)
        hm2 = lop3("kLutAndOr", ap, q, bm)
        t1 = fma_sub(ap, x)
        t2 = fma_sub(t1, hm2)
        return vp, vm, fma_sub(am, t2), hm2
    if code == 3:
        x = lop3("kLutNor3", q, am, bm)
        vm = lop3("kLutNotAAndB", x, bp, bp)
        vp = lop3("kLutCOrAAndNotB", x, bp, bm)
        hm2 = lop3("kLutNotAAndC", x, ap, ap)
        t1 = fma_sub(ap, x)
        t2 = fma_sub(t1, hm2)
        return vp, vm, fma_sub(am, t2), hm2
    # code 2: a = (ap = [a != 0], am = [a == -1]), b = (bp = [b == +1], bm = [b != 0])
    vm = lop3("kLutAndOr", bp, q, am)
    s = lop3("kLutOrNor", bm, q, am)
    u = fma_sub(bp, vm)
    vnz = fma_sub(s, u)
    hnz = lop3("kLutXor3", ap, bm, vnz)
    return vnz, vm, lop3("kLutAndOrN", hnz, am, ap), hnz


def sliced_distances(code, text: bytes, pat: bytes):
    """D[m][m] of the 32 windows text[b : b + m], b = 0..31, by the row-major sweep of the kernel"""
    m = len(pat)
    plus2 = M32 if code == 2 else 0  # second plane of a +1 delta (cell_plus_second_plane)
    hp = [M32] * m
    hm = [plus2] * m
    for i in range(m):
        ap, am = M32, 0  # D[i][0] - D[i-1][0] = +1 (the same in every encoding)
        for j in range(m):
            q = 0
            for b in range(32):
                if text[b + j] == pat[i]:
                    q |= 1 << b
            ap, am, hp[j], hm[j] = cell(code, q, ap, am, hp[j], hm[j])
    out = []
    for b in range(32):
        d = m
        for j in range(m):
            plus = (hp[j] >> b) & 1
            minus = ((hm[j] & ~hp[j]) >> b) & 1 if code == 2 else (hm[j] >> b) & 1
            d += plus - minus  # last row: D[m][m] = m + sum_j (D[m][j] - D[m][j-1]) - ... (horizontal deltas)
        out.append(d)
    return out


def test_header_defines_the_tables_the_cells_use():
    for name in ("kLutOr3", "kLutAndOr", "kLutOrNor", "kLutNor3", "kLutOrAndN", "kLutXor3", "kLutAndOrN", "kLutNotAAndB",
                 "kLutNotAAndC", "kLutCOrAAndNotB"):
        assert name in LUT, name


@pytest.mark.parametrize("code", [0, 1, 2, 3])
@pytest.mark.parametrize("alphabet", [b"ACGT", b"AC", b"A"])
def test_sliced_cell_codes_reproduce_levenshtein(code, alphabet):
    rng = np.random.default_rng(100 * code + len(alphabet))
    for m in (1, 2, 5, 17, 33):
        text = bytes(alphabet[i] for i in rng.integers(0, len(alphabet), size=32 + m))
        pat = bytearray(text[3:3 + m]) if rng.integers(0, 2) else bytearray(alphabet[i] for i in rng.integers(0, len(alphabet), size=m))
        if m > 2:
            pat[int(rng.integers(0, m))] = alphabet[int(rng.integers(0, len(alphabet)))]
        got = sliced_distances(code, text, bytes(pat))
        want = [oracle.levenshtein(bytes(pat), text[b:b + m]) for b in range(32)]
        assert got == want, (code, m)
