#!/usr/bin/env python
"""Generate tests/golden/{fixtures.npz,golden.json,lev_cases.npz} from the REFERENCE itself.

Run in the build container only (needs /root/reference and oracle/_ref, see oracle/Makefile):

    make -C oracle ref && python tests/golden/make_golden.py

Every expected value below is the stdout of the reference's own `apm_sequential`
(src/sequential.c + src/utils.c compiled unmodified) or the return value of the reference's
`levenshtein()` (src/utils.c:76-99) called through oracle/_ref/libapm_ref.so.  The inputs are the
reference's fixture texts under dna/ (stored here compressed, as test vectors) and patterns from its
test scripts (scripts/basic_test.batch:10, scripts/run_tests:31,51) plus the BASELINE.json configs.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = os.environ.get("APM_REFERENCE", "/root/reference")
DNA = os.path.join(REF, "dna")


def rd(name: str) -> bytes:
    with open(os.path.join(DNA, name), "rb") as f:
        return f.read()


def run_ref(k: int, text_file: str, patterns: list[bytes]) -> tuple[list[int], float]:
    """Run the reference binary; patterns go through argv exactly as the reference's scripts do."""
    argv = [oracle.REF_BINARY.encode(), str(k).encode(), os.path.join(DNA, text_file).encode()] + patterns
    t0 = time.time()
    out = subprocess.run(argv, check=True, capture_output=True).stdout
    dt = time.time() - t0
    # a pattern may contain '\n', so parse counts with a regex anchored on the fixed phrase
    counts = [int(x) for x in re.findall(rb">: (-?\d+)\n", out)]
    assert len(counts) == len(patterns), out[-400:]
    return counts, dt


def config2_patterns(text: bytes) -> list[bytes]:
    """BASELINE.json configs[1] / SURVEY.md section 8d config 2: 64 patterns of length 32 cut from the
    text at offset (p*2053) mod (N-32), with p mod 4 substitutions at fixed positions."""
    n = len(text)
    pats = []
    for p in range(64):
        off = (p * 2053) % (n - 32)
        b = bytearray(text[off:off + 32])
        for s in range(p % 4):
            pos = (7 + 11 * s) % 32
            c = b[pos]
            b[pos] = b"CGTA"[b"ACGT".index(c)] if c in b"ACGT" else ord("A")
        assert 0 not in b
        pats.append(bytes(b))
    return pats


def main() -> None:
    assert os.path.exists(oracle.REF_BINARY), "build oracle/_ref first (make -C oracle ref)"
    texts = {name: rd(name + ".fa") for name in
             ["easy", "small_chrY", "small_chrY_x100", "small_chrY_medium"]}
    line = {n: rd(f"line_{n}.fa") for n in ["5", "10", "20", "1131", "20783", "non_existent"]}
    x100 = texts["small_chrY_x100"]
    bigger = rd("small_chrY_bigger.fa").replace(b"\n", b"")
    m64 = bigger[1000:1064]

    cases = []

    def add(name, text, k, pats, note=""):
        counts, dt = run_ref(k, text + ".fa", pats)
        cases.append({"name": name, "text": text, "k": k,
                      "patterns_latin1": [p.decode("latin1") for p in pats],
                      "expected": counts, "source": "reference apm_sequential", "note": note,
                      "ref_seconds": round(dt, 3)})
        print(f"{name}: k={k} -> {counts[:8]}{'...' if len(counts) > 8 else ''} ({dt:.1f}s)", flush=True)

    # BASELINE.json configs[0] == scripts/basic_test.batch:10 == README.md:55-63
    add("config1_readme", "small_chrY_x100", 0,
        [line["non_existent"]] + [line["20783"]] * 5, "README.md:58-63 publishes 0,4,4,4,4,4")
    # scripts/run_tests:31 and :51 inputs
    for k in (0, 1, 2):
        add(f"easy_k{k}", "easy", k, [b"123", b"456", b"78934"], "scripts/run_tests:31")
    add("run_tests_complex", "small_chrY_x100", 0,
        [line["10"], line["20"], line["non_existent"]] * 2, "scripts/run_tests:51")
    for k in (1, 2, 3, 5, 10):
        add(f"x100_k{k}", "small_chrY_x100", k,
            [line["10"], line["20"], line["20783"], line["non_existent"]])
    for k in (0, 1, 2, 4, 8, 16, 25, 49, 50, 60):
        add(f"small_k{k}", "small_chrY", k,
            [line["10"], line["20"], b"ACGT", b"A", b"T", b"TT", b"TTTT", line["5"], line["1131"]])
    add("x100_m64_k4", "small_chrY_x100", 4, [m64], "m=64: two 32-bit words")
    add("medium_mixedcase_k1", "small_chrY_medium", 1, [line["10"], line["5"], b"acgtn", b"NNNN"],
        "text holds lowercase and N")
    # BASELINE.json configs[1]
    add("config2", "small_chrY_x100", 2, config2_patterns(x100),
        "64 x m=32, k=2; some patterns contain a newline byte")
    # m = 200 (7 words) on the small text, k = 10 (configs[3] shape), pattern cut from x100 w/o newlines
    flat = x100.replace(b"\n", b"")
    p200 = bytearray(flat[5000:5200])
    for pos in (3, 50, 97, 150, 199):
        p200[pos] = ord("A") if p200[pos] != ord("A") else ord("C")
    add("small_m200_k10", "small_chrY", 10, [bytes(p200), flat[100:300]], "m=200: seven 32-bit words")

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"),
                        **{k: np.frombuffer(v, dtype=np.uint8) for k, v in texts.items()})

    # Known-answer vectors for levenshtein() itself (utils.c:76-99 via oracle/_ref/libapm_ref.so)
    rng = np.random.default_rng(560)
    alph = np.frombuffer(b"ACGT\nNacgt", dtype=np.uint8)
    lens, a_all, b_all, d_all = [], [], [], []
    for i in range(3000):
        ln = int(rng.integers(1, 201)) if i % 3 else int(rng.integers(1, 40))
        na = 2 + int(rng.integers(0, 4)) if i % 2 else len(alph)
        a = alph[rng.integers(0, na, ln)]
        b = a.copy()
        if i % 5:  # related strings: a few edits
            for _ in range(int(rng.integers(0, 8))):
                b[int(rng.integers(0, ln))] = alph[int(rng.integers(0, na))]
            if i % 7 == 0 and ln > 3:
                b = np.roll(b, 1)
        else:
            b = alph[rng.integers(0, na, ln)]
        d = oracle.ref_levenshtein(a.tobytes(), b.tobytes(), ln)
        lens.append(ln); a_all.append(a); b_all.append(b); d_all.append(d)
    np.savez_compressed(os.path.join(HERE, "lev_cases.npz"),
                        lens=np.array(lens, dtype=np.int32), dist=np.array(d_all, dtype=np.int32),
                        a=np.concatenate(a_all), b=np.concatenate(b_all))
    print("wrote golden.json, fixtures.npz, lev_cases.npz")


if __name__ == "__main__":
    main()
