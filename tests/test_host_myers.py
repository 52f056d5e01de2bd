"""CPU check of the bit-parallel recurrence used by the CUDA kernel (host build of csrc/apm_myers.cuh's
myers_step / add_chain / score, see tests/host_myers_check.cpp) against the oracle and the reference's
levenshtein() known answers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle
from tests.golden_util import lev_cases

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_host_myers_check.so")


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "host_myers_check.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                        "-Wno-unknown-pragmas", "-o", SO, src], check=True)
    L = C.CDLL(SO)
    L.host_myers_distance.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    return L


def test_reference_known_answers(hm):
    for a, b, ln, d in lev_cases():
        assert hm.host_myers_distance(a, b, ln) == d


def test_random_all_word_counts(hm):
    rng = np.random.default_rng(11)
    for ln in list(range(1, 70)) + [95, 96, 97, 127, 128, 129, 160, 191, 192, 193, 200, 224, 225, 255, 256]:
        for na in (1, 2, 4):
            a = rng.integers(65, 65 + na, ln, dtype=np.uint8).tobytes()
            b = rng.integers(65, 65 + na, ln, dtype=np.uint8).tobytes()
            assert hm.host_myers_distance(a, b, ln) == oracle.levenshtein(a, b, ln), (ln, na)
        # carries that ripple through every word: identical strings, and a single early edit
        a = bytes([65 + (i % 3) for i in range(ln)])
        assert hm.host_myers_distance(a, a, ln) == 0
        b = bytes([90]) + a[1:]
        assert hm.host_myers_distance(a, b, ln) == oracle.levenshtein(a, b, ln)
