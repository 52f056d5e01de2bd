#!/usr/bin/env python
"""Exploration bench (not the contract bench.py): int-ALU peak microbenchmarks and slab timings of the
count kernel for several (m, k, P, rblock, tile) combinations.  Writes JSON lines to stdout."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))

import torch  # noqa: E402

import apm_b200  # noqa: E402
from apm_b200.synth import TEXT_SEED, make_patterns  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    props = torch.cuda.get_device_properties(dev)
    print(json.dumps({"gpu": props.name, "sms": props.multi_processor_count}), flush=True)
    peaks = {}
    for kind, name in ((0, "lop3+iadd3"), (1, "lop3"), (3, "lop3+imad")):
        ops, sec = apm_b200.int_peak(kind)
        peaks[name] = ops
        print(json.dumps({"int_peak": name, "Tops": ops / 1e12, "sec": sec,
                          "lanes_per_clk_per_sm_at_1965MHz": ops / props.multi_processor_count / 1.965e9}), flush=True)
    n = 256 << 20
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    combos = []
    for mode in ("direct", "band"):
      for cell in ("lop3", "fma3", "fma", "fma3r"):
        combos.append((64, 4, 256, "4", "512", 16 << 20, "0", "auto", mode, cell))
        combos.append((200, 10, 64, "1", "512", 4 << 20, "0", "auto", mode, cell))
        combos.append((32, 2, 256, "4", "512", 32 << 20, "0", "auto", mode, cell))
        combos.append((50, 0, 256, "4", "512", 16 << 20, "0", "auto", mode, cell))
        combos.append((1000, 16, 16, "1", "512", 2 << 20, "0", "auto", mode, cell))
    for m, k, P, rb, tile, slab, var, kern, mode, cell in combos:
        apm_b200.set_option("cell", cell)
        apm_b200.set_option("rblock", rb)
        apm_b200.set_option("tile", tile)
        apm_b200.set_option("variant", var)
        apm_b200.set_option("kernel", kern)
        apm_b200.set_option("mode", mode)
        pats, _, _ = make_patterns(TEXT_SEED, n, P, m, 7)
        with apm_b200.Plan(pats, k) as plan:
            def step(i):
                a = (i * slab) % (n - slab - m)
                plan.count_device(text.data_ptr(), 0, n, n, a, a + slab, st)
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for i in range(reps):
                step(2 + i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nw = (m + 31) // 32
            cells = slab * P * m * m
            ops = slab * P * m * nw * 10
            print(json.dumps({"m": m, "k": k, "P": P, "kernel": kern, "mode": mode, "cell": cell, "rblock": rb, "tile": tile, "variant": var, "slab": slab, "ms": ms,
                              "GCUPS": cells / ms / 1e6, "Tiops": ops / ms / 1e9,
                              "frac_of_lop3_iadd3_peak": ops / (ms * 1e-3) / peaks["lop3+iadd3"]}), flush=True)


if __name__ == "__main__":
    main()
