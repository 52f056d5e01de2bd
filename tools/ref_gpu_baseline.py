#!/usr/bin/env python
"""The reference's OWN CUDA kernel (src/patterns_over_ranks.cu ComputeMatches, compiled unmodified for sm_100a into
oracle/_ref/libapm_refgpu.so) timed on the B200 through its extern "C" entry points invoke_kernel /
write_kernel_result, next to this repo's kernels on the same input.  A reported baseline ("reference GPU" line of
SURVEY.md 8d), not an oracle: its match counter is incremented without atomics (patterns_over_ranks.cu:67-69)."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import apm_b200
from oracle import oracle
# The reference kernel mallocs its DP column from the device heap in every thread (patterns_over_ranks.cu:31); the
# default 8 MB heap is exhausted by the ~300k resident threads of a B200 (the NULL is not checked), so the heap limit of
# the primary context is raised first.
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12")
assert rt.cudaDeviceSetLimit(C.c_int(2), C.c_size_t(4 << 30)) == 0  # cudaLimitMallocHeapSize
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libapm_refgpu.so"))
L.invoke_kernel.restype = C.c_void_p
L.invoke_kernel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
L.write_kernel_result.argtypes = [C.c_void_p, C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
text = oracle.synth_text(0x5EED0001, 0, n).tobytes()
tb = C.create_string_buffer(text, n)
out = []
for m, k in ((50, 0), (64, 4)):
    pat = text[12345:12345 + m]
    pb = C.create_string_buffer(pat, m)
    ours = {}
    for mode in ("direct", "filter"):
        apm_b200.set_option("mode", mode)
        apm_b200.count_matches(text, [pat], k)
        t0 = time.perf_counter()
        got = apm_b200.count_matches(text, [pat], k)
        ours[mode] = time.perf_counter() - t0
    best = None
    for rep in range(3):
        local = C.c_int(0)
        t0 = time.perf_counter()
        h = L.invoke_kernel(C.addressof(tb), n, C.addressof(pb), m, k, C.byref(local))
        L.write_kernel_result(C.byref(local), h)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cells = float(n) * m * m
    out.append({"n_bytes": n, "m": m, "k": k, "reference_gpu_matches": local.value, "our_matches": got[0],
                "reference_gpu_ms": best * 1e3, "reference_gpu_GCUPS": cells / best / 1e9,
                "ours_direct_ms": ours["direct"] * 1e3, "ours_direct_GCUPS": cells / ours["direct"] / 1e9,
                "ours_filter_ms": ours["filter"] * 1e3,
                "note": "one pattern, whole call incl. H2D (the reference API copies the text per pattern)"})
    print(json.dumps(out[-1]), flush=True)
