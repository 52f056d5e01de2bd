import os, sys, json
sys.path[:0]=["/root/repo","/root/repo/inf560-approximate-pattern-matching_b200"]
import torch, apm_b200
from apm_b200.synth import TEXT_SEED, make_patterns
n = 64 << 20
text = torch.empty(n, dtype=torch.uint8, device="cuda")
apm_b200.synth_text_device(text.data_ptr(), TEXT_SEED, 0, n); torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
for m in (48, 50, 56, 64, 72, 96, 128, 136, 192, 200, 256):
    P = 64; slab = (4 << 20) * 64 // m
    pats,_,_ = make_patterns(TEXT_SEED, n, P, m, 7)
    with apm_b200.Plan(pats, 4) as plan:
        for i in range(2): plan.count_device(text.data_ptr(), 0, n, n, 0, slab, st)
        torch.cuda.synchronize()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3): plan.count_device(text.data_ptr(), 0, n, n, i*1000, i*1000+slab, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/3
    print(m, "%.1f TCUPS" % (slab*P*m*m/ms/1e9), flush=True)
