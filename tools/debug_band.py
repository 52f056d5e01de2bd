import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf560-approximate-pattern-matching_b200"))
import torch, apm_b200
from tests.golden_util import fixtures, cases
from oracle import oracle
FX = fixtures(); C = [c for c in cases() if c['name'] == 'x100_k5'][0]
text = FX[C['text']]; pat = C['patterns'][1]; n = len(text)
dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
def counts(mode, a, b, k):
    apm_b200.set_option("mode", mode)
    with apm_b200.Plan([pat], k) as plan:
        plan.count_device(dev.data_ptr(), 0, n, n, a, b)
        return plan.read_counts()[0]
for j in (132395, 132394):
    for k in (4, 5, 6, 7):
        row = []
        for s in (0, 1, 2, 3, 5, 6, 31, 32, 33, 64, 4095, 4096):
            a = j - s
            row.append((s, counts("direct", a, j + 1, k) - counts("direct", a, j, k), counts("band", a, j + 1, k) - counts("band", a, j, k)))
        print("j", j, "k", k, "true", oracle.levenshtein(pat, text[j:j+50], 50), row)
