#!/usr/bin/env python
"""Instruction mix of the innermost hot loop of a kernel in libapm_b200.so (from cuobjdump -sass).
usage: sass_mix.py <mangled-substring> [min_body]"""
import collections, re, subprocess, sys
lib = "inf560-approximate-pattern-matching_b200/libapm_b200.so"
pat = sys.argv[1]
minbody = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    best = None
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:.*?)0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                body = [x for x in ins if tgt <= x[0] <= a]
                if len(body) >= minbody and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        print(name, "no loop found"); continue
    def op(x):
        t = re.sub(r"@!?U?P\d+\s+", "", x).split()[0]
        parts = t.split(".")
        if parts[0] == "IMAD" and len(parts) > 1 and parts[1] in ("WIDE", "HI", "MOV", "IADD", "SHL", "X", "U32"):
            return "IMAD." + parts[1]
        return parts[0]
    c = collections.Counter(op(x[1]) for x in best)
    alu = sum(v for k, v in c.items() if k in ("LOP3", "SHF", "IADD3", "LEA", "PRMT", "VIADD", "ISETP", "SEL", "MOV", "POPC") )
    fma = sum(v for k, v in c.items() if k.startswith("IMAD"))
    print(name, "loop", len(best), "ALU", alu, "FMA", fma, dict(c.most_common(14)))
