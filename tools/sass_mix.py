#!/usr/bin/env python
"""Small SASS helper: per-function instruction mix, loop bodies (back-edges) and their pipe mix.
usage: python tools_sass.py <lib.so> <function-substring>"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("==", name, len(ins), "instructions")
    def mix(sub):
        c = collections.Counter()
        for a, t in sub:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
        return c
    def show(c):
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        tot = sum(c.values())
        print("   total", tot, "fp64-pipe", fp64, " ".join(f"{k}:{v}" for k, v in c.most_common(18)))
    show(mix(ins))
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [(x, y) for x, y in ins if tgt <= x <= a]
            print(f" loop {tgt:#x}..{a:#x}: {len(body)} instr")
            show(mix(body))
            # sub-split by forward branches spanning big regions
