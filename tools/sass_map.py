#!/usr/bin/env python
"""Instruction-category map of the largest loop of one kernel, in chunks of N instructions (where are the FP64 ops, the
shared / local memory accesses, the calls).  usage: python tools/sass_map.py obj function_substring [chunk]"""
import collections, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [x for x in funcs if sys.argv[2] in x.split("\n", 1)[0]][0]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 60
ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)]
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
lo, hi = max(loops, key=lambda l: l[1] - l[0])
seg = [(a, t) for a, t in ins if lo <= a <= hi]
print(f"loop {lo:#x}..{hi:#x}: {len(seg)} instructions")
cat = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
for k in range(0, len(seg), chunk):
    c = collections.Counter(cat(t) for a, t in seg[k:k + chunk])
    fp = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
    mv = c["IMAD"] + c["MOV"]
    print(f"{seg[k][0]:#7x} fp64={fp:2d} LDS={c['LDS']:2d} STS={c['STS']:2d} LDL={c['LDL']:2d} STL={c['STL']:2d} LDC={c['LDC'] + c['LDCU']:2d} BRA={c['BRA']:2d} CALL={c['CALL']} IMAD/MOV={mv:2d} SEL={c['SEL'] + c['FSEL']:2d}")
