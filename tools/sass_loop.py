#!/usr/bin/env python
"""Instruction mix of the outermost loop(s) of one kernel, from SASS.
usage: python tools/sass_loop.py obj_or_so function_substring"""
import collections, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [x for x in funcs if sys.argv[2] in x.split("\n", 1)[0]][0]
ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)]
print("function:", f.split("\n", 1)[0][:100], "instructions:", len(ins))
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
loops.sort(key=lambda l: l[0] - l[1])
for lo, hi in loops[:6]:
    body = [t for a, t in ins if lo <= a <= hi]
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(f"loop {lo:#x}..{hi:#x}: {len(body)} instr;", ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
