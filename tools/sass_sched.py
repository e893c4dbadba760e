#!/usr/bin/env python
"""Decode the scheduling control bits of a kernel's SASS (stall count = bits 105..108 of each 128-bit instruction,
yield = bit 109, wait mask = bits 116..121) and print a region with them.
usage: python tools/sass_sched.py <lib.so> <function-substring> <start-hex> <end-hex>"""
import re, subprocess, sys
lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if pat not in f.split("\n", 1)[0]:
        continue
    lines = f.split("\n")
    total = 0; n = 0; fp64 = 0
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            addr = int(m.group(1), 16)
            if m2 and lo <= addr <= hi:
                hiw = int(m2.group(1), 16)
                stall = (hiw >> 41) & 0xf
                yld = (hiw >> 45) & 1
                wbar = (hiw >> 46) & 7
                rbar = (hiw >> 49) & 7
                wmask = (hiw >> 52) & 0x3f
                total += max(stall, 1); n += 1
                op = m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1]
                if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"): fp64 += 1
                print(f"{addr:6x} s{stall:2d} {'Y' if yld else ' '} w{wbar if wbar<7 else '-'} r{rbar if rbar<7 else '-'} m{wmask:02x}  {m.group(2)[:90]}")
            i += 2
        else:
            i += 1
    print(f"-- {n} instr, sum of stall counts {total}, fp64-pipe instr {fp64}")
