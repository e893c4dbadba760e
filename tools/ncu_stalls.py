#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall reasons (totals) and the hottest instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); samples = 0; execd = 0
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[ix["# Samples"]] or 0); samples += s
    execd += int(r[ix["Instructions Executed"]] or 0)
    for h in stalls:
        tot[h] += int(r[ix[h]] or 0)
    data.append((s, r[ix["Address"]], r[ix["Source"]], int(r[ix["Instructions Executed"]] or 0), {h: int(r[ix[h]] or 0) for h in stalls}))
print("samples", samples, "warp-instructions", execd)
for h, v in tot.most_common():
    if v: print(f"  {h:26s} {v:9d} {100.0*v/max(samples,1):6.2f}%")
print("hottest instructions:")
for s, a, src, ex, st in sorted(data, reverse=True)[:top]:
    why = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"  {s:7d} {a[-6:]} x{ex:<10d} {src[:70]:70s} {why}")
