#!/usr/bin/env python
"""Stall samples of an ncu report by address region (chunks of N instructions) with instruction counts per warp-step, thread
activity and the top stall reasons; then the hottest instructions.  usage: python tools/ncu_regions.py rep warp_steps [chunk] [top]"""
import subprocess, csv, io, collections, sys
rep = sys.argv[1]; wsteps = float(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 128; top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    data.append((int(r[ix["Address"]][-8:], 16), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), float(r[ix["Avg. Threads Executed"]] or 0), r[ix["Source"]], {h: int(r[ix[h]] or 0) for h in stalls}))
base = min(d[0] for d in data); tot = sum(d[1] for d in data)
print("kernel", rows[0][1][:90], "samples", tot)
agg = collections.OrderedDict()
for a, s, ex, at, t, st in sorted(data):
    d = agg.setdefault((a - base) // (16 * chunk), [0, 0, 0.0, collections.Counter()])
    d[0] += s; d[1] += ex; d[2] += at * ex
    for h, v in st.items(): d[3][h[6:]] += v
for k, d in agg.items():
    if d[0] > tot * 0.004:
        why = ", ".join(f"{h}:{100 * v / d[0]:.0f}%" for h, v in d[3].most_common(3))
        print(f"{k * 16 * chunk:#8x} samples {100 * d[0] / tot:5.1f}%  instr/warp-step {d[1] / wsteps:7.1f} threads {d[2] / max(d[1], 1):5.1f}  {why}")
print("hottest instructions:")
for a, s, ex, at, t, st in sorted(data, key=lambda d: -d[1])[:top]:
    why = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    print(f"  {a - base:#7x} {100 * s / tot:5.2f}% x{ex / wsteps:6.3f} thr{at:5.1f} {t[:64]:64s} {why}")
