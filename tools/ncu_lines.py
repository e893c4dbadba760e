#!/usr/bin/env python
"""Stall samples of one ncu report by SOURCE LINE (and by inlined function), using the line table of the cubin the report was
taken from.  usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_mangled_substring [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    data.append((int(r[ix["Address"]][-8:], 16), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), float(r[ix["Avg. Threads Executed"]] or 0), r[ix["Source"]]))
base = min(d[0] for d in data)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
sec = re.split(r"\n\s*\.text\.", dis)
body = [x for x in sec if key in x.split("\n", 1)[0]][0]
cur = None; off2 = {}
for ln in body.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m: off2[int(m.group(1), 16)] = cur
tot = sum(d[1] for d in data); totex = sum(d[2] for d in data)
byline = collections.Counter(); exline = collections.Counter(); outer = collections.Counter(); exouter = collections.Counter()
for a, s, ex, at, t in data:
    k = off2.get(a - base)
    byline[k[:2] if k else None] += s; exline[k[:2] if k else None] += ex
    o = (k[2][-1] if k[2] else k[:2]) if k else None   # outermost call site = line of the kernel body
    outer[o] += s; exouter[o] += ex
print(f"samples {tot}  warp-instructions {totex}")
print("-- by line of the outermost (kernel-level) call site")
for k, v in outer.most_common(top): print(f"  {str(k):34s} {100 * v / tot:5.1f}%  instr {100 * exouter[k] / totex:5.1f}%")
print("-- by innermost line")
for k, v in byline.most_common(top): print(f"  {str(k):34s} {100 * v / tot:5.1f}%  instr {100 * exline[k] / totex:5.1f}%")
