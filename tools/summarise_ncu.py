#!/usr/bin/env python
"""Summarise ncu reports into profiles/: python tools/summarise_ncu.py <out.json> <label>=<rep.ncu-rep> ...
Also writes <out>_<label>_stalls.txt from the source page (needs -lineinfo + --import-source on)."""
import csv, io, json, os, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'sm__cycles_elapsed.avg.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
out_path = sys.argv[1]
out = {}
here = os.path.dirname(os.path.abspath(__file__))
for arg in sys.argv[2:]:
    label, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    out[label] = {"kernel": d.get("Kernel Name"), **{k: {"value": d.get(k), "unit": u.get(k)} for k in KEYS if k in d}}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = out_path.replace(".json", f"_{label}_src.csv")
    open(tmp, "w").write(src)
    st = subprocess.run([sys.executable, os.path.join(here, "ncu_stalls.py"), tmp, "12"], capture_output=True, text=True).stdout
    open(out_path.replace(".json", f"_{label}_stalls.txt"), "w").write(st)
    os.remove(tmp)
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1)[:600])
