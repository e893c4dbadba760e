#!/usr/bin/env python
"""Throughput of the full-semantics (flex) kernel: hold below 2 cm/s + one biquad stage on the P input and the D output,
per-instance sine commands (the configuration of bench.py's extra.general_variant_nc8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

n, k = (1 << 20), 1000
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
for nc in (8, 4):
    for name, edit in (("hold+1P+1D", lambda c: (setattr(c, "velocity_epsilon", 0.02), setattr(c.vel_pid, "p_cascade", 1), setattr(c.vel_pid, "d_cascade", 1))),
                       ("hold only", lambda c: setattr(c, "velocity_epsilon", 0.02)),
                       ("launch values, independent", None)):
        cfg = cb.default_config(nc)
        if edit:
            edit(cfg)
        with cb.CdprBatch(cfg, n) as g:
            if edit is None:
                g.set_independent(True)
            g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
            ms = []
            for _ in range(3):
                g.step(k); ms.append(g.last_kernel_ms)
            t = float(np.mean(ms[1:]))
            print(f"NC={nc} {name:28s} variant={g.kernel_variant} {n * k / (t * 1e-3):.3e} instance-steps/s  ({t:.1f} ms per 2^20 x 1000)", flush=True)
