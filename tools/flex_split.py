#!/usr/bin/env python
"""Where the flex kernel's time goes: the same hold-capable code with and without hold transitions, with and without filters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

n, k = (1 << 20), 1000
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
nc = int(os.environ.get("NC", "8"))
for name, eps, pc, dc in (("hold code, never holds", 1e-12, 0, 0), ("hold 2cm/s", 0.02, 0, 0), ("always holds", 1.0, 0, 0),
                          ("never holds +1P1D", 1e-12, 1, 1), ("hold 2cm/s +1P1D", 0.02, 1, 1), ("no hold code +1P1D", -1.0, 1, 1)):
    cfg = cb.default_config(nc)
    cfg.velocity_epsilon = eps
    cfg.vel_pid.p_cascade = pc; cfg.vel_pid.d_cascade = dc
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(3):
            g.step(k); ms.append(g.last_kernel_ms)
        t = float(np.mean(ms[1:]))
        print(f"NC={nc} LANES={os.environ.get('CDPR_FLEX_LANES','auto')} {name:28s} variant={g.kernel_variant} {n * k / (t * 1e-3):.3e}  ({t:.1f} ms)", flush=True)
