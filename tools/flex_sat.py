#!/usr/bin/env python
"""How often do the clamps fire in the flex benchmark configurations?  (oracle-free: counts |effort| at the limits from the GPU run)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
n = 1 << 14
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
for pc, dc in ((0, 0), (1, 1)):
    cfg = cb.default_config(8); cfg.velocity_epsilon = 1e-12
    cfg.vel_pid.p_cascade = pc; cfg.vel_pid.d_cascade = dc
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        hits = 0; tot = 0
        for k in range(60):
            g.step(17)
            eff = g.pid_terms()[:, :, 4]
            hits += int((np.abs(eff) >= min(cfg.vel_pid.cmd_limit, cfg.effort_limit) - 1e-9).sum()); tot += eff.size
        print("cascades", pc, dc, "cmd_limit", cfg.vel_pid.cmd_limit, "effort_limit", cfg.effort_limit, "saturated samples", hits, "of", tot)
