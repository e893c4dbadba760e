#!/usr/bin/env python
"""Throughput of the full-semantics (flex) kernels at the headline size (2^20 instances x 1000 steps per launch, per-instance sine
commands of the C3 workload): the launch values with independent robots, velocity hold below 2 cm/s, hold + one biquad stage on
the P input and the D output (the reference's filter constants: a loop that saturates two steps out of three), two D stages
(k_step_flex, the classic kernel), and the same hold-capable code when no robot ever crosses the band.
CDPR_FLEX_CLASSIC=1 runs k_step_flex everywhere (the round-2 kernel) for comparison."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

n, k = (1 << 20), 1000
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
def cfg_edit(eps=None, p=0, d=0):
    def f(c):
        if eps is not None: c.velocity_epsilon = eps
        c.vel_pid.p_cascade, c.vel_pid.d_cascade = p, d
    return f
CASES = (("launch values, independent", None, True), ("hold below 2 cm/s", cfg_edit(0.02), False), ("hold + 1 P + 1 D stage", cfg_edit(0.02, 1, 1), False),
         ("hold + 1 P + 2 D stages", cfg_edit(0.02, 1, 2), False), ("hold code, band never crossed", cfg_edit(1e-12), False),
         ("1 P + 1 D stage, no hold", cfg_edit(None, 1, 1), False))
for nc in (8, 4):
    for name, edit, indep in CASES:
        cfg = cb.default_config(nc)
        if edit:
            edit(cfg)
        with cb.CdprBatch(cfg, n) as g:
            if indep:
                g.set_independent(True)
            g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
            ms = []
            for _ in range(3):
                g.step(k); ms.append(g.last_kernel_ms)
            t = float(np.mean(ms[1:]))
            print(f"NC={nc} {name:30s} {n * k / (t * 1e-3):.3e} instance-steps/s  ({t:6.1f} ms per 2^20 x 1000)  {g.kernel_detail}", flush=True)
