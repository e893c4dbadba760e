#!/usr/bin/env python
"""SURVEY H2(c) -- two lanes per instance, four cables each -- measured through its lower bound.
A 2-lane step kernel at NC=8 does, per lane, exactly what the NC=4 kernel does per thread (4 cables of kinematics + force law,
one full platform update -- replicated in both lanes), PLUS the exchange of the six wrench components (12 SHFL + 6 DADD per
lane-step).  So the time of the NC=4 kernel on 2 x 2^20 threads is a lower bound of the 2-lane kernel's time on 2^20 8-cable
robots.  This script measures both sides on the same GPU; the flex kernel's LANES template (step_flex.cuh) is the direct
measurement of the same trade in a latency-bound kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

k = 1000
def run(nc, n):
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
    with cb.CdprBatch(cb.default_config(nc), n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(4):
            g.step(k); ms.append(g.last_kernel_ms)
        return float(np.mean(ms[1:]))
t8 = run(8, 1 << 20)
t4 = run(4, 1 << 21)
print(f"one lane per robot : k_step_fast<NC=8> on 2^20 robots x {k} steps: {t8:.2f} ms")
print(f"two lanes, lower bound: k_step_fast<NC=4> on 2^21 threads x {k} steps: {t4:.2f} ms  (no wrench exchange yet)")
print(f"=> a 2-lane kernel is at least {100 * (t4 / t8 - 1):.1f} % slower than the one-lane kernel it would replace")
