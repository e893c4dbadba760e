import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
ng = 1 << 18
amp, freq, phase, pose7, twist6 = wl.c3_instances(ng, 1)
for nc in (8, 4):
    gcfg = cb.default_config(nc)
    gcfg.velocity_epsilon = 0.02; gcfg.vel_pid.p_cascade = 1; gcfg.vel_pid.d_cascade = 1
    with cb.CdprBatch(gcfg, ng) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(3):
            g.step(200); ms.append(g.last_kernel_ms)
        print("general NC", nc, g.kernel_variant, ms, "rate", ng * 200 / (ms[-1] * 1e-3))
