import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
n, k = 1 << 18, 200
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
for nc in (4, 8):
    cfg = cb.default_config(nc); cfg.leg_model = 1
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(3):
            g.step(k); ms.append(g.last_kernel_ms)
        t = float(np.mean(ms[1:]))
        print(f"legs NC={nc}: {n*k/(t*1e-3):.3e} instance-steps/s ({t:.1f} ms per 2^18 x 200) {g.kernel_detail}", flush=True)
