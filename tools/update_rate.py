#!/usr/bin/env python
"""Plugin-style operation (one cdpr_update per physics step, N = 1 robot, host buffers, synchronous) for the kernel variants:
updates per second with read-back."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
for name, edit in (("launch values (fast kernel)", None), ("hold below 2 cm/s", lambda c: setattr(c, "velocity_epsilon", 0.02)),
                   ("hold + 1 P + 1 D stage", lambda c: (setattr(c, "velocity_epsilon", 0.02), setattr(c.vel_pid, "p_cascade", 1), setattr(c.vel_pid, "d_cascade", 1)))):
    for nc in (4, 8):
        cfg = cb.default_config(nc)
        if edit: edit(cfg)
        with cb.CdprBatch(cfg, 1) as g:
            axes = np.full((1, nc), 0.03, dtype=np.float32)
            for k in range(200): g.update(axes if k % 10 == 0 else None)
            reps = 3000
            t0 = time.perf_counter()
            for k in range(reps): g.update(axes if k % 10 == 0 else None)
            dt = (time.perf_counter() - t0) / reps
            print(f"NC={nc} {name:28s} {1.0 / dt:9.0f} updates/s ({dt * 1e6:5.1f} us)  {g.kernel_detail}", flush=True)
