#!/usr/bin/env python
"""Numbers behind the parity claims (run on a B200): CUDA path vs CPU oracle on the same seeded inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
from oracle import binding as ob
from helpers import to_oracle_config, state_rel_err

for nc in (4, 8):
    n = 512
    cfg = cb.default_config(nc)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
    g = cb.CdprBatch(cfg, n); g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
    o = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    worst = 0.0
    for s in range(40):
        g.step(1); o.step(1)
        worst = max(worst, state_rel_err(*g.platform_state(), *o.platform_state()))
    print(f"NC={nc}: max relative state error over each of the first 40 steps (1 step per launch): {worst:.2e}")
    done = 40
    for k in (1000, 5000, 20000):
        g.step(k - done); o.step(k - done); done = k
        print(f"NC={nc}: relative state error after {k} steps: {state_rel_err(*g.platform_state(), *o.platform_state()):.2e}")
    if ob.ref_available():
        g2 = cb.CdprBatch(cfg, 64); g2.set_platform_state(pose7[:64], twist6[:64]); g2.set_sine_cmd(amp[:64], freq[:64], phase[:64])
        o2 = ob.Batch(to_oracle_config(cfg), 64, pose7[:64], twist6[:64], amp[:64], freq[:64], phase[:64])
        g2.step(1000); o2.step_reference_forcelaw(1000)
        print(f"NC={nc}: vs the reference's own force law (oracle L0) after 1000 steps: {state_rel_err(*g2.platform_state(), *o2.platform_state()):.2e}"
              "  (the reference's absolute-time D-term noise)")
        g2.close()
    g.close()
