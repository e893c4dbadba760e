#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitizer_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

for nc in (4, 8):
    n = 333
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
    for general in (False, True):
        cfg = cb.default_config(nc)
        if general:
            cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = 1; cfg.vel_pid.d_cascade = 1
        with cb.CdprBatch(cfg, n) as g:
            g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
            g.step(1); g.step(37)
            g.set_position_cmd(np.zeros((n, nc), dtype=np.float32)); g.step(20)
            g.set_effort_cmd(np.full((n, nc), 4.0)); g.step(5)
            g.set_velocity_cmd(np.full((n, nc), 0.01, dtype=np.float32)); g.step(30)
            g.platform_state(); g.joint_states(); g.pid_state()
            blob = g.get_state(); g.set_state(blob); g.reset(); g.step(3)
            g.ik(*wl.c2_poses(100, 0))
            if not general:
                g.rollout(3, 111, wl.c5_rollouts(111, 4, nc), 5, [0, 0, 0.3], 0.1, pose7[:3], twist6[:3])
            print("ok", nc, g.kernel_variant)

# the saturation paths of the fast kernel: optimistic body -> out-of-line exact pass -> inline-clamping body and back
for nc in (4, 8):
    n = 192
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 2)
    cfg = cb.default_config(nc)
    for pid in (cfg.vel_pid, cfg.pos_pid):
        pid.i_limit, pid.cmd_limit = 0.4, 5.0
    cfg.effort_limit = 4.0
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)   # uniform-target layout
        g.step(150)
        v = np.random.default_rng(0).uniform(-0.3, 0.3, (n, nc)).astype(np.float32)
        g.set_velocity_cmd(v); g.step(150)                                         # per-cable-target layout
        g.rollout(3, 64, wl.c5_rollouts(64, 4, nc) * 10.0, 5, [0, 0, 0.3], 0.1, pose7[:3], twist6[:3])
        g.platform_state()
        print("ok saturating", nc, g.kernel_variant)
