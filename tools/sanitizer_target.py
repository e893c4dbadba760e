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

# round 2: independent robots (masked commands, per-instance modes), the leg model, plugin-style updates, both kinematics
# kernels, the flex kernel with 1 and 2 lanes per robot, the catch-all kernel
import torch
for nc in (4, 8):
    n = 77                                                  # ragged: the padded instances of the last block run too
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 3)
    rng = np.random.default_rng(1)
    for legs in (0, 1):
        cfg = cb.default_config(nc)
        cfg.leg_model = legs
        with cb.CdprBatch(cfg, n) as g:
            g.set_independent(True)
            g.set_platform_state(pose7, twist6)
            third = np.arange(n) % 3
            g.step(5)
            g.set_velocity_cmd(rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32), mask=third == 0)
            g.set_position_cmd(rng.uniform(-0.02, 0.02, (n, nc)).astype(np.float32), mask=third == 1)
            g.set_effort_cmd(rng.uniform(2, 6, (n, nc)), mask=third == 2)
            g.step(30); g.modes(); g.pid_terms()
            for k in range(12):
                g.update(rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32) if k % 5 == 0 else None)
            blob = g.get_state(); g.set_state(blob)
            buf = torch.zeros((3, 13, n), dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
            g.set_snapshots(10, buf.data_ptr(), 3); g.step(30); g.synchronize()
            print("ok independent", nc, "legs", legs, g.kernel_variant)
    for lanes in ("1", "2"):                                 # hold + filters with one and two lanes per robot
        os.environ["CDPR_FLEX_LANES"] = lanes
        cfg = cb.default_config(nc)
        cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = 1; cfg.vel_pid.d_cascade = 2
        with cb.CdprBatch(cfg, n) as g:
            g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
            g.step(1); g.step(300)
            g.rollout(7, 11, wl.c5_rollouts(11, 4, nc), 5, [0, 0, 0.3], 0.1, pose7[:7], twist6[:7])
            print("ok flex lanes", lanes, nc, g.kernel_detail)
    os.environ.pop("CDPR_FLEX_LANES", None)
    # k_step_flexr (0, 1, 2 stages; hold transitions through the on-chip and the HBM gap fit: commands that toggle the
    # hold band faster and slower than 11 steps; independent robots) and, with CDPR_FLEX_CLASSIC, k_step_flex on the same
    for classic in ("0", "1"):
        os.environ["CDPR_FLEX_CLASSIC"] = classic
        for pc, dc in ((0, 0), (1, 1), (1, 2)):
            cfg = cb.default_config(nc)
            cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = pc; cfg.vel_pid.d_cascade = dc
            with cb.CdprBatch(cfg, n) as g:
                g.set_independent(True)
                g.set_platform_state(pose7, twist6)
                mv = rng.uniform(0.03, 0.06, (n, nc)).astype(np.float32); st = rng.uniform(-0.01, 0.01, (n, nc)).astype(np.float32)
                st[:, 0] = mv[:, 0]
                for period in (14, 3, 5, 2, 13, 30):
                    for rep in range(4):
                        g.set_velocity_cmd(mv if rep % 2 == 0 else st, mask=rng.random(n) < 0.7)
                        g.step(period)
                g.set_sine_cmd(amp, freq, phase); g.step(120)
                blob = g.get_state(); g.set_state(blob); g.step(7); g.pid_terms(); g.update(None)
                print("ok hold toggling", nc, g.kernel_detail)
    os.environ.pop("CDPR_FLEX_CLASSIC", None)
    for eps in (-0.001, 0.02):                               # the square-wave publisher: fast kernel / hold in the dead band
        cfg = cb.default_config(nc)
        cfg.velocity_epsilon = eps; cfg.sine_publish_hz = 10.0
        with cb.CdprBatch(cfg, n) as g:
            g.set_platform_state(pose7, twist6); g.set_square_velocity_cmd(amp, freq * 10.0, phase)
            g.step(1); g.step(450)
            print("ok square publisher", nc, g.kernel_detail)
    cfg = cb.default_config(nc)
    cfg.vel_pid.cmd_limit = 0.0; cfg.vel_pid.d_buffer_length = 5; cfg.vel_pid.d_degree = 1
    with cb.CdprBatch(cfg, n) as g:                          # catch-all kernel
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase); g.step(60)
        g.update(None)
        print("ok", nc, g.kernel_variant)
    for npose in (4096, 4097):                               # pair kernel (even) and per-pose kernel (odd)
        p7, t6 = wl.c2_poses(npose, 0)
        st = np.ascontiguousarray(np.concatenate([p7[:, :3], p7[:, 6:7], p7[:, 3:6], t6], axis=1).T)
        d_in = torch.from_numpy(st).cuda(); d_out = torch.empty((nc, 8, npose), dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
        with cb.CdprBatch(cb.default_config(nc), 1) as g:
            g.ik_device(npose, d_in.data_ptr(), d_out.data_ptr()); g.synchronize()
    print("ok ik", nc)
