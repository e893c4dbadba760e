#!/usr/bin/env python
"""Minimal launcher for ncu: a few launches of the step kernel (or the IK kernel) on resident state.
usage: python tools/profile_target.py [--nc 8] [--instances 1048576] [--sim-steps 1000] [--launches 3] [--ik]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--nc", type=int, default=8)
ap.add_argument("--instances", type=int, default=1 << 20)
ap.add_argument("--sim-steps", type=int, default=1000)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--ik", action="store_true")
ap.add_argument("--cmd-limit", type=float, default=None, help="shrink the command clamp so that steps saturate")
ap.add_argument("--i-limit", type=float, default=None)
ap.add_argument("--sine-hz", type=float, default=None, help="publisher rate of the in-kernel sine generator")
ap.add_argument("--mode", default="sine", choices=["sine", "position", "velocity"], help="sine publisher | per-cable position targets | per-cable velocity targets")
ap.add_argument("--eps", type=float, default=None, help="velocityEpsilon (>= 0 enables hold: the flex kernel)")
ap.add_argument("--p-cascade", type=int, default=0)
ap.add_argument("--d-cascade", type=int, default=0)
ap.add_argument("--independent", action="store_true", help="per-instance modes: the flex kernel on the launch configuration")
a = ap.parse_args()
cfg = cb.default_config(a.nc)
if a.eps is not None: cfg.velocity_epsilon = a.eps
cfg.vel_pid.p_cascade, cfg.vel_pid.d_cascade = a.p_cascade, a.d_cascade
if a.sine_hz: cfg.sine_publish_hz = a.sine_hz
for pid in (cfg.vel_pid, cfg.pos_pid):
    if a.cmd_limit is not None: pid.cmd_limit = a.cmd_limit
    if a.i_limit is not None: pid.i_limit = a.i_limit
if a.ik:
    import torch
    if a.instances == 1 << 20: a.instances = 65536       # config 2 at its stated size unless asked otherwise
    pose7, twist6 = wl.c2_poses(a.instances, 0)
    st = np.concatenate([pose7[:, :3], pose7[:, 6:7], pose7[:, 3:6], twist6], axis=1).T.copy()
    d_in = torch.from_numpy(st).cuda(); d_out = torch.empty((a.nc, 8, a.instances), dtype=torch.float64, device="cuda")
    with cb.CdprBatch(cfg, 1) as g:
        for _ in range(a.launches):
            g.ik_device(a.instances, d_in.data_ptr(), d_out.data_ptr()); g.synchronize()
            print("ik ms", g.last_kernel_ms)
else:
    amp, freq, phase, pose7, twist6 = wl.c3_instances(a.instances, 1)
    with cb.CdprBatch(cfg, a.instances) as g:
        if a.independent: g.set_independent(True)
        print("variant", g.kernel_variant)
        g.set_platform_state(pose7, twist6)
        rng = np.random.default_rng(3)
        if a.mode == "sine": g.set_sine_cmd(amp, freq, phase)
        elif a.mode == "position": g.set_position_cmd(rng.uniform(-0.02, 0.02, (a.instances, a.nc)).astype(np.float32))
        else: g.set_velocity_cmd(rng.uniform(-0.05, 0.05, (a.instances, a.nc)).astype(np.float32))
        for _ in range(a.launches):
            g.step(a.sim_steps); g.synchronize()
            print("step ms", g.last_kernel_ms, "rate", a.instances * a.sim_steps / g.last_kernel_ms * 1e3)
