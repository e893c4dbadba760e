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

# ---- round 2 additions ---------------------------------------------------------------------------------------------
def general_cfg(cfg):
    cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = 1; cfg.vel_pid.d_cascade = 2

# (a) 20,000 steps at NC=8: sliding-moment form (production) vs plain FIR form of the D-term, both against the oracle
n, nc, k = 256, 8, 20000
cfg = cb.default_config(nc)
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
o = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
marks = (1000, 5000, 10000, 20000)
ostates = []
done = 0
for m in marks:
    o.step(m - done); done = m
    ostates.append(o.platform_state())
runs = {}
for form in ("moments", "fir"):
    g = cb.CdprBatch(cfg, n)
    if form == "fir": g.set_option(cb.api.OPT_DTERM_FIR, 1)
    g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
    done, errs, states = 0, [], []
    for m, os_ in zip(marks, ostates):
        g.step(m - done); done = m
        st = g.platform_state(); states.append(st)
        errs.append(state_rel_err(*st, *os_))
    runs[form] = states
    print(f"NC=8 long run, D-term as {form:8s}: state error vs oracle after {marks} steps: " + ", ".join(f"{e:.2e}" for e in errs))
    g.close()
print("NC=8 long run, moments vs FIR (GPU vs GPU): " + ", ".join(f"{state_rel_err(*a, *b):.2e}" for a, b in zip(runs["moments"], runs["fir"])))

# (b) d_gain = 0: GPU vs the reference's own compiled force law (no D-term noise left)
if ob.ref_available():
    for nc in (4, 8):
        cfg = cb.default_config(nc); cfg.vel_pid.d_gain = 0.0; cfg.pos_pid.d_gain = 0.0
        a_, f_, p_, po_, tw_ = wl.c3_instances(64, 23)
        g = cb.CdprBatch(cfg, 64); g.set_platform_state(po_, tw_); g.set_sine_cmd(a_, f_, p_); g.step(1500)
        o = ob.Batch(to_oracle_config(cfg), 64, po_, tw_, a_, f_, p_); o.step_reference_forcelaw(1500)
        print(f"NC={nc}, d_gain = 0: GPU vs the reference's compiled Pid.cpp/JointForceCalculator.cpp after 1500 steps: {state_rel_err(*g.platform_state(), *o.platform_state()):.2e}")
        g.close()

# (c) full-semantics kernels (hold + filters) and leg model
def hold_1p1d(cfg):
    cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = 1; cfg.vel_pid.d_cascade = 1
def hold_only(cfg):
    cfg.velocity_epsilon = 0.02
def general_cfg_classic(cfg):
    general_cfg(cfg); os.environ["CDPR_FLEX_CLASSIC"] = "1"
for nc, edit, what in [(nc, e, w) for nc in (4, 8) for e, w in ((hold_only, "hold 2 cm/s"), (hold_1p1d, "hold 2 cm/s, 1 P + 1 D biquad stage"),
                                                              (general_cfg, "hold 2 cm/s, 1 P + 2 D biquad stages"), (general_cfg_classic, "hold 2 cm/s, 1 P + 2 D biquad stages"))]:
    cfg = cb.default_config(nc); edit(cfg)
    a_, f_, p_, po_, tw_ = wl.c3_instances(256, 61)
    g = cb.CdprBatch(cfg, 256); g.set_platform_state(po_, tw_); g.set_sine_cmd(a_, f_, p_)
    o = ob.Batch(to_oracle_config(cfg), 256, po_, tw_, a_, f_, p_)
    worst = 0.0
    for s in range(40):
        g.step(1); o.step(1); worst = max(worst, state_rel_err(*g.platform_state(), *o.platform_state()))
    os.environ.pop("CDPR_FLEX_CLASSIC", None)
    line = f"NC={nc} flex ({what}): first 40 steps {worst:.2e}"
    done = 40
    for m in (1000, 3000):
        g.step(m - done); o.step(m - done); done = m
        line += f"; after {m}: {state_rel_err(*g.platform_state(), *o.platform_state()):.2e}"
    print(line + f"  [{g.kernel_detail}]")
    g.close()
for nc in (4, 8):
    cfg = cb.default_config(nc); cfg.leg_model = 1
    a_, f_, p_, po_, tw_ = wl.c3_instances(256, 95)
    g = cb.CdprBatch(cfg, 256); g.set_platform_state(po_, tw_); g.set_sine_cmd(a_, f_, p_)
    o = ob.Batch(to_oracle_config(cfg), 256, po_, tw_, a_, f_, p_)
    g.step(1000); o.step(1000)
    e = state_rel_err(*g.platform_state(), *o.platform_state())
    g0 = cb.CdprBatch(cb.default_config(nc), 256); g0.set_platform_state(po_, tw_); g0.set_sine_cmd(a_, f_, p_); g0.step(1000)
    dp = np.max(np.abs(g0.platform_state()[0][:, :3] - g.platform_state()[0][:, :3]))
    dv = np.max(np.abs(g0.platform_state()[1][:, :3] - g.platform_state()[1][:, :3]))
    print(f"NC={nc} leg model (N2): GPU vs oracle after 1000 steps {e:.2e}; effect of the legs on the trajectory (vs reduced model, 1 s): "
          f"max |dp| = {dp:.2e} m, max |dv| = {dv:.2e} m/s")
    g.close(); g0.close()
