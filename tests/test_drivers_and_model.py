"""CPU-only: the restated command drivers against the constants in the reference's driver sources, and the
description -> config loader against the reference's cube.yaml / cube.sdf literals (all via tests/golden)."""
import json
import math
import os

import numpy as np

import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import drivers, model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_constants.json")))


def test_driver_defaults_match_reference_sources():
    d = G["drivers"]
    s = drivers.SineVelocity()
    assert (s.amp, s.freq, s.publish_hz) == (d["sinevelocitytest"]["cVelocityAmplitude"], d["sinevelocitytest"]["cVelocityFrequency"], d["sinevelocitytest"]["cPublishFrequency"])
    v = drivers.SquareVelocity()
    assert (v.amp, v.freq, v.publish_hz) == (d["squarevelocitytest"]["cVelocityAmplitude"], d["squarevelocitytest"]["cVelocityFrequency"], d["squarevelocitytest"]["cPublishFrequency"])
    p = drivers.SquarePosition()
    assert (p.amp, p.bias, p.freq, p.publish_hz) == (d["squarepositiontest"]["cPositionAmplitude"], d["squarepositiontest"]["cPositionBias"],
                                                     d["squarepositiontest"]["cPositionFrequency"], d["squarepositiontest"]["cPublishFrequency"])


def test_driver_waveforms():
    v = drivers.SquareVelocity()
    vals = [float(v.publish()) for _ in range(200)]                 # one 20 s period at 10 Hz
    assert set(np.round(vals, 6)) == {0.0, 0.06, -0.06}
    assert vals[0] == 0.0 and vals[25] == np.float32(0.06)          # dead band until |sin| >= sqrt(1/2): t = 2.5 s
    assert abs(v.time - 20.0) < 1e-9
    p = drivers.SquarePosition()
    pv = [float(p.publish()) for _ in range(100)]
    assert pv[0] == np.float32(0.05) and pv[60] == np.float32(-0.05)   # copysign(amp, sin(0)) = +amp, as in the source
    s = drivers.SineVelocity()
    sv = [s.publish() for _ in range(3)]
    assert sv[0] == 0.0 and sv[1] == np.float32(0.05 * math.sin(0.01 * 0.1 * 2 * math.pi)) and sv[1].dtype == np.float32


def test_config_from_reference_yaml_matches_sdf(built_lib):
    desc = G["cube_yaml"]
    stale = model.config_from_description(desc)
    assert list(stale.home_pos) == [0.0, 0.0, 2.0]                  # cube.yaml is stale ...
    cfg = model.config_from_description(desc, home_xyz=G["platform_pose"][:3])   # ... cube.sdf:310 is authoritative
    ref = cb.default_config(4)
    assert bytes(cfg) == bytes(ref)
    assert cfg.cable_damping == desc["joints"]["actuated"]["damping"] and cfg.effort_limit == desc["joints"]["actuated"]["effort"]


def test_launch_params_round_trip(built_lib):
    cfg = cb.default_config(4)
    cfg.vel_pid.p_gain = 1.0; cfg.pos_pid.i_gain = 2.0; cfg.velocity_epsilon = 9.0
    model.apply_launch_params(cfg, G["launch_params"])
    assert bytes(cfg) == bytes(cb.default_config(4))
    # the commented Ziegler-Nichols alternative (launch/cdpr_gazebo.launch:40-45)
    model.apply_launch_params(cfg, {"positionControllerP": 20.0, "positionControllerI": 100, "positionControllerD": 2.666666666})
    assert (cfg.pos_pid.p_gain, cfg.pos_pid.i_gain, cfg.pos_pid.d_gain) == (20.0, 100.0, 2.666666666)


def test_eight_cable_description(built_lib):
    desc = json.loads(json.dumps(G["cube_yaml"]))
    desc["points"] = desc["points"] + [{"frame": [p["frame"][0], p["frame"][1], 0.0], "platform": p["platform"]} for p in desc["points"]]
    cfg = model.config_from_description(desc, home_xyz=[0, 0, 0.3])
    assert bytes(cfg) == bytes(cb.default_config(8))                # SURVEY.md App. A.2 synthetic extension


def test_dterm_fir_weights_against_numpy_least_squares(built_lib):
    """Host logic of the step kernel's D-term: the FIR weights equal the derivative-at-the-end row of the least-squares
    polynomial fit (numpy), for every degree / window length; for degree <= 2 they are an exact quadratic in the sample
    position (the sliding-moment form); the reference's launch values give the classic 11-point end-point weights."""
    import cdpr_simulation_b200 as cb
    cfg = cb.default_config(4)
    dt = 0.001
    for deg, ln in [(1, 2), (1, 5), (2, 11), (2, 3), (3, 9), (4, 32), (2, 20), (0, 4)]:
        pid = cb.PidParams.from_buffer_copy(bytes(cfg.vel_pid))
        pid.d_degree, pid.d_buffer_length = deg, ln
        fir, quad, is_quad = cb.dterm_weights(pid, dt)
        t = (np.arange(ln) - (ln - 1)) * dt                       # window-relative times, newest = 0
        if deg == 0:
            assert np.all(fir == 0)
            continue
        V = np.vander(t / (dt * (ln - 1)), deg + 1, increasing=True)
        row = np.linalg.pinv(V)[1] / (dt * (ln - 1))              # d/dt of the fitted polynomial at t = 0
        assert np.max(np.abs(fir - row)) < 1e-9 * np.max(np.abs(row)), (deg, ln)
        assert abs(fir.sum()) < 1e-9 * np.abs(fir).sum()          # a constant signal has zero derivative
        assert abs(fir @ t - 1.0) < 1e-10                         # a unit ramp has derivative 1
        assert is_quad == (deg <= 2)
        if is_quad:
            p = np.arange(1, ln + 1)
            assert np.max(np.abs(quad[0] + quad[1] * p + quad[2] * p * p - fir)) < 1e-12 * np.max(np.abs(fir))
    pid = cb.PidParams.from_buffer_copy(bytes(cfg.vel_pid))
    fir, _, _ = cb.dterm_weights(pid, dt)
    # 11-point quadratic fit, derivative at the last point: closed form (3 j^2 ... ) / h -- check two known entries
    ref = np.linalg.pinv(np.vander(np.arange(-10.0, 1.0), 3, increasing=True))[1] / dt
    assert np.allclose(fir, ref, rtol=1e-10)


def test_sliding_window_recursion_of_the_step_kernel(built_lib):
    """The step kernel carries (S0, S1, Kd*D) per cable instead of re-reading the 11-sample window (step_fast.cuh):
    the recursion below is its update, statement by statement.  Algebra: exact in rational terms; in doubles it must
    stay within ~1e-12 of the FIR over the 64 steps between two re-summations, and a re-summation must land on the
    FIR again.  (Drift grows like n^2.5, which is why the kernel re-sums every 64 steps and not every 1000.)"""
    cfg = cb.default_config(4)
    dt, ln, kd = cfg.dt, 11, 1.0
    pid = cb.PidParams.from_buffer_copy(bytes(cfg.vel_pid))
    fir, (a, b, c), is_quad = cb.dterm_weights(pid, dt)
    assert is_quad and len(fir) == ln
    dk = (kd * (a + ln * b + ln * ln * c), kd * (c - b), -2.0 * kd * c, -kd * a)   # y_new, S0, S1, y_old (api.cu fill_args)
    rng = np.random.default_rng(11)
    t = np.arange(400) * dt
    y = 0.05 * np.sin(2 * np.pi * 0.7 * t + 0.3) + 1e-3 * rng.standard_normal(t.size)   # velocity-error-like signal
    ring = list(y[:ln])
    p = np.arange(1, ln + 1, dtype=float)
    def resum(w):
        w = np.asarray(w)
        s0 = 0.0; s1 = 0.0; s2 = 0.0
        for j in range(ln):                      # same order as resync_moments
            s0 += w[j]; s1 = p[j] * w[j] + s1; s2 = p[j] * p[j] * w[j] + s2
        return s0, s1, kd * (a * s0 + (b * s1 + c * s2))
    s0, s1, kdd = resum(ring)
    assert abs(kdd - kd * float(fir @ np.asarray(ring))) < 1e-12 * np.abs(fir).sum() * 0.05
    worst = 0.0
    for n in range(ln, t.size):
        e, y_old = y[n], ring.pop(0)
        ring.append(e)
        kdd = dk[3] * y_old + (dk[2] * s1 + (dk[1] * s0 + (dk[0] * e + kdd)))     # old S0, S1 on the right
        s1 = ln * e + (s1 - s0)
        s0 = (s0 + e) - y_old
        exact = kd * float(fir @ np.asarray(ring))
        worst = max(worst, abs(kdd - exact))
        if (n - ln + 1) % 64 == 0:               # kResync
            s0, s1, kdd = resum(ring)
            assert abs(kdd - exact) < 1e-12
    assert worst < 2e-11, worst                  # D itself is O(0.1 .. 1) here


def test_oracle_square_publisher_follows_the_driver_source():
    """The checker's in-loop square publisher (what the kernels' CDPR_OPT_PUBLISHER_SHAPE = 1 is compared with) against the
    restated squarevelocitytest loop (P/src/squarevelocitytest.cpp:19-33): same float32 value at every one of 250 publishes,
    both plateaus and the dead band visited."""
    from oracle import binding as ob
    from cdpr_simulation_b200 import drivers
    b = ob.Batch(ob.default_config(4), 1, amp=[0.06], freq=[0.05], phase=[0.0])
    b.publisher(1, 10.0)
    d = drivers.SquareVelocity()
    seen = set()
    for _ in range(250):
        v = d.publish()
        b.step(1)
        vt = b.targets()[0][0, 0]
        assert np.float32(vt) == v
        seen.add(float(np.sign(vt)))
        b.step(99)
    assert seen == {-1.0, 0.0, 1.0}
