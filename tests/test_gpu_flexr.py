"""GPU parity of the rebuilt full-semantics kernel (k_step_flexr, step_flexr.cuh): hold through the
position Pid (JointForceCalculator.cpp:72-82) with at most one biquad stage per filter (Pid.cpp:27-44), its event-driven hot
loop, the on-chip fits of windows that span a gap (Pid.cpp:193-247) and their HBM fallback -- against the oracle per step, and
bit for bit across launch splits and checkpoints.  test_gpu_flex.py / test_gpu_parity*.py cover k_step_flex (two or more
stages) with the same scenarios."""
import os

import numpy as np
import pytest

import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
from oracle import binding as ob
from helpers import to_oracle_config, state_rel_err
from test_gpu_parity import make_pair
from test_gpu_parity_r2 import _ran_pid

pytestmark = pytest.mark.gpu


def hold_cfg(cfg):
    cfg.velocity_epsilon = 0.02       # hold below 2 cm/s


def hold_1p1d_cfg(cfg):
    cfg.velocity_epsilon = 0.02
    cfg.vel_pid.p_cascade = 1         # the reference's filter constants (launch:27-32) with one stage switched on
    cfg.vel_pid.d_cascade = 1


def stable_filters_cfg(cfg):
    cfg.velocity_epsilon = 0.02
    cfg.vel_pid.p_cascade = 1; cfg.vel_pid.p_cutoff = 0.3   # a loop that does not live on its clamps
    cfg.pos_pid.p_cascade = 1; cfg.pos_pid.p_cutoff = 0.3   # the same coefficients on both Pids
    cfg.vel_pid.d_cascade = 1; cfg.vel_pid.d_cutoff = 0.25


def _check(gpu, orc, tol, tag):
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    e = state_rel_err(pg, tg, po, to)
    assert e < tol, (tag, e)
    for a, b in zip(gpu.joint_states(), orc.joint_states()):
        assert np.max(np.abs(a - b)) < tol * max(1.0, np.max(np.abs(b))), tag
    return e


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("edit", [hold_cfg, hold_1p1d_cfg, stable_filters_cfg])
def test_flexr_per_step_and_long_run(built_lib, nc, edit):
    """Sine commands that cross the hold band: each of the first 60 steps at 1e-9 (joint states, platform state, the `pid`
    topic), then 2000 more."""
    cfg, gpu, orc = make_pair(nc, 200, seed=91, cfg_edit=edit)
    assert gpu.kernel_detail.startswith("flexr:"), gpu.kernel_detail
    held = 0
    k_prev = None
    for step in range(1, 61):
        gpu.step(1); orc.step(1)
        _check(gpu, orc, 1e-9, f"step {step}")
        k, _ = _ran_pid(orc, cfg)
        held += int(np.sum(k == 1))
        if step >= 3:
            # topic `pid` (Pid.cpp:140-141,167): a Pid publishes from its second update on, so the last message of a cable
            # that changed Pid in this step is still the other Pid's -- compare where the same Pid ran twice in a row
            same = (k == k_prev)
            terms, ot = gpu.pid_terms(), orc.pid_terms()
            idx = np.indices(k.shape)
            sel = ot[idx[0], idx[1], k]
            for j, name in enumerate(("pTerm", "iTerm", "dTerm")):
                scale = max(1.0, float(np.max(np.abs(sel[..., j]))))
                assert np.max(np.abs(terms[..., j] - sel[..., j])[same]) < 1e-9 * scale, (step, name)
            eff = orc.last_outputs()[3]
            assert np.max(np.abs(terms[..., 4] - eff)) < 1e-9 * max(1.0, float(np.max(np.abs(eff)))), (step, "applied force")
        k_prev = k
    assert held > 0, "the test is meant to exercise hold"
    for k in (140, 800, 1000):
        gpu.step(k); orc.step(k)
        # over thousands of steps 1-ulp differences of the gap-spanning fits grow with the closed loop (see test_gpu_parity_r2)
        _check(gpu, orc, 2e-7, f"after {k} more")
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("edit", [hold_cfg, hold_1p1d_cfg])
def test_flexr_launch_split_and_checkpoint_bitwise(built_lib, nc, edit):
    """K steps in one launch == the same steps in uneven launches (which cut hot runs, gap windows and hold transitions at
    arbitrary places) == a run resumed from a checkpoint, bit for bit."""
    runs = [make_pair(nc, 180, seed=92, cfg_edit=edit)[1] for _ in range(3)]
    a, b, c = runs
    assert a.kernel_detail.startswith("flexr:")
    a.step(1511)
    splits = (1, 2, 9, 10, 11, 12, 13, 1, 100, 3, 500, 7, 600, 242)
    assert sum(splits) == 1511
    for k in splits:
        b.step(k)
    c.step(337)
    blob = c.get_state()
    c.step(41)
    c.set_state(blob)
    assert c.step_count == 337
    c.step(1511 - 337)
    # (the state blobs themselves may differ: the telemetry columns of a Pid are written by the last step of a launch only)
    for more in (0, 1, 96):  # ... and again after more steps: state that is not observable now would show later
        for g in runs:
            if more:
                g.step(more)
        pa, ta = a.platform_state()
        for other in (b, c):
            po, to = other.platform_state()
            assert np.array_equal(pa, po) and np.array_equal(ta, to)
            for x, y in zip(a.joint_states(), other.joint_states()):
                assert np.array_equal(x, y)
            assert np.array_equal(a.pid_terms(), other.pid_terms())
    for g in runs:
        g.close()


@pytest.mark.parametrize("nc", [4, 8])
def test_flexr_rapid_hold_toggling_uses_both_gap_fits(built_lib, nc):
    """Velocity commands that enter and leave the hold band every few steps: a Pid goes back to sleep before it has 11 fresh
    samples, so its window is stale twice over and the fit runs from the HBM ring (time stamps loaded); slower toggling
    afterwards takes the on-chip fit (stamps from one value).  Every step against the oracle, per-robot commands."""
    n = 96
    cfg = cb.default_config(nc)
    hold_1p1d_cfg(cfg)
    _, _, _, pose7, twist6 = wl.c3_instances(n, 93)
    gpu = cb.CdprBatch(cfg, n)
    gpu.set_independent(True)
    assert gpu.kernel_detail.startswith("flexr:")
    gpu.set_platform_state(pose7, twist6)
    orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6)
    rng = np.random.default_rng(5)
    moving = rng.uniform(0.03, 0.06, (n, nc)).astype(np.float32) * rng.choice([-1.0, 1.0], (n, nc)).astype(np.float32)
    still = rng.uniform(-0.01, 0.01, (n, nc)).astype(np.float32)
    still[:, 0] = moving[:, 0]  # cable 0 never holds: the cables of a robot do not all switch together
    step = 0
    for period in (14, 3, 5, 2, 7, 12, 30, 25):
        for rep in range(4):
            cmd = moving if rep % 2 == 0 else still
            mask = rng.random(n) < 0.7  # not every robot gets every command
            gpu.set_velocity_cmd(cmd, mask=mask); orc.velocity_cmd_masked(cmd, mask)
            for _ in range(period):
                gpu.step(1); orc.step(1); step += 1
                _check(gpu, orc, 1e-9, f"step {step} (period {period})")
    gpu.close()


def test_flexr_different_filter_coefficients_fall_back_to_the_classic_kernel(built_lib):
    """One coefficient set per filter is a precondition of k_step_flexr; two Pids with different stages run k_step_flex."""
    def edit(cfg):
        hold_1p1d_cfg(cfg)
        cfg.pos_pid.p_cascade = 1; cfg.pos_pid.p_cutoff = 0.2
    cfg, gpu, orc = make_pair(4, 64, seed=94, cfg_edit=edit)
    assert gpu.kernel_detail.startswith("flex:"), gpu.kernel_detail
    for k in (1, 1, 12, 50, 300):
        gpu.step(k); orc.step(k)
        _check(gpu, orc, 1e-9 if k < 100 else 1e-7, f"after {k}")
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
def test_classic_flex_kernel_on_the_same_configuration(built_lib, nc):
    """CDPR_FLEX_CLASSIC=1 keeps k_step_flex for a configuration k_step_flexr would take (A/B runs): both match the oracle, and
    they agree with each other at rounding level."""
    os.environ["CDPR_FLEX_CLASSIC"] = "1"
    try:
        cfg, classic, orc = make_pair(nc, 128, seed=95, cfg_edit=hold_1p1d_cfg)
    finally:
        del os.environ["CDPR_FLEX_CLASSIC"]
    _, regs, _ = make_pair(nc, 128, seed=95, cfg_edit=hold_1p1d_cfg)
    assert classic.kernel_detail.startswith("flex:") and regs.kernel_detail.startswith("flexr:")
    for k in (1, 5, 20, 74, 400):
        classic.step(k); regs.step(k); orc.step(k)
        _check(classic, orc, 1e-8, f"classic after {k}")
        _check(regs, orc, 1e-8, f"registers after {k}")
    classic.close(); regs.close()


def test_flexr_snapshots_rollouts_and_update(built_lib):
    """The other entry points on top of k_step_flexr: decimated snapshots equal stepwise states (the hot runs stop at a
    snapshot), rollout costs match the oracle, cdpr_update publishes like the plugin."""
    import torch
    n, every, k = 100, 20, 120
    _, a, _ = make_pair(8, n, seed=96, cfg_edit=hold_cfg)
    _, b, _ = make_pair(8, n, seed=96, cfg_edit=hold_cfg)
    buf = torch.zeros((k // every, 13, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    a.set_snapshots(every, buf.data_ptr(), buf.shape[0])
    a.step(k); a.synchronize()
    snaps = buf.cpu().numpy()
    for s in range(k // every):
        b.step(every)
        pose, twist = b.platform_state()
        assert np.array_equal(snaps[s, 0:3].T, pose[:, 0:3]) and np.array_equal(snaps[s, 7:13].T, twist)
    a.close(); b.close()
    nc, n_robots, n_seq, n_cmd, spc = 8, 2, 16, 5, 10
    cfg = cb.default_config(nc)
    hold_1p1d_cfg(cfg)
    cmds = wl.c5_rollouts(n_seq, n_cmd, nc)
    _, _, _, pose7, twist6 = wl.c3_instances(n_robots, 12)
    target, lam = np.array([0.0, 0.0, 0.32]), 0.05
    with cb.CdprBatch(cfg, n_robots * n_seq) as g:
        assert g.kernel_detail.startswith("flexr:")
        cost = g.rollout(n_robots, n_seq, cmds, spc, target, lam, pose7, twist6)
    ocost = np.zeros(n_robots * n_seq)
    o = ob.Batch(to_oracle_config(cfg), n_robots * n_seq, np.repeat(pose7, n_seq, axis=0), np.repeat(twist6, n_seq, axis=0))
    for c in range(n_cmd):
        o.velocity_cmd(np.tile(cmds[:, c, :], (n_robots, 1)))
        for _ in range(spc):
            o.step(1)
            pose, twist = o.platform_state()
            ocost += np.sum((pose[:, :3] - target) ** 2, axis=1) + lam * np.sum(twist[:, 3:] ** 2, axis=1)
    assert np.max(np.abs(cost - ocost) / ocost) < 1e-9


def test_flexr_full_size_properties(built_lib):
    """2^20 eight-cable robots with hold + one stage per filter, at a size the oracle cannot follow: finite state, unit
    quaternions, the platform inside the frame, robots that hold and robots that do not; the first 2048 instances are bit for bit
    what a 2048-instance batch computes (no cross-talk between robots, blocks or lanes), and cutting the launch changes nothing."""
    n, k = 1 << 20, 400
    cfg = cb.default_config(8)
    hold_1p1d_cfg(cfg)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
    with cb.CdprBatch(cfg, n) as g:
        assert g.kernel_detail.startswith("flexr:")
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        g.step(k)
        pose, twist = g.platform_state()
        terms = g.pid_terms()
    assert np.all(np.isfinite(pose)) and np.all(np.isfinite(twist))
    assert np.max(np.abs(np.linalg.norm(pose[:, 3:], axis=1) - 1.0)) < 1e-14
    assert np.all(np.abs(pose[:, :2]) < 0.3) and np.all((pose[:, 2] > 0.0) & (pose[:, 2] < 0.6))
    with cb.CdprBatch(cfg, 2048) as a, cb.CdprBatch(cfg, 2048) as b:
        for h in (a, b):
            h.set_platform_state(pose7[:2048], twist6[:2048]); h.set_sine_cmd(amp[:2048], freq[:2048], phase[:2048])
        a.step(k)
        for part in (1, 13, 200, k - 214):
            b.step(part)
        pa, ta = a.platform_state(); pb, tb = b.platform_state()
        assert np.array_equal(pa, pose[:2048]) and np.array_equal(ta, twist[:2048])
        assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
        assert np.array_equal(a.pid_terms()[..., 4], terms[:2048, :, 4])


def _square_pair(nc, n, seed, eps):
    """squarevelocitytest.cpp inside the kernel and inside the oracle: 10 Hz publisher, per-instance amplitude / frequency / phase
    (frequencies far above the driver's 0.05 Hz so that a short run crosses the dead band many times)."""
    cfg = cb.default_config(nc)
    cfg.velocity_epsilon = eps
    cfg.sine_publish_hz = 10.0            # squarevelocitytest.cpp:6
    rng = np.random.default_rng(seed)
    amp, freq, phase = rng.uniform(0.03, 0.06, n), rng.uniform(0.3, 1.5, n), rng.uniform(0.0, 2 * np.pi, n)
    _, _, _, pose7, twist6 = wl.c3_instances(n, seed)
    gpu = cb.CdprBatch(cfg, n)
    gpu.set_platform_state(pose7, twist6)
    gpu.set_square_velocity_cmd(amp, freq, phase)
    orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    orc.publisher(1, cfg.sine_publish_hz)
    return cfg, gpu, orc


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("eps", [-0.001, 0.02])
def test_square_velocity_publisher_in_kernel(built_lib, nc, eps):
    """The reference's second command driver (P/src/squarevelocitytest.cpp:19-33) run by the kernels' own publisher: +-amp
    outside the dead band, 0 inside -- where, with velocityEpsilon >= 0, the cables hold through the position Pid
    (JointForceCalculator.cpp:72-82).  Launch values: the fast kernel; with the hold band: k_step_flexr."""
    cfg, gpu, orc = _square_pair(nc, 160, 97, eps)
    assert gpu.kernel_variant == ("fast" if eps < 0 else "flex")
    seen = set()
    for k in [1] * 5 + [95, 1, 99, 1, 1, 98, 100, 100, 300, 700]:
        gpu.step(k); orc.step(k)
        _check(gpu, orc, 1e-9 if gpu.step_count <= 500 else 1e-7, f"after {gpu.step_count}")
        seen.update(np.unique(np.sign(orc.targets()[0][:, 0])).tolist())
    assert seen == {-1.0, 0.0, 1.0}, "the run is meant to visit both plateaus and the dead band"
    if eps >= 0:
        k, _ = _ran_pid(orc, cfg)
        assert np.any(k == 1) and np.any(k == 0), "some cables hold, some move"
    gpu.close()


@pytest.mark.parametrize("eps", [-0.001, 0.02])
def test_square_velocity_publisher_launch_split_bitwise(built_lib, eps):
    _, a, _ = _square_pair(8, 128, 98, eps)
    _, b, _ = _square_pair(8, 128, 98, eps)
    a.step(1203)
    for k in (1, 99, 100, 3, 250, 7, 743):
        b.step(k)
    pa, ta = a.platform_state(); pb, tb = b.platform_state()
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
    for x, y in zip(a.joint_states(), b.joint_states()):
        assert np.array_equal(x, y)
    a.close(); b.close()
