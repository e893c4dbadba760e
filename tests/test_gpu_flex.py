"""GPU parity of the on-chip full-semantics kernel ("flex", step_flex.cuh): independent robots (per-instance modes and
command latches, CdprGazeboPlugin.cpp:67-83,206-219), hold and biquad cascades, launch-split and checkpoint invariance."""
import numpy as np
import pytest

import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
from oracle import binding as ob
from helpers import to_oracle_config, state_rel_err
from test_gpu_parity import make_pair, general_cfg

pytestmark = pytest.mark.gpu


def _pair_independent(nc, n, seed, cfg_edit=None, sine=False):
    cfg = cb.default_config(nc)
    if cfg_edit:
        cfg_edit(cfg)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed)
    gpu = cb.CdprBatch(cfg, n)
    gpu.set_independent(True)
    assert gpu.kernel_variant == "flex"
    gpu.set_platform_state(pose7, twist6)
    if sine:
        gpu.set_sine_cmd(amp, freq, phase)
        orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    else:
        orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6)
    return cfg, gpu, orc


def _check(gpu, orc, tol, tag):
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    e = state_rel_err(pg, tg, po, to)
    assert e < tol, (tag, e)
    for a, b in zip(gpu.joint_states(), orc.joint_states()):
        assert np.max(np.abs(a - b)) < tol * max(1.0, np.max(np.abs(b))), tag


@pytest.mark.parametrize("nc", [4, 8])
def test_flex_on_the_launch_configuration_matches_the_oracle_per_step(built_lib, nc):
    """The reference's launch values (no hold, no filters) through the flex kernel: every one of the first 40 steps at 1e-9,
    then 1000 more."""
    cfg, gpu, orc = _pair_independent(nc, 200, seed=81, sine=True)
    for step in range(1, 41):
        gpu.step(1); orc.step(1)
        _check(gpu, orc, 1e-9, f"step {step}")
    gpu.step(1000); orc.step(1000)
    _check(gpu, orc, 1e-9, "after 1040")
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
def test_independent_robots_modes_and_commands(built_lib, nc):
    """Robots of one batch in different modes, commanded at different steps: thirds of the batch get velocity, position and
    effort commands; later a subset is re-commanded while the rest keeps running."""
    n = 150
    cfg, gpu, orc = _pair_independent(nc, n, seed=82)
    rng = np.random.default_rng(4)
    third = np.arange(n) % 3
    v = rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32)
    p = rng.uniform(-0.02, 0.02, (n, nc)).astype(np.float32)
    f = rng.uniform(2.0, 6.0, (n, nc))
    gpu.step(12); orc.step(12); _check(gpu, orc, 1e-9, "position hold after Load")
    gpu.set_velocity_cmd(v, mask=third == 0); orc.velocity_cmd_masked(v, third == 0)
    gpu.step(7); orc.step(7); _check(gpu, orc, 1e-9, "first third in velocity")
    gpu.set_position_cmd(p, mask=third == 1); orc.position_cmd_masked(p, third == 1)
    gpu.set_effort_cmd(f, mask=third == 2); orc.effort_cmd_masked(f, third == 2)
    gpu.step(30); orc.step(30); _check(gpu, orc, 1e-9, "three modes side by side")
    assert np.array_equal(gpu.modes(), orc.targets()[2][:, 0].astype(np.int32))
    assert set(np.unique(gpu.modes())) == {0, 1, 2}
    # both messages pending for some robots in the same update: velocity first, position wins (.cpp:206-219)
    both = (np.arange(n) % 5) == 0
    gpu.set_velocity_cmd(v[::-1].copy(), mask=both); gpu.set_position_cmd(p[::-1].copy(), mask=both)
    orc.velocity_cmd_masked(v[::-1].copy(), both); orc.position_cmd_masked(p[::-1].copy(), both)
    gpu.step(25); orc.step(25); _check(gpu, orc, 1e-9, "velocity + position pending")
    # effort robots back to velocity: the velocity Pid is reset on the mode change
    gpu.set_velocity_cmd(v, mask=third == 2); orc.velocity_cmd_masked(v, third == 2)
    for k in (1, 1, 11, 200):
        gpu.step(k); orc.step(k)
    _check(gpu, orc, 1e-8, "force -> velocity")
    assert np.array_equal(gpu.modes(), orc.targets()[2][:, 0].astype(np.int32))
    gpu.close()


def test_masked_commands_need_independent_robots(built_lib):
    cfg, gpu, _ = make_pair(4, 40, sine=False)
    assert gpu.kernel_variant == "fast"
    before = gpu.get_state()
    with pytest.raises(cb.CdprError) as e:
        gpu.set_velocity_cmd(np.zeros((40, 4), dtype=np.float32), mask=np.ones(40))
    assert e.value.code == cb.api.ERR_UNSUPPORTED
    assert np.array_equal(before, gpu.get_state())
    gpu.step(3)
    with pytest.raises(cb.CdprError):
        gpu.set_independent(True)          # only before the first step
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
def test_flex_launch_split_and_checkpoint_bitwise(built_lib, nc):
    """Hold / release cycles with filters: K steps in one launch == the same steps in uneven launches == a run resumed from a
    checkpoint taken in the middle of a hold transition, bit for bit."""
    _, a, _ = make_pair(nc, 180, seed=83, cfg_edit=general_cfg)
    _, b, _ = make_pair(nc, 180, seed=83, cfg_edit=general_cfg)
    _, c, _ = make_pair(nc, 180, seed=83, cfg_edit=general_cfg)
    assert a.kernel_variant == "flex"
    a.step(1237)
    for k in (1, 2, 10, 11, 13, 100, 500, 600):
        b.step(k)
    c.step(333)
    blob = c.get_state()
    c.step(50)
    c.set_state(blob)
    assert c.step_count == 333
    c.step(1237 - 333)
    pa, ta = a.platform_state()
    for other in (b, c):
        po, to = other.platform_state()
        assert np.array_equal(pa, po) and np.array_equal(ta, to)
        for x, y in zip(a.joint_states(), other.joint_states()):
            assert np.array_equal(x, y)
        assert np.array_equal(a.pid_terms(), other.pid_terms())
    a.close(); b.close(); c.close()


def test_flex_snapshots_and_rollouts(built_lib):
    """The flex kernel behind the other entry points: decimated snapshots equal stepwise states; rollout costs match the oracle."""
    import torch
    n, every, k = 100, 20, 120
    _, a, _ = make_pair(4, n, seed=84, cfg_edit=general_cfg)
    _, b, _ = make_pair(4, n, seed=84, cfg_edit=general_cfg)
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
    nc, n_robots, n_seq, n_cmd, spc = 4, 2, 16, 5, 10
    cfg = cb.default_config(nc)
    general_cfg(cfg)
    cmds = wl.c5_rollouts(n_seq, n_cmd, nc)
    _, _, _, pose7, twist6 = wl.c3_instances(n_robots, 11)
    target, lam = np.array([0.0, 0.0, 0.31]), 0.1
    with cb.CdprBatch(cfg, n_robots * n_seq) as g:
        assert g.kernel_variant == "flex"
        cost = g.rollout(n_robots, n_seq, cmds, spc, target, lam, pose7, twist6)
    ocost = np.zeros(n_robots * n_seq)
    o = ob.Batch(to_oracle_config(cfg), n_robots * n_seq, np.repeat(pose7, n_seq, axis=0), np.repeat(twist6, n_seq, axis=0))
    for cc in range(n_cmd):
        o.velocity_cmd(np.tile(cmds[:, cc, :], (n_robots, 1)))
        for _ in range(spc):
            o.step(1)
            pose, twist = o.platform_state()
            ocost += np.sum((pose[:, :3] - target) ** 2, axis=1) + lam * np.sum(twist[:, 3:] ** 2, axis=1)
    assert np.max(np.abs(cost - ocost) / ocost) < 1e-9


def _short_window(cfg):
    cfg.vel_pid.d_buffer_length = 5
    cfg.vel_pid.d_degree = 1


@pytest.mark.parametrize("variant,edit", [("fast", None), ("flex", general_cfg), ("general", _short_window)])
def test_update_publishes_like_the_plugin(built_lib, variant, edit):
    """cdpr_update = CdprGazeboPlugin::update in one call: messages in, one step, and the plugin's own pairing out -- joint
    position / velocity and platform pose / twist as read at the update (before the step integrates) with the effort applied
    in it (CdprGazeboPlugin.cpp:248-280) -- for all three kernel variants, against the oracle's record of the last update."""
    n, nc = 37, 4
    cfg, gpu, orc = make_pair(nc, n, seed=91, cfg_edit=edit, sine=False)
    assert gpu.kernel_variant == variant
    rng = np.random.default_rng(6)
    for step in range(1, 61):
        v = p = None
        if step % 10 == 1:
            v = rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32)
            orc.velocity_cmd(v)
        if step == 35:
            p = rng.uniform(-0.02, 0.02, (n, nc)).astype(np.float32)
            orc.position_cmd(p)
        pose_pre, twist_pre = orc.platform_state()
        orc.step(1)
        jpos, jvel, _, eff = orc.last_outputs()
        gpos, gvel, geff, gpose, gtwist = gpu.update(v, p)
        tol = 1e-9
        assert state_rel_err(gpose, gtwist, pose_pre, twist_pre) < tol, step
        for a, b in ((gpos, jpos), (gvel, jvel), (geff, eff)):
            assert np.max(np.abs(a - b)) < tol * max(1.0, np.max(np.abs(b))), step
    # and the state after 60 updates is what 60 plain steps give
    _check(gpu, orc, 1e-9, "after 60 updates")
    with pytest.raises(cb.CdprError) as e:
        gpu.update(np.zeros((n, nc + 1), dtype=np.float32))
    assert e.value.code == cb.api.ERR_BAD_LENGTH
    gpu.close()


def _legs(cfg):
    cfg.leg_model = 1


def _legs_hold_filters(cfg):
    cfg.leg_model = 1
    general_cfg(cfg)


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("edit", [_legs, _legs_hold_filters])
def test_leg_model_matches_the_oracle(built_lib, nc, edit):
    """SURVEY.md 8(f) N2 -- leg link masses / inertias and passive joint damping (cube.sdf:344-518) as a configuration-dependent
    6x6 mass matrix solved every step (csrc/legs.cuh): every one of the first 30 steps and 600 more against the CPU checker.
    The model itself is parity-unpinned (no Gazebo/ODE); this pins the CUDA path to its written definition."""
    cfg, gpu, orc = make_pair(nc, 120, seed=95, cfg_edit=edit)
    assert gpu.kernel_variant == "flex"
    for step in range(1, 31):
        gpu.step(1); orc.step(1)
        _check(gpu, orc, 1e-9, f"step {step}")
    gpu.step(600); orc.step(600)
    _check(gpu, orc, 1e-9 if edit is _legs else 1e-8, "after 630")
    # the legs matter: the same run without them ends somewhere else
    cfg0, g0, _ = make_pair(nc, 120, seed=95, cfg_edit=(general_cfg if edit is _legs_hold_filters else None))
    g0.step(630)
    d = np.max(np.abs(g0.platform_state()[0][:, :3] - gpu.platform_state()[0][:, :3]))
    assert d > 1e-7, d
    gpu.close(); g0.close()


def test_leg_model_launch_split_bitwise_and_unsupported_shapes(built_lib):
    _, a, _ = make_pair(4, 90, seed=96, cfg_edit=_legs)
    _, b, _ = make_pair(4, 90, seed=96, cfg_edit=_legs)
    a.step(257)
    for k in (1, 6, 50, 200):
        b.step(k)
    pa, ta = a.platform_state(); pb, tb = b.platform_state()
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
    a.close(); b.close()
    cfg = cb.default_config(4)
    cfg.leg_model = 1
    cfg.vel_pid.d_buffer_length = 5; cfg.vel_pid.d_degree = 1      # a shape only the HBM catch-all kernel runs
    with pytest.raises(cb.CdprError) as e:
        cb.CdprBatch(cfg, 8)
    assert e.value.code == cb.api.ERR_UNSUPPORTED


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_flex_random_command_sequences_against_the_oracle(built_lib, seed):
    """Fuzz of the rare paths: random configurations (hold band, cascades on either Pid, feed-forward, limits that bite),
    random sequences of masked velocity / position / effort commands landing at random steps -- hold entered and left from
    every mode, Pids reset while live or asleep, windows spanning several gaps, launches of random length -- against the
    oracle after every segment, and bitwise against the same sequence cut into different launches."""
    rng = np.random.default_rng(1000 + seed)
    nc = int(rng.choice([4, 8]))
    n = int(rng.integers(40, 90))

    def edit(cfg):
        cfg.velocity_epsilon = float(rng.choice([-0.001, 0.0, 0.01, 0.03]))
        cfg.vel_pid.p_cascade = int(rng.integers(0, 3)); cfg.vel_pid.d_cascade = int(rng.integers(0, 3))
        cfg.pos_pid.p_cascade = int(rng.integers(0, 2)); cfg.pos_pid.d_cascade = int(rng.integers(0, 2))
        cfg.vel_pid.forward_gain = float(rng.choice([0.0, 1.5]))
        if rng.random() < 0.5:
            for pid in (cfg.vel_pid, cfg.pos_pid):
                pid.i_limit, pid.cmd_limit = 0.5, 6.0
            cfg.effort_limit = 5.0
        if rng.random() < 0.3:
            cfg.vel_pid.d_degree = cfg.pos_pid.d_degree = int(rng.choice([1, 3]))
        cfg.leg_model = int(rng.random() < 0.25)

    cfg = cb.default_config(nc)
    edit(cfg)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 200 + seed)
    use_sine = bool(rng.random() < 0.5)
    ops = []
    for _ in range(14):
        kind = rng.choice(["vel", "pos", "eff", "step", "step"])
        mask = rng.random(n) < rng.choice([0.3, 0.7, 1.0])
        if kind == "vel":
            ops.append(("vel", (rng.uniform(-0.05, 0.05, (n, nc)) * (rng.random((n, nc)) < 0.8)).astype(np.float32), mask))
        elif kind == "pos":
            ops.append(("pos", rng.uniform(-0.02, 0.02, (n, nc)).astype(np.float32), mask))
        elif kind == "eff":
            ops.append(("eff", rng.uniform(2.0, 6.0, (n, nc)), mask))
        ops.append(("step", int(rng.choice([1, 2, 9, 10, 11, 12, 13, 37, 120])), None))

    def play(split):
        g = cb.CdprBatch(cfg, n)
        g.set_independent(True)
        assert g.kernel_variant == "flex"
        g.set_platform_state(pose7, twist6)
        if use_sine:
            g.set_sine_cmd(amp, freq, phase)
        o = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, *( (amp, freq, phase) if use_sine else ()))
        for kind, val, mask in ops:
            if kind == "vel":
                g.set_velocity_cmd(val, mask=mask); o.velocity_cmd_masked(val, mask)
            elif kind == "pos":
                g.set_position_cmd(val, mask=mask); o.position_cmd_masked(val, mask)
            elif kind == "eff":
                g.set_effort_cmd(val, mask=mask); o.effort_cmd_masked(val, mask)
            else:
                if split and val > 3:
                    g.step(1); g.step(val - 3); g.step(2)
                else:
                    g.step(val)
                o.step(val)
                _check(g, o, 2e-8, (seed, kind, val))
                assert np.array_equal(g.modes(), o.targets()[2][:, 0].astype(np.int32))
        out = (g.platform_state(), g.joint_states(), g.pid_terms(), g.get_state(), g.modes())
        g.close()
        return out

    a, b = play(False), play(True)
    assert np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1])
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x, y)
    # topic `pid`: the applied force always; the terms wherever a Pid PUBLISHED in the final step.  A Pid publishes from its
    # second update on (Pid.cpp:123-126), and the telemetry columns are written by the last step of a launch only, so for a cable
    # that primes in the final step (zero force) or is in Force mode they still show an earlier launch boundary.
    assert np.array_equal(a[2][..., 4], b[2][..., 4])
    published = (a[2][..., 4] != 0.0) & (a[4] != 0)[:, None]
    assert np.array_equal(a[2][published], b[2][published])
