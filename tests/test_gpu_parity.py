"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp64 per-step state <= 1e-9 relative; bounded divergence over
1000 steps; bit-exact instance indexing and cable ordering."""
import numpy as np
import pytest

import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
from oracle import binding as ob
from helpers import to_oracle_config, state_rel_err

pytestmark = pytest.mark.gpu

PER_STEP_TOL = 1e-9
DIVERGENCE_TOL_1000 = 1e-9   # measured 6e-12 (NC=4), 2e-11 (NC=8)


def make_pair(nc, n, seed=1, cfg_edit=None, sine=True):
    cfg = cb.default_config(nc)
    if cfg_edit:
        cfg_edit(cfg)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed)
    gpu = cb.CdprBatch(cfg, n)
    gpu.set_platform_state(pose7, twist6)
    if sine:
        gpu.set_sine_cmd(amp, freq, phase)
        orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    else:
        orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6)
    return cfg, gpu, orc


@pytest.mark.parametrize("nc", [4, 8])
def test_ik_matches_oracle(built_lib, nc):
    cfg = cb.default_config(nc)
    pose7, twist6 = wl.c2_poses(4099, seed=0)  # ragged: not a multiple of the block size
    with cb.CdprBatch(cfg, 1) as gpu:
        ln, lr, w = gpu.ik(pose7, twist6)
    oln, olr, ow = ob.ik(to_oracle_config(cfg), pose7, twist6)
    assert np.max(np.abs(ln - oln) / oln) < 1e-14
    assert np.max(np.abs(lr - olr)) < 1e-14
    assert np.max(np.abs(w - ow)) < 1e-14


@pytest.mark.parametrize("nc", [4, 8])
def test_per_step_state_parity_sine(built_lib, nc):
    """Every one of the first 40 steps (priming, window fill, first D-term samples), one step per launch."""
    cfg, gpu, orc = make_pair(nc, 257)
    assert gpu.kernel_variant == "fast"
    worst = 0.0
    for step in range(1, 41):
        gpu.step(1); orc.step(1)
        pg, tg = gpu.platform_state(); po, to = orc.platform_state()
        worst = max(worst, state_rel_err(pg, tg, po, to))
        jg = gpu.joint_states(); jo = orc.joint_states()
        for a, b in zip(jg, jo):
            assert np.max(np.abs(a - b)) < 1e-9 * max(1.0, np.max(np.abs(b))), f"joint states differ at step {step}"
    assert worst < PER_STEP_TOL, worst
    gpu.close()


@pytest.mark.parametrize("nc,k", [(4, 1000), (8, 1000)])
def test_1000_step_divergence_bounded(built_lib, nc, k):
    """One persistent launch of 1000 steps against 1000 oracle steps."""
    cfg, gpu, orc = make_pair(nc, 300)
    gpu.step(k); orc.step(k)
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    err = state_rel_err(pg, tg, po, to)
    assert err < DIVERGENCE_TOL_1000, err
    assert gpu.step_count == k and abs(gpu.sim_time - k * cfg.dt) < 1e-12
    gpu.close()


def test_launch_split_invariance(built_lib):
    """K steps in one launch == the same K steps in uneven launches, bitwise (state round-trips through HBM)."""
    _, a, _ = make_pair(4, 200)
    _, b, _ = make_pair(4, 200)
    a.step(137)
    for k in (1, 2, 10, 11, 13, 100):
        b.step(k)
    pa, ta = a.platform_state(); pb, tb = b.platform_state()
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
    for x, y in zip(a.joint_states(), b.joint_states()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.pid_state(), b.pid_state())
    a.close(); b.close()


def test_instance_indexing_and_cable_order_bit_exact(built_lib):
    """Permuting the instances permutes the outputs bit-for-bit; a cable's column only depends on its own index."""
    n, nc = 333, 4
    cfg = cb.default_config(nc)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 3)
    perm = np.random.default_rng(0).permutation(n)
    outs = []
    for order in (np.arange(n), perm):
        with cb.CdprBatch(cfg, n) as g:
            g.set_platform_state(pose7[order], twist6[order])
            g.set_sine_cmd(amp[order], freq[order], phase[order])
            g.step(64)
            outs.append((g.platform_state(), g.joint_states()))
    (p0, t0), j0 = outs[0]
    (p1, t1), j1 = outs[1]
    assert np.array_equal(p0[perm], p1) and np.array_equal(t0[perm], t1)
    for a, b in zip(j0, j1):
        assert np.array_equal(a[perm], b)
    # cable order: per-cable velocity commands, oracle agrees column by column
    axes = np.tile(np.array([0.01, -0.02, 0.03, -0.04], dtype=np.float32), (n, 1))
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6)
        g.set_velocity_cmd(axes)
        g.step(50)
        jg = g.joint_states()
    o = ob.Batch(to_oracle_config(cfg), n, pose7, twist6)
    o.velocity_cmd(axes); o.step(50)
    jo = o.joint_states()
    for a, b in zip(jg, jo):
        assert np.max(np.abs(a - b)) < 1e-9


def test_position_force_modes_and_switches(built_lib):
    """Position hold after Load, then velocity, then position again, then a raw effort command."""
    n, nc = 130, 4
    cfg, gpu, orc = make_pair(nc, n, sine=False)
    rng = np.random.default_rng(5)
    def check(tag):
        pg, tg = gpu.platform_state(); po, to = orc.platform_state()
        e = state_rel_err(pg, tg, po, to)
        assert e < 1e-8, (tag, e)
        for a, b in zip(gpu.joint_states(), orc.joint_states()):
            assert np.max(np.abs(a - b)) < 1e-8, tag
    gpu.step(25); orc.step(25); check("position hold after Load")
    v = rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32)
    gpu.set_velocity_cmd(v); orc.velocity_cmd(v)
    gpu.step(40); orc.step(40); check("velocity")
    p = rng.uniform(-0.02, 0.02, (n, nc)).astype(np.float32)
    gpu.set_position_cmd(p); orc.position_cmd(p)
    gpu.step(40); orc.step(40); check("position")
    # both pending in the same step: velocity is applied first, position wins (CdprGazeboPlugin.cpp:206-219)
    gpu.set_velocity_cmd(v); gpu.set_position_cmd(p); orc.velocity_cmd(v); orc.position_cmd(p)
    gpu.step(15); orc.step(15); check("velocity+position")
    f = rng.uniform(2.0, 6.0, (n, nc))
    gpu.set_effort_cmd(f); orc.effort_cmd(f)
    gpu.step(30); orc.step(30); check("force")
    gpu.set_velocity_cmd(v); orc.velocity_cmd(v)
    gpu.step(30); orc.step(30); check("force->velocity")
    gpu.close()


def test_malformed_command_is_dropped(built_lib):
    n = 64
    cfg, gpu, orc = make_pair(4, n, sine=False)
    gpu.step(5)
    before = gpu.get_state()
    with pytest.raises(cb.CdprError) as ei:
        gpu.set_velocity_cmd(np.zeros((n, 3), dtype=np.float32))
    assert ei.value.code == cb.api.ERR_BAD_LENGTH
    with pytest.raises(cb.CdprError):
        gpu.set_position_cmd(np.zeros((n, 5), dtype=np.float32))
    assert np.array_equal(before, gpu.get_state())
    gpu.close()


def general_cfg(cfg):
    cfg.velocity_epsilon = 0.02       # hold below 2 cm/s: both Pids live, non-uniform D windows
    cfg.vel_pid.p_cascade = 1         # biquad on the P input
    cfg.vel_pid.d_cascade = 2         # two biquads on the D output


@pytest.mark.parametrize("nc", [4, 8])
def test_general_variant_hold_and_filters(built_lib, nc):
    cfg, gpu, orc = make_pair(nc, 150, cfg_edit=general_cfg)
    assert gpu.kernel_variant == "flex"
    worst = 0.0
    for k in (1, 1, 1, 7, 20, 70, 400, 1500):   # sine commands cross the epsilon band several times
        gpu.step(k); orc.step(k)
        pg, tg = gpu.platform_state(); po, to = orc.platform_state()
        worst = max(worst, state_rel_err(pg, tg, po, to))
    assert worst < 1e-7, worst
    gpu.close()


def test_general_variant_cmd_limit_zero_and_short_window(built_lib):
    def edit(cfg):
        cfg.vel_pid.cmd_limit = 0.0       # Pid.cpp:175-184: mCmd frozen, anti-windup accumulates
        cfg.vel_pid.d_buffer_length = 5
        cfg.vel_pid.d_degree = 1
    cfg, gpu, orc = make_pair(4, 100, cfg_edit=edit)
    assert gpu.kernel_variant == "general"
    gpu.step(300); orc.step(300)
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    assert state_rel_err(pg, tg, po, to) < 1e-7
    gpu.close()


def test_checkpoint_resume_bitwise(built_lib):
    _, a, _ = make_pair(8, 140)
    a.step(77)
    blob = a.get_state()
    a.step(200)
    ref = a.platform_state()
    _, b, _ = make_pair(8, 140)
    b.set_state(blob)
    assert b.step_count == 77
    b.step(200)
    out = b.platform_state()
    assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
    a.close(); b.close()


def test_snapshots_match_stepwise_states(built_lib):
    import torch
    n, every, k = 190, 25, 200
    _, a, _ = make_pair(4, n)
    buf = torch.zeros((k // every, 13, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()   # the handle runs on its own stream: buffers handed to it must be ready
    a.set_snapshots(every, buf.data_ptr(), buf.shape[0])
    a.step(k)
    a.synchronize()
    assert a.snapshot_count == k // every
    snaps = buf.cpu().numpy()
    _, b, _ = make_pair(4, n)
    for s in range(k // every):
        b.step(every)
        pose, twist = b.platform_state()
        assert np.array_equal(snaps[s, 0:3].T, pose[:, 0:3])
        assert np.array_equal(snaps[s, 3].T, pose[:, 6]) and np.array_equal(snaps[s, 4:7].T, pose[:, 3:6])
        assert np.array_equal(snaps[s, 7:13].T, twist)
    a.close(); b.close()


def test_shard_equivalence_bitwise(built_lib):
    """1 handle over all instances == concatenation of 3 contiguous shards."""
    n, nc = 1000, 8
    cfg = cb.default_config(nc)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 7)
    def run(lo, hi):
        with cb.CdprBatch(cfg, hi - lo) as g:
            g.set_platform_state(pose7[lo:hi], twist6[lo:hi]); g.set_sine_cmd(amp[lo:hi], freq[lo:hi], phase[lo:hi])
            g.step(120)
            return g.platform_state()
    full = run(0, n)
    parts = [run(*wl.shard_range(n, r, 3)) for r in range(3)]
    assert np.array_equal(full[0], np.concatenate([p[0] for p in parts]))
    assert np.array_equal(full[1], np.concatenate([p[1] for p in parts]))


def test_rollout_costs_match_oracle(built_lib):
    import torch
    nc, n_robots, n_seq, n_cmd, spc = 4, 3, 40, 6, 10
    cfg = cb.default_config(nc)
    cmds = wl.c5_rollouts(n_seq, n_cmd, nc)
    _, _, _, pose7, twist6 = wl.c3_instances(n_robots, 11)
    target, lam = np.array([0.0, 0.0, 0.31]), 0.1
    with cb.CdprBatch(cfg, n_robots * n_seq) as g:
        dev = torch.zeros(n_seq, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        cost = g.rollout(n_robots, n_seq, cmds, spc, target, lam, pose7, twist6, dev_cost_seq=dev.data_ptr())
        g.synchronize()
        cost_seq = dev.cpu().numpy()
    ocost = np.zeros(n_robots * n_seq)
    o = ob.Batch(to_oracle_config(cfg), n_robots * n_seq, np.repeat(pose7, n_seq, axis=0), np.repeat(twist6, n_seq, axis=0))
    for c in range(n_cmd):
        o.velocity_cmd(np.tile(cmds[:, c, :], (n_robots, 1)))
        for _ in range(spc):
            o.step(1)
            pose, twist = o.platform_state()
            ocost += np.sum((pose[:, :3] - target) ** 2, axis=1) + lam * np.sum(twist[:, 3:] ** 2, axis=1)
    assert np.max(np.abs(cost - ocost) / ocost) < 1e-9
    assert np.max(np.abs(cost_seq - ocost.reshape(n_robots, n_seq).sum(axis=0)) / cost_seq) < 1e-9


def test_full_size_properties(built_lib):
    """At a size the oracle cannot follow: unit quaternions, finite state, platform stays inside the frame,
    static-equilibrium tension near m g / (4 * 0.617802) for slow commands."""
    n = 1 << 20                       # BASELINE.json configs[2]: 2^20 instances x 1000 steps
    cfg = cb.default_config(4)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        g.step(1000)
        pose, twist = g.platform_state()
        _, _, eff = g.joint_states()
    assert np.all(np.isfinite(pose)) and np.all(np.isfinite(twist))
    assert np.max(np.abs(np.linalg.norm(pose[:, 3:], axis=1) - 1.0)) < 1e-14
    assert np.all(np.abs(pose[:, :2]) < 0.3) and np.all((pose[:, 2] > 0.0) & (pose[:, 2] < 0.6))
    assert abs(np.median(eff.sum(axis=1)) - 4 * 3.9657) < 1.5
    # the first 512 instances of the full-size batch are exactly what a 512-instance batch computes (no cross-talk)
    with cb.CdprBatch(cfg, 512) as g:
        g.set_platform_state(pose7[:512], twist6[:512]); g.set_sine_cmd(amp[:512], freq[:512], phase[:512])
        g.step(1000)
        p2, t2 = g.platform_state()
    assert np.array_equal(p2, pose[:512]) and np.array_equal(t2, twist[:512])


def test_cpp_host_shim_runs_the_sine_driver(built_lib):
    """The plugin-shaped C++ shim (Load / callbacks / update / publish*) driven by the restated sinevelocitytest loop,
    against the oracle running the same schedule.  The shim publishes like the plugin: the pose BEFORE step k with the
    effort OF step k."""
    import subprocess
    from cdpr_simulation_b200 import build as b
    exe = b.build_host()
    out = subprocess.run([exe, "3", "600", "0"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    rows = np.array([[float(x) for x in line.split()] for line in out])
    assert rows.shape == (6, 5)
    cfg = ob.default_config(4)
    o = ob.Batch(cfg, 1, amp=[0.05], freq=[0.1], phase=[0.0])
    done = 0
    for k, x, y, z, eff in rows:
        o.step(int(k) - 1 - done)
        pose, _ = o.platform_state()
        o.step(1)
        done = int(k)
        effort = o.last_outputs()[3]
        assert abs(pose[0, 2] - z) < 1e-9 * 0.3 and abs(pose[0, 0] - x) < 1e-12 and abs(pose[0, 1] - y) < 1e-12
        assert abs(effort[0, 0] - eff) < 1e-8


@pytest.mark.parametrize("n", [1, 2, 127, 129])
def test_ragged_batch_sizes(built_lib, n):
    """Batches that do not fill a block (the reference's own case is ONE robot): tail threads never store."""
    cfg, gpu, orc = make_pair(4, n, seed=21)
    gpu.step(90); orc.step(90)
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    assert pg.shape == (n, 7) and state_rel_err(pg, tg, po, to) < PER_STEP_TOL
    gpu.close()


def test_reference_single_robot_config1(built_lib):
    """BASELINE.json configs[0]: ONE 4-cable robot from cdpr_gazebo.launch under sinevelocitytest's default command
    (0.05 m/s, 0.1 Hz), 3 s of simulated time."""
    cfg = cb.default_config(4)
    with cb.CdprBatch(cfg, 1) as g:
        g.set_sine_cmd(0.05, 0.1, 0.0)
        o = ob.Batch(to_oracle_config(cfg), 1, amp=[0.05], freq=[0.1], phase=[0.0])
        for _ in range(3):
            g.step(1000); o.step(1000)
            pg, tg = g.platform_state(); po, to = o.platform_state()
            assert state_rel_err(pg, tg, po, to) < DIVERGENCE_TOL_1000
        assert abs(g.sim_time - 3.0) < 1e-12


def test_reset_restores_post_load_state(built_lib):
    _, a, _ = make_pair(4, 70)
    first = None
    for _ in range(2):
        a.step(123)
        out = a.platform_state()
        if first is None:
            first = out
            a.reset()
            amp, freq, phase, pose7, twist6 = wl.c3_instances(70, 1)
            a.set_platform_state(pose7, twist6)
    assert np.array_equal(first[0], out[0]) and np.array_equal(first[1], out[1])
    a.close()


class _OracleTarget:
    """Adapter so cdpr_simulation_b200.drivers.run can play a driver against the oracle."""
    def __init__(self, batch): self.b = batch
    def set_velocity_cmd(self, axes): self.b.velocity_cmd(axes)
    def set_position_cmd(self, axes): self.b.position_cmd(axes)
    def step(self, k): self.b.step(k)


@pytest.mark.parametrize("driver_name,eps,variant", [("SquareVelocity", -0.001, "fast"), ("SquareVelocity", 0.001, "flex"),
                                                     ("SquarePosition", -0.001, "fast"), ("SineVelocity", -0.001, "fast")])
def test_reference_drivers_against_oracle(built_lib, driver_name, eps, variant):
    """The reference's three manual test drivers (square velocity with dead band -> hold when eps > 0, square position,
    sine velocity) played headless against the CUDA batch and the oracle."""
    from cdpr_simulation_b200 import drivers
    n, nc, steps = 96, 4, 3200
    def edit(cfg): cfg.velocity_epsilon = eps
    cfg, gpu, orc = make_pair(nc, n, seed=13, cfg_edit=edit, sine=False)
    assert gpu.kernel_variant == variant
    drivers.run(gpu, getattr(drivers, driver_name)(), n, nc, steps, cfg.dt)
    drivers.run(_OracleTarget(orc), getattr(drivers, driver_name)(), n, nc, steps, cfg.dt)
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    assert state_rel_err(pg, tg, po, to) < 1e-7
    for a, b in zip(gpu.joint_states(), orc.joint_states()):
        assert np.max(np.abs(a - b)) < 1e-7
    gpu.close()


def _inertia_general(cfg):
    for k, v in enumerate([1.0, 2.0, 3.0, 0.1, 0.05, -0.02]):
        cfg.inertia[k] = v
    for c in range(cfg.n_cables):
        cfg.platform_anchor[c][2] = 0.004 * (c - 1.5)      # anchors off the platform's xy plane

def _inertia_diag(cfg):
    cfg.inertia[0], cfg.inertia[1], cfg.inertia[2] = 0.8, 1.7, 2.9

def _inertia_diag_bz(cfg):
    _inertia_diag(cfg)
    cfg.platform_anchor[2][2] = -0.01

def _iso_bz(cfg):
    cfg.platform_anchor[1][2] = 0.02

def _unpaired(cfg):
    # the 8-cable cube pairs cables c and c + 4 (same platform anchor, frame anchors above each other): break one pair, so the
    # kernel without the paired-anchor form runs; and give the pairs different heights otherwise
    if cfg.n_cables == 8:
        cfg.frame_anchor[5][0] += 0.01
        cfg.frame_anchor[6][2] = 0.05

def _paired_uneven(cfg):
    if cfg.n_cables == 8:
        cfg.frame_anchor[6][2] = 0.05; cfg.frame_anchor[7][2] = -0.02     # still pairs, each with its own height difference


@pytest.mark.parametrize("edit", [_inertia_general, _inertia_diag, _inertia_diag_bz, _iso_bz, _unpaired, _paired_uneven])
@pytest.mark.parametrize("nc", [4, 8])
def test_robot_constant_specialisations(built_lib, nc, edit):
    """Every compile-time specialisation of the step kernel (general / diagonal / isotropic inertia, anchors in or
    off the platform plane) against the oracle, with an initial spin so the gyroscopic term matters."""
    cfg = cb.default_config(nc)
    edit(cfg)
    n = 160
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 17)
    twist6[:, 3:] = np.random.default_rng(3).uniform(-1.0, 1.0, (n, 3))
    gpu = cb.CdprBatch(cfg, n)
    gpu.set_platform_state(pose7, twist6); gpu.set_sine_cmd(amp, freq, phase)
    orc = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    assert gpu.kernel_variant == "fast"
    gpu.step(250); orc.step(250)
    pg, tg = gpu.platform_state(); po, to = orc.platform_state()
    assert state_rel_err(pg, tg, po, to) < 1e-8
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("n", [65536, (1 << 19) + 77])
def test_ik_device_soa(built_lib, nc, n):
    """The device (SoA) kinematics sweep at config-2 size and at a ragged large size."""
    import torch
    cfg = cb.default_config(nc)
    pose7, twist6 = wl.c2_poses(n, seed=0)
    st = np.ascontiguousarray(np.concatenate([pose7[:, :3], pose7[:, 6:7], pose7[:, 3:6], twist6], axis=1).T)
    d_in = torch.from_numpy(st).cuda(); d_out = torch.empty((nc, 8, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with cb.CdprBatch(cfg, 1) as g:
        g.ik_device(n, d_in.data_ptr(), d_out.data_ptr()); g.synchronize()
    out = d_out.cpu().numpy()
    sub = slice(0, 20000)
    oln, olr, ow = ob.ik(to_oracle_config(cfg), pose7[sub], twist6[sub])
    assert np.max(np.abs(out[:, 0, sub].T - oln) / oln) < 1e-14 and np.max(np.abs(out[:, 1, sub].T - olr)) < 1e-14
    assert np.max(np.abs(np.transpose(out[:, 2:8, sub], (2, 0, 1)) - ow)) < 1e-14
    tail = slice(n - 500, n)
    oln, _, _ = ob.ik(to_oracle_config(cfg), pose7[tail], twist6[tail])
    assert np.max(np.abs(out[:, 0, tail].T - oln) / oln) < 1e-14


def test_api_edge_cases(built_lib):
    import torch
    n = 50
    cfg, g, _ = make_pair(4, n)
    g.step(0)
    assert g.step_count == 0
    with pytest.raises(cb.CdprError):
        g.step(-1)
    # snapshot buffer smaller than the number of snapshots produced: the first `capacity` are written, the rest dropped
    buf = torch.full((2, 13, n), -7.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    g.set_snapshots(10, buf.data_ptr(), 2)
    g.step(55); g.synchronize()
    assert g.snapshot_count == 2 and not bool((buf == -7.0).any())
    # a checkpoint of another shape is refused and leaves the state alone
    other = cb.CdprBatch(cfg, n + 1)
    before = g.platform_state()
    with pytest.raises(cb.CdprError):
        g.set_state(other.get_state())
    after = g.platform_state()
    assert np.array_equal(before[0], after[0])
    other.close(); g.close()


def test_async_mode_and_external_stream_match_sync(built_lib):
    import torch
    n = 300
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 4)
    cfg = cb.default_config(8)
    with cb.CdprBatch(cfg, n) as a:
        a.set_platform_state(pose7, twist6); a.set_sine_cmd(amp, freq, phase); a.step(77)
        ref = a.platform_state(); refj = a.joint_states()
    s = torch.cuda.Stream()
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    p_in = [pin(x) for x in (pose7, twist6, amp, freq, phase)]
    outs = [torch.empty(sh, dtype=torch.float64).pin_memory() for sh in ((n, 7), (n, 6), (n, 8), (n, 8), (n, 8))]
    with cb.CdprBatch(cfg, n) as b:
        b.set_stream(s.cuda_stream); b.set_async(True)
        b.set_platform_state(p_in[0].numpy(), p_in[1].numpy()); b.set_sine_cmd(p_in[2].numpy(), p_in[3].numpy(), p_in[4].numpy())
        b.step(77)
        b.platform_state((outs[0].numpy(), outs[1].numpy())); b.joint_states(tuple(o.numpy() for o in outs[2:]))
        b.synchronize()
    assert np.array_equal(outs[0].numpy(), ref[0]) and np.array_equal(outs[1].numpy(), ref[1])
    for o, r in zip(outs[2:], refj):
        assert np.array_equal(o.numpy(), r)


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nc", [4, 8])
def test_gpu_against_the_references_own_force_law(built_lib, nc):
    """CUDA path vs the reduced model driven by the REFERENCE's compiled Pid.cpp / JointForceCalculator.cpp (oracle L0).
    The only difference allowed is the reference's own D-term conditioning noise (absolute-time normal equations,
    SURVEY.md F5): ~1e-7 on the pose after 1.5 s, against 1e-14 when the same fit is done in window-relative time."""
    n = 48
    cfg = cb.default_config(nc)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 23)
    with cb.CdprBatch(cfg, n) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        g.step(1500)
        pg, tg = g.platform_state()
        _, _, eg = g.joint_states()
    o = ob.Batch(to_oracle_config(cfg), n, pose7, twist6, amp, freq, phase)
    o.step_reference_forcelaw(1500)
    po, to = o.platform_state()
    eo = o.last_outputs()[3]
    assert np.max(np.abs(pg - po)) < 1e-6 and np.max(np.abs(tg - to)) < 1e-4
    assert np.max(np.abs(eg - eo)) < 1e-2          # forces: the D gain multiplies the reference's derivative noise
    # and it is NOT trivially satisfied: the platforms moved by centimetres
    assert np.max(np.abs(pg[:, :3] - pose7[:, :3])) > 1e-2


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("limits", [(0.5, 100.0, 100.0), (100.0, 3.0, 100.0), (0.3, 1.5, 100.0), (100.0, 100.0, 2.0), (100.0, 6.0, 4.0)])
def test_fast_variant_saturation_paths(built_lib, nc, limits):
    """Integral clamp with back-calculation, command clamp + anti-windup (which may exceed the clamp, Pid.cpp:181-184)
    and the joint's effort truncation, all inside the FAST kernel: small limits so that every path fires."""
    i_limit, cmd_limit, effort = limits
    def edit(cfg):
        for pid in (cfg.vel_pid, cfg.pos_pid):
            pid.i_limit, pid.cmd_limit = i_limit, cmd_limit
        cfg.effort_limit = effort
    cfg, gpu, orc = make_pair(nc, 140, seed=31, cfg_edit=edit)
    assert gpu.kernel_variant == "fast"
    saw_sat = False
    for k in (3, 30, 200, 700):
        gpu.step(k); orc.step(k)
        pg, tg = gpu.platform_state(); po, to = orc.platform_state()
        assert state_rel_err(pg, tg, po, to) < 1e-8, (limits, k)
        eff = orc.last_outputs()[3]
        saw_sat = saw_sat or bool(np.any(np.abs(eff) >= min(cmd_limit, effort) * 0.999))
        for a, b in zip(gpu.joint_states(), orc.joint_states()):
            assert np.max(np.abs(a - b)) < 1e-8
    if min(cmd_limit, effort) < 10:
        assert saw_sat, "the test is meant to saturate"
    gpu.close()


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("limits", [(0.3, 1.5, 100.0), (100.0, 6.0, 4.0), (0.5, 100.0, 100.0)])
def test_saturated_steps_do_not_depend_on_the_launch_split(built_lib, nc, limits):
    """The hot loop handles saturation optimistically (one vote per step, out-of-line exact pass when it fires); the
    first-steps and last-step bodies clamp inline.  Both must give the same bits, so how K steps are cut into launches
    cannot show in the state, the integrals or the telemetry."""
    i_limit, cmd_limit, effort = limits
    def edit(cfg):
        for pid in (cfg.vel_pid, cfg.pos_pid):
            pid.i_limit, pid.cmd_limit = i_limit, cmd_limit
        cfg.effort_limit = effort
    _, a, _ = make_pair(nc, 160, seed=32, cfg_edit=edit)
    _, b, _ = make_pair(nc, 160, seed=32, cfg_edit=edit)
    a.step(300)
    for k in (1, 2, 10, 11, 13, 100, 163):
        b.step(k)
    pa, ta = a.platform_state(); pb, tb = b.platform_state()
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
    for x, y in zip(a.joint_states(), b.joint_states()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.pid_state(), b.pid_state())
    a.close(); b.close()


@pytest.mark.parametrize("nc", [4, 8])
@pytest.mark.parametrize("shape", ["per_cable_velocity", "position", "cubic_fir_feedforward"])
def test_saturation_with_per_cable_commands(built_lib, nc, shape):
    """The same saturation paths in the layouts that keep per-cable targets (and feed-forward terms) in shared memory,
    in Position mode, and with the FIR form of the D-term (cubic fit): against the oracle, and bitwise against a
    different launch split."""
    def edit(cfg):
        for pid in (cfg.vel_pid, cfg.pos_pid):
            pid.i_limit, pid.cmd_limit = 0.4, 5.0
            if shape == "cubic_fir_feedforward":
                pid.d_degree, pid.forward_gain = 3, 2.5
        cfg.effort_limit = 4.0
    n = 150
    rng = np.random.default_rng(77)
    v = rng.uniform(-0.3, 0.3, (n, nc)).astype(np.float32)
    p = rng.uniform(-0.05, 0.05, (n, nc)).astype(np.float32)
    runs = []
    for split in ((260,), (1, 12, 47, 200)):
        cfg, gpu, orc = make_pair(nc, n, seed=41, cfg_edit=edit, sine=False)
        assert gpu.kernel_variant == "fast"
        if shape == "position":
            gpu.step(20); orc.step(20)
            gpu.set_position_cmd(p); orc.position_cmd(p)
        else:
            gpu.set_velocity_cmd(v); orc.velocity_cmd(v)
        for k in split:
            gpu.step(k); orc.step(k)
        pg, tg = gpu.platform_state(); po, to = orc.platform_state()
        assert state_rel_err(pg, tg, po, to) < 1e-8, (shape, split)
        for a, b in zip(gpu.joint_states(), orc.joint_states()):
            assert np.max(np.abs(a - b)) < 1e-8
        assert np.any(np.abs(orc.last_outputs()[3]) >= 4.0 * 0.999), "the test is meant to saturate"
        runs.append((pg, tg, gpu.joint_states(), gpu.pid_state()))
        gpu.close()
    a, b = runs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y)
