"""CPU-only property tests of the oracle's reduced model: the invariants the GPU path is then
held to at sizes the oracle cannot follow."""
import numpy as np

from cdpr_simulation_b200 import workloads as wl
from oracle import binding as ob


def quat_mul(a, b):  # x y z w
    ax, ay, az, aw = a.T; bx, by, bz, bw = b.T
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=1)


def test_structure_matrix_is_the_length_jacobian():
    """q_dot = W^T [v; w]  <=>  dL/dt matches a central finite difference of the lengths along the twist."""
    for nc in (4, 8):
        cfg = ob.default_config(nc)
        pose7, twist6 = wl.c2_poses(500, seed=3)
        ln, lr, w = ob.ik(cfg, pose7, twist6)
        assert np.allclose(np.linalg.norm(w[..., :3], axis=-1), 1.0, atol=1e-14)
        assert np.allclose(lr, -np.einsum("ncd,nd->nc", w, twist6), atol=1e-15)
        h = 1e-6
        def moved(sign):
            p = pose7.copy()
            p[:, :3] += sign * h * twist6[:, :3]
            dq = np.concatenate([0.5 * sign * h * twist6[:, 3:], np.zeros((len(p), 1))], axis=1)
            q = pose7[:, 3:] + quat_mul(dq, pose7[:, 3:])
            p[:, 3:] = q / np.linalg.norm(q, axis=1, keepdims=True)
            return ob.ik(cfg, p, twist6)[0]
        fd = (moved(+1) - moved(-1)) / (2 * h)
        assert np.max(np.abs(fd - lr)) < 1e-8


def test_quaternion_stays_unit_and_free_fall_without_cables():
    cfg = ob.default_config(4)
    cfg.vel_pid.p_gain = cfg.vel_pid.i_gain = cfg.vel_pid.d_gain = 0.0
    cfg.pos_pid.p_gain = cfg.pos_pid.i_gain = cfg.pos_pid.d_gain = 0.0
    cfg.cable_damping = 0.0
    _, _, _, pose7, twist6 = wl.c3_instances(16, seed=2)
    twist6[:, 3:] = 0.3
    b = ob.Batch(cfg, 16, pose7, twist6)
    b.step(200)
    pose, twist = b.platform_state()
    assert np.max(np.abs(np.linalg.norm(pose[:, 3:], axis=1) - 1)) < 1e-15
    t = 200 * cfg.dt
    assert np.allclose(twist[:, 2], -9.8 * t, atol=1e-12)                       # semi-implicit Euler: v = g t exactly
    assert np.allclose(pose[:, 2], pose7[:, 2] - 9.8 * cfg.dt ** 2 * 200 * 201 / 2, atol=1e-12)
    assert np.allclose(twist[:, 3:], 0.3, atol=1e-13)                           # isotropic inertia: no gyroscopic torque


def test_angular_momentum_conserved_with_anisotropic_inertia():
    cfg = ob.default_config(4)
    for p in (cfg.vel_pid, cfg.pos_pid):
        p.p_gain = p.i_gain = p.d_gain = 0.0
    cfg.cable_damping = 0.0
    cfg.gravity[2] = 0.0
    cfg.inertia[0], cfg.inertia[1], cfg.inertia[2], cfg.inertia[3] = 1.0, 2.0, 3.0, 0.1
    twist6 = np.zeros((1, 6)); twist6[0, 3:] = [0.4, -0.2, 0.7]
    pose7 = np.array([[0, 0, 0.3, 0, 0, 0, 1.0]])
    b = ob.Batch(cfg, 1, pose7, twist6)
    I = np.array([[1.0, 0.1, 0], [0.1, 2.0, 0], [0, 0, 3.0]])
    def ang_mom():
        pose, tw = b.platform_state()
        x, y, z, w = pose[0, 3:]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        return R @ I @ R.T @ tw[0, 3:]
    l0 = ang_mom()
    b.step(2000)
    assert np.linalg.norm(ang_mom() - l0) / np.linalg.norm(l0) < 2e-3            # first-order integrator drift over 2 s


def test_hold_freezes_position_target_at_last_position():
    """eps > 0: a velocity command inside the band is served by the position Pid on the latched position
    (JointForceCalculator.cpp:72-82)."""
    cfg = ob.default_config(4)
    cfg.velocity_epsilon = 0.01
    b = ob.Batch(cfg, 1)
    b.velocity_cmd(np.full((1, 4), 0.03, dtype=np.float32)); b.step(300)
    pos_before = b.joint_states()[0].copy()
    b.velocity_cmd(np.full((1, 4), 0.001, dtype=np.float32))                   # inside the band: hold
    # the position Pid takes over un-primed (no integral yet): the platform sags, then is pulled back to the
    # position latched at the last velocity-mode update -- it does NOT follow the 1 mm/s command
    err = []
    for _ in range(10):
        b.step(1000)
        err.append(np.max(np.abs(b.joint_states()[0] - pos_before)))
    assert err[0] > 5e-3 and err[-1] < 1e-3 and all(x > y for x, y in zip(err[1:], err[2:]))
    _, _, mode = b.targets()
    assert np.all(mode == 2)


def test_equilibrium_tension_under_position_hold():
    cfg = ob.default_config(4)
    b = ob.Batch(cfg, 1)
    b.step(18000)                                  # the hold settles with a ~2.6 s time constant
    _, _, eff = b.joint_states()
    assert np.allclose(eff, 3.9657, atol=1e-3)     # m g / (4 * 0.617802), SURVEY.md 8(c)
    pose, _ = b.platform_state()
    assert abs(pose[0, 2] - 0.3) < 1e-4
