"""CPU-only: the leg-fidelity model of the CPU checker (parity-unpinned, no Gazebo/ODE here).
(1) its constants against the SDF literals; (2) its closed-form joint rates and mass matrix against an INDEPENDENT numerical
derivation -- the explicit link chain of cube.sdf:344-518 posed for two nearby platform poses and differenced; (3) limits:
massless legs give back the reduced model, damping dissipates, without damping the energy of the free system is kept."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_constants.json")))


def test_leg_constants_match_the_sdf():
    assert G["sdf_version"] == "1.4"                      # joint axes are expressed in the model frame
    cfg = ob.default_config(4)
    assert cfg.leg_model == 0                             # reduced model unless asked for
    for i, c in enumerate(G["cables"]):
        for nm, link in c["leg_links"].items():
            assert link["mass"] == cfg.leg_link_mass and link["inertia"] == [cfg.leg_link_inertia] * 3 + [0.0] * 3, nm
        for nm, j in c["leg_joints"].items():
            assert j["damping"] == cfg.passive_damping, nm
        ax = np.array(c["leg_joints"]["rev_X"]["axis"])
        assert np.allclose(ax / np.linalg.norm(ax), list(cfg.leg_axis_frame[i]), atol=2e-6)
        assert c["leg_joints"]["rev_X"]["parent"] == "frame" and c["leg_joints"]["rev_Y"]["child"] == f"virt_Y{i}"
        assert c["leg_joints"]["rev_Zpf"]["axis"] == list(cfg.leg_axis_cable[i]) == [0.0, 0.0, 1.0]
        assert c["leg_joints"]["rev_Zpf"]["parent"] == f"cable{i}" and c["leg_joints"]["rev_Zpf"]["child"] == f"virt_Ypf{i}"
        assert c["leg_joints"]["rev_Xpf"]["axis"] == list(cfg.leg_axis_platform[i]) == [1.0, 0.0, 0.0]
        assert c["leg_joints"]["rev_Xpf"]["parent"] == "platform" and c["leg_joints"]["rev_Ypf"]["parent"] == f"virt_Xpf{i}"
        assert c["upper"] == cfg.slider_upper and c["lower"] == cfg.slider_lower and c["velocity"] == cfg.slider_velocity_limit
        # centre of mass of the cable link: l/2 from the platform anchor along the leg (cube.sdf:344)
        pp = np.array(c["platform_anchor_link_pose"][:3]); fp = np.array(c["frame_anchor_link_pose"][:3])
        u = (fp - pp) / np.linalg.norm(fp - pp)
        assert np.allclose(np.array(c["leg_links"]["cable"]["pose"][:3]), pp + cfg.leg_cable_com * u, atol=1e-6)
        assert np.allclose(c["leg_links"]["virt_X"]["pose"][:3], fp) and np.allclose(c["leg_links"]["virt_Ypf"]["pose"][:3], pp)


# ---- independent derivation: pose every link of a leg explicitly, difference two nearby configurations ------------------
def quat_R(q):  # w x y z
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot(axis, ang):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def leg_links(cfg, i, p, Rp):
    """[(com, R)] of virt_X, virt_Y, cable, virt_Ypf, virt_Xpf for platform position p, rotation Rp: the joint angles are
    solved from the loop closure, then each link is posed by composing joint rotations from its own side of the chain."""
    A = np.array(list(cfg.frame_anchor[i])); b = np.array(list(cfg.platform_anchor[i])); x0 = np.array(list(cfg.leg_axis_frame[i]))
    home_p = np.array(list(cfg.home_pos)); B0 = home_p + b
    u0 = (A - B0) / np.linalg.norm(A - B0)
    y0 = np.cross(u0, x0); y0 /= np.linalg.norm(y0)               # rev_Y axis at home (cube.sdf:419: second column of the leg frame)
    B = p + Rp @ b
    u = (A - B) / np.linalg.norm(A - B)
    # universal joint: u = Rot(x0, thx) Rot(y0, thy) u0 ; solve thy from the component along x0, thx from the rest
    thy = np.arcsin(u @ x0)                                       # Rot(y0, thy) u0 = u0 cos + x0 sin; Rot(x0, .) keeps the x0 component
    w1 = rot(y0, thy) @ u0
    perp = lambda v: v - (v @ x0) * x0
    a, c = perp(w1), perp(u)
    thx = np.arctan2(np.cross(a, c) @ x0, a @ c)
    R_vx = rot(x0, thx)
    R_leg = R_vx @ rot(y0, thy)                                    # virt_Y and the cable link (prismatic: same orientation)
    assert np.allclose(R_leg @ u0, u, atol=1e-9)
    # gimbal: R_leg Rz(phi) Ry(psy) Rx(psx) = Rp, axes z / y / x of the model frame at home (cube.sdf:462-512)
    Gm = R_leg.T @ Rp
    psy = -np.arcsin(Gm[2, 0]); phi = np.arctan2(Gm[1, 0], Gm[0, 0]); psx = np.arctan2(Gm[2, 1], Gm[2, 2])
    Rz, Ry = rot(np.array([0, 0, 1.0]), phi), rot(np.array([0, 1.0, 0]), psy)
    assert np.allclose(Rz @ Ry @ rot(np.array([1.0, 0, 0]), psx), Gm, atol=1e-9)
    C = B + cfg.leg_cable_com * u
    return [(A, R_vx), (A, R_leg), (C, R_leg), (B, R_leg @ Rz), (B, R_leg @ Rz @ Ry)], np.array([thx, thy, phi, psy, psx])


def advance(p, q, v, w, h):
    """pose after time h at constant twist (first order, like the integrator)"""
    dq = 0.5 * np.array([-w @ q[1:], q[0] * w[0] + w[1] * q[3] - w[2] * q[2], q[0] * w[1] - w[0] * q[3] + w[2] * q[1], q[0] * w[2] + w[0] * q[2] - w[1] * q[1]])
    qn = q + h * dq
    return p + h * v, qn / np.linalg.norm(qn)


@pytest.mark.parametrize("nc", [4, 8])
def test_leg_joint_rates_and_mass_matrix_against_the_differenced_link_chain(nc):
    cfg = ob.default_config(nc)
    cfg.leg_model = 1
    rng = np.random.default_rng(5)
    for trial in range(4):
        p = np.array([0.0, 0.0, 0.3]) + rng.uniform(-0.05, 0.05, 3)
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(-0.15, 0.15)
        q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        v, w = rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.5, 0.5, 3)
        pose7 = np.concatenate([p, q[1:], q[:1]])[None]; twist6 = np.concatenate([v, w])[None]
        b = ob.Batch(cfg, 1, pose7, twist6)
        M = b.legs_mass_matrix()
        assert np.allclose(M, M.T, atol=1e-15) and np.all(np.linalg.eigvalsh(M) > 0.9)
        xi = np.concatenate([v, w])
        h = 1e-6
        pm, qm = advance(p, q, v, w, -h); pp, qp = advance(p, q, v, w, +h)
        ke = 0.5 * cfg.mass * (v @ v) + 0.5 * w @ (quat_R(q) @ np.diag(list(cfg.inertia)[:3]) @ quat_R(q).T) @ w
        for i in range(nc):
            lm, am = leg_links(cfg, i, pm, quat_R(qm)); lp, ap = leg_links(cfg, i, pp, quat_R(qp))
            rates_fd = (ap - am) / (2 * h)
            rates = b.legs_joint_rates(i)                     # thx, thy, phi, psy, psx of the closed form
            # joint sign conventions are free (only squares enter): compare magnitudes
            assert np.allclose(np.abs(rates), np.abs(rates_fd), rtol=2e-5, atol=2e-8), (i, rates, rates_fd)
            for (cm, Rm), (cp_, Rp_) in zip(lm, lp):
                vcom = (cp_ - cm) / (2 * h)
                W = (Rp_ - Rm) / (2 * h) @ (0.5 * (Rp_ + Rm)).T   # [w]x
                wl = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) / 2
                ke += 0.5 * cfg.leg_link_mass * (vcom @ vcom) + 0.5 * cfg.leg_link_inertia * (wl @ wl)
        assert abs(0.5 * xi @ M @ xi - ke) < 2e-6 * ke, (0.5 * xi @ M @ xi, ke)
        assert abs(b.legs_energy()[0] - 0.5 * xi @ M @ xi) < 1e-15 * max(1.0, ke)


def _run(cfg, n_steps, pose7=None, twist6=None, force=None):
    b = ob.Batch(cfg, 1, pose7, twist6)
    if force is not None:
        b.effort_cmd(np.full((1, cfg.n_cables), force))
    b.step(n_steps)
    return b


def test_massless_legs_reduce_to_the_reduced_model():
    base = ob.default_config(4)
    legs = ob.default_config(4)
    legs.leg_model = 1; legs.leg_link_mass = 0.0; legs.leg_link_inertia = 0.0; legs.passive_damping = 0.0
    amp, freq, phase = [0.05], [0.1], [0.3]
    pose7 = np.array([[0.01, -0.02, 0.31, 0.01, 0.02, -0.01, 1.0]]); pose7[0, 3:] /= np.linalg.norm(pose7[0, 3:])
    a = ob.Batch(base, 1, pose7, None, amp, freq, phase); b = ob.Batch(legs, 1, pose7, None, amp, freq, phase)
    a.step(400); b.step(400)
    for x, y in zip(a.platform_state(), b.platform_state()):
        assert np.max(np.abs(x - y)) < 1e-12
    # and the real leg constants DO change the trajectory: a few per cent of extra inertia
    legs.leg_link_mass = legs.leg_link_inertia = 1e-3; legs.passive_damping = 0.01
    c = ob.Batch(legs, 1, pose7, None, amp, freq, phase)
    c.step(400)
    diff = np.max(np.abs(c.platform_state()[0][:, :3] - a.platform_state()[0][:, :3]))
    assert 1e-7 < diff < 1e-2, diff


def test_energy_free_motion_and_dissipation():
    """No gravity, zero cable force (Force mode), no joint damping: the kinetic energy 1/2 xi^T M(x) xi of platform + legs
    is an invariant of the exact dynamics.  With the links' velocity-product terms in the model it is kept to the
    integrator's order over 200 steps at |v| = 0.2 m/s, |w| = 0.5 rad/s -- 1.5e-7 relative with the SDF's leg constants and
    3e-6 with 20x heavier legs (without those terms: 8.5e-4 and 1.4e-2).  With the passive damping on, the same run loses
    energy at every step."""
    def run(scale, damping, steps=200):
        c = ob.default_config(4)
        c.leg_model = 1; c.cable_damping = 0.0; c.passive_damping = damping
        c.leg_link_mass = 1e-3 * scale; c.leg_link_inertia = 1e-3 * scale
        for k in range(3):
            c.gravity[k] = 0.0
        b = ob.Batch(c, 1, None, np.array([[0.1, -0.05, 0.2, 0.3, -0.2, 0.4]]))
        b.effort_cmd(np.zeros((1, 4)))
        e = []
        for _ in range(steps):
            e.append(b.legs_energy()[0])
            b.step(1)
        return np.array(e)
    e1, e20 = run(1, 0.0), run(20, 0.0)
    d1, d20 = abs(e1[-1] - e1[0]) / e1[0], abs(e20[-1] - e20[0]) / e20[0]
    assert d1 < 2e-6 and d20 < 2e-5, (d1, d20)
    ed = run(1, 0.5)
    assert np.all(np.diff(ed) < 0.0) and ed[-1] < 0.9 * ed[0]


def test_total_energy_under_gravity_with_legs():
    """Free fall with legs attached (zero cable force, no damping): kinetic + potential energy of platform and leg links
    changes only by the semi-implicit integrator's own first-order term, -1/2 g^2 h^2 per step and unit of falling mass."""
    c = ob.default_config(4)
    c.leg_model = 1; c.cable_damping = 0.0; c.passive_damping = 0.0
    b = ob.Batch(c, 1, None, np.array([[0.05, 0.0, 0.1, 0.0, 0.0, 0.0]]))
    b.effort_cmd(np.zeros((1, 4)))
    e = []
    for _ in range(100):
        ke, pe = b.legs_energy()
        e.append(ke + pe)
        b.step(1)
    per_step = np.diff(np.array(e))
    expect = -0.5 * 9.8 ** 2 * c.dt ** 2 * 1.05             # ~5 % of extra falling mass from the legs
    assert np.all(np.abs(per_step - expect) < 0.15 * abs(expect)), (per_step[:3], expect)
