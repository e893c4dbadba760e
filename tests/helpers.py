"""Shared test helpers: product config <-> oracle config, error metrics."""
import numpy as np

from oracle import binding as ob
import cdpr_simulation_b200 as cb

PID_FIELDS = ["forward_gain", "p_gain", "i_gain", "d_gain", "d_degree", "d_buffer_length", "i_limit", "cmd_limit",
              "p_cutoff", "p_quality", "p_cascade", "d_cutoff", "d_quality", "d_cascade"]


def to_oracle_config(cfg: cb.Config) -> ob.Config:
    """Field-by-field copy (the two structs are declared independently)."""
    o = ob.Config()
    o.n_cables = cfg.n_cables
    for c in range(cb.api.MAX_CABLES):
        for k in range(3):
            o.frame_anchor[c][k] = cfg.frame_anchor[c][k]
            o.platform_anchor[c][k] = cfg.platform_anchor[c][k]
    for k in range(3):
        o.home_pos[k] = cfg.home_pos[k]
        o.gravity[k] = cfg.gravity[k]
    for k in range(4):
        o.home_quat[k] = cfg.home_quat[k]
    for k in range(6):
        o.inertia[k] = cfg.inertia[k]
    o.mass, o.cable_damping, o.effort_limit, o.dt = cfg.mass, cfg.cable_damping, cfg.effort_limit, cfg.dt
    for name in ("vel_pid", "pos_pid"):
        for f in PID_FIELDS:
            setattr(getattr(o, name), f, getattr(getattr(cfg, name), f))
    o.velocity_epsilon = cfg.velocity_epsilon
    for f in ("leg_model", "leg_link_mass", "leg_link_inertia", "leg_cable_com", "passive_damping", "slider_lower", "slider_upper", "slider_velocity_limit"):
        setattr(o, f, getattr(cfg, f))
    for c in range(cb.api.MAX_CABLES):
        for k in range(3):
            o.leg_axis_frame[c][k] = cfg.leg_axis_frame[c][k]
            o.leg_axis_cable[c][k] = cfg.leg_axis_cable[c][k]
            o.leg_axis_platform[c][k] = cfg.leg_axis_platform[c][k]
    o.derive_absolute_time = 0
    return o


def group_rel_err(a, b, floor):
    """max over instances of |a-b| / max(|b|, floor), vector norms over the last axis."""
    a = np.asarray(a); b = np.asarray(b)
    num = np.linalg.norm(a - b, axis=-1)
    den = np.maximum(np.linalg.norm(b, axis=-1), floor)
    return float(np.max(num / den))


def state_rel_err(pose_a, twist_a, pose_b, twist_b):
    """Relative error of the platform state, per group: position, quaternion, linear and angular velocity.
    Floors: 1 mm/s and 1 mrad/s for the velocities (they cross zero), none needed for pose."""
    return max(
        group_rel_err(pose_a[:, :3], pose_b[:, :3], 1e-3),
        group_rel_err(pose_a[:, 3:], pose_b[:, 3:], 1.0),
        group_rel_err(twist_a[:, :3], twist_b[:, :3], 1e-3),
        group_rel_err(twist_a[:, 3:], twist_b[:, 3:], 1e-3),
    )
