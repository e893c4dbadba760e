"""Robot description -> cdpr_config.

Restates the geometry conventions of the reference's YAML -> SDF generator (sdf/gen_cdpr.py:101-125, input
sdf/cube.yaml): `points[i].frame` is the frame anchor a_i, `points[i].platform` the platform anchor b_i in platform
coordinates, `platform.position` the home pose, `joints.actuated` the prismatic joint's damping / effort.  Only the
numbers the hot path consumes are read; emitting SDF is out of scope."""
from __future__ import annotations

import math

from .api import Config, default_config, MAX_CABLES


def _rpy_to_quat(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2), math.cos(y / 2), math.sin(y / 2)
    return (cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy)


def config_from_description(desc: dict, home_xyz=None) -> Config:
    """desc: the parsed YAML (same keys as sdf/cube.yaml).  home_xyz overrides platform.position.xyz -- the
    reference's cube.yaml says z = 2 while the authoritative cube.sdf places the platform at z = 0.3."""
    pts = desc["points"]
    if not 1 <= len(pts) <= MAX_CABLES:
        raise ValueError("invalid joint count")
    cfg = default_config(4)
    cfg.n_cables = len(pts)
    for i in range(MAX_CABLES):
        for k in range(3):
            cfg.frame_anchor[i][k] = float(pts[i]["frame"][k]) if i < len(pts) else 0.0
            cfg.platform_anchor[i][k] = float(pts[i]["platform"][k]) if i < len(pts) else 0.0
    plat = desc["platform"]
    xyz = home_xyz if home_xyz is not None else plat["position"]["xyz"]
    for k in range(3):
        cfg.home_pos[k] = float(xyz[k])
    q = _rpy_to_quat(*[float(a) for a in plat["position"].get("rpy", [0, 0, 0])])
    for k in range(4):
        cfg.home_quat[k] = q[k]
    cfg.mass = float(plat["mass"])
    inertia = [float(v) for v in plat.get("inertia", [1, 1, 1, 0, 0, 0])]   # ixx iyy izz ixy ixz iyz (gen_cdpr.py BuildInertial order)
    for k in range(6):
        cfg.inertia[k] = inertia[k]
    act = desc.get("joints", {}).get("actuated", {})
    cfg.cable_damping = float(act.get("damping", cfg.cable_damping))
    cfg.effort_limit = float(act.get("effort", cfg.effort_limit))
    cfg.slider_velocity_limit = float(act.get("velocity", cfg.slider_velocity_limit))
    cfg.passive_damping = float(desc.get("joints", {}).get("passive", {}).get("damping", cfg.passive_damping))
    set_leg_axes(cfg)
    return cfg


def set_leg_axes(cfg: Config) -> Config:
    """Leg joint axes of the generator (sdf/gen_cdpr.py:113-125,152,171,209,225,237): the leg frame is the rotation about
    z x u_fp that takes z onto the frame -> platform direction u_fp; rev_X turns about its first column; the platform-side
    gimbal axes are written as plain "0 0 1" / "1 0 0" (model frame in SDF 1.4).  Same arithmetic, in the same order, as
    cdpr_config_default."""
    for c in range(MAX_CABLES):
        for k in range(3):
            cfg.leg_axis_frame[c][k] = cfg.leg_axis_cable[c][k] = cfg.leg_axis_platform[c][k] = 0.0
    for c in range(cfg.n_cables):
        ufp = [cfg.home_pos[k] + cfg.platform_anchor[c][k] - cfg.frame_anchor[c][k] for k in range(3)]
        n = 0.0
        for k in range(3):
            n += ufp[k] * ufp[k]
        n = math.sqrt(n)
        ufp = [x / n for x in ufp]
        ax = [-ufp[1], ufp[0], 0.0]
        sn, cs = math.sqrt(ax[0] * ax[0] + ax[1] * ax[1]), ufp[2]
        if sn > 0.0:
            ax[0] /= sn; ax[1] /= sn
        cfg.leg_axis_frame[c][0] = cs + ax[0] * ax[0] * (1.0 - cs)
        cfg.leg_axis_frame[c][1] = ax[2] * sn + ax[1] * ax[0] * (1.0 - cs)
        cfg.leg_axis_frame[c][2] = -ax[1] * sn + ax[2] * ax[0] * (1.0 - cs)
        cfg.leg_axis_cable[c][2] = 1.0
        cfg.leg_axis_platform[c][0] = 1.0
    return cfg


def config_from_yaml(path: str, home_xyz=None) -> Config:
    import yaml
    with open(path) as f:
        return config_from_description(yaml.safe_load(f), home_xyz)


def apply_launch_params(cfg: Config, params: dict) -> Config:
    """ROS parameters of launch/cdpr_gazebo.launch:17-39 (names of CdprGazeboPlugin.h:32-54, without the namespace)."""
    v, p = cfg.vel_pid, cfg.pos_pid
    m = {"velocityControllerForward": (v, "forward_gain"), "velocityControllerP": (v, "p_gain"), "velocityControllerI": (v, "i_gain"),
         "velocityControllerD": (v, "d_gain"), "velocityControllerDdegree": (v, "d_degree"), "velocityControllerDbuffer": (v, "d_buffer_length"),
         "velocityControllerMaxI": (v, "i_limit"), "velocityControllerMaxCmd": (v, "cmd_limit"),
         "velocityControllerPcutoff": (v, "p_cutoff"), "velocityControllerPquality": (v, "p_quality"), "velocityControllerPcascade": (v, "p_cascade"),
         "velocityControllerDcutoff": (v, "d_cutoff"), "velocityControllerDquality": (v, "d_quality"), "velocityControllerDcascade": (v, "d_cascade"),
         "positionControllerP": (p, "p_gain"), "positionControllerI": (p, "i_gain"), "positionControllerD": (p, "d_gain"),
         "positionControllerDdegree": (p, "d_degree"), "positionControllerDbuffer": (p, "d_buffer_length"),
         "positionControllerMaxI": (p, "i_limit"), "positionControllerMaxCmd": (p, "cmd_limit")}
    for name, value in params.items():
        if name == "velocityEpsilon":
            cfg.velocity_epsilon = float(value)
        elif name in m:
            obj, field = m[name]
            setattr(obj, field, int(value) if field in ("d_degree", "d_buffer_length", "p_cascade", "d_cascade") else float(value))
    p.forward_gain = 0.0            # CdprGazeboPlugin.cpp:123
    p.p_cascade = p.d_cascade = 0   # CdprGazeboPlugin.cpp:133
    return cfg
