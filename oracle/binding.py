"""ctypes binding of the CPU oracle (oracle/liborc.so, oracle/_ref/libcdpr_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_CABLES = 8


class PidParams(C.Structure):
    _fields_ = [
        ("forward_gain", C.c_double), ("p_gain", C.c_double), ("i_gain", C.c_double), ("d_gain", C.c_double),
        ("d_degree", C.c_int32), ("d_buffer_length", C.c_int32),
        ("i_limit", C.c_double), ("cmd_limit", C.c_double),
        ("p_cutoff", C.c_double), ("p_quality", C.c_double), ("p_cascade", C.c_int32),
        ("d_cutoff", C.c_double), ("d_quality", C.c_double), ("d_cascade", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("n_cables", C.c_int32),
        ("frame_anchor", (C.c_double * 3) * MAX_CABLES),
        ("platform_anchor", (C.c_double * 3) * MAX_CABLES),
        ("home_pos", C.c_double * 3),
        ("home_quat", C.c_double * 4),
        ("mass", C.c_double),
        ("inertia", C.c_double * 6),
        ("gravity", C.c_double * 3),
        ("cable_damping", C.c_double),
        ("effort_limit", C.c_double),
        ("dt", C.c_double),
        ("vel_pid", PidParams), ("pos_pid", PidParams),
        ("velocity_epsilon", C.c_double),
        ("leg_model", C.c_int32),
        ("leg_link_mass", C.c_double), ("leg_link_inertia", C.c_double), ("leg_cable_com", C.c_double), ("passive_damping", C.c_double),
        ("leg_axis_frame", (C.c_double * 3) * MAX_CABLES),
        ("leg_axis_cable", (C.c_double * 3) * MAX_CABLES),
        ("leg_axis_platform", (C.c_double * 3) * MAX_CABLES),
        ("slider_lower", C.c_double), ("slider_upper", C.c_double), ("slider_velocity_limit", C.c_double),
        ("derive_absolute_time", C.c_int32),
    ]


def build(force: bool = False) -> None:
    """Compile liborc.so (always possible) and _ref/libcdpr_ref.so (only where /root/reference exists); again whenever a
    source is newer than what was built from it."""
    def mtime(*parts):
        f = os.path.join(HERE, *parts)
        return os.path.getmtime(f) if os.path.exists(f) else 0.0
    src = max(mtime("cdpr_oracle.c"), mtime("cdpr_oracle.h"))
    have_ref_src = os.path.isdir("/root/reference/src/cdpr_gazebo/src")
    stale = mtime("liborc.so") < src or (have_ref_src and mtime("_ref", "libcdpr_ref.so") < max(src, mtime("ref_harness.cpp")))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def _opt(a, dtype=np.float64):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.ctypes.data_as(C.c_void_p)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liborc.so"))
        L.orc_config_default.argtypes = [C.POINTER(Config), C.c_int]
        L.orc_sizeof_robot.restype = C.c_int
        L.orc_batch_init.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Config)] + [C.c_void_p] * 5
        L.orc_batch_step.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        L.orc_batch_publisher.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_double]
        L.orc_batch_velocity_cmd.argtypes = [C.c_void_p, C.c_int64, _fp]
        L.orc_batch_position_cmd.argtypes = [C.c_void_p, C.c_int64, _fp]
        L.orc_batch_effort_cmd.argtypes = [C.c_void_p, C.c_int64, _dp]
        for f in (L.orc_robot_velocity_cmd, L.orc_robot_position_cmd, L.orc_robot_effort_cmd):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_batch_platform_state.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
        L.orc_batch_joint_states.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp]
        L.orc_batch_ik.argtypes = [C.POINTER(Config), C.c_int64, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_home_lengths.argtypes = [C.POINTER(Config), _dp]
        L.orc_time_double.argtypes = [C.c_int32, C.c_int32]
        L.orc_time_double.restype = C.c_double
        L.orc_batch_last_outputs.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp]
        L.orc_batch_pid_terms.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.orc_legs_mass_matrix.argtypes = [C.c_void_p, _dp]
        L.orc_legs_kinetic_energy.argtypes = [C.c_void_p]; L.orc_legs_kinetic_energy.restype = C.c_double
        L.orc_legs_potential_energy.argtypes = [C.c_void_p]; L.orc_legs_potential_energy.restype = C.c_double
        L.orc_legs_joint_rates.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_batch_targets.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp]
        L.orc_pid_new.argtypes = [C.POINTER(PidParams), C.c_int]; L.orc_pid_new.restype = C.c_void_p
        L.orc_pid_free.argtypes = [C.c_void_p]
        L.orc_pid_reset.argtypes = [C.c_void_p]
        L.orc_pid_update.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]; L.orc_pid_update.restype = C.c_double
        L.orc_pid_derive.argtypes = [C.c_void_p, C.c_double, C.c_double]; L.orc_pid_derive.restype = C.c_double
        L.orc_pid_derive_abs.argtypes = [C.c_void_p, C.c_double, C.c_double]; L.orc_pid_derive_abs.restype = C.c_double
        L.orc_pid_get.argtypes = [C.c_void_p, _dp]
        L.orc_cable_new.argtypes = [C.POINTER(Config)]; L.orc_cable_new.restype = C.c_void_p
        L.orc_cable_free.argtypes = [C.c_void_p]
        L.orc_cable_mode.argtypes = [C.c_void_p]; L.orc_cable_mode.restype = C.c_int
        L.orc_cable_last_position.argtypes = [C.c_void_p]; L.orc_cable_last_position.restype = C.c_double
        L.orc_cable_set_position_target.argtypes = [C.c_void_p, C.c_double]
        L.orc_cable_set_velocity_target.argtypes = [C.c_void_p, C.c_double]
        L.orc_cable_set_force.argtypes = [C.c_void_p, C.c_double]
        L.orc_cable_update.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double]; L.orc_cable_update.restype = C.c_double
        _lib = L
    return _lib


def ref_available() -> bool:
    build()
    return os.path.exists(os.path.join(HERE, "_ref", "libcdpr_ref.so"))


def ref():
    """The reference's own Pid.cpp/JointForceCalculator.cpp behind the shim harness (oracle L0)."""
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(os.path.join(HERE, "_ref", "libcdpr_ref.so"))
        R.ref_plugin_create.argtypes = [C.c_int, C.POINTER(PidParams), C.POINTER(PidParams), C.c_double]
        R.ref_plugin_create.restype = C.c_void_p
        R.ref_plugin_destroy.argtypes = [C.c_void_p]
        R.ref_plugin_set_time.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        R.ref_plugin_set_joint.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        R.ref_plugin_velocity_cmd.argtypes = [C.c_void_p, _fp]
        R.ref_plugin_position_cmd.argtypes = [C.c_void_p, _fp]
        R.ref_plugin_effort_cmd.argtypes = [C.c_void_p, _dp]
        R.ref_plugin_update_cable.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        R.ref_plugin_update_cable.restype = C.c_double
        R.ref_pid_create.argtypes = [C.POINTER(PidParams)]
        R.ref_pid_create.restype = C.c_void_p
        R.ref_pid_destroy.argtypes = [C.c_void_p]
        R.ref_pid_reset.argtypes = [C.c_void_p]
        R.ref_pid_update.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        R.ref_pid_update.restype = C.c_double
        R.ref_pid_derive.argtypes = [C.c_void_p, C.c_double, C.c_double]
        R.ref_pid_derive.restype = C.c_double
        R.ref_pid_update_terms.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
        R.ref_pid_update_terms.restype = C.c_double
        R.ref_batch_step.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        _ref = R
    return _ref


def default_config(n_cables: int = 4) -> Config:
    cfg = Config()
    lib().orc_config_default(C.byref(cfg), n_cables)
    return cfg


class Batch:
    """n oracle robots in one flat buffer; every robot starts in the plugin's post-Load state."""

    def __init__(self, cfg: Config, n: int, pose7=None, twist6=None, amp=None, freq=None, phase=None):
        L = lib()
        self.cfg, self.n, self.nc = cfg, int(n), int(cfg.n_cables)
        self._buf = np.zeros(self.n * L.orc_sizeof_robot(), dtype=np.uint8)
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None
                      for a in (pose7, twist6, amp, freq, phase)]
        L.orc_batch_init(self.ptr, self.n, C.byref(cfg), *[_opt(a) for a in self._keep])

    @property
    def ptr(self):
        return self._buf.ctypes.data_as(C.c_void_p)

    def step(self, k: int = 1, threads: int = 0):
        lib().orc_batch_step(self.ptr, self.n, int(k), int(threads))

    def step_reference_forcelaw(self, k: int = 1, threads: int = 0):
        """Same reduced model, but the force law is the reference's own compiled code (L0).
        Only valid from the post-Load state (a fresh reference plugin is created per call)."""
        ref().ref_batch_step(self.ptr, self.n, int(k), int(threads))

    def publisher(self, shape: int, publish_hz: float):
        """Wave form (0 sinevelocitytest, 1 squarevelocitytest) and rate of the command publisher of every robot."""
        lib().orc_batch_publisher(self.ptr, self.n, int(shape), float(publish_hz))

    def velocity_cmd(self, axes):
        lib().orc_batch_velocity_cmd(self.ptr, self.n, np.ascontiguousarray(axes, dtype=np.float32).reshape(self.n, self.nc))

    def position_cmd(self, axes):
        lib().orc_batch_position_cmd(self.ptr, self.n, np.ascontiguousarray(axes, dtype=np.float32).reshape(self.n, self.nc))

    def effort_cmd(self, force):
        lib().orc_batch_effort_cmd(self.ptr, self.n, np.ascontiguousarray(force, dtype=np.float64).reshape(self.n, self.nc))

    def _robot_ptr(self, i: int):
        return C.c_void_p(self._buf.ctypes.data + i * lib().orc_sizeof_robot())

    def velocity_cmd_masked(self, axes, mask):
        """The message reaches only robots with mask[i] set (each plugin instance latches its own, CdprGazeboPlugin.cpp:67-74)."""
        axes = np.ascontiguousarray(axes, dtype=np.float32).reshape(self.n, self.nc)
        for i in np.flatnonzero(np.asarray(mask)):
            lib().orc_robot_velocity_cmd(self._robot_ptr(int(i)), axes[i].ctypes.data_as(C.c_void_p), self.nc)

    def position_cmd_masked(self, axes, mask):
        axes = np.ascontiguousarray(axes, dtype=np.float32).reshape(self.n, self.nc)
        for i in np.flatnonzero(np.asarray(mask)):
            lib().orc_robot_position_cmd(self._robot_ptr(int(i)), axes[i].ctypes.data_as(C.c_void_p), self.nc)

    def effort_cmd_masked(self, force, mask):
        force = np.ascontiguousarray(force, dtype=np.float64).reshape(self.n, self.nc)
        for i in np.flatnonzero(np.asarray(mask)):
            lib().orc_robot_effort_cmd(self._robot_ptr(int(i)), force[i].ctypes.data_as(C.c_void_p), self.nc)

    def legs_mass_matrix(self, i: int = 0):
        M = np.empty((6, 6)); lib().orc_legs_mass_matrix(self._robot_ptr(i), M); return M

    def legs_energy(self, i: int = 0):
        """(kinetic, potential) energy of platform + leg links of robot i (leg model)."""
        return lib().orc_legs_kinetic_energy(self._robot_ptr(i)), lib().orc_legs_potential_energy(self._robot_ptr(i))

    def legs_joint_rates(self, leg: int, i: int = 0):
        out = np.empty(5); lib().orc_legs_joint_rates(self._robot_ptr(i), leg, out); return out

    def platform_state(self):
        pose = np.empty((self.n, 7)); twist = np.empty((self.n, 6))
        lib().orc_batch_platform_state(self.ptr, self.n, pose, twist)
        return pose, twist

    def joint_states(self):
        pos = np.empty((self.n, self.nc)); vel = np.empty((self.n, self.nc)); eff = np.empty((self.n, self.nc))
        lib().orc_batch_joint_states(self.ptr, self.n, pos, vel, eff)
        return pos, vel, eff

    def last_outputs(self):
        """(joint_pos, joint_vel, pid_force, effort) as seen by the plugin in the LAST update() call."""
        out = [np.empty((self.n, self.nc)) for _ in range(4)]
        lib().orc_batch_last_outputs(self.ptr, self.n, *out)
        return out

    def targets(self):
        """(velocity_target, position_target, mode) per cable, as latched in the force calculators."""
        out = [np.empty((self.n, self.nc)) for _ in range(3)]
        lib().orc_batch_targets(self.ptr, self.n, *out)
        return out

    def pid_terms(self):
        """[n][nc][2 pids: vel,pos][p_term, i_term(pre-clamp), d_term, i_err, cmd, d_err]"""
        out = np.empty((self.n, self.nc, 2, 6))
        lib().orc_batch_pid_terms(self.ptr, self.n, out)
        return out


def ik(cfg: Config, pose7, twist6, threads: int = 0):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64); twist6 = np.ascontiguousarray(twist6, dtype=np.float64)
    n, nc = pose7.shape[0], int(cfg.n_cables)
    ln = np.empty((n, nc)); lr = np.empty((n, nc)); w = np.empty((n, nc, 6))
    lib().orc_batch_ik(C.byref(cfg), n, pose7, twist6, ln, lr, w, threads)
    return ln, lr, w


def home_lengths(cfg: Config):
    out = np.zeros(MAX_CABLES)
    lib().orc_home_lengths(C.byref(cfg), out)
    return out[: cfg.n_cables].copy()
