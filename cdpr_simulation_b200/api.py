"""ctypes binding of libcdpr_b200.so -- the host-side mirror of the reference plugin's interface
for the hot path (cable commands in, joint states and platform pose/twist out).

Names follow the reference: topics `jointVelocities` / `jointPositions` (CdprGazeboPlugin.h:24-25)
become `set_velocity_cmd` / `set_position_cmd`; `jointStates` / `platformPose` (:26,28) become
`joint_states()` / `platform_state()`.  All arithmetic happens in the CUDA library; this module
only moves pointers.  There is no CPU fallback: importing works anywhere, creating a batch needs
a B200 and the built extension.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MAX_CABLES = 8
MODE_FORCE, MODE_POSITION, MODE_VELOCITY = 0, 1, 2
OPT_DTERM_FIR, OPT_KERNEL_TIMING, OPT_INDEPENDENT, OPT_PUBLISHER_SHAPE = 1, 2, 3, 4
PUBLISHER_SINE, PUBLISHER_SQUARE_VELOCITY = 0, 1

OK = 0
ERR_BAD_ARG, ERR_BAD_CABLE_COUNT, ERR_BAD_LENGTH, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7

# every symbol include/cdpr_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "cdpr_config_default", "cdpr_create", "cdpr_destroy", "cdpr_reset", "cdpr_last_error", "cdpr_set_stream", "cdpr_synchronize", "cdpr_set_async",
    "cdpr_set_option", "cdpr_set_velocity_cmd", "cdpr_set_position_cmd", "cdpr_set_effort_cmd", "cdpr_set_sine_cmd",
    "cdpr_set_velocity_cmd_masked", "cdpr_set_position_cmd_masked", "cdpr_set_effort_cmd_masked", "cdpr_get_modes",
    "cdpr_step", "cdpr_update", "cdpr_step_count", "cdpr_sim_time",
    "cdpr_get_joint_states", "cdpr_get_platform_state", "cdpr_set_platform_state", "cdpr_get_pid_state", "cdpr_get_pid_terms",
    "cdpr_state_bytes", "cdpr_get_state", "cdpr_set_state",
    "cdpr_set_snapshots", "cdpr_set_snapshot_peers", "cdpr_set_snapshot_multicast", "cdpr_snapshot_count",
    "cdpr_ik", "cdpr_ik_device", "cdpr_rollout",
    "cdpr_comm_create", "cdpr_comm_destroy", "cdpr_comm_size", "cdpr_comm_last_error", "cdpr_comm_attach_gather", "cdpr_comm_gather_buffer",
    "cdpr_comm_step", "cdpr_comm_allreduce",
    "cdpr_dterm_weights", "cdpr_padded_instances", "cdpr_device_platform_state",
    "cdpr_measure_fp64_tflops", "cdpr_last_kernel_ms", "cdpr_launch_count", "cdpr_kernel_variant", "cdpr_kernel_detail",
]


class PidParams(C.Structure):
    """gazebo::common::Pid::PidParameters (Pid.h:70-81)."""
    _fields_ = [
        ("forward_gain", C.c_double), ("p_gain", C.c_double), ("i_gain", C.c_double), ("d_gain", C.c_double),
        ("d_degree", C.c_int32), ("d_buffer_length", C.c_int32),
        ("i_limit", C.c_double), ("cmd_limit", C.c_double),
        ("p_cutoff", C.c_double), ("p_quality", C.c_double), ("p_cascade", C.c_int32),
        ("d_cutoff", C.c_double), ("d_quality", C.c_double), ("d_cascade", C.c_int32),
    ]


class Config(C.Structure):
    """cdpr_config of include/cdpr_b200.h."""
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
        ("sine_publish_hz", C.c_double),
    ]


class CdprError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cdpr error {code}: {msg}")
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """dlopen the in-tree extension; raises if it has not been built (never falls back to anything)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        raise ImportError(f"{_build.LIB} is missing: run `python -m cdpr_simulation_b200.build` (nvcc, sm_100a)")
    L = C.CDLL(_build.LIB)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    L.cdpr_config_default.argtypes = [C.POINTER(Config), C.c_int]
    L.cdpr_create.argtypes = [C.POINTER(Config), i64, C.c_int, C.POINTER(vp)]
    L.cdpr_destroy.argtypes = [vp]
    L.cdpr_reset.argtypes = [vp]
    L.cdpr_last_error.argtypes = [vp]; L.cdpr_last_error.restype = C.c_char_p
    L.cdpr_set_stream.argtypes = [vp, vp]
    L.cdpr_synchronize.argtypes = [vp]
    L.cdpr_set_async.argtypes = [vp, C.c_int]
    L.cdpr_set_option.argtypes = [vp, C.c_int, i64]
    for f in (L.cdpr_set_velocity_cmd, L.cdpr_set_position_cmd, L.cdpr_set_effort_cmd):
        f.argtypes = [vp, vp, i64, C.c_int]
    for f in (L.cdpr_set_velocity_cmd_masked, L.cdpr_set_position_cmd_masked, L.cdpr_set_effort_cmd_masked):
        f.argtypes = [vp, vp, vp, i64, C.c_int]
    L.cdpr_get_modes.argtypes = [vp, vp]
    L.cdpr_set_sine_cmd.argtypes = [vp, vp, vp, vp, i64]
    L.cdpr_step.argtypes = [vp, i64]
    L.cdpr_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.cdpr_step_count.argtypes = [vp]; L.cdpr_step_count.restype = i64
    L.cdpr_sim_time.argtypes = [vp]; L.cdpr_sim_time.restype = dbl
    L.cdpr_get_joint_states.argtypes = [vp, vp, vp, vp]
    L.cdpr_get_platform_state.argtypes = [vp, vp, vp]
    L.cdpr_set_platform_state.argtypes = [vp, vp, vp]
    L.cdpr_get_pid_state.argtypes = [vp, vp]
    L.cdpr_get_pid_terms.argtypes = [vp, vp]
    L.cdpr_state_bytes.argtypes = [vp]; L.cdpr_state_bytes.restype = C.c_size_t
    L.cdpr_get_state.argtypes = [vp, vp, C.c_size_t]
    L.cdpr_set_state.argtypes = [vp, vp, C.c_size_t]
    L.cdpr_set_snapshots.argtypes = [vp, i64, vp, i64]
    L.cdpr_set_snapshot_peers.argtypes = [vp, i64, C.POINTER(vp), C.c_int, i64, i64, i64]
    L.cdpr_set_snapshot_multicast.argtypes = [vp, i64, vp, i64, i64, i64]
    L.cdpr_snapshot_count.argtypes = [vp]; L.cdpr_snapshot_count.restype = i64
    L.cdpr_ik.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.cdpr_ik_device.argtypes = [vp, i64, vp, vp]
    L.cdpr_rollout.argtypes = [vp, i64, i64, vp, vp, vp, i64, i64, C.POINTER(dbl * 3), dbl, vp, vp]
    L.cdpr_comm_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.cdpr_comm_destroy.argtypes = [vp]
    L.cdpr_comm_size.argtypes = [vp]
    L.cdpr_comm_last_error.argtypes = [vp]; L.cdpr_comm_last_error.restype = C.c_char_p
    L.cdpr_comm_attach_gather.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64]
    L.cdpr_comm_gather_buffer.argtypes = [vp, C.c_int]; L.cdpr_comm_gather_buffer.restype = vp
    L.cdpr_comm_step.argtypes = [vp, C.POINTER(vp), i64]
    L.cdpr_comm_allreduce.argtypes = [vp, C.POINTER(vp), i64]
    L.cdpr_dterm_weights.argtypes = [C.POINTER(PidParams), dbl, vp, vp, C.POINTER(C.c_int)]
    L.cdpr_padded_instances.argtypes = [vp]; L.cdpr_padded_instances.restype = i64
    L.cdpr_device_platform_state.argtypes = [vp]; L.cdpr_device_platform_state.restype = vp
    L.cdpr_measure_fp64_tflops.argtypes = [C.c_int, C.c_int]; L.cdpr_measure_fp64_tflops.restype = dbl
    L.cdpr_last_kernel_ms.argtypes = [vp]; L.cdpr_last_kernel_ms.restype = C.c_float
    L.cdpr_launch_count.argtypes = [vp]; L.cdpr_launch_count.restype = i64
    L.cdpr_kernel_variant.argtypes = [vp]; L.cdpr_kernel_variant.restype = C.c_char_p
    L.cdpr_kernel_detail.argtypes = [vp]; L.cdpr_kernel_detail.restype = C.c_char_p
    _lib = L
    return L


def default_config(n_cables: int = 4) -> Config:
    """Reference constants (sdf/cube.sdf, launch/cdpr_gazebo.launch); 8 = synthetic 8-cable extension."""
    cfg = Config()
    rc = load().cdpr_config_default(C.byref(cfg), n_cables)
    if rc != OK:
        raise CdprError(rc, "cdpr_config_default")
    return cfg


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


class CdprBatch:
    """N independent robots stepped by the CUDA library; one instance per GPU / process."""

    def __init__(self, cfg: Config | None = None, n_instances: int = 1, device: int = 0, n_cables: int = 4):
        self._L = load()
        self.cfg = cfg if cfg is not None else default_config(n_cables)
        self.n, self.nc, self.device = int(n_instances), int(self.cfg.n_cables), int(device)
        self._h = C.c_void_p()
        rc = self._L.cdpr_create(C.byref(self.cfg), self.n, self.device, C.byref(self._h))
        if rc != OK:
            msg = self._L.cdpr_last_error(None).decode()
            self._h = C.c_void_p()
            raise CdprError(rc, msg)

    # -- plumbing --------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != OK:
            raise CdprError(rc, self._L.cdpr_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.cdpr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        """Back to the post-Load state (world reset)."""
        self._ck(self._L.cdpr_reset(self._h))

    def set_stream(self, cuda_stream: int):
        self._ck(self._L.cdpr_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_async(self, on: bool = True):
        """Host-buffer calls only enqueue; keep the (pinned) buffers alive until synchronize()."""
        self._ck(self._L.cdpr_set_async(self._h, 1 if on else 0))

    def set_option(self, option: int, value: int):
        self._ck(self._L.cdpr_set_option(self._h, int(option), int(value)))

    def synchronize(self):
        self._ck(self._L.cdpr_synchronize(self._h))

    # -- commands: topics jointVelocities / jointPositions, JointForceCalculator::setForce -----
    def _cmd(self, fn, axes, mask, dtype):
        axes = np.ascontiguousarray(axes, dtype=dtype)
        n_axes = axes.shape[-1] if axes.ndim > 1 else axes.size // max(self.n, 1)
        if mask is not None:
            mask = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
            if mask.shape != (self.n,):
                raise ValueError("mask must have one entry per instance")
        self._ck(fn(self._h, _ptr(axes), _ptr(mask), axes.size // max(n_axes, 1), n_axes))

    def set_velocity_cmd(self, axes, mask=None):
        """Topic jointVelocities; mask (one flag per robot) addresses a subset (independent robots only)."""
        self._cmd(self._L.cdpr_set_velocity_cmd_masked, axes, mask, np.float32)

    def set_position_cmd(self, axes, mask=None):
        self._cmd(self._L.cdpr_set_position_cmd_masked, axes, mask, np.float32)

    def set_effort_cmd(self, force, mask=None):
        self._cmd(self._L.cdpr_set_effort_cmd_masked, force, mask, np.float64)

    def set_independent(self, on: bool = True):
        """Every robot gets its own mode and command latch (CDPR_OPT_INDEPENDENT); before the first step."""
        self.set_option(OPT_INDEPENDENT, 1 if on else 0)

    def modes(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int32)
        self._ck(self._L.cdpr_get_modes(self._h, _ptr(out)))
        return out

    def set_square_velocity_cmd(self, amp, freq=None, phase=None):
        """squarevelocitytest.cpp run inside the kernel (+-amp outside the dead band |sin| >= sqrt(0.5), else 0; the driver
        publishes at 10 Hz: Config.sine_publish_hz); per-instance amp / freq / phase."""
        self.set_option(OPT_PUBLISHER_SHAPE, PUBLISHER_SQUARE_VELOCITY)
        self.set_sine_cmd(amp, freq, phase, _keep_shape=True)

    def set_sine_cmd(self, amp, freq=None, phase=None, _keep_shape=False):
        """sinevelocitytest.cpp run inside the kernel; per-instance amp / freq / phase."""
        if not _keep_shape:
            self.set_option(OPT_PUBLISHER_SHAPE, PUBLISHER_SINE)
        if amp is None:
            self._ck(self._L.cdpr_set_sine_cmd(self._h, None, None, None, 0))
            return
        def col(a):
            if a is None:
                return None
            a = np.asarray(a)
            if a.dtype == np.float64 and a.shape == (self.n,) and a.flags.c_contiguous:
                return a                      # used in place (async mode: the caller keeps it alive)
            return np.broadcast_to(_f64(a), (self.n,)).copy()
        cols = [col(amp), col(freq), col(phase)]
        self._keep = cols                     # temporaries must outlive an enqueued copy
        self._ck(self._L.cdpr_set_sine_cmd(self._h, _ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2]), self.n))

    # -- stepping ----------------------------------------------------------------------------
    def step(self, k_steps: int = 1):
        self._ck(self._L.cdpr_step(self._h, int(k_steps)))

    def update(self, vel_axes=None, pos_axes=None):
        """One plugin update (CdprGazeboPlugin::update): latch the messages given, one physics step, and the plugin's own
        publish: (position, velocity, effort, pose7, twist6) with position / velocity / pose / twist as read at this update
        (before the step integrates) and the effort applied in it.  Returns views of buffers reused by the next call."""
        if not hasattr(self, "_upd"):
            self._upd = (np.empty((self.n, self.nc)), np.empty((self.n, self.nc)), np.empty((self.n, self.nc)), np.empty((self.n, 7)), np.empty((self.n, 6)))
            self._upd_ptr = [_ptr(a) for a in self._upd]
        v = None if vel_axes is None else np.ascontiguousarray(vel_axes, dtype=np.float32)
        p = None if pos_axes is None else np.ascontiguousarray(pos_axes, dtype=np.float32)
        if (v is not None and v.size != self.n * self.nc) or (p is not None and p.size != self.n * self.nc):
            raise CdprError(ERR_BAD_LENGTH, "command length != instances x cable count: dropped")
        self._ck(self._L.cdpr_update(self._h, _ptr(v), _ptr(p), *self._upd_ptr))
        return self._upd

    @property
    def step_count(self) -> int:
        return int(self._L.cdpr_step_count(self._h))

    @property
    def sim_time(self) -> float:
        return float(self._L.cdpr_sim_time(self._h))

    # -- outputs: topics jointStates / platformPose ----------------------------------------
    def joint_states(self, out=None):
        pos, vel, eff = out if out is not None else (np.empty((self.n, self.nc)) for _ in range(3))
        self._ck(self._L.cdpr_get_joint_states(self._h, _ptr(pos), _ptr(vel), _ptr(eff)))
        return pos, vel, eff

    def platform_state(self, out=None):
        pose, twist = out if out is not None else (np.empty((self.n, 7)), np.empty((self.n, 6)))
        self._ck(self._L.cdpr_get_platform_state(self._h, _ptr(pose), _ptr(twist)))
        return pose, twist

    def set_platform_state(self, pose7=None, twist6=None):
        pose7 = None if pose7 is None else _f64(pose7, (self.n, 7))
        twist6 = None if twist6 is None else _f64(twist6, (self.n, 6))
        self._ck(self._L.cdpr_set_platform_state(self._h, _ptr(pose7), _ptr(twist6)))

    def pid_state(self):
        out = np.empty((self.n, self.nc, 6))
        self._ck(self._L.cdpr_get_pid_state(self._h, _ptr(out)))
        return out

    def pid_terms(self):
        """Topic `pid` for every cable: [n][nc][5] = pTerm, pre-clamp iTerm, dTerm, desired, applied force (last step)."""
        out = np.empty((self.n, self.nc, 5))
        self._ck(self._L.cdpr_get_pid_terms(self._h, _ptr(out)))
        return out

    # -- checkpoint --------------------------------------------------------------------------
    def get_state(self) -> np.ndarray:
        blob = np.empty(self._L.cdpr_state_bytes(self._h), dtype=np.uint8)
        self._ck(self._L.cdpr_get_state(self._h, _ptr(blob), blob.size))
        return blob

    def set_state(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._ck(self._L.cdpr_set_state(self._h, _ptr(blob), blob.size))

    # -- snapshots ---------------------------------------------------------------------------
    def set_snapshots(self, every: int, dev_ptr: int | None, capacity: int):
        self._ck(self._L.cdpr_set_snapshots(self._h, int(every), C.c_void_p(dev_ptr) if dev_ptr else None, int(capacity)))

    def set_snapshot_peers(self, every: int, peer_ptrs, instance_offset: int, total_instances: int, capacity: int):
        """Fused all-gather: snapshots go straight into every rank's gather buffer (NVLink-mapped pointers)."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._ck(self._L.cdpr_set_snapshot_peers(self._h, int(every), arr, len(peer_ptrs), int(instance_offset), int(total_instances), int(capacity)))

    def set_snapshot_multicast(self, every: int, multicast_ptr: int, instance_offset: int, total_instances: int, capacity: int):
        """Fused all-gather through one NVLS multicast address (multimem.st)."""
        self._ck(self._L.cdpr_set_snapshot_multicast(self._h, int(every), C.c_void_p(int(multicast_ptr)), int(instance_offset), int(total_instances), int(capacity)))

    @property
    def snapshot_count(self) -> int:
        return int(self._L.cdpr_snapshot_count(self._h))

    # -- kinematics only ---------------------------------------------------------------------
    def ik(self, pose7, twist6):
        pose7 = _f64(pose7); twist6 = _f64(twist6)
        n = pose7.shape[0]
        ln = np.empty((n, self.nc)); lr = np.empty((n, self.nc)); w = np.empty((n, self.nc, 6))
        self._ck(self._L.cdpr_ik(self._h, n, _ptr(pose7), _ptr(twist6), _ptr(ln), _ptr(lr), _ptr(w)))
        return ln, lr, w

    def ik_device(self, n: int, dev_state13: int, dev_out: int):
        self._ck(self._L.cdpr_ik_device(self._h, int(n), C.c_void_p(dev_state13), C.c_void_p(dev_out)))

    # -- rollouts ----------------------------------------------------------------------------
    def rollout(self, n_robots: int, n_seq: int, cmds, steps_per_cmd: int, target_pos, lam: float, pose7=None, twist6=None,
                dev_cost_seq: int | None = None, want_host_cost: bool = True):
        cmds = np.ascontiguousarray(cmds, dtype=np.float32).reshape(n_seq, -1, self.nc)
        pose7 = None if pose7 is None else _f64(pose7, (n_robots, 7))
        twist6 = None if twist6 is None else _f64(twist6, (n_robots, 6))
        tgt = (C.c_double * 3)(*[float(x) for x in target_pos])
        cost = np.empty(self.n) if want_host_cost else None
        self._ck(self._L.cdpr_rollout(self._h, n_robots, n_seq, _ptr(pose7), _ptr(twist6), _ptr(cmds), cmds.shape[1], int(steps_per_cmd),
                                      C.byref(tgt), float(lam), C.c_void_p(dev_cost_seq) if dev_cost_seq else None, _ptr(cost)))
        return cost

    # -- raw device access / measurement -------------------------------------------------------
    @property
    def padded_instances(self) -> int:
        return int(self._L.cdpr_padded_instances(self._h))

    @property
    def device_platform_state(self) -> int:
        return int(self._L.cdpr_device_platform_state(self._h))

    @property
    def last_kernel_ms(self) -> float:
        return float(self._L.cdpr_last_kernel_ms(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.cdpr_launch_count(self._h))

    @property
    def kernel_variant(self) -> str:
        return self._L.cdpr_kernel_variant(self._h).decode()

    @property
    def kernel_detail(self) -> str:
        """Which instance of the variant runs (tests, tuning logs)."""
        return self._L.cdpr_kernel_detail(self._h).decode()


def dterm_weights(pid: PidParams, dt: float):
    """(fir[len], quadratic[3], is_quadratic) of the D-term for uniform time stamps; host only."""
    fir = np.zeros(int(pid.d_buffer_length)); quad = np.zeros(3); flag = C.c_int(0)
    rc = load().cdpr_dterm_weights(C.byref(pid), float(dt), _ptr(fir), _ptr(quad), C.byref(flag))
    if rc != OK:
        raise CdprError(rc, "cdpr_dterm_weights")
    return fir, quad, bool(flag.value)


def measure_fp64_tflops(device: int = 0, iters: int = 8192) -> float:
    return float(load().cdpr_measure_fp64_tflops(device, iters))
