"""Synthetic inputs of SURVEY.md 8(d) (C2, C3/C4, C5), numpy only, seeded: the same arrays feed the
CUDA path, the CPU checker in the tests and the CPU baseline in bench.py."""
from __future__ import annotations

import numpy as np

HOME_POS = np.array([0.0, 0.0, 0.3])  # sdf/cube.sdf:310


def _axis_angle_quat(axis, angle):
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    s = np.sin(0.5 * angle)[..., None]
    return np.concatenate([axis * s, np.cos(0.5 * angle)[..., None]], axis=-1)  # x y z w


def c2_poses(n: int, seed: int = 0):
    """C2: poses inside the 0.6 m frame, tilt <= 30 deg, twist U(+-0.1 m/s, +-0.5 rad/s)."""
    g = np.random.default_rng(seed)
    pos = np.stack([g.uniform(-0.15, 0.15, n), g.uniform(-0.15, 0.15, n), g.uniform(0.1, 0.5, n)], axis=1)
    quat = _axis_angle_quat(g.normal(size=(n, 3)), g.uniform(0.0, np.deg2rad(30.0), n))
    twist = np.concatenate([g.uniform(-0.1, 0.1, (n, 3)), g.uniform(-0.5, 0.5, (n, 3))], axis=1)
    return np.ascontiguousarray(np.concatenate([pos, quat], axis=1)), np.ascontiguousarray(twist)


def c3_instances(n: int, seed: int = 1):
    """C3/C4: per-instance sine command (amp, freq, phase) and an initial pose near home (+-1 cm, +-2 deg), at rest."""
    g = np.random.default_rng(seed)
    amp = g.uniform(0.01, 0.06, n)
    freq = g.uniform(0.05, 0.2, n)
    phase = g.uniform(0.0, 2.0 * np.pi, n)
    pos = HOME_POS + g.uniform(-0.01, 0.01, (n, 3))
    quat = _axis_angle_quat(g.normal(size=(n, 3)), g.uniform(-np.deg2rad(2.0), np.deg2rad(2.0), n))
    pose7 = np.ascontiguousarray(np.concatenate([pos, quat], axis=1))
    twist6 = np.zeros((n, 6))
    return amp, freq, phase, pose7, twist6


def c5_rollouts(n_seq: int, n_cmd: int, n_cables: int, seed: int = 2):
    """C5: piecewise-constant float32 velocity commands ~ N(0, 0.03^2) m/s, [n_seq][n_cmd][NC]."""
    g = np.random.default_rng(seed)
    return g.normal(0.0, 0.03, (n_seq, n_cmd, n_cables)).astype(np.float32)


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous instance range of `rank`: global id = offset + local id."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
