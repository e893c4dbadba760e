"""CPU-only: (1) the oracle and the product's default constants against the reference's numeric literals
(tests/golden/reference_constants.json, extracted by tests/golden/make_golden.py from sdf/cube.sdf, the launch file
and the driver sources); (2) the C-ABI library loads and exports every symbol include/cdpr_b200.h declares."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import cdpr_simulation_b200 as cb
from oracle import binding as ob
from helpers import to_oracle_config, PID_FIELDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_constants.json")))


def test_wire_count_and_topology():
    assert G["wire_count"] == 4 and len(G["cables"]) == 4            # CdprGazeboPlugin.h:20
    assert G["n_links"] == 22 and G["n_joints"] == 24
    for i, c in enumerate(G["cables"]):
        assert c["child"] == f"cable{i}" and c["parent"] == f"virt_Y{i}"


@pytest.mark.parametrize("make", ["oracle", "product"])
def test_default_config_matches_reference_literals(make, built_lib):
    cfg = ob.default_config(4) if make == "oracle" else cb.default_config(4)
    assert cfg.n_cables == G["wire_count"]
    pp = np.array(G["platform_pose"])
    assert list(cfg.home_pos) == list(pp[:3]) and np.all(pp[3:] == 0) and list(cfg.home_quat) == [1, 0, 0, 0]
    assert cfg.mass == G["platform_mass"] and list(cfg.inertia) == G["platform_inertia"]
    for i, c in enumerate(G["cables"]):
        assert list(cfg.frame_anchor[i]) == c["frame_anchor_link_pose"][:3]
        b_world = np.array(c["platform_anchor_link_pose"][:3])
        assert np.allclose(np.array(list(cfg.platform_anchor[i])), b_world - pp[:3], atol=1e-15)
        assert cfg.cable_damping == c["damping"] and cfg.effort_limit == c["effort"]
    lp = G["launch_params"]
    v, p = cfg.vel_pid, cfg.pos_pid
    assert (v.forward_gain, v.p_gain, v.i_gain, v.d_gain) == (lp["velocityControllerForward"], lp["velocityControllerP"], lp["velocityControllerI"], lp["velocityControllerD"])
    assert (v.d_degree, v.d_buffer_length, v.i_limit, v.cmd_limit) == (lp["velocityControllerDdegree"], lp["velocityControllerDbuffer"], lp["velocityControllerMaxI"], lp["velocityControllerMaxCmd"])
    assert (v.p_cutoff, v.p_quality, v.p_cascade, v.d_cutoff, v.d_quality, v.d_cascade) == (
        lp["velocityControllerPcutoff"], lp["velocityControllerPquality"], lp["velocityControllerPcascade"],
        lp["velocityControllerDcutoff"], lp["velocityControllerDquality"], lp["velocityControllerDcascade"])
    assert (p.forward_gain, p.p_gain, p.i_gain, p.d_gain) == (0.0, lp["positionControllerP"], lp["positionControllerI"], lp["positionControllerD"])
    assert (p.d_degree, p.d_buffer_length, p.i_limit, p.cmd_limit) == (lp["positionControllerDdegree"], lp["positionControllerDbuffer"], lp["positionControllerMaxI"], lp["positionControllerMaxCmd"])
    assert p.p_cascade == 0 and p.d_cascade == 0                      # CdprGazeboPlugin.cpp:133
    assert cfg.velocity_epsilon == lp["velocityEpsilon"]
    if make == "product":
        assert cfg.sine_publish_hz == G["drivers"]["sinevelocitytest"]["cPublishFrequency"]


def test_product_and_oracle_defaults_agree(built_lib):
    for nc in (4, 8):
        a, b = to_oracle_config(cb.default_config(nc)), ob.default_config(nc)
        assert bytes(a) == bytes(b)


def rpy_from_matrix(R):  # static x-y-z ('sxyz') Euler angles, as transformations.euler_from_matrix
    cy = np.hypot(R[0, 0], R[1, 0])
    return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arctan2(-R[2, 0], cy), np.arctan2(R[1, 0], R[0, 0])])


def axis_angle_matrix(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def test_home_pose_ik_against_sdf_leg_literals():
    """Oracle IK at the home pose vs the leg geometry baked into cube.sdf (6 significant digits):
    prismatic axis = 0.15 * u_i (cube.sdf:434), cable link pose = pp - (l/2) * unit(pp - fp) and the leg's Euler
    angles (gen_cdpr.py:113-125; cube.sdf:344), slider limits = +-l/2 (cube.sdf:436-437)."""
    cfg = ob.default_config(4)
    pose7 = np.array([[0, 0, 0.3, 0, 0, 0, 1.0]]); twist6 = np.zeros((1, 6))
    ln, lr, w = ob.ik(cfg, pose7, twist6)
    L0 = ob.home_lengths(cfg)
    assert np.allclose(ln[0], L0, atol=0) and np.allclose(L0, 0.4855924, atol=5e-8) and np.all(lr == 0)
    half_l = 0.5 * np.linalg.norm([0.6, 0.6, 0.6])
    for i, c in enumerate(G["cables"]):
        u = w[0, i, :3]
        axis = np.array(c["prismatic_axis"])
        assert abs(np.linalg.norm(axis) - 0.15) < 1e-6
        assert np.allclose(u, axis / np.linalg.norm(axis), atol=2e-6)     # +q shortens the cable: axis = platform -> frame
        assert abs(c["upper"] - half_l) < 1e-8 and c["lower"] == -c["upper"]
        pp = np.array(c["platform_anchor_link_pose"][:3])
        assert np.allclose(np.array(c["cable_link_pose"][:3]), pp + half_l * u, atol=1e-6)
        u_fp = -u
        z = np.array([0.0, 0.0, 1.0])
        R = axis_angle_matrix(np.cross(z, u_fp), np.arctan2(np.linalg.norm(np.cross(z, u_fp)), u_fp @ z))
        assert np.allclose(rpy_from_matrix(R), c["cable_link_pose"][3:], atol=1e-6)
        assert np.allclose(R[:, 2], u_fp, atol=1e-12)
    # static equilibrium implied by those literals: 4 equal tensions carry m*g
    tension = G["platform_mass"] * 9.8 / np.sum(w[0, :, 2])
    assert abs(tension - 3.9657) < 1e-4


def test_sine_command_schedule_matches_driver_source():
    """sinevelocitytest.cpp:33-49 restated: float32 axes, accumulated publisher time, 10 physics steps per command."""
    d = G["drivers"]["sinevelocitytest"]
    cfg = ob.default_config(4)
    b = ob.Batch(cfg, 1, amp=[d["cVelocityAmplitude"]], freq=[d["cVelocityFrequency"]], phase=[0.0])
    t, expect = 0.0, []
    for k in range(30):
        expect.append(float(np.float32(d["cVelocityAmplitude"] * np.sin(t * d["cVelocityFrequency"] * 2 * np.pi))))
        t += 1.0 / d["cPublishFrequency"]
    for step in range(300):
        b.step(1)
        vt, _, mode = b.targets()
        assert np.all(vt[0] == expect[step // 10]), step          # same float32-rounded value on all 4 cables
        assert np.all(mode[0] == 2)                               # Velocity mode from the first command on
    assert abs(t - 0.3) < 1e-15


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "cdpr_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cdpr_[a-z0-9_]+)\s*\(", body))
    assert declared == set(cb.EXPORTS), declared ^ set(cb.EXPORTS)
    lib = C.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert C.sizeof(cb.Config) == 8 + 8 * (24 + 24 + 3 + 4 + 1 + 6 + 3 + 3) + 2 * C.sizeof(cb.PidParams) + 8 + 8 + 8 * (4 + 72 + 3) + 8


def test_create_fails_loudly_without_a_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CdprError) as e:
        cb.CdprBatch(cb.default_config(4), 16)
    assert e.value.code == cb.api.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_create_argument_checks(built_lib):
    L = cb.load()
    h = C.c_void_p()
    cfg = cb.default_config(4)
    cfg.n_cables = 9
    assert L.cdpr_create(C.byref(cfg), 4, 0, C.byref(h)) == cb.api.ERR_BAD_CABLE_COUNT   # "invalid joint count"
    cfg = cb.default_config(4)
    cfg.dt = 0.0012345678912
    assert L.cdpr_create(C.byref(cfg), 4, 0, C.byref(h)) == cb.api.ERR_BAD_ARG
    bad = cb.Config()
    assert L.cdpr_config_default(C.byref(bad), 0) == cb.api.ERR_BAD_ARG
    cfg = cb.default_config(4)
    assert L.cdpr_create(C.byref(cfg), 0, 0, C.byref(h)) == cb.api.ERR_BAD_ARG          # empty batch
    assert L.cdpr_create(C.byref(cfg), 1 << 40, 0, C.byref(h)) == cb.api.ERR_BAD_ARG    # beyond the 2^31 index range
    cfg.vel_pid.d_buffer_length = 40
    assert L.cdpr_create(C.byref(cfg), 4, 0, C.byref(h)) == cb.api.ERR_BAD_ARG          # window longer than CDPR_MAX_DBUF
    assert not h.value


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "cdpr_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_roofline_numerator_is_frozen_and_bounds_what_the_kernel_executes(built_lib):
    """bench.py's roofline numerator is the FROZEN algorithmic count (1028 @ NC=8, 636 @ NC=4); the FP64 work
    the built kernel's hot loop really issues (SASS, FMA = 2) is reported beside it and may only be smaller -- so deleting an
    instruction raises `frac` instead of lowering the numerator."""
    import shutil
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from cdpr_simulation_b200 import flops
    assert bench.flops_per_instance_step(8) == 1028 and bench.flops_per_instance_step(4) == 636
    assert flops.frozen_flops_per_instance_step(8, "diag") == 980 and flops.frozen_ik_flops_per_pose(8) == 438
    assert sum(flops.PER_CABLE.values()) == 98 and flops.ik_bytes_per_pose(8) == 616 and flops.ik_bytes_per_pose(4) == 360
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    for nc in (4, 8):
        executed = bench.executed_flops_per_instance_step(nc)
        assert executed is not None and 0 < executed <= bench.flops_per_instance_step(nc), (nc, executed)
    committed = json.load(open(os.path.join(root, "profiles", "hot_loop_flops.json")))
    for nc in (4, 8):   # the committed copy (used where cuobjdump is missing) must describe the library that was just built
        assert committed[str(nc)]["executed_flops"] == bench.executed_flops_per_instance_step(nc)
