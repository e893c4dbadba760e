"""Frozen algorithmic work of one CDPR instance-step -- the roofline NUMERATOR.

FROZEN: these numbers describe the ALGORITHM (the reduced model of SURVEY.md App. C with the window-relative FIR form of
the D-term, which is what the CPU checker restates), not any kernel.  They do not move when the kernel gets cheaper, so
`roofline.frac` rises when instructions are deleted.  What the built kernel really executes is reported beside it as
`roofline.executed_flops` (SASS count of the hot loop, tools/hot_loop_flops.py) and may only ever be SMALLER.

Convention: FMA = 2 flop, add/sub/mul = 1, sqrt = div = 1; compares, selects, moves and float32
conversions = 0.  Per-term derivation, each against the statement of the restated step it counts (`O:` = the CPU
checker's cdpr_*.c line ranges under the repository's test infrastructure; `R:` = /root/reference/src/cdpr_gazebo):

per cable (98)
  kinematics 52   O: kinematics_eval        r = R b              9 mul + 6 add            15
                                            d = a - p - r        6 sub                      6
                                            L, 1/L, u = d/L      3 mul 2 add, sqrt, div, 3 mul  10
                                            r x u                6 mul 3 sub                9
                                            rate = u.v + (r x u).w   6 mul 5 add           11
                                            q = L0 - L           1 sub                      1
  force law 32    O: pid_update, R: Pid.cpp:127-187
                                            f = Kf des, e = des - act, P = Kp e             3
                                            Ierr += dt e, I = Ki Ierr                       3
                                            D = Kd derr, cmd = f + P + I + D                4
                                            derr = FIR over the 11-sample window (uniform stamps; R: Pid.cpp:193-247)
                                                 11 mul + 10 add + span scale              22
                                            clamps / anti-windup: compares; arithmetic only on saturated steps   0
  wrench 14       O: robot_step             tau = eff - c rate   1 mul 1 sub                2
                                            F += tau u, M += tau (r x u)   6 mul 6 add     12
platform (244 general inertia | 196 diagonal inertia), O: robot_step, SURVEY.md App. C.6
  quaternion -> R                                                                          30
  v += h F / m, p += h v                    3 div/mul + 3 mul 3 add ; 3 mul 3 add       10 + 6
  I_w^-1 = R I_b^-1 R^T applied             two 3x3 products + mat-vec                     75   (diagonal I_b: 39)
  gyroscopic torque w x (R I_b R^T w)                                                      57   (diagonal I_b: 45)
  w += h I_w^-1 M                                                                          21
  q += h/2 (0,w) (x) q, renormalise         16 mul 12 add, 4 mul 3 add, sqrt, div, 4 mul   45
"""
from __future__ import annotations

PER_CABLE = {"kinematics": 52, "force_law": 32, "wrench": 14}          # 98
PLATFORM = {"general": 244, "diag": 196}
IK_PER_CABLE, IK_PER_POSE = 51, 30                                      # config 2: kinematics without q = L0 - L; R once


def frozen_flops_per_instance_step(nc: int, inertia: str = "general") -> int:
    """1028 @ NC=8, 636 @ NC=4 (general inertia)."""
    return PLATFORM[inertia] + nc * sum(PER_CABLE.values())


def frozen_ik_flops_per_pose(nc: int) -> int:
    """438 @ NC=8, 234 @ NC=4."""
    return IK_PER_POSE + nc * IK_PER_CABLE


def ik_bytes_per_pose(nc: int) -> int:
    """13 doubles in, (L, dL/dt, W[6]) per cable out = 616 B @ NC=8, 360 B @ NC=4."""
    return 104 + 64 * nc


assert frozen_flops_per_instance_step(8) == 1028 and frozen_flops_per_instance_step(4) == 636
assert frozen_ik_flops_per_pose(8) == 438 and frozen_ik_flops_per_pose(4) == 234
