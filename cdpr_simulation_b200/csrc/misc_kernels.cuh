// misc_kernels.cuh -- K1 (batched inverse kinematics + structure matrix) and the small layout /
// bookkeeping kernels around the step kernels.
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace cdpr {

struct IkArgs {
  RobotConsts rc;
  int nc;
  long long n;
  // SoA path
  const double *state13;  // [13][n]
  double *out;            // [NC][8][n]: L, dL/dt, u xyz, (r x u) xyz
  // AoS (reference layout) path
  const double *pose7, *twist6;      // [n][7] x y z qx qy qz qw ; [n][6]
  double *length, *length_rate, *wmat;  // [n][NC], [n][NC], [n][NC][6]
};

// K1: per pose, the prismatic-joint read-backs of the reference (Joint::Position / GetVelocity,
// JointForceCalculator.cpp:68,76) expressed as cable length and rate, plus the 6 x NC structure
// matrix column W_c = [u_c ; r_c x u_c].  HBM-bound: 104 B in, 64*NC B out.
template <int NC, bool AOS>
__global__ void __launch_bounds__(256) k_ik(const __grid_constant__ IkArgs A) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  FastState S;
  if (AOS) {
    const double *ps = A.pose7 + 7 * i, *tw = A.twist6 + 6 * i;
    S.px = ps[0]; S.py = ps[1]; S.pz = ps[2]; S.qx = ps[3]; S.qy = ps[4]; S.qz = ps[5]; S.qw = ps[6];
    S.vx = tw[0]; S.vy = tw[1]; S.vz = tw[2]; S.wx = tw[3]; S.wy = tw[4]; S.wz = tw[5];
  } else {
    const double *p = A.state13 + i;
    const long long n = A.n;
    S.px = __ldcs(p); S.py = __ldcs(p + n); S.pz = __ldcs(p + 2 * n);
    S.qw = __ldcs(p + 3 * n); S.qx = __ldcs(p + 4 * n); S.qy = __ldcs(p + 5 * n); S.qz = __ldcs(p + 6 * n);
    S.vx = __ldcs(p + 7 * n); S.vy = __ldcs(p + 8 * n); S.vz = __ldcs(p + 9 * n);
    S.wx = __ldcs(p + 10 * n); S.wy = __ldcs(p + 11 * n); S.wz = __ldcs(p + 12 * n);
  }
  const Rot R = make_rot(S);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const double bx = A.rc.b[c][0], by = A.rc.b[c][1], bz = A.rc.b[c][2];
    const double rx = fma(R.r00, bx, fma(R.r01, by, R.r02 * bz));
    const double ry = fma(R.r10, bx, fma(R.r11, by, R.r12 * bz));
    const double rz = fma(R.r20, bx, fma(R.r21, by, R.r22 * bz));
    const double dx = (A.rc.a[c][0] - S.px) - rx, dy = (A.rc.a[c][1] - S.py) - ry, dz = (A.rc.a[c][2] - S.pz) - rz;
    const double l2 = fma(dx, dx, fma(dy, dy, dz * dz));
    const double len = sqrt(l2);
    const double il = 1.0 / len;
    const double ux = dx * il, uy = dy * il, uz = dz * il;
    const double cx = fma(ry, uz, -(rz * uy)), cy = fma(rz, ux, -(rx * uz)), cz = fma(rx, uy, -(ry * ux));
    const double qd = fma(ux, S.vx, fma(uy, S.vy, fma(uz, S.vz, fma(cx, S.wx, fma(cy, S.wy, cz * S.wz)))));
    if (AOS) {
      A.length[i * NC + c] = len;
      A.length_rate[i * NC + c] = -qd;
      double *w = A.wmat + (i * NC + c) * 6;
      w[0] = ux; w[1] = uy; w[2] = uz; w[3] = cx; w[4] = cy; w[5] = cz;
    } else {
      double *o = A.out + (long long)c * 8 * A.n + i;
      const long long n = A.n;
      __stcs(o, len); __stcs(o + n, -qd); __stcs(o + 2 * n, ux); __stcs(o + 3 * n, uy); __stcs(o + 4 * n, uz);
      __stcs(o + 5 * n, cx); __stcs(o + 6 * n, cy); __stcs(o + 7 * n, cz);
    }
  }
}

// K1, device (SoA) path for even n: one thread = (cable c, poses 2j and 2j+1).  Compared with one thread
// per pose this puts NC times as many warps in flight (65,536 poses: 110 warps per SM instead of 14 -- the sweep is a
// 6 us kernel, so memory-level parallelism is what it lives on), every access is a 16-byte double2, and a warp's store
// covers 512 contiguous bytes of one output column.  The 13 state columns are re-read by the NC cable blocks of a pose
// range; they stay in L2 (6.8 MB at 65,536 poses), so DRAM still sees them once -- which is why the host picks this kernel
// only while the state fits comfortably in L2 (<= 2^19 poses; measured: 10.1 us against 15.5 us for one thread per pose at
// 65,536 poses, NC = 8) and the one-thread-per-pose k_ik beyond (96 % of the copy bandwidth at 4.2 M poses, where this
// one drops to 71 % because the re-reads reach DRAM).  Same arithmetic as k_ik.
struct IkOne { double len, rate, ux, uy, uz, cx, cy, cz; };
__device__ __forceinline__ IkOne ik_one(const RobotConsts &rc, int c, const FastState &S) {
  const Rot R = make_rot(S);
  const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
  const double rx = fma(R.r00, bx, fma(R.r01, by, R.r02 * bz));
  const double ry = fma(R.r10, bx, fma(R.r11, by, R.r12 * bz));
  const double rz = fma(R.r20, bx, fma(R.r21, by, R.r22 * bz));
  const double dx = (rc.a[c][0] - S.px) - rx, dy = (rc.a[c][1] - S.py) - ry, dz = (rc.a[c][2] - S.pz) - rz;
  const double l2 = fma(dx, dx, fma(dy, dy, dz * dz));
  IkOne o;
  o.len = sqrt(l2);
  const double il = 1.0 / o.len;
  o.ux = dx * il; o.uy = dy * il; o.uz = dz * il;
  o.cx = fma(ry, o.uz, -(rz * o.uy)); o.cy = fma(rz, o.ux, -(rx * o.uz)); o.cz = fma(rx, o.uy, -(ry * o.ux));
  o.rate = -fma(o.ux, S.vx, fma(o.uy, S.vy, fma(o.uz, S.vz, fma(o.cx, S.wx, fma(o.cy, S.wy, o.cz * S.wz)))));
  return o;
}
__global__ void __launch_bounds__(256) k_ik_pair(const __grid_constant__ IkArgs A) {
  const int c = blockIdx.y;
  const long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= A.n) return;
  const long long n = A.n;
  const double *p = A.state13 + i;
  double2 v[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) v[k] = __ldg(reinterpret_cast<const double2 *>(p + k * n));
  FastState S0, S1;
  S0.px = v[0].x; S0.py = v[1].x; S0.pz = v[2].x; S0.qw = v[3].x; S0.qx = v[4].x; S0.qy = v[5].x; S0.qz = v[6].x;
  S0.vx = v[7].x; S0.vy = v[8].x; S0.vz = v[9].x; S0.wx = v[10].x; S0.wy = v[11].x; S0.wz = v[12].x;
  S1.px = v[0].y; S1.py = v[1].y; S1.pz = v[2].y; S1.qw = v[3].y; S1.qx = v[4].y; S1.qy = v[5].y; S1.qz = v[6].y;
  S1.vx = v[7].y; S1.vy = v[8].y; S1.vz = v[9].y; S1.wx = v[10].y; S1.wy = v[11].y; S1.wz = v[12].y;
  const IkOne a = ik_one(A.rc, c, S0), b = ik_one(A.rc, c, S1);
  double *o = A.out + (long long)c * 8 * n + i;
  __stcs(reinterpret_cast<double2 *>(o), make_double2(a.len, b.len));
  __stcs(reinterpret_cast<double2 *>(o + n), make_double2(a.rate, b.rate));
  __stcs(reinterpret_cast<double2 *>(o + 2 * n), make_double2(a.ux, b.ux));
  __stcs(reinterpret_cast<double2 *>(o + 3 * n), make_double2(a.uy, b.uy));
  __stcs(reinterpret_cast<double2 *>(o + 4 * n), make_double2(a.uz, b.uz));
  __stcs(reinterpret_cast<double2 *>(o + 5 * n), make_double2(a.cx, b.cx));
  __stcs(reinterpret_cast<double2 *>(o + 6 * n), make_double2(a.cy, b.cy));
  __stcs(reinterpret_cast<double2 *>(o + 7 * n), make_double2(a.cz, b.cz));
}

// state after CdprGazeboPlugin::Load: platform at home and at rest; every cable in Position mode,
// target 0, both Pids reset (wasLast = false, missing = bufferLength); everything else is memset 0.
__global__ void k_init_state(DevLayout L, RobotConsts rc, double hx, double hy, double hz, double qw, double qx, double qy, double qz,
                             unsigned vel_len, unsigned pos_len, int general) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.np) return;
  double *p = L.plat + i;
  p[0] = hx; p[L.np] = hy; p[2 * L.np] = hz;
  p[3 * L.np] = qw; p[4 * L.np] = qx; p[5 * L.np] = qy; p[6 * L.np] = qz;
  for (int k = 7; k < 13; ++k) p[k * L.np] = 0.0;
  // control word: fast variant = wasLast bits 0-1, missing counters in bytes 1-2; general variant: see step_general.cuh
  const unsigned ctl0 = general ? ((vel_len << 2) | (pos_len << 8)) : ((vel_len << 8) | (pos_len << 16));
  for (int c = 0; c < L.nc; ++c) L.ctl[(long long)c * L.np + i] = ctl0;
  L.ictl[i] = (unsigned)MODE_POSITION;  // CdprGazeboPlugin.cpp:154, no command pending
}

// Pid::reset (Pid.cpp:100-115) for pid k of every cable of every instance; mLastTime is kept.
__global__ void k_reset_pid(DevLayout L, int k, unsigned len_k, int general) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.np) return;
  for (int c = 0; c < L.nc; ++c) {
    L.pid[pid_off(L, c, k, PID_P_ERR) + i] = 0.0;
    L.pid[pid_off(L, c, k, PID_I_ERR) + i] = 0.0;
    L.pid[pid_off(L, c, k, PID_D_ERR) + i] = 0.0;
    L.pid[pid_off(L, c, k, PID_CMD) + i] = 0.0;
    if (L.mom)
      for (int mm = 0; mm < 3; ++mm) L.mom[mom_off(L, c, k, mm) + i] = 0.0;
    for (int j = 0; j < L.len; ++j) {
      L.win_y[win_off(L, c, k, j) + i] = 0.0;
      if (L.win_x) L.win_x[win_off(L, c, k, j) + i] = 0.0;
    }
    if (L.filt)
      for (int pd = 0; pd < 2; ++pd)
        for (int s = 0; s < L.casc; ++s)
          for (int f = 0; f < 4; ++f) L.filt[filt_off(L, c, k, pd, s, f) + i] = 0.0;
    unsigned ctl = L.ctl[(long long)c * L.np + i];
    if (general) {  // wasLast = false, missing = bufferLength, ring head = 0
      ctl &= ~((1u << k) | (0x3fu << (2 + 6 * k)) | (0x1fu << (14 + 5 * k)));
      ctl |= len_k << (2 + 6 * k);
    } else {
      ctl &= ~((1u << k) | (0xffu << (8 + 8 * k)));
      ctl |= len_k << (8 + 8 * k);
    }
    L.ctl[(long long)c * L.np + i] = ctl;
  }
}

// Joy.axes float32 [n][nc] (or float64) -> one cab field, widened to double (CdprGazeboPlugin.cpp:208,216)
template <typename T>
__global__ void k_scatter_cab(DevLayout L, int field, const T *aos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  for (int c = 0; c < L.nc; ++c) L.cab[cab_off(L, c, field) + i] = (double)aos[i * L.nc + c];
}

// The same for the flex variant, where every instance latches its own commands: only instances with mask[i] != 0 (all
// when mask is null) receive the message; `pending_bit` marks it for the instance's next update (bit 2 velocity, bit 3
// position, CdprGazeboPlugin.cpp:67-83); force_mode: JointForceCalculator::setForce switches the mode at once (.h:92-95)
template <typename T>
__global__ void k_scatter_cab_masked(DevLayout L, int field, const T *aos, const unsigned char *mask, unsigned pending_bit, int force_mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  if (mask && !mask[i]) return;
  for (int c = 0; c < L.nc; ++c) L.cab[cab_off(L, c, field) + i] = (double)aos[i * L.nc + c];
  unsigned w = L.ictl[i] | pending_bit;
  if (force_mode) w = (w & ~3u) | (unsigned)MODE_FORCE;
  L.ictl[i] = w;
}

// UpdateMode of every instance (flex variant)
__global__ void k_modes(DevLayout L, int *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  out[i] = (int)(L.ictl[i] & 3u);
}

__global__ void k_pack_platform(DevLayout L, double *pose7, double *twist6) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  FastState S;
  load_plat(L, i, S);
  if (pose7) {  // publishPlatformState: position, then orientation x y z w (CdprGazeboPlugin.cpp:263-269)
    double *o = pose7 + 7 * i;
    o[0] = S.px; o[1] = S.py; o[2] = S.pz; o[3] = S.qx; o[4] = S.qy; o[5] = S.qz; o[6] = S.qw;
  }
  if (twist6) {
    double *o = twist6 + 6 * i;
    o[0] = S.vx; o[1] = S.vy; o[2] = S.vz; o[3] = S.wx; o[4] = S.wy; o[5] = S.wz;
  }
}

// src_n > 0: instance i takes row (i / rep) of the source (rollouts: every sequence of a robot
// starts from that robot's state); rep = 1 for a plain set.
__global__ void k_unpack_platform(DevLayout L, const double *pose7, const double *twist6, long long rep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  const long long r = i / rep;
  double *p = L.plat + i;
  const long long np = L.np;
  if (pose7) {
    const double *s = pose7 + 7 * r;
    p[0] = s[0]; p[np] = s[1]; p[2 * np] = s[2];
    p[3 * np] = s[6]; p[4 * np] = s[3]; p[5 * np] = s[4]; p[6 * np] = s[5];
  }
  if (twist6) {
    const double *s = twist6 + 6 * r;
    for (int k = 0; k < 6; ++k) p[(7 + k) * np] = s[k];
  }
}

// publishJointStates (CdprGazeboPlugin.cpp:248-256): Position(), GetVelocity(0), GetForce(0)
__global__ void k_joint_states(DevLayout L, RobotConsts rc, double *pos, double *vel, double *eff) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  FastState S;
  load_plat(L, i, S);
  const Rot R = make_rot(S);
  for (int c = 0; c < L.nc; ++c) {
    const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
    const double rx = fma(R.r00, bx, fma(R.r01, by, R.r02 * bz));
    const double ry = fma(R.r10, bx, fma(R.r11, by, R.r12 * bz));
    const double rz = fma(R.r20, bx, fma(R.r21, by, R.r22 * bz));
    const double dx = (rc.a[c][0] - S.px) - rx, dy = (rc.a[c][1] - S.py) - ry, dz = (rc.a[c][2] - S.pz) - rz;
    const double len = sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
    const double ux = dx / len, uy = dy / len, uz = dz / len;
    const double cx = fma(ry, uz, -(rz * uy)), cy = fma(rz, ux, -(rx * uz)), cz = fma(rx, uy, -(ry * ux));
    if (pos) pos[i * L.nc + c] = rc.home_len[c] - len;
    if (vel) vel[i * L.nc + c] = fma(ux, S.vx, fma(uy, S.vy, fma(uz, S.vz, fma(cx, S.wx, fma(cy, S.wy, cz * S.wz)))));
    if (eff) eff[i * L.nc + c] = L.cab[cab_off(L, c, CAB_EFFORT) + i];
  }
}

// [n][nc][6] = pid_force, p_err, i_err, d_err, cmd (of the Pid the mode runs), mode
__global__ void k_pid_state(DevLayout L, int mode, int flex, double *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  if (flex) mode = (int)(L.ictl[i] & 3u);
  int k = (mode == MODE_POSITION) ? PID_POS : PID_VEL;
  for (int c = 0; c < L.nc; ++c) {
    if (flex) {  // the Pid that ran last on this cable (bits 28-29 of the control word, step_flex.cuh)
      const unsigned live = (L.ctl[(long long)c * L.np + i] >> 28) & 3u;
      if (live) k = (int)live - 1;
    }
    double *o = out + (i * L.nc + c) * 6;
    o[0] = L.cab[cab_off(L, c, CAB_PID_FORCE) + i];
    o[1] = L.pid[pid_off(L, c, k, PID_P_ERR) + i];
    o[2] = L.pid[pid_off(L, c, k, PID_I_ERR) + i];
    o[3] = L.pid[pid_off(L, c, k, PID_D_ERR) + i];
    o[4] = L.pid[pid_off(L, c, k, PID_CMD) + i];
    o[5] = (double)mode;
  }
}

// topic "pid" for every cable (the reference publishes cable 0 only, CdprGazeboPlugin.cpp:223-235):
// [n][nc][5] = pTerm, iTerm before its clamp, dTerm, desired (Pid.cpp:140-141,159,167), applied force = Joint::GetForce
__global__ void k_pid_terms(DevLayout L, double *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  for (int c = 0; c < L.nc; ++c) {
    double *o = out + (i * L.nc + c) * 5;
    o[0] = L.cab[cab_off(L, c, CAB_TERM_P) + i];
    o[1] = L.cab[cab_off(L, c, CAB_TERM_I) + i];
    o[2] = L.cab[cab_off(L, c, CAB_TERM_D) + i];
    o[3] = L.cab[cab_off(L, c, CAB_DESIRED) + i];
    o[4] = L.cab[cab_off(L, c, CAB_EFFORT) + i];
  }
}

// cost_seq[s] = sum over robots r (ascending: deterministic) of cost[r * n_seq + s]
__global__ void k_reduce_cost_seq(const double *cost, long long n_robots, long long n_seq, double *cost_seq) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seq) return;
  double acc = 0.0;
  for (long long r = 0; r < n_robots; ++r) acc += cost[r * n_seq + s];
  cost_seq[s] = acc;
}

// FP64 roofline denominator: 8 independent DFMA chains per thread, no memory traffic
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, b = 1e-9;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
      a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace cdpr
