// physics.cuh -- device functions shared by the step kernels: rotation matrix and the rigid-body
// update of the reduced model (stands in for Gazebo/ODE's world step).
#pragma once
#include "common.cuh"

namespace cdpr {

struct FastState {
  double px, py, pz, qw, qx, qy, qz, vx, vy, vz, wx, wy, wz;
};
struct Rot {
  double r00, r01, r02, r10, r11, r12, r20, r21, r22;
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// 1/sqrt(x) for normal, positive x without the special-case branch of CUDA's rsqrt(): MUFU.RSQ64H
// seed (rel. error < 2^-22) and one third-order Newton step y += y*e*(1/2 + 3/8 e), e = 1 - x*y*y,
// which leaves ~2^-66 of method error, i.e. the result is within ~1 ulp.
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// rotation matrix of the platform (unit quaternion w x y z)
__device__ __forceinline__ Rot make_rot(const FastState &S) {
  // 17 FP64 instructions: doubled components, then one FMA per off-diagonal entry
  const double x2 = S.qx + S.qx, y2 = S.qy + S.qy, z2 = S.qz + S.qz;
  const double wx2 = S.qw * x2, wy2 = S.qw * y2, wz2 = S.qw * z2;
  const double ax = fma(-x2, S.qx, 1.0), ay = fma(-y2, S.qy, 1.0);
  Rot R;
  R.r00 = fma(-z2, S.qz, ay); R.r01 = fma(x2, S.qy, -wz2); R.r02 = fma(x2, S.qz, wy2);
  R.r10 = fma(x2, S.qy, wz2); R.r11 = fma(-z2, S.qz, ax); R.r12 = fma(y2, S.qz, -wx2);
  R.r20 = fma(x2, S.qz, -wy2); R.r21 = fma(y2, S.qz, wx2); R.r22 = fma(-y2, S.qy, ax);
  return R;
}

__device__ __forceinline__ void load_plat(const DevLayout &L, long long i, FastState &S) {
  const double *p = L.plat + i;
  const long long np = L.np;
  S.px = p[0]; S.py = p[np]; S.pz = p[2 * np];
  S.qw = p[3 * np]; S.qx = p[4 * np]; S.qy = p[5 * np]; S.qz = p[6 * np];
  S.vx = p[7 * np]; S.vy = p[8 * np]; S.vz = p[9 * np];
  S.wx = p[10 * np]; S.wy = p[11 * np]; S.wz = p[12 * np];
}
__device__ __forceinline__ void store_plat(double *p, long long stride, const FastState &S) {
  p[0] = S.px; p[stride] = S.py; p[2 * stride] = S.pz;
  p[3 * stride] = S.qw; p[4 * stride] = S.qx; p[5 * stride] = S.qy; p[6 * stride] = S.qz;
  p[7 * stride] = S.vx; p[8 * stride] = S.vy; p[9 * stride] = S.vz;
  p[10 * stride] = S.wx; p[11 * stride] = S.wy; p[12 * stride] = S.wz;
}

// Compile-time specialisations of the robot constants (detected at cdpr_create, never assumed):
//   SPEC_DIAG  body inertia is diagonal (products of inertia are 0), as in sdf/cube.sdf:331-338
//   SPEC_ISO   ... and ixx == iyy == izz: the gyroscopic torque vanishes and I_w^-1 is a scalar
//   SPEC_BZ0   every platform anchor has b_z == 0 (anchors in the platform's xy plane, cube.yaml:21-29)
//   SPEC_NOFF  the live Pid has no feed-forward gain (velocityControllerForward = 0, launch:19)
//   SPEC_UTGT  every cable has the same target (the sine publisher writes one value to all axes, sinevelocitytest.cpp:36-38)
//   SPEC_PAIR  cables c and c + NC/2 leave the SAME platform anchor for frame anchors that differ in z only (an upper and a
//              lower frame corner above each other: the 8-cable cube of SURVEY.md App. A.2): their kinematics share d_x, d_y,
//              g_x, g_y and everything built from them (cable_kin_pair)
enum { SPEC_DIAG = 1, SPEC_ISO = 2, SPEC_BZ0 = 4, SPEC_NOFF = 8, SPEC_UTGT = 16, SPEC_PAIR = 32 };

// NVLS multicast store: one store, replicated by the NVSwitch into the mapped buffer of every rank
__device__ __forceinline__ void mc_store(double *p, double v) { asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void store_plat_multicast(double *p, long long stride, const FastState &S) {
  mc_store(p, S.px); mc_store(p + stride, S.py); mc_store(p + 2 * stride, S.pz);
  mc_store(p + 3 * stride, S.qw); mc_store(p + 4 * stride, S.qx); mc_store(p + 5 * stride, S.qy); mc_store(p + 6 * stride, S.qz);
  mc_store(p + 7 * stride, S.vx); mc_store(p + 8 * stride, S.vy); mc_store(p + 9 * stride, S.vz);
  mc_store(p + 10 * stride, S.wx); mc_store(p + 11 * stride, S.wy); mc_store(p + 12 * stride, S.wz);
}

// pose update with the NEW velocities (App. C.6): p += h v, q += h 1/2 (0, w) (x) q, renormalise
__device__ __forceinline__ void integrate_pose(const RobotConsts &rc, FastState &S) {
  S.px = fma(rc.h, S.vx, S.px); S.py = fma(rc.h, S.vy, S.py); S.pz = fma(rc.h, S.vz, S.pz);
  const double hx = rc.half_h * S.wx, hy = rc.half_h * S.wy, hz = rc.half_h * S.wz;
  const double nw = fma(-hx, S.qx, fma(-hy, S.qy, fma(-hz, S.qz, S.qw)));
  const double nx = fma(hx, S.qw, fma(hy, S.qz, fma(-hz, S.qy, S.qx)));
  const double ny = fma(-hx, S.qz, fma(hy, S.qw, fma(hz, S.qx, S.qy)));
  const double nz = fma(hx, S.qy, fma(-hy, S.qx, fma(hz, S.qw, S.qz)));
  const double inv = rsqrt_nr(fma(nw, nw, fma(nx, nx, fma(ny, ny, nz * nz))));
  S.qw = nw * inv; S.qx = nx * inv; S.qy = ny * inv; S.qz = nz * inv;
}

// (fx..fz, mx..mz) = net force / torque about the COM in frame axes, gravity included
template <int SPEC>
__device__ __forceinline__ void rigid_body_step(const RobotConsts &rc, FastState &S, const Rot &R, double fx, double fy, double fz,
                                                double mx, double my, double mz) {
  const double r00 = R.r00, r01 = R.r01, r02 = R.r02, r10 = R.r10, r11 = R.r11, r12 = R.r12, r20 = R.r20, r21 = R.r21, r22 = R.r22;
  // ---- rigid-body step, ODE order (a9): I_w = R I_b R^T, explicit gyroscopic torque
  double alx, aly, alz;
  if (SPEC & SPEC_ISO) {
    // I_b = k * identity: w x (I w) = 0 and I_w^-1 = 1/k
    alx = rc.ib_inv[0] * mx; aly = rc.ib_inv[0] * my; alz = rc.ib_inv[0] * mz;
  } else {
    // angular part in the body frame: alpha = R I_b^-1 (R^T M - w_b x (I_b w_b)); identical to ODE's world-frame
    // form M - w x (R I_b R^T w) because rotations preserve cross products
    const double wbx = fma(r00, S.wx, fma(r10, S.wy, r20 * S.wz));
    const double wby = fma(r01, S.wx, fma(r11, S.wy, r21 * S.wz));
    const double wbz = fma(r02, S.wx, fma(r12, S.wy, r22 * S.wz));
    double lbx, lby, lbz;
    if (SPEC & SPEC_DIAG) {
      lbx = rc.ib[0] * wbx; lby = rc.ib[1] * wby; lbz = rc.ib[2] * wbz;
    } else {
      lbx = fma(rc.ib[0], wbx, fma(rc.ib[3], wby, rc.ib[4] * wbz));
      lby = fma(rc.ib[3], wbx, fma(rc.ib[1], wby, rc.ib[5] * wbz));
      lbz = fma(rc.ib[4], wbx, fma(rc.ib[5], wby, rc.ib[2] * wbz));
    }
    const double mbx = fma(r00, mx, fma(r10, my, r20 * mz)) - fma(wby, lbz, -(wbz * lby));
    const double mby = fma(r01, mx, fma(r11, my, r21 * mz)) - fma(wbz, lbx, -(wbx * lbz));
    const double mbz = fma(r02, mx, fma(r12, my, r22 * mz)) - fma(wbx, lby, -(wby * lbx));
    double abx, aby, abz;
    if (SPEC & SPEC_DIAG) {
      abx = rc.ib_inv[0] * mbx; aby = rc.ib_inv[1] * mby; abz = rc.ib_inv[2] * mbz;
    } else {
      abx = fma(rc.ib_inv[0], mbx, fma(rc.ib_inv[3], mby, rc.ib_inv[4] * mbz));
      aby = fma(rc.ib_inv[3], mbx, fma(rc.ib_inv[1], mby, rc.ib_inv[5] * mbz));
      abz = fma(rc.ib_inv[4], mbx, fma(rc.ib_inv[5], mby, rc.ib_inv[2] * mbz));
    }
    alx = fma(r00, abx, fma(r01, aby, r02 * abz));
    aly = fma(r10, abx, fma(r11, aby, r12 * abz));
    alz = fma(r20, abx, fma(r21, aby, r22 * abz));

  }
  S.vx = fma(rc.h_over_m, fx, S.vx); S.vy = fma(rc.h_over_m, fy, S.vy); S.vz = fma(rc.h_over_m, fz, S.vz);
  S.wx = fma(rc.h, alx, S.wx); S.wy = fma(rc.h, aly, S.wy); S.wz = fma(rc.h, alz, S.wz);
  integrate_pose(rc, S);
}

// ---- inverse kinematics of one cable (a7).  With g = a - p (anchor seen from the platform origin): d = g - R b = L u,
// and r x u = (g - d) x u = g x u because d is parallel to u -- so neither r nor u is formed:
//   m = g x d = L (r x u),  joint rate = (d.v + m.w) / L,  and the wrench scales d and m by tension / L.
struct CableKin { double dx, dy, dz, cx, cy, cz, il, qd, qp; };
template <int SPEC, bool WANT_QP>
__device__ __forceinline__ CableKin cable_kin(const RobotConsts &rc, const FastState &S, const Rot &R, int c) {
  CableKin k;
  const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
  const double gx = rc.a[c][0] - S.px, gy = rc.a[c][1] - S.py, gz = rc.a[c][2] - S.pz;
  k.dx = fma(-R.r00, bx, fma(-R.r01, by, (SPEC & SPEC_BZ0) ? gx : fma(-R.r02, bz, gx)));
  k.dy = fma(-R.r10, bx, fma(-R.r11, by, (SPEC & SPEC_BZ0) ? gy : fma(-R.r12, bz, gy)));
  k.dz = fma(-R.r20, bx, fma(-R.r21, by, (SPEC & SPEC_BZ0) ? gz : fma(-R.r22, bz, gz)));
  const double l2 = fma(k.dx, k.dx, fma(k.dy, k.dy, k.dz * k.dz));
  k.il = rsqrt_nr(l2);
  k.cx = fma(gy, k.dz, -(gz * k.dy)); k.cy = fma(gz, k.dx, -(gx * k.dz)); k.cz = fma(gx, k.dy, -(gy * k.dx));
  k.qd = (fma(k.dx, S.vx, fma(k.dy, S.vy, k.dz * S.vz)) + fma(k.cx, S.wx, fma(k.cy, S.wy, k.cz * S.wz))) * k.il;
  k.qp = 0.0;
  if (WANT_QP) k.qp = fma(-l2, k.il, rc.home_len[c]);  // explicit: ptxas must not get to choose between mul + sub and fma per call site
  return k;
}

// SPEC_PAIR: the two cables of a pair (c, c2 = c + NC/2) in one go.  g2 = g + (0, 0, dz0), d2 = d + (0, 0, dz0) with
// dz0 = a2_z - a_z, so of the second cable's 31 FP64 instructions only 17 remain:
//   d2.d2 = (dx^2 + dy^2) + dz2^2,  (g2 x d2)_z = (g x d)_z,  d2.v = (dx vx + dy vy) + dz2 vz,  (g2 x d2).w shares (g x d)_z wz.
template <int SPEC, bool WANT_QP>
__device__ __forceinline__ void cable_kin_pair(const RobotConsts &rc, const FastState &S, const Rot &R, int c, int c2, double dz0, CableKin &k, CableKin &k2) {
  const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
  const double gx = rc.a[c][0] - S.px, gy = rc.a[c][1] - S.py, gz = rc.a[c][2] - S.pz;
  k.dx = fma(-R.r00, bx, fma(-R.r01, by, (SPEC & SPEC_BZ0) ? gx : fma(-R.r02, bz, gx)));
  k.dy = fma(-R.r10, bx, fma(-R.r11, by, (SPEC & SPEC_BZ0) ? gy : fma(-R.r12, bz, gy)));
  k.dz = fma(-R.r20, bx, fma(-R.r21, by, (SPEC & SPEC_BZ0) ? gz : fma(-R.r22, bz, gz)));
  const double sxy = fma(k.dx, k.dx, k.dy * k.dy);
  const double dvxy = fma(k.dx, S.vx, k.dy * S.vy);
  k.cz = fma(gx, k.dy, -(gy * k.dx));
  const double czw = k.cz * S.wz;
  // first cable
  const double l2 = fma(k.dz, k.dz, sxy);
  k.il = rsqrt_nr(l2);
  k.cx = fma(gy, k.dz, -(gz * k.dy)); k.cy = fma(gz, k.dx, -(gx * k.dz));
  k.qd = (fma(k.dz, S.vz, dvxy) + fma(k.cx, S.wx, fma(k.cy, S.wy, czw))) * k.il;
  k.qp = 0.0;
  if (WANT_QP) k.qp = fma(-l2, k.il, rc.home_len[c]);  // explicit: ptxas must not get to choose between mul + sub and fma per call site
  // second cable: same platform anchor, frame anchor dz0 higher
  const double gz2 = gz + dz0;
  k2.dx = k.dx; k2.dy = k.dy; k2.dz = k.dz + dz0;
  const double l22 = fma(k2.dz, k2.dz, sxy);
  k2.il = rsqrt_nr(l22);
  k2.cx = fma(gy, k2.dz, -(gz2 * k.dy)); k2.cy = fma(gz2, k.dx, -(gx * k2.dz)); k2.cz = k.cz;
  k2.qd = (fma(k2.dz, S.vz, dvxy) + fma(k2.cx, S.wx, fma(k2.cy, S.wy, czw))) * k2.il;
  k2.qp = 0.0;
  if (WANT_QP) k2.qp = fma(-l22, k2.il, rc.home_len[c2]);
}

// rare path (every snap_every steps), kept out of line so the hot loop's register allocation does not see it
static __device__ __noinline__ void write_snapshot(const StepArgs &A, FastState S, long long o) {
  if (A.snap_multimem) {
    store_plat_multicast(A.snap_peers[0] + o, A.snap_stride, S);
  } else {
    for (int p = 0; p < A.n_snap_peers; ++p)  // plain stores; peer buffers are NVLink-mapped device memory
      store_plat(A.snap_peers[p] + o, A.snap_stride, S);
  }
}

// publishPlatformState (CdprGazeboPlugin.cpp:258-280) of the state the update sees: position, orientation x y z w, twist
__device__ __forceinline__ void publish_platform(const StepArgs &A, const FastState &S, long long i) {
  if (A.pub_pose) {
    double *o = A.pub_pose + 7 * i;
    o[0] = S.px; o[1] = S.py; o[2] = S.pz; o[3] = S.qx; o[4] = S.qy; o[5] = S.qz; o[6] = S.qw;
  }
  if (A.pub_twist) {
    double *o = A.pub_twist + 6 * i;
    o[0] = S.vx; o[1] = S.vy; o[2] = S.vz; o[3] = S.wx; o[4] = S.wy; o[5] = S.wz;
  }
}
// publishJointStates (CdprGazeboPlugin.cpp:248-256): Position(), GetVelocity(0) at this update, GetForce(0) of this step
__device__ __forceinline__ void publish_joint(const StepArgs &A, int nc, int c, double qp, double qd, double eff, long long i) {
  if (A.pub_pos) A.pub_pos[i * nc + c] = qp;
  if (A.pub_vel) A.pub_vel[i * nc + c] = qd;
  if (A.pub_eff) A.pub_eff[i * nc + c] = eff;
}

}  // namespace cdpr
