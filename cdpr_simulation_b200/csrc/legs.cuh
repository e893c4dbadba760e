// legs.cuh -- leg fidelity: the five links of every UPS leg of sdf/cube.sdf:344-518 and the viscous
// damping of their passive joints, as a configuration-dependent generalised mass matrix and extra generalised forces on
// the 6-DOF platform.  PARITY UNPINNED like the rest of the rigid-body model (Gazebo/ODE are not available); the model is
// the one written out in the CPU checker (derivation there), which this file follows formula by formula.
//
//   chain:  frame -rev_X- virt_X -rev_Y- virt_Y -cable (prismatic)- cable -rev_Zpf- virt_Ypf -rev_Ypf- virt_Xpf -rev_Xpf- platform
//   leg triad (fixed in virt_Y / cable):  e2 = (u x x0)/c,  e1 = e2 x u,  u;   s = u.x0,  c = sqrt(1 - s^2)
//   rates for the anchor velocity v_B = v + w x r:  thy = -(e1.v_B)/L,  thx = (e2.v_B)/(L c),  w_leg = thx x0 + thy e2
//   gimbal (a3 fixed in the cable link, a1 in the platform, a2 = (a3 x a1)/|a3 x a1|), D = w_leg - w, t = a3.a1:
//       psy = D.a2,  psx = (D.a1 - t D.a3)/(1 - t^2),  phi = psx t - D.a3
//   kinetic energy of a leg = 1/2 |y|^2,  y = [sqrt(I) thx | sqrt(2I) w_leg | sqrt(I)(w_leg + phi a3) | sqrt(I)(w + psx a1) |
//                                               sqrt(m) v_c | sqrt(2m) v_B],  v_c = v_B - (l_c/L)(v_B - u (u.v_B))
//   M(x) = diag(m, m, m, R I_b R^T) + sum_legs Jy^T Jy;   Q += sum_legs [ m_l g.(v_c + 2 v_B) columns - Jz^T z - Jy^T (J'y xi) ],
//   z = sqrt(c_p) * rates;  J'y xi = d/dt y(x(t), xi held): the links' velocity-product terms, a one-sided difference of y
//   along the motion over kLegBiasDt (each link obeys m a = f, I alpha = tau with a = J xi' + J' xi; isotropic inertia, so
//   the links have no gyroscopic torque of their own)
//   step:  M(x_n) (xi+ - xi)/h = Q_cables + Q_gravity + Q_passive - Jy^T J'y xi - gyro(platform), then the pose update of App. C.6.
// A slow path by construction (a 6x6 system per instance and step, about 10,000 FP64 instructions at 8 cables): out of line, a
// rolled loop over the legs; inside a leg everything is unrolled (the unit twists behind the Jacobian are constants then).
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace cdpr {

struct LegGeom {
  double r[3], u[3], e1[3], e2[3], a1[3], a2[3], a3[3], x0[3];
  double L, c, t;
  double iL, iLc, i1t2;  // 1/L, 1/(L c), 1/(1 - t^2): the rates divide by them eight times per leg and step
};

__device__ __forceinline__ double dot3d(const double *a, const double *b) { return fma(a[0], b[0], fma(a[1], b[1], a[2] * b[2])); }
__device__ __forceinline__ void cross3d(const double *a, const double *b, double *o) {
  o[0] = fma(a[1], b[2], -(a[2] * b[1])); o[1] = fma(a[2], b[0], -(a[0] * b[2])); o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}

constexpr double kLegBiasDt = 1e-6;  // s: step of the one-sided difference behind the velocity-product terms

// geometry of leg i for platform position p and rotation R
__device__ inline void leg_geometry(const RobotConsts &rc, int i, const double *p, const double (*R)[3], LegGeom &g) {
  double d[3];
  for (int k = 0; k < 3; ++k) g.r[k] = fma(R[k][0], rc.b[i][0], fma(R[k][1], rc.b[i][1], R[k][2] * rc.b[i][2]));
  for (int k = 0; k < 3; ++k) d[k] = rc.a[i][k] - p[k] - g.r[k];
  g.L = sqrt(dot3d(d, d));
  g.iL = 1.0 / g.L;
  for (int k = 0; k < 3; ++k) { g.u[k] = d[k] * g.iL; g.x0[k] = rc.leg_x0[i][k]; }
  const double s = dot3d(g.u, g.x0);
  g.c = sqrt(1.0 - s * s);
  const double ic = 1.0 / g.c;
  g.iLc = g.iL * ic;
  for (int k = 0; k < 3; ++k) g.e1[k] = (g.x0[k] - s * g.u[k]) * ic;
  cross3d(g.u, g.e1, g.e2);
  for (int k = 0; k < 3; ++k) g.a3[k] = fma(rc.leg_alpha[i][0], g.e1[k], fma(rc.leg_alpha[i][1], g.e2[k], rc.leg_alpha[i][2] * g.u[k]));
  for (int k = 0; k < 3; ++k) g.a1[k] = fma(R[k][0], rc.leg_a1[i][0], fma(R[k][1], rc.leg_a1[i][1], R[k][2] * rc.leg_a1[i][2]));
  g.t = dot3d(g.a3, g.a1);
  double nrm[3];
  cross3d(g.a3, g.a1, nrm);
  const double omt = 1.0 - g.t * g.t;
  g.i1t2 = 1.0 / omt;
  const double inn = rsqrt(omt);
  for (int k = 0; k < 3; ++k) g.a2[k] = nrm[k] * inn;
}

__device__ inline void leg_rates(const RobotConsts &rc, const LegGeom &g, const double *v, const double *w, double *y, double *z) {
  double wr[3], vB[3];
  cross3d(w, g.r, wr);
  for (int k = 0; k < 3; ++k) vB[k] = v[k] + wr[k];
  const double thy = -dot3d(g.e1, vB) * g.iL;
  const double thx = dot3d(g.e2, vB) * g.iLc;
  double wleg[3], D[3];
  for (int k = 0; k < 3; ++k) { wleg[k] = fma(thx, g.x0[k], thy * g.e2[k]); D[k] = wleg[k] - w[k]; }
  const double psy = dot3d(D, g.a2);
  const double Da3 = dot3d(D, g.a3);
  const double psx = (dot3d(D, g.a1) - g.t * Da3) * g.i1t2;
  const double phi = psx * g.t - Da3;
  const double uv = dot3d(g.u, vB), lam = rc.leg_lc * g.iL;
  y[0] = rc.leg_sI * thx;
  for (int k = 0; k < 3; ++k) {
    y[1 + k] = rc.leg_s2I * wleg[k];
    y[4 + k] = rc.leg_sI * fma(phi, g.a3[k], wleg[k]);
    y[7 + k] = rc.leg_sI * fma(psx, g.a1[k], w[k]);
    y[10 + k] = rc.leg_sm * (vB[k] - lam * (vB[k] - g.u[k] * uv));
    y[13 + k] = rc.leg_s2m * vB[k];
  }
  z[0] = rc.leg_sc * thx; z[1] = rc.leg_sc * thy; z[2] = rc.leg_sc * phi; z[3] = rc.leg_sc * psy; z[4] = rc.leg_sc * psx;
}

// (fx..mz) = cable wrench + platform gravity about the COM in frame axes (what rigid_body_step takes)
static __device__ __noinline__ FastState legs_step(const StepArgs &A, FastState S, double fx, double fy, double fz, double mx, double my, double mz) {
  const RobotConsts &rc = A.rc;
  const Rot Rr = make_rot(S);
  const double R[3][3] = {{Rr.r00, Rr.r01, Rr.r02}, {Rr.r10, Rr.r11, Rr.r12}, {Rr.r20, Rr.r21, Rr.r22}};
  const double Ib[3][3] = {{rc.ib[0], rc.ib[3], rc.ib[4]}, {rc.ib[3], rc.ib[1], rc.ib[5]}, {rc.ib[4], rc.ib[5], rc.ib[2]}};
  const double p[3] = {S.px, S.py, S.pz}, v[3] = {S.vx, S.vy, S.vz}, w[3] = {S.wx, S.wy, S.wz};
  double M[6][6];
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) M[a][b] = 0.0;
  const double mass = rc.mass;
  for (int k = 0; k < 3; ++k) M[k][k] = mass;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0.0;
      for (int j = 0; j < 3; ++j)
        for (int l = 0; l < 3; ++l) s = fma(R[a][j] * Ib[j][l], R[b][l], s);
      M[3 + a][3 + b] = s;
    }
  // explicit gyroscopic torque of the platform, w x (I_w w)
  double Lw[3], gy[3];
  for (int a = 0; a < 3; ++a) Lw[a] = fma(M[3 + a][3], w[0], fma(M[3 + a][4], w[1], M[3 + a][5] * w[2]));
  cross3d(w, Lw, gy);
  double Q[6] = {fx, fy, fz, mx - gy[0], my - gy[1], mz - gy[2]};
  const double grav[3] = {rc.grav[0], rc.grav[1], rc.grav[2]};
  // the pose a moment later at the current twist (first order, like the integrator): behind the velocity-product terms
  double p1[3], R1[3][3];
  {
    FastState S1 = S;
    const double hb = 0.5 * kLegBiasDt;
    const double hx = hb * S.wx, hy = hb * S.wy, hz = hb * S.wz;
    const double nw = fma(-hx, S.qx, fma(-hy, S.qy, fma(-hz, S.qz, S.qw)));
    const double nx = fma(hx, S.qw, fma(hy, S.qz, fma(-hz, S.qy, S.qx)));
    const double ny = fma(-hx, S.qz, fma(hy, S.qw, fma(hz, S.qx, S.qy)));
    const double nz = fma(hx, S.qy, fma(-hy, S.qx, fma(hz, S.qw, S.qz)));
    const double inv = 1.0 / sqrt(fma(nw, nw, fma(nx, nx, fma(ny, ny, nz * nz))));
    S1.qw = nw * inv; S1.qx = nx * inv; S1.qy = ny * inv; S1.qz = nz * inv;
    const Rot Q1 = make_rot(S1);
    R1[0][0] = Q1.r00; R1[0][1] = Q1.r01; R1[0][2] = Q1.r02; R1[1][0] = Q1.r10; R1[1][1] = Q1.r11; R1[1][2] = Q1.r12;
    R1[2][0] = Q1.r20; R1[2][1] = Q1.r21; R1[2][2] = Q1.r22;
    for (int k = 0; k < 3; ++k) p1[k] = fma(kLegBiasDt, v[k], p[k]);
  }
#pragma unroll 1
  for (int i = 0; i < A.L.nc; ++i) {
    LegGeom g, g1;
    leg_geometry(rc, i, p, R, g);
    leg_geometry(rc, i, p1, R1, g1);
    double Jy[16][6], Jz[5][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {  // unrolled: the unit twists are constants, most of leg_rates folds away
      double ev[3] = {0.0, 0.0, 0.0}, ew[3] = {0.0, 0.0, 0.0}, y[16], z[5];
      if (k < 3) ev[k] = 1.0; else ew[k - 3] = 1.0;
      leg_rates(rc, g, ev, ew, y, z);
#pragma unroll
      for (int j = 0; j < 16; ++j) Jy[j][k] = y[j];
#pragma unroll
      for (int j = 0; j < 5; ++j) Jz[j][k] = z[j];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fma(Jy[j][a], Jy[j][b], acc);
        M[a][b] += acc;
      }
    double y[16], z[5], y1[16], z1[5];
    leg_rates(rc, g, v, w, y, z);
    leg_rates(rc, g1, v, w, y1, z1);
#pragma unroll
    for (int j = 0; j < 16; ++j) y1[j] = (y1[j] - y[j]) * (1.0 / kLegBiasDt);  // J'y xi
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double damp = 0.0, gr = 0.0, bias = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j) damp = fma(Jz[j][k], z[j], damp);
#pragma unroll
      for (int j = 0; j < 3; ++j) gr = fma(grav[j], fma(rc.leg_sm, Jy[10 + j][k], rc.leg_s2m * Jy[13 + j][k]), gr);
#pragma unroll
      for (int j = 0; j < 16; ++j) bias = fma(Jy[j][k], y1[j], bias);
      Q[k] += gr - damp - bias;
    }
  }
  // Cholesky of the upper triangle (M = U^T U), two triangular solves
  double U[6][6];
  for (int j = 0; j < 6; ++j) {
    double sdiag = M[j][j];
    for (int k = 0; k < j; ++k) sdiag = fma(-U[k][j], U[k][j], sdiag);
    U[j][j] = sqrt(sdiag);
    for (int c2 = j + 1; c2 < 6; ++c2) {
      double t = M[j][c2];
      for (int k = 0; k < j; ++k) t = fma(-U[k][c2], U[k][j], t);
      U[j][c2] = t / U[j][j];
    }
  }
  double yv[6], acc[6];
  for (int a = 0; a < 6; ++a) {
    double sv = Q[a];
    for (int k = 0; k < a; ++k) sv = fma(-U[k][a], yv[k], sv);
    yv[a] = sv / U[a][a];
  }
  for (int a = 5; a >= 0; --a) {
    double sv = yv[a];
    for (int k = a + 1; k < 6; ++k) sv = fma(-U[a][k], acc[k], sv);
    acc[a] = sv / U[a][a];
  }
  S.vx = fma(rc.h, acc[0], S.vx); S.vy = fma(rc.h, acc[1], S.vy); S.vz = fma(rc.h, acc[2], S.vz);
  S.wx = fma(rc.h, acc[3], S.wx); S.wy = fma(rc.h, acc[4], S.wy); S.wz = fma(rc.h, acc[5], S.wz);
  integrate_pose(rc, S);
  return S;
}

}  // namespace cdpr
