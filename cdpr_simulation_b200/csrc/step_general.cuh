// step_general.cuh -- the catch-all variant of the K-step kernel: full semantics of
// JointForceCalculator (three modes, velocity hold through the position Pid on
// |target| <= velocityEpsilon, JointForceCalculator.cpp:59-96) and gazebo::common::Pid
// (biquad cascades on the P input and D output, Pid.cpp:27-44,133,157; cmdLimit == 0 quirk,
// Pid.cpp:175-184; non-uniform time stamps in the D window, Pid.cpp:193-247).
//
// One thread per instance; the platform state lives in registers for the K steps, the per-cable
// controller state is read-modify-written in HBM/L2 every step (coalesced: consecutive threads,
// consecutive addresses); the D windows are rings there, so a push is one store and the fit one pass.  This variant is HBM/L2-bound, not FP64-bound; the reference's launch
// configuration never needs it (see step_fast.cuh).
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace cdpr {

// Pid::CascadeFilter::update (Pid.cpp:38-44) over BiQuad::process (Filter.h:152-165)
static __device__ inline double cascade_update(const DevLayout &L, int c, int k, int pd, int stages, const double *co, double x, long long i) {
  double out = x;
  for (int s = 0; s < stages; ++s) {
    double *x1 = L.filt + filt_off(L, c, k, pd, s, 0) + i, *x2 = L.filt + filt_off(L, c, k, pd, s, 1) + i;
    double *y1 = L.filt + filt_off(L, c, k, pd, s, 2) + i, *y2 = L.filt + filt_off(L, c, k, pd, s, 3) + i;
    const double y0 = co[0] * out + co[1] * *x1 + co[2] * *x2 - co[3] * *y1 - co[4] * *y2;
    *x2 = *x1; *x1 = out; *y2 = *y1; *y1 = y0;
    out = y0;
  }
  return out;
}

// Control word of the general variant, one per (instance, cable):
//   bit 0 / 1      velocity / position Pid: mWasLastTime
//   bits 2-7, 8-13 mDbufferMissing of the two Pids (0..32)
//   bits 14-18, 19-23 ring head of the two Pids = slot of the OLDEST window sample = next slot to overwrite
// The D windows of this variant are rings in HBM (no shifting through memory, Pid.cpp:194-199 restated).
__host__ __device__ inline unsigned gctl_pack(unsigned vel_len, unsigned pos_len) { return (vel_len << 2) | (pos_len << 8); }
__device__ __forceinline__ unsigned gctl_missing(unsigned ctl, int k) { return (ctl >> (2 + 6 * k)) & 0x3fu; }
__device__ __forceinline__ unsigned gctl_head(unsigned ctl, int k) { return (ctl >> (14 + 5 * k)) & 0x1fu; }
__device__ __forceinline__ unsigned gctl_set(unsigned ctl, int k, unsigned missing, unsigned head) {
  ctl &= ~((0x3fu << (2 + 6 * k)) | (0x1fu << (14 + 5 * k)));
  return ctl | (missing << (2 + 6 * k)) | (head << (14 + 5 * k));
}

// normal equations -> derivative coefficient: Gaussian elimination with partial pivoting by compare-and-swap (every index static)
template <int D>
__device__ __forceinline__ double ls_solve(const double (&sx)[2 * D + 1], const double (&sy)[D + 1], double inv_span) {
  constexpr int M = D + 1;
  double A[M][M + 1];
#pragma unroll
  for (int r = 0; r < M; ++r) {
#pragma unroll
    for (int q = 0; q < M; ++q) A[r][q] = sx[r + q];
    A[r][M] = sy[r];
  }
#pragma unroll
  for (int col = 0; col < M; ++col) {
#pragma unroll
    for (int r = col + 1; r < M; ++r) {  // partial pivoting by compare-and-swap keeps every index static
      const bool sw = fabs(A[r][col]) > fabs(A[col][col]);
#pragma unroll
      for (int q = col; q <= M; ++q) {
        const double a = A[col][q], b = A[r][q];
        A[col][q] = sw ? b : a;
        A[r][q] = sw ? a : b;
      }
    }
    if (A[col][col] == 0.0) return 0.0;
    const double inv = 1.0 / A[col][col];
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      const double f = A[r][col] * inv;
#pragma unroll
      for (int q = col + 1; q <= M; ++q) A[r][q] = fma(-f, A[col][q], A[r][q]);
    }
  }
  double coef[M];
#pragma unroll
  for (int r = M - 1; r >= 1; --r) {  // coef[0] is not needed
    double s = A[r][M];
#pragma unroll
    for (int q = r + 1; q < M; ++q) s = fma(-A[r][q], coef[q], s);
    coef[r] = s / A[r][r];
  }
  return coef[1] * inv_span;
}

// Derivative at `now` of the degree-D least-squares polynomial through the window (Pid.cpp:203-212 + 219-247), fitted
// in window-relative, span-scaled time (same polynomial as the reference's absolute-time fit, but well conditioned):
// one pass over the ring accumulates the normal equations, Gaussian elimination with partial pivoting solves them.
// LEN > 0: window length known at compile time -- the sample loop is fully unrolled, so all 2 LEN loads are in flight at once
// (the flex kernel's gap fit is bound by exactly that latency)
template <int D, int LEN = 0>
__device__ __forceinline__ double ls_derivative(const DevLayout &L, int c, int k, int len, unsigned oldest, double now, long long i) {
  constexpr int M = D + 1;
  if (LEN > 0) {
    double xs[LEN > 0 ? LEN : 1], ys[LEN > 0 ? LEN : 1];
#pragma unroll
    for (int j = 0; j < LEN; ++j) { xs[j] = L.win_x[win_off(L, c, k, j) + i]; ys[j] = L.win_y[win_off(L, c, k, j) + i]; }
    double oldest_x = xs[0];
#pragma unroll
    for (int j = 1; j < LEN; ++j) oldest_x = ((unsigned)j == oldest) ? xs[j] : oldest_x;
    const double span = now - oldest_x, inv_span = 1.0 / span;
    double sx[2 * D + 1], sy[M];
#pragma unroll
    for (int p = 0; p <= 2 * D; ++p) sx[p] = 0.0;
#pragma unroll
    for (int p = 0; p < M; ++p) sy[p] = 0.0;
#pragma unroll
    for (int j = 0; j < LEN; ++j) {
      const double x = (xs[j] - now) * inv_span;
      double pw = 1.0;
#pragma unroll
      for (int p = 0; p <= 2 * D; ++p) {
        sx[p] += pw;
        if (p < M) sy[p] = fma(pw, ys[j], sy[p]);
        pw *= x;
      }
    }
    return ls_solve<D>(sx, sy, inv_span);
  }
  const double span = now - L.win_x[win_off(L, c, k, (int)oldest) + i];
  const double inv_span = 1.0 / span;
  double sx[2 * D + 1], sy[M];
#pragma unroll
  for (int p = 0; p <= 2 * D; ++p) sx[p] = 0.0;
#pragma unroll
  for (int p = 0; p < M; ++p) sy[p] = 0.0;
#pragma unroll 4
  for (int j = 0; j < len; ++j) {
    const double x = (L.win_x[win_off(L, c, k, j) + i] - now) * inv_span;
    const double y = L.win_y[win_off(L, c, k, j) + i];
    double pw = 1.0;
#pragma unroll
    for (int p = 0; p <= 2 * D; ++p) {
      sx[p] += pw;
      if (p < M) sy[p] = fma(pw, y, sy[p]);
      pw *= x;
    }
  }
  return ls_solve<D>(sx, sy, inv_span);
}

// Pid::derive (Pid.cpp:193-217): overwrite the oldest sample, then the fit once the window is full
template <int DMAX>
__device__ __forceinline__ double derive_general(const StepArgs &A, const DevLayout &L, const PidConsts &pc, int c, int k, unsigned &ctl,
                                                 double value, double now, long long i) {
  const int len = pc.len;
  unsigned head = gctl_head(ctl, k), missing = gctl_missing(ctl, k);
  L.win_x[win_off(L, c, k, (int)head) + i] = now;
  L.win_y[win_off(L, c, k, (int)head) + i] = value;
  head = (head + 1u == (unsigned)len) ? 0u : head + 1u;
  missing -= (missing > 0u) ? 1u : 0u;
  ctl = gctl_set(ctl, k, missing, head);
  if (missing != 0u) return 0.0;
  if (pc.degree < 1) return 0.0;  // degree 0: derivative of a constant
  // Common case: the window holds the last `len` CONSECUTIVE steps (span == (len-1) dt): the fit is the fixed FIR of
  // step_fast.cuh -- one pass over the error ring, no time stamps, no solve.  Gaps (hold phases, stale windows) take
  // the general fit below.
  const double span = now - L.win_x[win_off(L, c, k, (int)head) + i];
  if (fabs(span - (len - 1) * A.rc.h) < 0.25 * A.rc.h) {
    double d0 = 0.0, d1 = 0.0;
    int sl = (int)head;  // oldest sample = logical position 0
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
      const double y = L.win_y[win_off(L, c, k, sl) + i];
      if (j & 1) d1 = fma(A.fir2[k][j], y, d1); else d0 = fma(A.fir2[k][j], y, d0);
      sl = (sl + 1 == len) ? 0 : sl + 1;
    }
    return d0 + d1;
  }
  // DMAX bounds the degrees compiled into this instance (the 5 x 6 system of degree 4 would set the register budget
  // of the common degree-2 case otherwise)
  if (pc.degree == 1) return ls_derivative<1>(L, c, k, len, head, now, i);
  if (pc.degree == 2) return ls_derivative<2>(L, c, k, len, head, now, i);
  if (DMAX >= 3 && pc.degree == 3) return ls_derivative<3>(L, c, k, len, head, now, i);
  if (DMAX >= 4 && pc.degree == 4) return ls_derivative<4>(L, c, k, len, head, now, i);
  return 0.0;  // degree 0: derivative of a constant
}

// Pid::update (Pid.cpp:122-191) on the state columns of (cable c, pid k)
template <int DMAX>
__device__ __forceinline__ double pid_update_general(const StepArgs &A, const DevLayout &L, const PidConsts &pc, int c, int k, unsigned &ctl, double desired,
                                                     double actual, double now, long long i) {
  double *last_time = L.pid + pid_off(L, c, k, PID_LAST_TIME) + i;
  double *cmdp = L.pid + pid_off(L, c, k, PID_CMD) + i;
  double cmd_out;
  if (!((ctl >> k) & 1u)) {
    ctl |= 1u << k;
    cmd_out = 0.0;
  } else {
    double *p_err = L.pid + pid_off(L, c, k, PID_P_ERR) + i;
    double *i_err = L.pid + pid_off(L, c, k, PID_I_ERR) + i;
    double *d_err = L.pid + pid_off(L, c, k, PID_D_ERR) + i;
    const double f_term = pc.kf * desired;
    const double error = desired - actual;
    const double dt = now - *last_time;
    const double pe = cascade_update(L, c, k, 0, pc.p_casc, pc.pf, error, i);
    *p_err = pe;
    const double p_term = pc.kp * pe;
    const double prev_ierr = *i_err;
    double ie = prev_ierr + dt * error;
    double i_term = pc.ki * ie;
    if (i_term > pc.i_max) { i_term = pc.i_max; ie = pc.i_max_over_ki; }
    else if (i_term < pc.i_min) { i_term = pc.i_min; ie = pc.i_min_over_ki; }
    double de;
    if (dt > 0.0) {
      const double derived = derive_general<DMAX>(A, L, pc, c, k, ctl, error, now, i);
      de = cascade_update(L, c, k, 1, pc.d_casc, pc.df, derived, i);
      *d_err = de;
    } else {
      de = *d_err;
    }
    const double d_term = pc.kd * de;
    const double cmd_raw = f_term + p_term + i_term + d_term;
    double cmd;
    if (pc.cmd_max > pc.cmd_min) cmd = clampd(cmd_raw, pc.cmd_min, pc.cmd_max);
    else cmd = *cmdp;  // cmdLimit == 0: mCmd keeps its previous value (Pid.cpp:175-179)
    if (cmd != cmd_raw) {
      ie = prev_ierr;
      cmd += dt * error * pc.ki;
    }
    *i_err = ie;
    cmd_out = cmd;
    // topic "pid": pTerm, iTerm before its clamp, dTerm, desired (Pid.cpp:140-141,159,167)
    L.cab[cab_off(L, c, CAB_TERM_P) + i] = p_term;
    L.cab[cab_off(L, c, CAB_TERM_I) + i] = pc.ki * (prev_ierr + dt * error);
    L.cab[cab_off(L, c, CAB_TERM_D) + i] = d_term;
    if (dt > 0.0) L.cab[cab_off(L, c, CAB_DESIRED) + i] = desired;
  }
  *cmdp = cmd_out;
  *last_time = now;
  return cmd_out;
}

#ifndef CDPR_GEN_BLOCKS
#define CDPR_GEN_BLOCKS 6
#endif
template <int DMAX>
__global__ void __launch_bounds__(kTpb, CDPR_GEN_BLOCKS) k_step_general(const __grid_constant__ StepArgs A) {
  const long long i = (long long)blockIdx.x * kTpb + threadIdx.x;
  if (i >= A.L.n) return;
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const long long np = L.np;
  const int nc = L.nc;
  FastState S;
  load_plat(L, i, S);
  double amp = 0.0, freq = 0.0, phase = 0.0;
  if (A.sine_on) { amp = L.sine[i]; freq = L.sine[np + i]; phase = L.sine[2 * np + i]; }
  const float *cmd_row = nullptr;
  if (A.cmd_table) cmd_row = A.cmd_table + (size_t)(i % A.n_seq) * A.n_cmd * nc;
  double cost = 0.0;
  int sec = A.sec0, nsec = A.nsec0;
  double sine_time = A.sine_time0;
  int sine_ctr = (int)(A.n0 % (A.sine_period > 0 ? A.sine_period : 1));
  int cmd_ctr = 0, cmd_idx = 0;
  long long snap_idx = A.snap_written0;
  long long snap_ctr = A.snap_every > 0 ? (A.n0 % A.snap_every) : 0;

  for (int s = 0; s < A.k_steps; ++s) {
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    const double now = time_double(sec, nsec);
    if (A.sine_on) {
      if (sine_ctr == 0) {
        const double arg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(sine_time, freq), 2.0), 3.14159265358979323846), phase);
        const double vel = publisher_value(A.pub_shape, amp, sin(arg));
        for (int c = 0; c < nc; ++c) L.cab[cab_off(L, c, CAB_VEL_TARGET) + i] = vel;
        sine_time = __dadd_rn(sine_time, A.sine_pub_dt);
      }
      sine_ctr = (sine_ctr + 1 == A.sine_period) ? 0 : sine_ctr + 1;
    }
    if (cmd_row) {
      if (cmd_ctr == 0 && cmd_idx < A.n_cmd) {
        for (int c = 0; c < nc; ++c) L.cab[cab_off(L, c, CAB_VEL_TARGET) + i] = (double)cmd_row[cmd_idx * nc + c];
        ++cmd_idx;
      }
      cmd_ctr = (cmd_ctr + 1 == A.steps_per_cmd) ? 0 : cmd_ctr + 1;
    }
    const Rot R = make_rot(S);
    double fx = rc.mg[0], fy = rc.mg[1], fz = rc.mg[2], mx = 0.0, my = 0.0, mz = 0.0;
    for (int c = 0; c < nc; ++c) {
      const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
      const double gx = rc.a[c][0] - S.px, gy = rc.a[c][1] - S.py, gz = rc.a[c][2] - S.pz;
      const double dx = fma(-R.r00, bx, fma(-R.r01, by, fma(-R.r02, bz, gx)));
      const double dy = fma(-R.r10, bx, fma(-R.r11, by, fma(-R.r12, bz, gy)));
      const double dz = fma(-R.r20, bx, fma(-R.r21, by, fma(-R.r22, bz, gz)));
      const double l2 = fma(dx, dx, fma(dy, dy, dz * dz));
      const double il = rsqrt_nr(l2);
      const double cx = fma(gy, dz, -(gz * dy)), cy = fma(gz, dx, -(gx * dz)), cz = fma(gx, dy, -(gy * dx));  // L (r x u), see step_fast.cuh
      const double qd = (fma(dx, S.vx, fma(dy, S.vy, dz * S.vz)) + fma(cx, S.wx, fma(cy, S.wy, cz * S.wz))) * il;
      const double qp = rc.home_len[c] - l2 * il;

      unsigned ctl = L.ctl[(long long)c * np + i];
      double *last_pos = L.cab + cab_off(L, c, CAB_LAST_POS) + i;
      // JointForceCalculator::update (.cpp:59-96): pick the Pid, its set point and its measurement, then ONE Pid update
      double force;
      if (A.mode == MODE_FORCE) {
        *last_pos = qp;
        force = L.cab[cab_off(L, c, CAB_FORCE_CMD) + i];
      } else {
        int k;
        double desired, actual;
        if (A.mode == MODE_VELOCITY) {
          const double vt = L.cab[cab_off(L, c, CAB_VEL_TARGET) + i];
          if (fabs(vt) > rc.vel_eps) { *last_pos = qp; k = PID_VEL; desired = vt; actual = qd; }
          else { k = PID_POS; desired = *last_pos; actual = qp; }  // hold the last position with the position Pid
        } else {
          *last_pos = qp; k = PID_POS; desired = L.cab[cab_off(L, c, CAB_POS_TARGET) + i]; actual = qp;
        }
        force = pid_update_general<DMAX>(A, L, A.pc[k], c, k, ctl, desired, actual, now, i);
      }
      L.ctl[(long long)c * np + i] = ctl;
      const double eff = (rc.effort_limit >= 0.0) ? clampd(force, -rc.effort_limit, rc.effort_limit) : force;
      L.cab[cab_off(L, c, CAB_EFFORT) + i] = eff;
      L.cab[cab_off(L, c, CAB_PID_FORCE) + i] = force;
      if (s == A.k_steps - 1) publish_joint(A, nc, c, qp, qd, eff, i);
      const double tl = fma(-rc.cdamp, qd, eff) * il;  // tension / L
      fx = fma(tl, dx, fx); fy = fma(tl, dy, fy); fz = fma(tl, dz, fz);
      mx = fma(tl, cx, mx); my = fma(tl, cy, my); mz = fma(tl, cz, mz);
    }
    if (s == A.k_steps - 1) publish_platform(A, S, i);
    if (rc.diag_inertia) rigid_body_step<SPEC_DIAG>(rc, S, R, fx, fy, fz, mx, my, mz);
    else rigid_body_step<0>(rc, S, R, fx, fy, fz, mx, my, mz);
    if (A.cost) {
      const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
      cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
    }
    if (A.snap_every > 0) {
      if (++snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (snap_idx < A.snap_capacity)
          for (int p = 0; p < A.n_snap_peers; ++p) store_plat(A.snap_peers[p] + snap_idx * 13 * A.snap_stride + A.snap_offset + i, A.snap_stride, S);
        ++snap_idx;
      }
    }
  }
  store_plat(L.plat + i, np, S);
  if (A.cost) A.cost[i] = cost;
}

}  // namespace cdpr
