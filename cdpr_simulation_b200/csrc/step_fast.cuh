// step_fast.cuh -- K2: the persistent K-step kernel (kinematics -> force law -> wrench -> rigid-body
// step), one thread per robot instance, state on chip for the whole launch.
//
// Covers reference rows a1,a3,a4,a5,a7,a8,a9,a11 of SURVEY.md 8(a) for the configurations in which
//   - velocity hold cannot trigger (velocityEpsilon < 0, launch/cdpr_gazebo.launch:18), so exactly one
//     Pid per cable is live (JointForceCalculator.cpp:71-89),
//   - no biquad stage is configured (cascade = 0, launch:29,32; CdprGazeboPlugin.cpp:133),
//   - cmdLimit != 0,
// which is the reference's launch configuration.  Everything else runs in step_general.cuh.
//
// On-chip residency:  platform state, integral errors, targets, flags    -> registers
//                     D-term error windows (Pid::mDbufferY, LEN per cable) -> shared memory,
//                     circular, [slot][cable][thread] so a warp reads 256 contiguous bytes
// The D-term is the reference's least-squares polynomial derivative (Pid.cpp:193-247) written as
// the equivalent fixed FIR over the window (uniform time stamps; weights from the host, fir[]).
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace cdpr {

// One physics step for one instance.  STEADY: every live Pid is primed and its window is full
// (Pid::mWasLastTime && mDbufferMissing == 0), so no flag logic is needed.
template <int NC, int LEN, bool STEADY, bool LAST>
__device__ __forceinline__ void fast_step(const StepArgs &A, FastState &S, double (&ierr)[NC], const double (&tgt)[NC],
                                          unsigned &primed, unsigned (&missing)[NC], double *__restrict__ win, int head,
                                          double dt, long long i) {
  const RobotConsts &rc = A.rc;
  const PidConsts &pc = A.live;
  const int mode = A.mode;
  const Rot R = make_rot(S);
  const double r00 = R.r00, r01 = R.r01, r02 = R.r02, r10 = R.r10, r11 = R.r11, r12 = R.r12, r20 = R.r20, r21 = R.r21, r22 = R.r22;

  double fx = rc.mg[0], fy = rc.mg[1], fz = rc.mg[2];
  double mx = 0.0, my = 0.0, mz = 0.0;

  // slot offsets of the samples by age (1 = previous step ... LEN-1 = oldest); warp-uniform
  int slot[LEN];
#pragma unroll
  for (int a = 0; a < LEN; ++a) {
    int s = head - a;
    s += (s < 0) ? LEN : 0;
    slot[a] = s * (NC * kTpb);
  }

#pragma unroll
  for (int c = 0; c < NC; ++c) {
    // ---- inverse kinematics (a7): r = R b, d = a - p - r, L, u, r x u, joint rate
    const double bx = rc.b[c][0], by = rc.b[c][1], bz = rc.b[c][2];
    const double rx = fma(r00, bx, fma(r01, by, r02 * bz));
    const double ry = fma(r10, bx, fma(r11, by, r12 * bz));
    const double rz = fma(r20, bx, fma(r21, by, r22 * bz));
    const double dx = (rc.a[c][0] - S.px) - rx, dy = (rc.a[c][1] - S.py) - ry, dz = (rc.a[c][2] - S.pz) - rz;
    const double l2 = fma(dx, dx, fma(dy, dy, dz * dz));
    const double il = rsqrt_nr(l2);
    const double len = l2 * il;
    const double ux = dx * il, uy = dy * il, uz = dz * il;
    const double cx = fma(ry, uz, -(rz * uy)), cy = fma(rz, ux, -(rx * uz)), cz = fma(rx, uy, -(ry * ux));
    const double qd = fma(ux, S.vx, fma(uy, S.vy, fma(uz, S.vz, fma(cx, S.wx, fma(cy, S.wy, cz * S.wz)))));
    const double qp = rc.home_len[c] - len;

    // ---- force law (a3, a4, a5)
    double force;
    if (mode == MODE_FORCE) {
      force = tgt[c];  // JointForceCalculator.cpp:67-70
      if (LAST) A.L.cab[cab_off(A.L, c, CAB_LAST_POS) + i] = qp;
    } else {
      const double e = tgt[c] - ((mode == MODE_VELOCITY) ? qd : qp);
      double *w = win + c * kTpb;
      if (STEADY || ((primed >> c) & 1u)) {  // Pid.cpp:127-187
        const double prev_ierr = ierr[c];
        double ie = fma(dt, e, prev_ierr);
        double iterm = pc.ki * ie;
        // i_min = -i_max, cmd_min = -cmd_max by construction (Pid.cpp:70-73): one compare per clamp
        const bool isat = fabs(iterm) > pc.i_max;
        iterm = isat ? copysign(pc.i_max, iterm) : iterm;
        ie = isat ? ((__double2hiint(iterm) < 0) ? pc.i_min_over_ki : pc.i_max_over_ki) : ie;
        // derive(): push, then LS derivative at `now` = FIR over the window (Pid.cpp:193-217)
        w[slot[0]] = e;
        double d0 = A.fir[LEN - 1] * e, d1 = 0.0;
#pragma unroll
        for (int a = 1; a < LEN; ++a) {
          if (a & 1) d1 = fma(A.fir[LEN - 1 - a], w[slot[a]], d1);
          else d0 = fma(A.fir[LEN - 1 - a], w[slot[a]], d0);
        }
        double derr = d0 + d1;
        if (!STEADY) {
          missing[c] -= (missing[c] > 0u) ? 1u : 0u;
          if (missing[c] != 0u) derr = 0.0;
        }
        const double cmd_raw = fma(pc.kd, derr, fma(pc.kp, e, pc.kf * tgt[c]) + iterm);
        // clamp + anti-windup (Pid.cpp:175-184): mCmd != cmd  <=>  |cmd| > cmdMax
        const bool csat = fabs(cmd_raw) > pc.cmd_max;
        const double cmd = csat ? fma(dt * e, pc.ki, copysign(pc.cmd_max, cmd_raw)) : cmd_raw;
        ie = csat ? prev_ierr : ie;
        ierr[c] = ie;
        force = cmd;
        if (LAST) {
          A.L.pid[pid_off(A.L, c, A.live_idx, PID_P_ERR) + i] = e;
          A.L.pid[pid_off(A.L, c, A.live_idx, PID_D_ERR) + i] = derr;
        }
      } else {  // first update after a reset: Pid.cpp:123-126
        primed |= 1u << c;
        force = 0.0;
      }
      if (LAST) {
        A.L.pid[pid_off(A.L, c, A.live_idx, PID_CMD) + i] = force;
        A.L.cab[cab_off(A.L, c, CAB_LAST_POS) + i] = qp;
      }
    }
    // ---- Joint::SetForce truncation, explicit joint damping, wrench (a8)
    const double eff = (fabs(force) > rc.effort_limit_abs) ? copysign(rc.effort_limit_abs, force) : force;
    if (LAST) {
      A.L.cab[cab_off(A.L, c, CAB_EFFORT) + i] = eff;
      A.L.cab[cab_off(A.L, c, CAB_PID_FORCE) + i] = force;
    }
    const double tau = fma(-rc.cdamp, qd, eff);
    fx = fma(tau, ux, fx); fy = fma(tau, uy, fy); fz = fma(tau, uz, fz);
    mx = fma(tau, cx, mx); my = fma(tau, cy, my); mz = fma(tau, cz, mz);
  }

  rigid_body_step(rc, S, R, fx, fy, fz, mx, my, mz);
}

template <int NC, int LEN>
__global__ void __launch_bounds__(kTpb, (NC <= 4) ? 4 : 2) k_step_fast(const __grid_constant__ StepArgs A) {
  extern __shared__ double win[];  // [LEN][NC][kTpb]
  const int tid = threadIdx.x;
  const long long gi = (long long)blockIdx.x * kTpb + tid;
  const bool valid = gi < A.L.n;
  const long long i = valid ? gi : (long long)A.L.n - 1;  // tail threads shadow the last instance, never store
  const long long np = A.L.np;
  const int live = A.live_idx;
  double *mywin = win + tid;

  FastState S;
  load_plat(A.L, i, S);
  double ierr[NC], tgt[NC];
  unsigned primed = 0, missing[NC];
  const int tgt_field = (A.mode == MODE_FORCE) ? CAB_FORCE_CMD : (A.mode == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
  bool steady = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ierr[c] = A.L.pid[pid_off(A.L, c, live, PID_I_ERR) + i];
    tgt[c] = A.L.cab[cab_off(A.L, c, tgt_field) + i];
    const unsigned ctl = A.L.ctl[(long long)c * np + i];
    primed |= ((ctl >> live) & 1u) << c;
    missing[c] = (ctl >> (8 + 8 * live)) & 0xffu;
    steady = steady && ((ctl >> live) & 1u) && missing[c] == 0u;
#pragma unroll
    for (int j = 0; j < LEN; ++j)  // logical j -> slot j; newest (j = LEN-1) sits at head = LEN-1
      mywin[(j * NC + c) * kTpb] = A.L.win_y[win_off(A.L, c, live, j) + i];
  }
  double amp = 0.0, freq = 0.0, phase = 0.0;
  if (A.sine_on) { amp = A.L.sine[i]; freq = A.L.sine[np + i]; phase = A.L.sine[2 * np + i]; }
  const float *cmd_row = nullptr;
  if (A.cmd_table) cmd_row = A.cmd_table + (size_t)(i % A.n_seq) * A.n_cmd * NC;
  double cost = 0.0;

  bool warp_steady = __all_sync(0xffffffffu, steady || A.mode == MODE_FORCE);
  int sec = A.sec0, nsec = A.nsec0, head = LEN - 1;
  double tprev = A.t0, sine_time = A.sine_time0;
  int sine_ctr = (int)(A.n0 % (A.sine_period > 0 ? A.sine_period : 1));
  int cmd_ctr = 0, cmd_idx = 0;
  long long n = A.n0;
  long long snap_idx = A.snap_written0;
  long long snap_ctr = A.snap_every > 0 ? (A.n0 % A.snap_every) : 0;

  for (int s = 0; s < A.k_steps; ++s) {
    // World::Step: simTime += dt, then the plugin callback
    ++n;
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    const double t = time_double(sec, nsec);
    const double dt = __dsub_rn(t, tprev);
    tprev = t;
    if (A.sine_on) {  // sinevelocitytest.cpp:35-38,48: float32 axes, accumulated publisher time
      if (sine_ctr == 0) {
        const double arg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(sine_time, freq), 2.0), 3.14159265358979323846), phase);
        const double vel = (double)(float)__dmul_rn(amp, sin(arg));
#pragma unroll
        for (int c = 0; c < NC; ++c) tgt[c] = vel;
        sine_time = __dadd_rn(sine_time, A.sine_pub_dt);
      }
      sine_ctr = (sine_ctr + 1 == A.sine_period) ? 0 : sine_ctr + 1;
    }
    if (cmd_row) {
      if (cmd_ctr == 0 && cmd_idx < A.n_cmd) {
#pragma unroll
        for (int c = 0; c < NC; ++c) tgt[c] = (double)cmd_row[cmd_idx * NC + c];
        ++cmd_idx;
      }
      cmd_ctr = (cmd_ctr + 1 == A.steps_per_cmd) ? 0 : cmd_ctr + 1;
    }
    head = (head + 1 == LEN) ? 0 : head + 1;
    if (s + 1 < A.k_steps) {
      if (warp_steady) {
        fast_step<NC, LEN, true, false>(A, S, ierr, tgt, primed, missing, mywin, head, dt, i);
      } else {
        fast_step<NC, LEN, false, false>(A, S, ierr, tgt, primed, missing, mywin, head, dt, i);
      }
    } else if (valid) {  // the last step also publishes effort / Pid telemetry columns
      if (warp_steady) fast_step<NC, LEN, true, true>(A, S, ierr, tgt, primed, missing, mywin, head, dt, i);
      else fast_step<NC, LEN, false, true>(A, S, ierr, tgt, primed, missing, mywin, head, dt, i);
    }
    if (!warp_steady) {
      bool st = true;
#pragma unroll
      for (int c = 0; c < NC; ++c) st = st && ((primed >> c) & 1u) && missing[c] == 0u;
      warp_steady = __all_sync(0xffffffffu, st);
    }
    if (A.cost) {
      const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
      cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
    }
    if (A.snap_every > 0) {
      if (++snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (valid && snap_idx < A.snap_capacity) {
          store_plat(A.snap + snap_idx * 13 * (long long)A.L.n + i, A.L.n, S);
        }
        ++snap_idx;
      }
    }
  }

  if (!valid) return;
  store_plat(A.L.plat + i, np, S);
  if (A.cost) A.cost[i] = cost;
  if (A.mode == MODE_FORCE) return;  // Force mode touches no Pid state
  // after the loop the newest sample sits in slot `head`; logical j lives in slot (head + 1 + j) % LEN
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    A.L.pid[pid_off(A.L, c, live, PID_I_ERR) + i] = ierr[c];
    A.L.pid[pid_off(A.L, c, live, PID_LAST_TIME) + i] = tprev;
    if (A.sine_on || A.cmd_table) A.L.cab[cab_off(A.L, c, CAB_VEL_TARGET) + i] = tgt[c];
    unsigned ctl = A.L.ctl[(long long)c * np + i];
    ctl &= ~((1u << live) | (0xffu << (8 + 8 * live)));
    ctl |= (((primed >> c) & 1u) << live) | (missing[c] << (8 + 8 * live));
    A.L.ctl[(long long)c * np + i] = ctl;
    for (int j = 0; j < LEN; ++j) {
      int sl = head + 1 + j;
      sl -= (sl >= LEN) ? LEN : 0;
      A.L.win_y[win_off(A.L, c, live, j) + i] = mywin[(sl * NC + c) * kTpb];
    }
  }
}

}  // namespace cdpr
