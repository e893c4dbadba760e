// step_flex.cuh -- K2'': the persistent K-step kernel with the FULL per-robot semantics of the plugin, state on chip.
//
// What the fast kernel (step_fast.cuh) leaves out, this one carries, still with one thread per robot instance and no
// per-step HBM traffic:
//   * velocity hold: below velocityEpsilon a cable runs the POSITION Pid on its latched position
//     (JointForceCalculator.cpp:72-82) -- the Pid that runs is chosen per cable and per step;
//   * biquad cascades on the P input and the D output of either Pid (Pid.cpp:27-44,133,157; Filter.h:152-165);
//   * per-INSTANCE mode and command latching: every robot has its own UpdateMode and its own pending velocity /
//     position command (CdprGazeboPlugin.cpp:67-83,206-219: each plugin instance latches its own messages), so the robots
//     of one batch can be commanded independently (cdpr_set_*_cmd_masked);
//   * the exact clamp chain of Pid::update, statement by statement (no algebraic rewrite; any sign of iGain).
//
// On-chip residency per instance:
//   registers       platform state (13), per cable: integral error of the LIVE Pid, control word
//   shared memory   per cable: the LIVE Pid's D-term window (ring of 11, slot = step index mod 11 as in the fast kernel),
//                   its biquad state, its last update time; the latched hold position; the target of the instance's mode
// "Live" = the Pid that ran in the previous step.  The other Pid of a cable sleeps in HBM.  When a cable changes Pid
// (hold begins / ends, mode switch) the live state is flushed and the other Pid's is loaded: a rare, per-thread branch.
//
// D-term windows and gaps.  Pid::derive keeps the last 11 (time, error) pairs it was GIVEN, so after a sleep the window
// of the Pid that wakes up spans a gap and the least-squares fit runs on non-uniform time stamps (Pid.cpp:193-247) until
// 11 new samples have pushed the stale ones out.  `fresh` counts the consecutive samples since the Pid went live:
//   fresh >= 11   the window is the last 11 steps: fixed FIR over the shared-memory ring (as in the fast kernel);
//   fresh <  11   the window still holds stale samples: the HBM ring of that Pid (time stamps + errors, head in the control
//                 word) is kept current -- one push per step -- and the fit is the general one-pass window-relative
//                 least squares of step_general.cuh over it.  At most 11 steps per wake-up.
// A flush with fresh >= 11 rewrites the whole HBM ring from the shared-memory ring (stamps recomputed from the step
// index); with fresh < 11 the HBM ring is already current.  Load and flush are exact inverses, so results do not depend
// on how the steps are cut into launches.
//
// Preconditions (cdpr_create picks this variant when they hold, else step_general.cuh): NC in {4, 8}; both windows 11
// samples long and fitted with the same degree (one FIR); cmdLimit != 0 for both Pids; the cascades fit in shared memory.
#pragma once
#include "common.cuh"
#include "physics.cuh"
#include "step_general.cuh"

namespace cdpr {

constexpr int kFlexLen = 11;

// control word, per (instance, cable): the general variant's layout (step_general.cuh) plus
//   bits 24-27  fresh: consecutive samples since the live Pid woke up, saturating at 11
//   bits 28-29  live: 0 none, 1 velocity Pid, 2 position Pid
__device__ __forceinline__ unsigned fctl_fresh(unsigned ctl) { return (ctl >> 24) & 0xfu; }
__device__ __forceinline__ unsigned fctl_live(unsigned ctl) { return (ctl >> 28) & 0x3u; }
__device__ __forceinline__ unsigned fctl_set_fresh(unsigned ctl, unsigned f) { return (ctl & ~(0xfu << 24)) | (f << 24); }
__device__ __forceinline__ unsigned fctl_set_live(unsigned ctl, unsigned l) { return (ctl & ~(0x3u << 28)) | (l << 28); }

// instance word ictl[i]: bits 0-1 UpdateMode, bit 2 velocity command pending, bit 3 position command pending
enum { ICTL_VEL_PENDING = 4u, ICTL_POS_PENDING = 8u };

// shared memory per block, in doubles: ring [11][NC][T], filters [NC][FS][T], last_pos [NC][T], target [NC][T],
// last_time [NC][T], sine [3][T]; FS = 4 * (P stages + D stages)
__host__ __device__ inline size_t flex_smem_doubles(int nc, int ps, int ds, int tpb) {
  return (size_t)tpb * ((size_t)kFlexLen * nc + (size_t)nc * 4 * (ps + ds) + 3 * (size_t)nc + 3);
}

// gazebo::common::Time of the step `back` steps before (sec, nsec)
__device__ __forceinline__ double stamp_back(int sec, int nsec, int dt_ns, int back) {
  long long ns = (long long)sec * 1000000000LL + nsec - (long long)back * dt_ns;
  return time_double((int)(ns / 1000000000LL), (int)(ns % 1000000000LL));
}

// ---- rare paths, out of line, scalar arguments only (nothing of the caller's register state escapes) ----------------
// Flush the live Pid `k` of cable `c` to HBM: integral, last update time, biquad state, and -- when the window is entirely
// fresh -- the window itself in logical order from the ring head on.
static __device__ __noinline__ void flex_flush(const StepArgs &A, int c, int k, unsigned ctl, double ierr, const double *ring, const double *filt,
                                               double last_time, int T, int nc, int slot_now, int sec, int nsec, long long i) {
  const DevLayout &L = A.L;
  L.pid[pid_off(L, c, k, PID_I_ERR) + i] = ierr;
  L.pid[pid_off(L, c, k, PID_LAST_TIME) + i] = last_time;
  const int ps = A.flex_ps, ds = A.flex_ds;
  for (int s = 0; s < ps; ++s)
    for (int f = 0; f < 4; ++f) L.filt[filt_off(L, c, k, 0, s, f) + i] = filt[(s * 4 + f) * T];
  for (int s = 0; s < ds; ++s)
    for (int f = 0; f < 4; ++f) L.filt[filt_off(L, c, k, 1, s, f) + i] = filt[((ps + s) * 4 + f) * T];
  if (fctl_fresh(ctl) >= (unsigned)kFlexLen) {
    int hd = (int)gctl_head(ctl, k);  // oldest slot of the HBM ring; unchanged by a full rewrite
    for (int j = 0; j < kFlexLen; ++j) {
      const int age = kFlexLen - 1 - j;
      int sl = slot_now - age;
      sl += (sl < 0) ? kFlexLen : 0;
      L.win_y[win_off(L, c, k, hd) + i] = ring[(sl * nc) * T];
      L.win_x[win_off(L, c, k, hd) + i] = stamp_back(sec, nsec, A.dt_ns, age);
      hd = (hd + 1 == kFlexLen) ? 0 : hd + 1;
    }
  }
}

// Wake Pid `k` of cable `c`: biquad state and last update time into shared memory; returns its integral error.
static __device__ __noinline__ double flex_wake(const StepArgs &A, int c, int k, double *filt, double *last_time, int T, long long i) {
  const DevLayout &L = A.L;
  const int ps = A.flex_ps, ds = A.flex_ds;
  for (int s = 0; s < ps; ++s)
    for (int f = 0; f < 4; ++f) filt[(s * 4 + f) * T] = L.filt[filt_off(L, c, k, 0, s, f) + i];
  for (int s = 0; s < ds; ++s)
    for (int f = 0; f < 4; ++f) filt[((ps + s) * 4 + f) * T] = L.filt[filt_off(L, c, k, 1, s, f) + i];
  *last_time = L.pid[pid_off(L, c, k, PID_LAST_TIME) + i];
  return L.pid[pid_off(L, c, k, PID_I_ERR) + i];
}

// The window of a live Pid whose samples are all fresh, HBM ring (logical order) -> shared-memory ring (step-aligned).
static __device__ __noinline__ void flex_load_window(const StepArgs &A, int c, int k, unsigned ctl, double *ring, int T, int nc, int slot_now, long long i) {
  const DevLayout &L = A.L;
  int hd = (int)gctl_head(ctl, k);
  for (int j = 0; j < kFlexLen; ++j) {
    const int age = kFlexLen - 1 - j;
    int sl = slot_now - age;
    sl += (sl < 0) ? kFlexLen : 0;
    ring[(sl * nc) * T] = L.win_y[win_off(L, c, k, hd) + i];
    hd = (hd + 1 == kFlexLen) ? 0 : hd + 1;
  }
}

// Pid::reset (Pid.cpp:100-115) of a SLEEPING Pid `k` of cable `c` (the live one is reset on chip by the caller);
// mLastTime is kept.  Returns the control word with wasLast cleared, missing = 11, ring head 0.
static __device__ __noinline__ unsigned flex_reset_sleeping(const StepArgs &A, int c, int k, unsigned ctl, long long i) {
  const DevLayout &L = A.L;
  L.pid[pid_off(L, c, k, PID_P_ERR) + i] = 0.0;
  L.pid[pid_off(L, c, k, PID_I_ERR) + i] = 0.0;
  L.pid[pid_off(L, c, k, PID_D_ERR) + i] = 0.0;
  L.pid[pid_off(L, c, k, PID_CMD) + i] = 0.0;
  if (L.filt)
    for (int pd = 0; pd < 2; ++pd)
      for (int s = 0; s < L.casc; ++s)
        for (int f = 0; f < 4; ++f) L.filt[filt_off(L, c, k, pd, s, f) + i] = 0.0;
  ctl &= ~(1u << k);
  return gctl_set(ctl, k, (unsigned)kFlexLen, 0u);
}

// least-squares derivative over the HBM ring of a window that spans a gap (degree from the Pid's parameters)
static __device__ __noinline__ double flex_gap_fit(const StepArgs &A, int c, int k, unsigned oldest, double now, long long i) {
  const int deg = A.pc[k].degree;
  if (deg == 1) return ls_derivative<1>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 2) return ls_derivative<2>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 3) return ls_derivative<3>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 4) return ls_derivative<4>(A.L, c, k, kFlexLen, oldest, now, i);
  return 0.0;
}

// Pid::CascadeFilter::update over shared-memory biquad state (Pid.cpp:38-44, Filter.h:152-165)
__device__ __forceinline__ double flex_cascade(double *st, int stages, const double *co0, const double *co1, bool second, double x, int T) {
  // the coefficients of the Pid that runs, picked value by value (a per-thread pointer into the kernel parameters
  // would turn every use into a generic load)
  const double a0 = second ? co1[0] : co0[0], a1 = second ? co1[1] : co0[1], a2 = second ? co1[2] : co0[2];
  const double b1 = second ? co1[3] : co0[3], b2 = second ? co1[4] : co0[4];
  double out = x;
  for (int s = 0; s < stages; ++s) {
    double *q = st + (s * 4) * T;
    const double x1 = q[0], x2 = q[T], y1 = q[2 * T], y2 = q[3 * T];
    const double y0 = a0 * out + a1 * x1 + a2 * x2 - b1 * y1 - b2 * y2;
    q[T] = x1; q[0] = out; q[3 * T] = y1; q[2 * T] = y0;
    out = y0;
  }
  return out;
}

#ifndef CDPR_FLEX_MAXTPB
#define CDPR_FLEX_MAXTPB 128
#endif

template <int NC>
__global__ void __launch_bounds__(CDPR_FLEX_MAXTPB) k_step_flex(const __grid_constant__ StepArgs A) {
  extern __shared__ double smem[];
  const int T = (int)blockDim.x, tid = (int)threadIdx.x;
  const long long i = (long long)blockIdx.x * T + tid;
  if (i >= A.L.n) return;  // no block-level synchronisation anywhere below
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const long long np = L.np;
  const int PS = A.flex_ps, DS = A.flex_ds, FS = 4 * (PS + DS);
  double *ring = smem + tid;                          // [11][NC][T]
  double *filt = smem + kFlexLen * NC * T + tid;      // [NC][FS][T]: P stages, then D stages; x1 x2 y1 y2 each
  double *lastp = filt + NC * FS * T;                 // [NC][T]  JointForceCalculator::mLastPosition
  double *tgt = lastp + NC * T;                       // [NC][T]  target of the instance's mode
  double *ltime = tgt + NC * T;                       // [NC][T]  Pid::mLastTime of the live Pid
  double *sinep = ltime + NC * T;                     // [3][T]

  FastState S;
  load_plat(L, i, S);
  unsigned ictl = L.ictl[i];
  int mode = (int)(ictl & 3u);
  const bool vel_pending0 = (ictl & ICTL_VEL_PENDING) != 0u, pos_pending0 = (ictl & ICTL_POS_PENDING) != 0u;
  const int head0 = (int)(A.n0 % kFlexLen);  // ring slot of the newest sample already in the windows
  double ierr[NC];
  unsigned ctl[NC];
  auto load_targets = [&](int m) {
    const int field = (m == MODE_FORCE) ? CAB_FORCE_CMD : (m == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
#pragma unroll
    for (int c = 0; c < NC; ++c) tgt[c * T] = L.cab[cab_off(L, c, field) + i];
  };
  load_targets(mode);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ctl[c] = L.ctl[(long long)c * np + i];
    lastp[c * T] = L.cab[cab_off(L, c, CAB_LAST_POS) + i];
    ierr[c] = 0.0;
    ltime[c * T] = 0.0;
    const unsigned live = fctl_live(ctl[c]);
    if (live != 0u) {
      ierr[c] = flex_wake(A, c, (int)live - 1, filt + c * FS * T, ltime + c * T, T, i);
      if (fctl_fresh(ctl[c]) >= (unsigned)kFlexLen) flex_load_window(A, c, (int)live - 1, ctl[c], ring + c * T, T, NC, head0, i);
    }
  }
  if (A.sine_on) {
#pragma unroll
    for (int m = 0; m < 3; ++m) sinep[m * T] = L.sine[m * np + i];
  }
  const float *cmd_row = nullptr;
  if (A.cmd_table) cmd_row = A.cmd_table + (size_t)(i % A.n_seq) * A.n_cmd * NC;
  double cost = 0.0;
  int sec = A.sec0, nsec = A.nsec0, head = head0;
  double sine_time = A.sine_time0;
  int sine_ctr = (int)(A.n0 % (A.sine_period > 0 ? A.sine_period : 1));
  int cmd_ctr = 0, cmd_idx = 0;
  long long snap_idx = A.snap_written0;
  long long snap_ctr = A.snap_every > 0 ? (A.n0 % A.snap_every) : 0;

  // Pid::reset of Pid k on every cable (setVelocityTarget / setPositionTarget on a mode change, JointForceCalculator.cpp:99-119)
  auto reset_pid = [&](int k) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (fctl_live(ctl[c]) == (unsigned)(k + 1)) {  // the live Pid: its state is on chip
        ierr[c] = 0.0;
        for (int f = 0; f < FS; ++f) filt[(c * FS + f) * T] = 0.0;
        ctl[c] = fctl_set_fresh(ctl[c], 0u);
      }
      ctl[c] = flex_reset_sleeping(A, c, k, ctl[c], i);  // HBM copy, wasLast, missing, ring head (also valid for the live one)
    }
  };

  for (int s = 0; s < A.k_steps; ++s) {
    const bool last = (s == A.k_steps - 1);
    // World::Step: simTime += dt, then the plugin callback
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    const double now = time_double(sec, nsec);
    head = (head + 1 == kFlexLen) ? 0 : head + 1;

    // ---- CdprGazeboPlugin::update, .cpp:206-219: velocity fan-out, then position fan-out
    bool vel_cmd = (s == 0) && vel_pending0;
    if (vel_cmd && mode != MODE_VELOCITY) load_targets(MODE_VELOCITY);
    if (A.sine_on) {
      if (sine_ctr == 0) {  // sinevelocitytest.cpp:35-38,48: float32 axes, accumulated publisher time
        const double arg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(sine_time, sinep[T]), 2.0), 3.14159265358979323846), sinep[2 * T]);
        const double vel = (double)(float)__dmul_rn(sinep[0], sin(arg));
#pragma unroll
        for (int c = 0; c < NC; ++c) tgt[c * T] = vel;
        sine_time = __dadd_rn(sine_time, A.sine_pub_dt);
        vel_cmd = true;
      }
      sine_ctr = (sine_ctr + 1 == A.sine_period) ? 0 : sine_ctr + 1;
    }
    if (cmd_row) {
      if (cmd_ctr == 0 && cmd_idx < A.n_cmd) {
#pragma unroll
        for (int c = 0; c < NC; ++c) tgt[c * T] = (double)cmd_row[cmd_idx * NC + c];
        ++cmd_idx;
        vel_cmd = true;
      }
      cmd_ctr = (cmd_ctr + 1 == A.steps_per_cmd) ? 0 : cmd_ctr + 1;
    }
    if (vel_cmd) {
      if (mode != MODE_VELOCITY) reset_pid(PID_VEL);
      mode = MODE_VELOCITY;
    }
    if (s == 0 && pos_pending0) {
      if (vel_cmd) {  // the velocity targets just latched must survive in HBM before the position targets replace them on chip
#pragma unroll
        for (int c = 0; c < NC; ++c) L.cab[cab_off(L, c, CAB_VEL_TARGET) + i] = tgt[c * T];
      }
      if (mode != MODE_POSITION) reset_pid(PID_POS);
      mode = MODE_POSITION;
      load_targets(MODE_POSITION);
    }

    const Rot R = make_rot(S);
    double fx = rc.mg[0], fy = rc.mg[1], fz = rc.mg[2], mx = 0.0, my = 0.0, mz = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const CableKin kin = cable_kin<0, true>(rc, S, R, c);
      // ---- JointForceCalculator::update (.cpp:59-96): which Pid runs, its set point and its measurement
      unsigned run = 0u;  // 0 none (Force mode), 1 velocity Pid, 2 position Pid
      double desired = 0.0, actual = 0.0, force = 0.0;
      const double target = tgt[c * T];
      if (mode == MODE_FORCE) {
        lastp[c * T] = kin.qp;
        force = target;
      } else if (mode == MODE_POSITION) {
        lastp[c * T] = kin.qp;
        run = 2u; desired = target; actual = kin.qp;
      } else if (fabs(target) > rc.vel_eps) {
        lastp[c * T] = kin.qp;
        run = 1u; desired = target; actual = kin.qd;
      } else {  // hold the last position with the position Pid
        run = 2u; desired = lastp[c * T]; actual = kin.qp;
      }
      unsigned w = ctl[c];
      if (fctl_live(w) != run) {  // rare: this cable changes Pid
        const unsigned live = fctl_live(w);
        int slot_prev = head - 1;
        slot_prev += (slot_prev < 0) ? kFlexLen : 0;
        // the ring's newest sample belongs to the PREVIOUS step (this step's has not been pushed yet)
        if (live != 0u)
          flex_flush(A, c, (int)live - 1, w, ierr[c], ring + c * T, filt + c * FS * T, ltime[c * T], T, NC, slot_prev, sec, nsec - A.dt_ns, i);
        if (run != 0u) ierr[c] = flex_wake(A, c, (int)run - 1, filt + c * FS * T, ltime + c * T, T, i);
        w = fctl_set_fresh(fctl_set_live(w, run), 0u);
      }
      if (run != 0u) {
        const int k = (int)run - 1;
        const bool pos = (k == PID_POS);
        if (!((w >> k) & 1u)) {  // first update after a reset: Pid.cpp:123-126
          w |= 1u << k;
          force = 0.0;
          if (last) L.pid[pid_off(L, c, k, PID_CMD) + i] = 0.0;
        } else {  // Pid.cpp:127-187
          const double kf = pos ? A.pc[1].kf : A.pc[0].kf, kp = pos ? A.pc[1].kp : A.pc[0].kp;
          const double ki = pos ? A.pc[1].ki : A.pc[0].ki, kd = pos ? A.pc[1].kd : A.pc[0].kd;
          const double i_max = pos ? A.pc[1].i_max : A.pc[0].i_max, i_min = pos ? A.pc[1].i_min : A.pc[0].i_min;
          const double c_max = pos ? A.pc[1].cmd_max : A.pc[0].cmd_max, c_min = pos ? A.pc[1].cmd_min : A.pc[0].cmd_min;
          const double f_term = kf * desired;
          const double e = desired - actual;
          const double dt = now - ltime[c * T];
          double pe = e;
          if (PS > 0) {
            const int st = pos ? A.pc[1].p_casc : A.pc[0].p_casc;
            if (st > 0) pe = flex_cascade(filt + c * FS * T, st, A.pc[0].pf, A.pc[1].pf, pos, e, T);
          }
          const double p_term = kp * pe;
          const double prev_ierr = ierr[c];
          double ie = fma(dt, e, prev_ierr);
          double i_term = ki * ie;
          const double i_term_pre = i_term;
          if (i_term > i_max) { i_term = i_max; ie = i_term / ki; }
          else if (i_term < i_min) { i_term = i_min; ie = i_term / ki; }
          // ---- derive (Pid.cpp:193-217): dt > 0 always (sim time advances every step)
          ring[(head * NC + c) * T] = e;
          unsigned fresh = fctl_fresh(w), missing = gctl_missing(w, k), hd = gctl_head(w, k);
          fresh += (fresh < (unsigned)kFlexLen) ? 1u : 0u;
          missing -= (missing > 0u) ? 1u : 0u;
          if (fresh < (unsigned)kFlexLen) {  // the window still holds older samples: keep the HBM ring current
            L.win_x[win_off(L, c, k, (int)hd) + i] = now;
            L.win_y[win_off(L, c, k, (int)hd) + i] = e;
            hd = (hd + 1u == (unsigned)kFlexLen) ? 0u : hd + 1u;
          }
          w = fctl_set_fresh(gctl_set(w, k, missing, hd), fresh);
          double derived = 0.0;
          if (missing == 0u && A.pc[0].degree >= 1) {  // both Pids fit the same degree in this variant
            if (fresh >= (unsigned)kFlexLen) {  // the last 11 steps: fixed FIR (weights oldest first)
              double d0 = A.fir[kFlexLen - 1] * e, d1 = 0.0;
#pragma unroll
              for (int a = 1; a < kFlexLen; ++a) {
                int sl = head - a;
                sl += (sl < 0) ? kFlexLen : 0;
                const double y = ring[(sl * NC + c) * T];
                if (a & 1) d1 = fma(A.fir[kFlexLen - 1 - a], y, d1); else d0 = fma(A.fir[kFlexLen - 1 - a], y, d0);
              }
              derived = d0 + d1;
            } else {
              derived = flex_gap_fit(A, c, k, hd, now, i);
            }
          }
          double de = derived;
          if (DS > 0) {
            const int st = pos ? A.pc[1].d_casc : A.pc[0].d_casc;
            if (st > 0) de = flex_cascade(filt + (c * FS + 4 * PS) * T, st, A.pc[0].df, A.pc[1].df, pos, derived, T);
          }
          const double d_term = kd * de;
          const double cmd_raw = f_term + p_term + i_term + d_term;
          double cmd = clampd(cmd_raw, c_min, c_max);  // cmdMax > cmdMin in this variant
          if (cmd != cmd_raw) {  // Pid.cpp:181-184
            ie = prev_ierr;
            cmd += dt * e * ki;
          }
          ierr[c] = ie;
          force = cmd;
          if (last) {
            L.pid[pid_off(L, c, k, PID_P_ERR) + i] = pe;
            L.pid[pid_off(L, c, k, PID_D_ERR) + i] = de;
            L.pid[pid_off(L, c, k, PID_CMD) + i] = cmd;
            L.cab[cab_off(L, c, CAB_TERM_P) + i] = p_term;
            L.cab[cab_off(L, c, CAB_TERM_I) + i] = i_term_pre;
            L.cab[cab_off(L, c, CAB_TERM_D) + i] = d_term;
            L.cab[cab_off(L, c, CAB_DESIRED) + i] = desired;
          }
        }
        ltime[c * T] = now;
      }
      ctl[c] = w;
      const double eff = (rc.effort_limit >= 0.0) ? clampd(force, -rc.effort_limit, rc.effort_limit) : force;
      if (last) {
        L.cab[cab_off(L, c, CAB_EFFORT) + i] = eff;
        L.cab[cab_off(L, c, CAB_PID_FORCE) + i] = force;
      }
      const double tl = fma(-rc.cdamp, kin.qd, eff) * kin.il;  // tension / L
      fx = fma(tl, kin.dx, fx); fy = fma(tl, kin.dy, fy); fz = fma(tl, kin.dz, fz);
      mx = fma(tl, kin.cx, mx); my = fma(tl, kin.cy, my); mz = fma(tl, kin.cz, mz);
    }
    if (rc.diag_inertia) rigid_body_step<SPEC_DIAG>(rc, S, R, fx, fy, fz, mx, my, mz);
    else rigid_body_step<0>(rc, S, R, fx, fy, fz, mx, my, mz);
    if (A.cost) {
      const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
      cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
    }
    if (A.snap_every > 0) {
      if (++snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (snap_idx < A.snap_capacity) write_snapshot(A, S, snap_idx * 13 * A.snap_stride + A.snap_offset + i);
        ++snap_idx;
      }
    }
  }

  // ---- back to HBM
  store_plat(L.plat + i, np, S);
  if (A.cost) A.cost[i] = cost;
  L.ictl[i] = (unsigned)mode;  // pending commands were consumed by the first step
  const int tgt_field = (mode == MODE_FORCE) ? CAB_FORCE_CMD : (mode == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const unsigned live = fctl_live(ctl[c]);
    if (live != 0u) flex_flush(A, c, (int)live - 1, ctl[c], ierr[c], ring + c * T, filt + c * FS * T, ltime[c * T], T, NC, head, sec, nsec, i);
    L.ctl[(long long)c * np + i] = ctl[c];
    L.cab[cab_off(L, c, CAB_LAST_POS) + i] = lastp[c * T];
    if (A.k_steps > 0) L.cab[cab_off(L, c, tgt_field) + i] = tgt[c * T];
  }
}

}  // namespace cdpr
