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
//   registers       platform state (13)
//   shared memory   per cable: the LIVE Pid's D-term window (ring of 11, slot = step index mod 11 as in the fast kernel),
//                   its biquad state, integral error and last update time; the latched hold position; the target of the
//                   instance's mode; the control words
// "Live" = the Pid that ran in the previous step.  The other Pid of a cable sleeps in HBM.  When a cable changes Pid
// (hold begins / ends, mode switch) the live state is flushed and the other Pid's is loaded.
//
// Two bodies, chosen per THREAD and per step from the thread's own state (so results never depend on warp-mates or on how
// the steps are cut into launches), bit-identical in their arithmetic (same inlined helpers):
//   hot body      every cable runs the Pid that is live, primed, with a full window of the last 11 steps: straight-line
//                 code, cables unrolled, no flags;
//   general step  anything else (a Pid wakes up or is being primed, a window still spans a gap, Force mode, the last step
//                 of a launch, which also publishes telemetry): one out-of-line function, a rolled loop over the cables.
//
// D-term windows and gaps.  Pid::derive keeps the last 11 (time, error) pairs it was GIVEN, so after a sleep the window
// of the Pid that wakes up spans a gap and the least-squares fit runs on non-uniform time stamps (Pid.cpp:193-247) until
// 11 new samples have pushed the stale ones out.  `fresh` counts the consecutive samples since the Pid went live:
//   fresh >= 11   the window is the last 11 steps: fixed FIR over the shared-memory ring (as in the fast kernel);
//   fresh <  11   the window still holds stale samples: the HBM ring of that Pid (time stamps + errors, head in the control
//                 word) is kept current -- one push per step -- and the fit is the general one-pass window-relative
//                 least squares of step_general.cuh over it.  At most 11 steps per wake-up.
// A flush with fresh >= 11 rewrites the whole HBM ring from the shared-memory ring (stamps recomputed from the step
// index); with fresh < 11 the HBM ring is already current.  Load and flush are exact inverses.
//
// Preconditions (cdpr_create picks this variant when they hold, else step_general.cuh): NC in {4, 8}; both windows 11
// samples long and fitted with the same degree (one FIR); cmdLimit != 0 for both Pids; the cascades fit in shared memory.
#pragma once
#include "common.cuh"
#include "physics.cuh"
#include "step_general.cuh"
#include "legs.cuh"

namespace cdpr {

constexpr int kFlexLen = 11;
// UNR = unroll factor of the hot body's cable loop (template parameter of the kernel).  The flex kernel is bound by latency
// and instruction fetch, not by FP64 issue (1-1.5 warps per scheduler; fully unrolled, 21 % of the stall samples were
// instruction-cache misses), so a SMALLER body is faster: measured at NC=8 (2^20 x 1000): steady 7.4e9 / 9.5e9 / 8.6e9 / 7.2e9
// instance-steps/s for UNR = 8 / 4 / 2 / 1, with hold transitions 2.7e9 / 3.4e9 / 4.1e9 / 4.1e9.  The host picks 4 when hold is
// impossible (velocityEpsilon < 0) and 2 otherwise.  The integrals live in shared memory (rolled loops need runtime indices).

// control word, per (instance, cable): the general variant's layout (step_general.cuh) plus
//   bits 24-27  fresh: consecutive samples since the live Pid woke up, saturating at 11
//   bits 28-29  live: 0 none, 1 velocity Pid, 2 position Pid
__device__ __forceinline__ unsigned fctl_fresh(unsigned ctl) { return (ctl >> 24) & 0xfu; }
__device__ __forceinline__ unsigned fctl_live(unsigned ctl) { return (ctl >> 28) & 0x3u; }
__device__ __forceinline__ unsigned fctl_set_fresh(unsigned ctl, unsigned f) { return (ctl & ~(0xfu << 24)) | (f << 24); }
__device__ __forceinline__ unsigned fctl_set_live(unsigned ctl, unsigned l) { return (ctl & ~(0x3u << 28)) | (l << 28); }
// "Pid k is live, primed, its window is the last 11 steps": live == k + 1, fresh == 11, wasLast(k), missing(k) == 0
__device__ __forceinline__ bool fctl_steady(unsigned ctl, int k) {
  const unsigned mask = (0x3u << 28) | (0xfu << 24) | (1u << k) | (0x3fu << (2 + 6 * k));
  const unsigned want = ((unsigned)(k + 1) << 28) | ((unsigned)kFlexLen << 24) | (1u << k);
  return (ctl & mask) == want;
}

// instance word ictl[i]: bits 0-1 UpdateMode, bit 2 velocity command pending, bit 3 position command pending
enum { ICTL_VEL_PENDING = 4u, ICTL_POS_PENDING = 8u };

// Shared memory of one block, as [field][TPB] columns of doubles (thread index fastest), then the 32-bit words.
// NF = biquad stages held per filter (P input and D output each): 0, 1 or 4.
template <int NC, int TPB, int NF>
struct FlexSmem {
  static constexpr int FS = 8 * NF;                 // per cable: NF P stages + NF D stages, x1 x2 y1 y2 each
  static constexpr int kRing = 0;                   // [11][NC]
  static constexpr int kFilt = kRing + kFlexLen * NC;  // [NC][FS]
  static constexpr int kLastp = kFilt + NC * FS;    // [NC]  JointForceCalculator::mLastPosition
  static constexpr int kTgt = kLastp + NC;          // [NC]  target of the instance's mode
  static constexpr int kLtime = kTgt + NC;          // [NC]  Pid::mLastTime of the live Pid
  static constexpr int kIerr = kLtime + NC;         // [NC]  Pid::mIerr of the live Pid
  static constexpr int kSine = kIerr + NC;          // [3]
  static constexpr int kWords = kSine + 3;          // 32-bit words from here: ctl[NC]
  static constexpr int kDoubles = kWords + (NC + 1) / 2;
  static constexpr size_t bytes = sizeof(double) * (size_t)kDoubles * TPB;
};

// gazebo::common::Time of the step `back` steps before (sec, nsec)
// (borrowing second by second: no 64-bit division; nsec may come in negative, as `nsec - dt_ns` of the previous step)
__device__ __forceinline__ double stamp_back(int sec, int nsec, int dt_ns, int back) {
  long long ns = (long long)nsec - (long long)back * dt_ns;
  while (ns < 0) { ns += 1000000000LL; --sec; }
  return time_double(sec, (int)ns);
}

// ---------------------------------------------------------------------------------------------------------------------
// arithmetic shared by the two bodies (always inlined, explicit roundings: the same bits wherever it is expanded)
// ---------------------------------------------------------------------------------------------------------------------
// the limits are symmetric by construction (Pid.cpp:70-73: max = |limit|, min = -|limit|), so only the upper ones travel
struct FlexGains { double kf, kp, ki, kd, i_max, i_max_over_ki, c_max; };
__device__ __forceinline__ FlexGains flex_gains(const StepArgs &A, bool pos) {
  FlexGains g;
  g.kf = pos ? A.pc[1].kf : A.pc[0].kf; g.kp = pos ? A.pc[1].kp : A.pc[0].kp;
  g.ki = pos ? A.pc[1].ki : A.pc[0].ki; g.kd = pos ? A.pc[1].kd : A.pc[0].kd;
  g.i_max = pos ? A.pc[1].i_max : A.pc[0].i_max;
  g.i_max_over_ki = pos ? A.pc[1].i_max_over_ki : A.pc[0].i_max_over_ki;
  g.c_max = pos ? A.pc[1].cmd_max : A.pc[0].cmd_max;
  return g;
}

// JointForceCalculator::update (.cpp:67-89) for the two Pid modes: which Pid runs (true = position Pid), its set point
// and its measurement; `lastp` is mLastPosition (updated unless the cable holds)
// HOLD = false: velocityEpsilon < 0, |target| > eps always holds, so the Pid follows the instance's mode alone
template <bool HOLD>
__device__ __forceinline__ bool flex_select(int mode, double target, double vel_eps, const CableKin &kin, double &lastp, double &desired, double &actual) {
  const bool pos_mode = (mode == MODE_POSITION);
  const bool hold = HOLD && !pos_mode && !(fabs(target) > vel_eps);
  const bool pos = pos_mode || hold;
  desired = hold ? lastp : target;
  actual = pos ? kin.qp : kin.qd;
  lastp = hold ? lastp : kin.qp;
  return pos;
}

// fixed FIR over the last 11 steps: `e` is this step's sample, ring slot (head - a) holds the one `a` steps back
template <int STRIDE>
__device__ __forceinline__ double flex_fir(const StepArgs &A, const double *ringc, int head, double e) {
  double d0 = __dmul_rn(A.fir[kFlexLen - 1], e), d1 = 0.0;
#pragma unroll
  for (int a = 1; a < kFlexLen; ++a) {
    int sl = head - a;
    sl += (sl < 0) ? kFlexLen : 0;
    const double y = ringc[sl * STRIDE];
    if (a & 1) d1 = fma(A.fir[kFlexLen - 1 - a], y, d1); else d0 = fma(A.fir[kFlexLen - 1 - a], y, d0);
  }
  return __dadd_rn(d0, d1);
}

// Pid::CascadeFilter::update over shared-memory biquad state (Pid.cpp:38-44, Filter.h:152-165); `stages` of the NF held
template <int TPB, int NF>
__device__ __forceinline__ double flex_cascade(double *st, int stages, const double *co0, const double *co1, bool second, double x) {
  // the coefficients of the Pid that runs, picked value by value (a per-thread pointer into the kernel parameters
  // would turn every use into a generic load)
  const double a0 = second ? co1[0] : co0[0], a1 = second ? co1[1] : co0[1], a2 = second ? co1[2] : co0[2];
  const double b1 = second ? co1[3] : co0[3], b2 = second ? co1[4] : co0[4];
  double out = x;
#pragma unroll
  for (int s = 0; s < NF; ++s) {
    if (s < stages) {
      double *q = st + (s * 4) * TPB;
      const double x1 = q[0], x2 = q[TPB], y1 = q[2 * TPB], y2 = q[3 * TPB];
      // a0 x + a1 x1 + a2 x2 - b1 y1 - b2 y2, left to right
      double y0 = __dmul_rn(a0, out);
      y0 = __dadd_rn(y0, __dmul_rn(a1, x1));
      y0 = __dadd_rn(y0, __dmul_rn(a2, x2));
      y0 = __dsub_rn(y0, __dmul_rn(b1, y1));
      y0 = __dsub_rn(y0, __dmul_rn(b2, y2));
      q[TPB] = x1; q[0] = out; q[3 * TPB] = y1; q[2 * TPB] = y0;
      out = y0;
    }
  }
  return out;
}

// Pid::update from the integral on (Pid.cpp:136-187), given the filtered P input `pe` and the filtered derivative `de`
struct FlexPidOut { double cmd, ierr, p_term, i_term_pre, d_term; };
__device__ __forceinline__ FlexPidOut flex_pid(const FlexGains &g, double desired, double e, double dt, double pe, double de, double prev_ierr) {
  FlexPidOut o;
  const double f_term = __dmul_rn(g.kf, desired);
  o.p_term = __dmul_rn(g.kp, pe);
  double ie = fma(dt, e, prev_ierr);
  double i_term = __dmul_rn(g.ki, ie);
  o.i_term_pre = i_term;
  if (i_term > g.i_max) { i_term = g.i_max; ie = g.i_max_over_ki; }       // iTerm / mIgain, precomputed (same division)
  else if (i_term < -g.i_max) { i_term = -g.i_max; ie = -g.i_max_over_ki; }
  o.d_term = __dmul_rn(g.kd, de);
  const double cmd_raw = __dadd_rn(__dadd_rn(__dadd_rn(f_term, o.p_term), i_term), o.d_term);
  double cmd = clampd(cmd_raw, -g.c_max, g.c_max);  // cmdMax > cmdMin in this variant
  if (cmd != cmd_raw) {  // Pid.cpp:181-184
    ie = prev_ierr;
    cmd = __dadd_rn(cmd, __dmul_rn(__dmul_rn(dt, e), g.ki));
  }
  o.cmd = cmd;
  o.ierr = ie;
  return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// rare paths: out of line, per-thread state reached through its shared-memory columns
// ---------------------------------------------------------------------------------------------------------------------
// Flush the live Pid `k` of cable `c` to HBM: integral, last update time, biquad state, and -- when the window is entirely
// fresh -- the window itself in logical order from the ring head on.  slot_now / (sec, nsec) = ring slot and time of the
// newest sample.
// (c = cable index within the lane: shared-memory columns; cg = c0 + c = cable index of the robot: HBM columns)
template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flex_flush(const StepArgs &A, double *sm, unsigned ctl, int c0, int c, int k, int slot_now, int sec, int nsec, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  const int cg = c0 + c;
  L.pid[pid_off(L, cg, k, PID_I_ERR) + i] = sm[(M::kIerr + c) * TPB];
  L.pid[pid_off(L, cg, k, PID_LAST_TIME) + i] = sm[(M::kLtime + c) * TPB];
  const double *filt = sm + (M::kFilt + c * M::FS) * TPB;
  for (int s = 0; s < A.flex_ps; ++s)
    for (int f = 0; f < 4; ++f) L.filt[filt_off(L, cg, k, 0, s, f) + i] = filt[(s * 4 + f) * TPB];
  for (int s = 0; s < A.flex_ds; ++s)
    for (int f = 0; f < 4; ++f) L.filt[filt_off(L, cg, k, 1, s, f) + i] = filt[((NF + s) * 4 + f) * TPB];
  if (fctl_fresh(ctl) >= (unsigned)kFlexLen) {
    int hd = (int)gctl_head(ctl, k);  // oldest slot of the HBM ring; unchanged by a full rewrite
    for (int j = 0; j < kFlexLen; ++j) {
      const int age = kFlexLen - 1 - j;
      int sl = slot_now - age;
      sl += (sl < 0) ? kFlexLen : 0;
      L.win_y[win_off(L, cg, k, hd) + i] = sm[(M::kRing + sl * CPL + c) * TPB];
      L.win_x[win_off(L, cg, k, hd) + i] = stamp_back(sec, nsec, A.dt_ns, age);
      hd = (hd + 1 == kFlexLen) ? 0 : hd + 1;
    }
  }
}

// Wake Pid `k` of cable `c`: biquad state, last update time and integral error into shared memory.
template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flex_wake(const StepArgs &A, double *sm, int c0, int c, int k, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  const int cg = c0 + c;
  double *filt = sm + (M::kFilt + c * M::FS) * TPB;
  for (int s = 0; s < A.flex_ps; ++s)
    for (int f = 0; f < 4; ++f) filt[(s * 4 + f) * TPB] = L.filt[filt_off(L, cg, k, 0, s, f) + i];
  for (int s = 0; s < A.flex_ds; ++s)
    for (int f = 0; f < 4; ++f) filt[((NF + s) * 4 + f) * TPB] = L.filt[filt_off(L, cg, k, 1, s, f) + i];
  sm[(M::kLtime + c) * TPB] = L.pid[pid_off(L, cg, k, PID_LAST_TIME) + i];
  sm[(M::kIerr + c) * TPB] = L.pid[pid_off(L, cg, k, PID_I_ERR) + i];
}

// The window of a live Pid, HBM ring (logical order) -> shared-memory ring (step-aligned: newest sample in slot_now).
template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flex_load_window(const StepArgs &A, double *sm, unsigned ctl, int c0, int c, int k, int slot_now, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  int hd = (int)gctl_head(ctl, k);
  for (int j = 0; j < kFlexLen; ++j) {
    const int age = kFlexLen - 1 - j;
    int sl = slot_now - age;
    sl += (sl < 0) ? kFlexLen : 0;
    sm[(M::kRing + sl * CPL + c) * TPB] = L.win_y[win_off(L, c0 + c, k, hd) + i];
    hd = (hd + 1 == kFlexLen) ? 0 : hd + 1;
  }
}

// Pid::reset (Pid.cpp:100-115) of Pid `k` on every cable of this instance (setVelocityTarget / setPositionTarget on a mode
// change, JointForceCalculator.cpp:99-119); mLastTime is kept.
template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flex_reset_pid(const StepArgs &A, double *sm, unsigned *sw, int c0, int k, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  for (int c = 0; c < CPL; ++c) {
    const int cg = c0 + c;
    unsigned ctl = sw[c * TPB];
    if (fctl_live(ctl) == (unsigned)(k + 1)) {  // the live Pid: its state is on chip
      sm[(M::kIerr + c) * TPB] = 0.0;
      for (int f = 0; f < M::FS; ++f) sm[(M::kFilt + c * M::FS + f) * TPB] = 0.0;
      ctl = fctl_set_fresh(ctl, 0u);
    }
    L.pid[pid_off(L, cg, k, PID_P_ERR) + i] = 0.0;
    L.pid[pid_off(L, cg, k, PID_I_ERR) + i] = 0.0;
    L.pid[pid_off(L, cg, k, PID_D_ERR) + i] = 0.0;
    L.pid[pid_off(L, cg, k, PID_CMD) + i] = 0.0;
    if (L.filt)
      for (int pd = 0; pd < 2; ++pd)
        for (int s = 0; s < L.casc; ++s)
          for (int f = 0; f < 4; ++f) L.filt[filt_off(L, cg, k, pd, s, f) + i] = 0.0;
    ctl &= ~((1u << k) | (1u << (30 + k)));  // wasLast; step_flexr.cuh's "slept on 11 consecutive steps" bit
    sw[c * TPB] = gctl_set(ctl, k, (unsigned)kFlexLen, 0u);  // wasLast cleared, missing = 11, ring head 0
  }
}

template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flex_load_targets(const StepArgs &A, double *sm, int c0, int mode, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const int field = (mode == MODE_FORCE) ? CAB_FORCE_CMD : (mode == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
  for (int c = 0; c < CPL; ++c) sm[(M::kTgt + c) * TPB] = A.L.cab[cab_off(A.L, c0 + c, field) + i];
}

// A velocity command reached this instance while it was not in Velocity mode: reset the velocity Pid, switch
template <int CPL, int TPB, int NF>
static __device__ __noinline__ int flex_enter_velocity(const StepArgs &A, double *sm, unsigned *sw, int c0, long long i) {
  flex_reset_pid<CPL, TPB, NF>(A, sm, sw, c0, PID_VEL, i);
  return MODE_VELOCITY;
}

// The commands latched before this launch, applied in the first step (CdprGazeboPlugin.cpp:206-219): velocity fan-out, then
// position fan-out.  `vel_event`: a velocity command of THIS step (sine publisher / command table) already wrote the targets.
template <int CPL, int TPB, int NF>
static __device__ __noinline__ int flex_apply_pending(const StepArgs &A, double *sm, unsigned *sw, int c0, int mode, bool vel_pending, bool pos_pending, bool vel_event, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  if (vel_pending && !vel_event && mode != MODE_VELOCITY) flex_load_targets<CPL, TPB, NF>(A, sm, c0, MODE_VELOCITY, i);
  if (vel_pending || vel_event) {
    if (mode != MODE_VELOCITY) flex_reset_pid<CPL, TPB, NF>(A, sm, sw, c0, PID_VEL, i);
    mode = MODE_VELOCITY;
  }
  if (pos_pending) {
    if (vel_event)  // the velocity targets just latched must survive in HBM before the position targets replace them on chip
      for (int c = 0; c < CPL; ++c) L.cab[cab_off(L, c0 + c, CAB_VEL_TARGET) + i] = sm[(M::kTgt + c) * TPB];
    if (mode != MODE_POSITION) flex_reset_pid<CPL, TPB, NF>(A, sm, sw, c0, PID_POS, i);
    mode = MODE_POSITION;
    flex_load_targets<CPL, TPB, NF>(A, sm, c0, MODE_POSITION, i);
  }
  return mode;
}

// least-squares derivative over the HBM ring of a window that spans a gap (degree from the Pid's parameters)
static __device__ __noinline__ double flex_gap_fit(const StepArgs &A, int c, int k, unsigned oldest, double now, long long i) {
  const int deg = A.pc[k].degree;
  if (deg == 1) return ls_derivative<1, kFlexLen>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 2) return ls_derivative<2, kFlexLen>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 3) return ls_derivative<3, kFlexLen>(A.L, c, k, kFlexLen, oldest, now, i);
  if (deg == 4) return ls_derivative<4, kFlexLen>(A.L, c, k, kFlexLen, oldest, now, i);
  return 0.0;
}

// Sum of one value over the LANES lanes of an instance (adjacent threads).  Pairwise, so every lane ends with the same bits.
template <int LANES>
__device__ __forceinline__ double lane_sum(double v) {
  if (LANES >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
  if (LANES >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
struct Wrench6 { double fx, fy, fz, mx, my, mz; };

// One physics step of THIS LANE's cables with every flag honoured: a rolled loop over the cables, state in the thread's
// shared-memory columns.  Returns the lane's share of the wrench (the lead lane's includes gravity).
template <int CPL, int TPB, int NF>
static __device__ __noinline__ Wrench6 flex_general_step(const StepArgs &A, FastState S, double *sm, unsigned *sw, int c0, bool lead, bool valid, int mode,
                                                         double now, int head, int sec, int nsec, bool last, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const Rot R = make_rot(S);
  Wrench6 W;
  W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
  W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
#pragma unroll 1
  for (int c = 0; c < CPL; ++c) {
    const int cg = c0 + c;
    const CableKin kin = cable_kin<0, true>(rc, S, R, cg);
    const double target = sm[(M::kTgt + c) * TPB];
    unsigned run = 0u;  // 0 none (Force mode), 1 velocity Pid, 2 position Pid
    double desired = 0.0, actual = 0.0, force = 0.0;
    if (mode == MODE_FORCE) {  // JointForceCalculator.cpp:67-70
      sm[(M::kLastp + c) * TPB] = kin.qp;
      force = target;
    } else {
      double lp = sm[(M::kLastp + c) * TPB];
      const bool pos = flex_select<true>(mode, target, rc.vel_eps, kin, lp, desired, actual);
      sm[(M::kLastp + c) * TPB] = lp;
      run = pos ? 2u : 1u;
    }
    unsigned w = sw[c * TPB];
    if (fctl_live(w) != run) {  // this cable changes Pid
      const unsigned live = fctl_live(w);
      int slot_prev = head - 1;
      slot_prev += (slot_prev < 0) ? kFlexLen : 0;
      // the ring's newest sample belongs to the PREVIOUS step (this step's has not been pushed yet)
      if (live != 0u) flex_flush<CPL, TPB, NF>(A, sm, w, c0, c, (int)live - 1, slot_prev, sec, nsec - A.dt_ns, i);
      if (run != 0u) flex_wake<CPL, TPB, NF>(A, sm, c0, c, (int)run - 1, i);
      w = fctl_set_fresh(fctl_set_live(w, run), 0u);
    }
    if (run != 0u) {
      const int k = (int)run - 1;
      const bool pos = (k == PID_POS);
      if (!((w >> k) & 1u)) {  // first update after a reset: Pid.cpp:123-126
        w |= 1u << k;
        force = 0.0;
        if (last) L.pid[pid_off(L, cg, k, PID_CMD) + i] = 0.0;
      } else {  // Pid.cpp:127-187
        const FlexGains g = flex_gains(A, pos);
        const double e = __dsub_rn(desired, actual);
        const double dt = __dsub_rn(now, sm[(M::kLtime + c) * TPB]);
        double pe = e;
        if (NF > 0) pe = flex_cascade<TPB, NF>(sm + (M::kFilt + c * M::FS) * TPB, pos ? A.pc[1].p_casc : A.pc[0].p_casc, A.pc[0].pf, A.pc[1].pf, pos, e);
        // ---- derive (Pid.cpp:193-217): dt > 0 always (sim time advances every step)
        sm[(M::kRing + head * CPL + c) * TPB] = e;
        unsigned fresh = fctl_fresh(w), missing = gctl_missing(w, k), hd = gctl_head(w, k);
        fresh += (fresh < (unsigned)kFlexLen) ? 1u : 0u;
        missing -= (missing > 0u) ? 1u : 0u;
        if (fresh < (unsigned)kFlexLen) {  // the window still holds older samples: keep the HBM ring current
          L.win_x[win_off(L, cg, k, (int)hd) + i] = now;
          L.win_y[win_off(L, cg, k, (int)hd) + i] = e;
          hd = (hd + 1u == (unsigned)kFlexLen) ? 0u : hd + 1u;
        }
        w = fctl_set_fresh(gctl_set(w, k, missing, hd), fresh);
        double derived = 0.0;
        if (missing == 0u && A.pc[0].degree >= 1) {  // both Pids fit the same degree in this variant
          if (fresh >= (unsigned)kFlexLen) derived = flex_fir<CPL * TPB>(A, sm + (M::kRing + c) * TPB, head, e);
          else derived = flex_gap_fit(A, cg, k, hd, now, i);
        }
        double de = derived;
        if (NF > 0) de = flex_cascade<TPB, NF>(sm + (M::kFilt + c * M::FS + 4 * NF) * TPB, pos ? A.pc[1].d_casc : A.pc[0].d_casc, A.pc[0].df, A.pc[1].df, pos, derived);
        const FlexPidOut o = flex_pid(g, desired, e, dt, pe, de, sm[(M::kIerr + c) * TPB]);
        sm[(M::kIerr + c) * TPB] = o.ierr;
        force = o.cmd;
        if (last) {
          L.pid[pid_off(L, cg, k, PID_P_ERR) + i] = pe;
          L.pid[pid_off(L, cg, k, PID_D_ERR) + i] = de;
          L.pid[pid_off(L, cg, k, PID_CMD) + i] = o.cmd;
          L.cab[cab_off(L, cg, CAB_TERM_P) + i] = o.p_term;
          L.cab[cab_off(L, cg, CAB_TERM_I) + i] = o.i_term_pre;
          L.cab[cab_off(L, cg, CAB_TERM_D) + i] = o.d_term;
          L.cab[cab_off(L, cg, CAB_DESIRED) + i] = desired;
        }
      }
      sm[(M::kLtime + c) * TPB] = now;
    }
    sw[c * TPB] = w;
    const double eff = (rc.effort_limit >= 0.0) ? clampd(force, -rc.effort_limit, rc.effort_limit) : force;
    if (last) {
      if (valid) publish_joint(A, L.nc, cg, kin.qp, kin.qd, eff, i);
      L.cab[cab_off(L, cg, CAB_EFFORT) + i] = eff;
      L.cab[cab_off(L, cg, CAB_PID_FORCE) + i] = force;
    }
    const double tl = __dmul_rn(fma(-rc.cdamp, kin.qd, eff), kin.il);  // tension / L
    W.fx = fma(tl, kin.dx, W.fx); W.fy = fma(tl, kin.dy, W.fy); W.fz = fma(tl, kin.dz, W.fz);
    W.mx = fma(tl, kin.cx, W.mx); W.my = fma(tl, kin.cy, W.my); W.mz = fma(tl, kin.cz, W.mz);
  }
  return W;
}

// LANES adjacent threads share one robot, CPL = NC / LANES cables each: every lane keeps a copy of the platform state, works
// the force law of its own cables out of its own shared-memory columns, the lanes' wrench shares are summed with
// __shfl_xor_sync and every lane integrates the same platform step (same bits).  More lanes = less shared memory and fewer
// registers per thread = more resident warps, which is what this latency-bound kernel needs; the price is the replicated
// platform update.  Instances beyond n (the padding of every column to a multiple of 128) run like any other and are
// only kept from writing to the caller's buffers: the shuffles need whole warps.
// Measured at NC=8, 2^20 x 1000 (instance-steps/s, LANES = 1 / 2 / 4): steady launch configuration 9.3e9 / 8.0e9 / 5.3e9; hold
// below 2 cm/s 3.8e9 / 5.2e9 / 4.0e9; hold + one P and one D biquad stage 2.2e9 / 3.8e9 / 3.3e9.  At NC=4 one lane wins
// everywhere.  The host therefore takes 2 lanes at 8 cables when hold is possible or filters are on, else 1.
template <int NC, int TPB, int NF, int UNR, int LANES>
__global__ void __launch_bounds__(TPB) k_step_flex(const __grid_constant__ StepArgs A) {
  constexpr int CPL = NC / LANES;
  constexpr bool kHold = (UNR < 4);  // the host launches UNR = 4 exactly when velocityEpsilon < 0 (api.cu): no cable can ever hold
  static_assert(CPL * LANES == NC && (LANES == 1 || LANES == 2 || LANES == 4), "lanes must divide the cables");
  using M = FlexSmem<CPL, TPB, NF>;
  extern __shared__ double smem[];
  const int tid = (int)threadIdx.x;
  const long long gt = (long long)blockIdx.x * TPB + tid;
  const long long i = gt / LANES;
  const int c0 = (int)(gt % LANES) * CPL;
  const bool lead = (c0 == 0);
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const long long np = L.np;
  const bool valid = i < L.n;  // i < np always: the grid covers the padded columns
  double *sm = smem + tid;
  unsigned *sw = reinterpret_cast<unsigned *>(smem + M::kWords * TPB) + tid;

  FastState S;
  load_plat(L, i, S);
  const unsigned ictl = L.ictl[i];
  int mode = (int)(ictl & 3u);
  const bool vel_pending0 = (ictl & ICTL_VEL_PENDING) != 0u, pos_pending0 = (ictl & ICTL_POS_PENDING) != 0u;
  const int head0 = (int)(A.n0 % kFlexLen);  // ring slot of the newest sample already in the windows
  flex_load_targets<CPL, TPB, NF>(A, sm, c0, mode, i);
#pragma unroll 1
  for (int c = 0; c < CPL; ++c) {
    const unsigned w = L.ctl[(long long)(c0 + c) * np + i];
    sw[c * TPB] = w;
    sm[(M::kLastp + c) * TPB] = L.cab[cab_off(L, c0 + c, CAB_LAST_POS) + i];
    sm[(M::kIerr + c) * TPB] = 0.0;
    sm[(M::kLtime + c) * TPB] = 0.0;
    const unsigned live = fctl_live(w);
    if (live != 0u) {
      flex_wake<CPL, TPB, NF>(A, sm, c0, c, (int)live - 1, i);
      // the HBM ring is current at a launch boundary whatever `fresh` is; its newest `fresh` samples are the consecutive
      // steps the FIR will need once 11 of them are there, the older ones land in slots that are overwritten before use
      flex_load_window<CPL, TPB, NF>(A, sm, w, c0, c, (int)live - 1, head0, i);
    }
  }
  if (A.sine_on) {
#pragma unroll
    for (int m = 0; m < 3; ++m) sm[(M::kSine + m) * TPB] = L.sine[m * np + i];
  }
  const float *cmd_row = nullptr;
  if (A.cmd_table) cmd_row = A.cmd_table + (size_t)(i % A.n_seq) * A.n_cmd * NC + c0;
  double cost = 0.0;
  int sec = A.sec0, nsec = A.nsec0, head = head0;
  double tprev = A.t0, sine_time = A.sine_time0;
  int sine_ctr = (int)(A.n0 % (A.sine_period > 0 ? A.sine_period : 1));
  int cmd_ctr = 0, cmd_idx = 0;
  long long snap_idx = A.snap_written0;
  long long snap_ctr = A.snap_every > 0 ? (A.n0 % A.snap_every) : 0;

  // every cable of this lane runs its live, primed Pid on a window of the last 11 steps?  Depends on the targets and the
  // control words only, so it is re-evaluated after a command event or a general step, not every step.
  auto steady_now = [&]() {
    if (mode == MODE_FORCE) return false;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const bool pos = (mode == MODE_POSITION) || (kHold && !(fabs(sm[(M::kTgt + c) * TPB]) > rc.vel_eps));
      ok = ok && fctl_steady(sw[c * TPB], pos ? PID_POS : PID_VEL);
    }
    return ok;
  };
  bool steady = false, recheck = true;
  bool hot = false;  // the previous step ran the hot body: mLastTime == tprev implicitly (written back when the run ends)

  for (int s = 0; s < A.k_steps; ++s) {
    const bool last = (s == A.k_steps - 1);
    // World::Step: simTime += dt, then the plugin callback
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    const double now = time_double(sec, nsec);
    head = (head + 1 == kFlexLen) ? 0 : head + 1;

    // ---- commands of this step (CdprGazeboPlugin::update, .cpp:206-219)
    bool vel_event = false;
    if (A.sine_on) {
      if (sine_ctr == 0) {  // sinevelocitytest.cpp:35-38,48: float32 axes, accumulated publisher time
        const double arg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(sine_time, sm[(M::kSine + 1) * TPB]), 2.0), 3.14159265358979323846), sm[(M::kSine + 2) * TPB]);
        const double vel = publisher_value(A.pub_shape, sm[M::kSine * TPB], sin(arg));
#pragma unroll
        for (int c = 0; c < CPL; ++c) sm[(M::kTgt + c) * TPB] = vel;
        sine_time = __dadd_rn(sine_time, A.sine_pub_dt);
        vel_event = true;
      }
      sine_ctr = (sine_ctr + 1 == A.sine_period) ? 0 : sine_ctr + 1;
    }
    if (cmd_row) {
      if (cmd_ctr == 0 && cmd_idx < A.n_cmd) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) sm[(M::kTgt + c) * TPB] = (double)cmd_row[cmd_idx * NC + c];
        ++cmd_idx;
        vel_event = true;
      }
      cmd_ctr = (cmd_ctr + 1 == A.steps_per_cmd) ? 0 : cmd_ctr + 1;
    }
    const bool pending = (s == 0) && (vel_pending0 || pos_pending0);
    if (pending || (vel_event && mode != MODE_VELOCITY)) {  // rare: a mode may change (the same way in every lane of the robot)
      if (hot) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) sm[(M::kLtime + c) * TPB] = tprev;
        hot = false;
      }
      mode = pending ? flex_apply_pending<CPL, TPB, NF>(A, sm, sw, c0, mode, vel_pending0, pos_pending0, vel_event, i)
                     : flex_enter_velocity<CPL, TPB, NF>(A, sm, sw, c0, i);
      recheck = true;
    }
    if (vel_event) recheck = true;
    if (recheck) { steady = steady_now(); recheck = false; }

    Wrench6 W;
    // (choosing the body per WARP -- all lanes general as soon as one needs it -- was measured: +3 % with hold transitions,
    // -20..-30 % in the steady configurations; the choice stays per thread)
    if (steady && !last) {
      // ================= hot body: straight-line, every cable of the lane on its live Pid =================
      hot = true;
      const double dt = __dsub_rn(now, tprev);  // == now - mLastTime of every live Pid
      const Rot R = make_rot(S);
      W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
      W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
#pragma unroll UNR
      for (int c = 0; c < CPL; ++c) {
        const CableKin kin = cable_kin<0, true>(rc, S, R, c0 + c);
        double lp = sm[(M::kLastp + c) * TPB], desired, actual;
        const bool pos = flex_select<kHold>(mode, sm[(M::kTgt + c) * TPB], rc.vel_eps, kin, lp, desired, actual);
        sm[(M::kLastp + c) * TPB] = lp;
        const FlexGains g = flex_gains(A, pos);  // HOLD = false: pos is the same for every cable, the selects leave the loop
        const double e = __dsub_rn(desired, actual);
        double pe = e;
        if (NF > 0) pe = flex_cascade<TPB, NF>(sm + (M::kFilt + c * M::FS) * TPB, pos ? A.pc[1].p_casc : A.pc[0].p_casc, A.pc[0].pf, A.pc[1].pf, pos, e);
        sm[(M::kRing + head * CPL + c) * TPB] = e;
        double derived = 0.0;
        if (A.pc[0].degree >= 1) derived = flex_fir<CPL * TPB>(A, sm + (M::kRing + c) * TPB, head, e);
        double de = derived;
        if (NF > 0) de = flex_cascade<TPB, NF>(sm + (M::kFilt + c * M::FS + 4 * NF) * TPB, pos ? A.pc[1].d_casc : A.pc[0].d_casc, A.pc[0].df, A.pc[1].df, pos, derived);
        const FlexPidOut o = flex_pid(g, desired, e, dt, pe, de, sm[(M::kIerr + c) * TPB]);
        sm[(M::kIerr + c) * TPB] = o.ierr;
        const double eff = (rc.effort_limit >= 0.0) ? clampd(o.cmd, -rc.effort_limit, rc.effort_limit) : o.cmd;
        const double tl = __dmul_rn(fma(-rc.cdamp, kin.qd, eff), kin.il);  // tension / L
        W.fx = fma(tl, kin.dx, W.fx); W.fy = fma(tl, kin.dy, W.fy); W.fz = fma(tl, kin.dz, W.fz);
        W.mx = fma(tl, kin.cx, W.mx); W.my = fma(tl, kin.cy, W.my); W.mz = fma(tl, kin.cz, W.mz);
      }
    } else {
      if (hot) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) sm[(M::kLtime + c) * TPB] = tprev;
        hot = false;
      }
      W = flex_general_step<CPL, TPB, NF>(A, S, sm, sw, c0, lead, valid, mode, now, head, sec, nsec, last, i);
      recheck = true;
    }
    // ---- the robot's wrench = sum over its lanes; then every lane integrates the same platform step
    W.fx = lane_sum<LANES>(W.fx); W.fy = lane_sum<LANES>(W.fy); W.fz = lane_sum<LANES>(W.fz);
    W.mx = lane_sum<LANES>(W.mx); W.my = lane_sum<LANES>(W.my); W.mz = lane_sum<LANES>(W.mz);
    if (last && lead && valid) publish_platform(A, S, i);
    if (rc.leg_model) {
      S = legs_step(A, S, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    } else {
      const Rot R = make_rot(S);
      if (rc.diag_inertia) rigid_body_step<SPEC_DIAG>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
      else rigid_body_step<0>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    }
    tprev = now;
    if (A.cost) {
      const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
      cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
    }
    if (A.snap_every > 0) {
      if (++snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (lead && valid && snap_idx < A.snap_capacity) write_snapshot(A, S, snap_idx * 13 * A.snap_stride + A.snap_offset + i);
        ++snap_idx;
      }
    }
  }
  // the last step always runs the general body, so the last update times are back in shared memory here

  // ---- back to HBM
  if (lead) {
    store_plat(L.plat + i, np, S);
    if (A.cost) A.cost[i] = cost;
    L.ictl[i] = (A.k_steps > 0) ? (unsigned)mode : ictl;  // pending commands are consumed by the first step
  }
  const int tgt_field = (mode == MODE_FORCE) ? CAB_FORCE_CMD : (mode == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
#pragma unroll 1
  for (int c = 0; c < CPL; ++c) {
    const unsigned w = sw[c * TPB];
    const unsigned live = fctl_live(w);
    if (live != 0u) flex_flush<CPL, TPB, NF>(A, sm, w, c0, c, (int)live - 1, head, sec, nsec, i);
    L.ctl[(long long)(c0 + c) * np + i] = w;
    L.cab[cab_off(L, c0 + c, CAB_LAST_POS) + i] = sm[(M::kLastp + c) * TPB];
    if (A.k_steps > 0) L.cab[cab_off(L, c0 + c, tgt_field) + i] = sm[(M::kTgt + c) * TPB];
  }
}

}  // namespace cdpr
