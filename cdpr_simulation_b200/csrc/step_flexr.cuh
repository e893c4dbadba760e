// step_flexr.cuh -- K2''r: the full-semantics kernel of step_flex.cuh rebuilt around what ncu showed about it.
//
// Same semantics, same HBM state and the same rare paths (flush / wake / reset / pending commands: the helpers of
// step_flex.cuh) as k_step_flex -- hold through the position Pid (JointForceCalculator.cpp:72-82), biquad cascades
// (Pid.cpp:27-44, Filter.h:152-165), the exact clamp chain of Pid::update (Pid.cpp:136-187), per-instance modes and command
// latches (CdprGazeboPlugin.cpp:67-83,206-219).  What changes is where the time went (ncu + SASS of k_step_flex, round 2):
//
//   * the hot body was 291 instructions per cable, 87 of them FP64: a rolled cable loop turns every robot constant, gain and
//     filter coefficient into an indexed constant load (48 LDC per cable) and every per-Pid value into a pair of selects
//     (30 FSEL per cable).  Here the per-Pid gains and the per-lane robot constants come from a block-shared table in shared
//     memory (broadcast LDS.128), the filter coefficients are constant-bank operands (one set per filter), which Pid a cable
//     runs (hold) and its set point are worked out when a command arrives, not every step, the FIR runs in ring-slot order
//     with rotated weights (no index arithmetic), and the clamp chain is branch-free.
//   * the steps run event-driven as in step_fast.cuh: a warp whose robots are all steady runs up to its next event in an
//     inner loop that holds the hot body and nothing else.
//   * the kernel turned out to be bound by INSTRUCTION FETCH whenever warps alternate between the hot loop and the full path
//     (both together exceed the instruction cache): the whole warp takes the full path in a step in which one robot needs
//     it, the cable loop is unrolled by 2 when hold is possible, the general step is inlined at its one call site.
//   * a Pid that woke up fitted its gap-spanning window out of HBM for 11 steps: 22 dependent-latency loads, normal equations
//     and 4 divisions per cable and step, with 1-2 threads of the warp active.  Here the shared-memory ring always holds the
//     live Pid's whole window (the stale samples sit in the slots the next pushes overwrite); when the stale part is a run of
//     consecutive steps (ctl bits 30/31, set when a Pid goes to sleep on a full window of fresh samples) its time stamps
//     follow from one value, so the fit needs no load at all, and its weights are shared by the cables that woke together.
//     Windows that are stale twice over keep the HBM fit of step_flex.cuh.  Wake-ups issue all their loads at once.
//   (Tried and measured slower: integrals and biquad state in registers -- the unrolled loop passes 255 registers and a spill
//   is an L2 round trip with the shared-memory carve-out at its maximum; optimistic clamps; 1 and 4 lanes per 8-cable robot.)
//
// Both bodies run the same inlined arithmetic helpers with explicit roundings, so
// which body ran never shows in the bits (GPU tests: bitwise launch-split and checkpoint invariance through hold transitions).
//
// Biquad slots: NF = 0, 1 or 2 stages per filter, ONE coefficient set per filter (constant-bank operands): every slot up to the
// larger stage count of the two Pids runs the arithmetic, and a select takes the output of the last stage the RUNNING Pid has
// (pe = e with an empty cascade, Pid.cpp:133) -- a slot beyond that is a don't-care that both bodies advance the same way.  Two
// Pids with different coefficients, more stages, or the leg model: k_step_flex.
#pragma once
#include "step_flex.cuh"

namespace cdpr {

// ctl bit 30 + k: the window Pid k took to sleep is 11 CONSECUTIVE steps ending at its last update time
constexpr unsigned kRunBit0 = 30;

template <int CPL, int TPB, int NF, int LANES>
struct FlexRSmem : FlexSmem<CPL, TPB, NF> {
  using B = FlexSmem<CPL, TPB, NF>;
  static constexpr int kDes = B::kDoubles;    // [CPL] set point of the Pid that runs: the target, or the latched hold position
  static constexpr int kStale = kDes + CPL;   // [CPL] time stamp of the newest STALE sample of a window that spans a gap
  static constexpr int kPerThread = kStale + CPL;
  // block-shared table behind the per-thread columns
  static constexpr int kRow = 8;              // kf kp ki kd i_max i_max/ki c_max min(c_max, effort limit): the gains of one Pid
  static constexpr int kTabCab = 2 * kRow;    // [LANES][CPL][7]: b xyz, a xyz, home length
  static constexpr int kTabDoubles = kTabCab + LANES * CPL * 7;
  static constexpr size_t bytes = sizeof(double) * ((size_t)kPerThread * TPB + kTabDoubles);
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// cable_kin (physics.cuh) on explicit constants: same expressions, same bits
__device__ __forceinline__ CableKin cable_kin_v(double bx, double by, double bz, double ax, double ay, double az, double home, const FastState &S, const Rot &R) {
  CableKin k;
  const double gx = ax - S.px, gy = ay - S.py, gz = az - S.pz;
  k.dx = fma(-R.r00, bx, fma(-R.r01, by, fma(-R.r02, bz, gx)));
  k.dy = fma(-R.r10, bx, fma(-R.r11, by, fma(-R.r12, bz, gy)));
  k.dz = fma(-R.r20, bx, fma(-R.r21, by, fma(-R.r22, bz, gz)));
  const double l2 = fma(k.dx, k.dx, fma(k.dy, k.dy, k.dz * k.dz));
  k.il = rsqrt_nr(l2);
  k.cx = fma(gy, k.dz, -(gz * k.dy)); k.cy = fma(gz, k.dx, -(gx * k.dz)); k.cz = fma(gx, k.dy, -(gy * k.dx));
  k.qd = (fma(k.dx, S.vx, fma(k.dy, S.vy, k.dz * S.vz)) + fma(k.cx, S.wx, fma(k.cy, S.wy, k.cz * S.wz))) * k.il;
  k.qp = fma(-l2, k.il, home);
  return k;
}

// BiQuad::process (Filter.h:152-165) on four values, y0 = a0 x + a1 x1 + a2 x2 - b1 y1 - b2 y2, then the shift.  Summed as
// (a0 x + a1 x1) + (a2 x2 - b1 y1 - b2 y2) with fused multiply-adds: the second group does not depend on the new input, so only
// three operations sit between x and y0 (the reference's left-to-right chain of five products and four sums is nine operations
// deep, and the two filters of a cable are in series with the D-term between them) -- 6 FP64 instructions instead of 9.  Same
// value up to the last rounding (the CUDA path is held to 1e-9 of the CPU checker per step, not to its bits); both bodies use it.
__device__ __forceinline__ double biquad_step(const double *co, double &x1, double &x2, double &y1, double &y2, double x) {
  const double hist = fma(-co[4], y2, fma(-co[3], y1, __dmul_rn(co[2], x2)));
  const double y0 = __dadd_rn(fma(co[1], x1, __dmul_rn(co[0], x)), hist);
  x2 = x1; x1 = x; y2 = y1; y1 = y0;
  return y0;
}

// The fixed FIR over the last 11 steps, summed in the order of the ring SLOTS (slot = step index mod 11) with the weights
// rotated to match: every tap is a load at an immediate offset and a constant-bank weight, no index arithmetic.  Split so
// that nothing waits for this step's sample e: the ten OLDER samples first (gw0 = firx + 10 - head is the weight of the
// sample in slot s, and 0 for slot `head` -- whether that slot still holds the sample that leaves the window or already e,
// it adds +0), then D = fir[10] e + that.  The loads precede the store of e, so they and their sums run under the
// kinematics; one operation sits between e and D (the kernel is bound by dependent FP64 chains, not by their count).
template <int STRIDE>
__device__ __forceinline__ double flexr_fir_older(const double *gw0, const double *ringc) {
  double d0 = __dmul_rn(gw0[0], ringc[0]), d1 = __dmul_rn(gw0[1], ringc[STRIDE]);
#pragma unroll
  for (int s = 2; s < kFlexLen; ++s) {
    const double y = ringc[s * STRIDE];
    if (s & 1) d1 = fma(gw0[s], y, d1); else d0 = fma(gw0[s], y, d0);
  }
  return __dadd_rn(d0, d1);
}

// Pid::CascadeFilter::update (Pid.cpp:38-44) over the NF stage slots of one filter in shared memory (x1 x2 y1 y2 per stage,
// one coefficient set): the slots up to `slots` (the larger stage count of the two Pids) all run, the output is the one of the
// last stage the RUNNING Pid has (`count`), the input itself if it has none.  RING_X: the first stage's x1, x2 are the ring's last
// two samples (the hot body; o1 / o2 their offsets) instead of its own columns.
template <int TPB, int NF, bool RING_X>
__device__ __forceinline__ double flexr_cascade(const double *co, double *q, int slots, int count, double x, const double *ringc = nullptr, int o1 = 0, int o2 = 0) {
  double out = x;
#pragma unroll
  for (int st = 0; st < NF; ++st) {
    if (st == 0 || st < slots) {  // (the first slot also runs for a filter without stages: count = 0 discards its output)
      double *p = q + (st * 4) * TPB;
      double x1, x2, y1 = p[2 * TPB], y2 = p[3 * TPB];
      if (RING_X && st == 0) { x1 = ringc[o1]; x2 = ringc[o2]; } else { x1 = p[0]; x2 = p[TPB]; }
      const double y0 = biquad_step(co, x1, x2, y1, y2, x);
      if (!(RING_X && st == 0)) { p[0] = x1; p[TPB] = x2; }
      p[2 * TPB] = y1; p[3 * TPB] = y2;
      x = y0;
      out = (count > st) ? y0 : out;
    }
  }
  return out;
}

// the same cascade on state already in registers (st = x1 x2 y1 y2 per stage): the hot body loads every piece of state at the
// top of a cable and stores at the bottom, so no load waits behind a store it cannot be proven independent of
template <int NF, bool RING_X>
__device__ __forceinline__ double flexr_cascade_regs(const double *co, double (&st)[NF > 0 ? 4 * NF : 1], int slots, int count, double x, double rx1 = 0.0, double rx2 = 0.0) {
  double out = x;
#pragma unroll
  for (int s = 0; s < NF; ++s) {
    if (s == 0 || s < slots) {
      double x1 = (RING_X && s == 0) ? rx1 : st[4 * s], x2 = (RING_X && s == 0) ? rx2 : st[4 * s + 1];
      const double y0 = biquad_step(co, x1, x2, st[4 * s + 2], st[4 * s + 3], x);
      if (!(RING_X && s == 0)) { st[4 * s] = x1; st[4 * s + 1] = x2; }
      x = y0;
      out = (count > s) ? y0 : out;
    }
  }
  return out;
}

__device__ __forceinline__ FlexGains flexr_gains(const double *row) {
  FlexGains g;
  g.kf = row[0]; g.kp = row[1]; g.ki = row[2]; g.kd = row[3]; g.i_max = row[4]; g.i_max_over_ki = row[5]; g.c_max = row[6];
  return g;
}

// flex_pid (step_flex.cuh) without a branch: the same operations and roundings with the clamp decisions as selects (the
// anti-windup correction is computed whether or not it is used).  A taken branch in the unrolled cable loop is an
// instruction-fetch bubble that two warps per scheduler cannot hide (ncu: 30-60 % of the hot loop's stall samples were
// `no_instruction`).
// v clamped to [-lim, lim], lim >= 0: one compare on |v| and the limit with v's sign (fmin / fmax cost five instructions
// each for their NaN rules; a NaN passes through here)
__device__ __forceinline__ double clamp_sym(double v, double lim) { return (fabs(v) > lim) ? copysign(lim, v) : v; }
// x with its sign flipped when s is negative (x may have either sign)
__device__ __forceinline__ double flip_sign_by(double x, double s) {
  return __hiloint2double(__double2hiint(x) ^ (__double2hiint(s) & (int)0x80000000), __double2loint(x));
}
// (... and arranged so that little is left to do once the last input, the filtered derivative, arrives: the anti-windup
// correction and the sum of the other three terms are ready by then, "the clamp changed the command" is the clamp's own
// compare, and Joint::SetForce's truncation is worked out for the clamped and the unclamped command side by side.)
struct FlexrPidOut { double cmd, eff, ierr, p_term, i_term_pre, d_term; };
__device__ __forceinline__ FlexrPidOut flexr_pid(const FlexGains &g, double eff_lim, double desired, double e, double dt, double pe, double de, double prev_ierr) {
  FlexrPidOut o;
  const double f_term = __dmul_rn(g.kf, desired);
  o.p_term = __dmul_rn(g.kp, pe);
  const double ie1 = fma(dt, e, prev_ierr);
  const double i_raw = __dmul_rn(g.ki, ie1);
  o.i_term_pre = i_raw;
  // Pid.cpp:143-150: iTerm > iMax -> iMax, mIerr = iMax / Ki; iTerm < iMin = -iMax -> -iMax, mIerr = -(iMax / Ki)
  const bool isat = fabs(i_raw) > g.i_max;
  const double i_term = isat ? copysign(g.i_max, i_raw) : i_raw;
  const double ie2 = isat ? flip_sign_by(g.i_max_over_ki, i_raw) : ie1;
  const double aw_corr = __dmul_rn(__dmul_rn(dt, e), g.ki);
  const double fpi = __dadd_rn(__dadd_rn(f_term, o.p_term), i_term);
  o.d_term = __dmul_rn(g.kd, de);
  const double cmd_raw = __dadd_rn(fpi, o.d_term);
  const bool aw = fabs(cmd_raw) > g.c_max;  // the clamp changes the command (Pid.cpp:175-184): frozen integral, corrected command
  const double cmd_sat = __dadd_rn(copysign(g.c_max, cmd_raw), aw_corr);
  const double eff_sat = clamp_sym(cmd_sat, eff_lim), eff_raw = clamp_sym(cmd_raw, eff_lim);
  o.cmd = aw ? cmd_sat : cmd_raw;
  o.eff = aw ? eff_sat : eff_raw;  // Joint::SetForce's truncation; eff_lim = +inf when it is off
  o.ierr = aw ? prev_ierr : ie2;
  return o;
}

// (sec, nsec) - back * dt_ns as a gazebo time stamp, without 64-bit divisions
__device__ __forceinline__ void stamp_dec(int &sec, int &nsec, int dt_ns) {
  nsec -= dt_ns;
  while (nsec < 0) { nsec += 1000000000; --sec; }
}

// Weights of the least-squares derivative: D = sum_a w[a] y[a] for samples with time stamps xs[a] (xs[10] the oldest), the
// derivative at `now` of the degree-DEG polynomial fitted through them (Pid.cpp:203-212 + 219-247).  The normal equations
// of ls_derivative (step_general.cuh: window-relative, span-scaled time; elimination with partial pivoting by
// compare-and-swap) solved for the row of the pseudo-inverse that gives the linear coefficient -- it depends on the STAMPS
// only, so the cables of a robot that woke up in the same step share one solve.
template <int DEG>
__device__ __forceinline__ void ls_weights11(const double (&xs)[kFlexLen], double now, double (&w)[kFlexLen]) {
  constexpr int M = DEG + 1;
  const double span = now - xs[kFlexLen - 1], inv_span = 1.0 / span;
  double x[kFlexLen], sx[2 * DEG + 1];
#pragma unroll
  for (int p = 0; p <= 2 * DEG; ++p) sx[p] = 0.0;
#pragma unroll
  for (int j = 0; j < kFlexLen; ++j) {
    x[j] = (xs[j] - now) * inv_span;
    double pw = 1.0;
#pragma unroll
    for (int p = 0; p <= 2 * DEG; ++p) { sx[p] += pw; pw *= x[j]; }
  }
  double G[M][M + 1];  // [X^T X | e_1]
#pragma unroll
  for (int r = 0; r < M; ++r) {
#pragma unroll
    for (int q = 0; q < M; ++q) G[r][q] = sx[r + q];
    G[r][M] = (r == 1) ? 1.0 : 0.0;
  }
  bool singular = false;
#pragma unroll
  for (int col = 0; col < M; ++col) {
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      const bool swp = fabs(G[r][col]) > fabs(G[col][col]);
#pragma unroll
      for (int q = col; q <= M; ++q) {
        const double a = G[col][q], b = G[r][q];
        G[col][q] = swp ? b : a;
        G[r][q] = swp ? a : b;
      }
    }
    singular = singular || (G[col][col] == 0.0);
    const double inv = 1.0 / G[col][col];
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      const double f = G[r][col] * inv;
#pragma unroll
      for (int q = col + 1; q <= M; ++q) G[r][q] = fma(-f, G[col][q], G[r][q]);
    }
  }
  double z[M];
#pragma unroll
  for (int r = M - 1; r >= 0; --r) {
    double t = G[r][M];
#pragma unroll
    for (int q = r + 1; q < M; ++q) t = fma(-G[r][q], z[q], t);
    z[r] = t / G[r][r];
  }
#pragma unroll
  for (int j = 0; j < kFlexLen; ++j) {
    double h = z[M - 1];
#pragma unroll
    for (int p = M - 2; p >= 0; --p) h = fma(h, x[j], z[p]);
    w[j] = singular ? 0.0 : h * inv_span;
  }
}

// Gap window on chip: ring position (head - a) holds the sample of age a; the newest `fresh` are the last steps, the others
// the run of consecutive steps that ended at `stale` (a gazebo time stamp: sec + nsec 1e-9 with nsec < 1e9, sec < 2^16).
// The weights of its least-squares derivative by sample age.
static __device__ __noinline__ void flexr_gap_weights(int degree, unsigned fresh, double stale, int sec, int nsec, int dt_ns, double now, double (&w)[kFlexLen]) {
  double xs[kFlexLen];
  int s = sec, ns = nsec;
#pragma unroll
  for (int a = 0; a < kFlexLen; ++a) {
    if ((unsigned)a == fresh) {  // from here on the stale run: its newest stamp back to integers (exact below 2^16 s)
      const double fl = floor(stale);
      s = (int)fl;
      ns = (int)__double2ll_rn(__dmul_rn(__dsub_rn(stale, fl), 1e9));
    }
    xs[a] = time_double(s, ns);
    stamp_dec(s, ns, dt_ns);
  }
  if (degree == 1) ls_weights11<1>(xs, now, w);
  else if (degree == 2) ls_weights11<2>(xs, now, w);
  else if (degree == 3) ls_weights11<3>(xs, now, w);
  else if (degree == 4) ls_weights11<4>(xs, now, w);
  else {
#pragma unroll
    for (int a = 0; a < kFlexLen; ++a) w[a] = 0.0;
  }
}

// flex_wake + flex_load_window of step_flex.cuh in one go, with every load issued before the first store: the state of a
// Pid that slept is cold in HBM, and loads that wait for each other cost a microsecond apiece with one or two threads of
// the warp active (this was most of the cost of a hold transition).  slot_now = ring slot of the window's newest sample.
template <int CPL, int TPB, int NF>
static __device__ __noinline__ void flexr_wake(const StepArgs &A, double *sm, unsigned ctl, int c0, int c, int k, int slot_now, long long i) {
  using M = FlexSmem<CPL, TPB, NF>;
  const DevLayout &L = A.L;
  const int cg = c0 + c;
  double wy[kFlexLen], fs[NF > 0 ? 8 * NF : 1];
  const double lt = L.pid[pid_off(L, cg, k, PID_LAST_TIME) + i], ie = L.pid[pid_off(L, cg, k, PID_I_ERR) + i];
#pragma unroll
  for (int j = 0; j < kFlexLen; ++j) wy[j] = L.win_y[win_off(L, cg, k, j) + i];
#pragma unroll
  for (int s = 0; s < NF; ++s) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      fs[s * 4 + f] = (s < A.flex_ps) ? L.filt[filt_off(L, cg, k, 0, s, f) + i] : 0.0;
      fs[(NF + s) * 4 + f] = (s < A.flex_ds) ? L.filt[filt_off(L, cg, k, 1, s, f) + i] : 0.0;
    }
  }
  sm[(M::kLtime + c) * TPB] = lt;
  sm[(M::kIerr + c) * TPB] = ie;
#pragma unroll
  for (int s = 0; s < NF; ++s) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (s < A.flex_ps) sm[(M::kFilt + c * M::FS + s * 4 + f) * TPB] = fs[s * 4 + f];
      if (s < A.flex_ds) sm[(M::kFilt + c * M::FS + (NF + s) * 4 + f) * TPB] = fs[(NF + s) * 4 + f];
    }
  }
  const int hd = (int)gctl_head(ctl, k);  // physical slot of the OLDEST sample
#pragma unroll
  for (int j = 0; j < kFlexLen; ++j) {
    int logical = j - hd;  // 0 = oldest
    logical += (logical < 0) ? kFlexLen : 0;
    int sl = slot_now - (kFlexLen - 1 - logical);
    sl += (sl < 0) ? kFlexLen : 0;
    sm[(M::kRing + sl * CPL + c) * TPB] = wy[j];
  }
}

#ifndef CDPR_FLEXR_GENERAL_INLINE
#define CDPR_FLEXR_GENERAL_INLINE __forceinline__  // inlined at its one call site: no by-value state through the stack (-6..-9 % with transitions)
#endif
#ifndef CDPR_FLEXR_UNR
#define CDPR_FLEXR_UNR 0  // tuning override of the unroll factor of the hot body's cable loop (0 = the measured choice below)
#endif
#ifndef CDPR_FLEXR_MINB0
#define CDPR_FLEXR_MINB0 1  // resident blocks asked for at two lanes without filter slots (caps the registers per thread); 10: +14 %
#endif
// One step of THIS LANE's cables with every flag honoured (the out-of-line body): flex_general_step of step_flex.cuh with
// the table-driven gains, the always-present biquad slots and the on-chip window of a Pid that woke up.
template <int CPL, int TPB, int NF, int LANES>
static __device__ CDPR_FLEXR_GENERAL_INLINE Wrench6 flexr_general_step(const StepArgs &A, const FastState &S, double *sm, unsigned *sw, const double *tab, int c0, bool lead, bool valid,
                                                          int mode, double now, int head, int sec, int nsec, bool last, long long i) {
  using M = FlexRSmem<CPL, TPB, NF, LANES>;
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const Rot R = make_rot(S);
  Wrench6 W;
  W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
  W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
  // A Pid that wakes up in this step: its state is cold in HBM -- ask for all of it now, for every cable of the lane, so
  // that the loads of flexr_wake find it on its way (they wait for each other cable by cable otherwise)
#pragma unroll 1
  for (int c = 0; c < CPL; ++c) {
    if (mode == MODE_FORCE) break;
    const bool pos = (mode == MODE_POSITION) || !(fabs(sm[(M::kTgt + c) * TPB]) > rc.vel_eps);
    const int k = pos ? PID_POS : PID_VEL;
    if (fctl_live(sw[c * TPB]) == (unsigned)(k + 1)) continue;
    const int cg = c0 + c;
    prefetch_l2(L.pid + pid_off(L, cg, k, PID_LAST_TIME) + i);
    prefetch_l2(L.pid + pid_off(L, cg, k, PID_I_ERR) + i);
    for (int j = 0; j < kFlexLen; ++j) prefetch_l2(L.win_y + win_off(L, cg, k, j) + i);
    if (NF > 0) {
      for (int f = 0; f < 4; ++f) {
        for (int st = 0; st < A.flex_ps; ++st) prefetch_l2(L.filt + filt_off(L, cg, k, 0, st, f) + i);
        for (int st = 0; st < A.flex_ds; ++st) prefetch_l2(L.filt + filt_off(L, cg, k, 1, st, f) + i);
      }
    }
  }
  // weights of the last gap fit of this step: cables whose windows have the same time stamps (they woke up together) share them
  double gw[kFlexLen];
  unsigned gw_fresh = 0xffffffffu;
  double gw_stale = 0.0;
#pragma unroll
  for (int a = 0; a < kFlexLen; ++a) gw[a] = 0.0;
#pragma unroll 1
  for (int c = 0; c < CPL; ++c) {
    const int cg = c0 + c;
    const CableKin kin = cable_kin_v(rc.b[cg][0], rc.b[cg][1], rc.b[cg][2], rc.a[cg][0], rc.a[cg][1], rc.a[cg][2], rc.home_len[cg], S, R);
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
      if (live != 0u) {
        // the ring's newest sample belongs to the PREVIOUS step (this step's has not been pushed yet)
        flex_flush<CPL, TPB, NF>(A, sm, w, c0, c, (int)live - 1, slot_prev, sec, nsec - A.dt_ns, i);
        // what it takes to sleep: 11 consecutive steps, or something older
        const unsigned bit = 1u << (kRunBit0 + live - 1u);
        w = (fctl_fresh(w) >= (unsigned)kFlexLen) ? (w | bit) : (w & ~bit);
      }
      if (run != 0u) {
        flexr_wake<CPL, TPB, NF>(A, sm, w, c0, c, (int)run - 1, slot_prev, i);
        sm[(M::kStale + c) * TPB] = sm[(M::kLtime + c) * TPB];  // its newest sample was pushed at its last update
      }
      w = fctl_set_fresh(fctl_set_live(w, run), 0u);
    }
    if (run != 0u) {
      const int k = (int)run - 1;
      const double *row = tab + k * M::kRow;
      if (!((w >> k) & 1u)) {  // first update after a reset: Pid.cpp:123-126
        w |= 1u << k;
        force = 0.0;
        if (last) L.pid[pid_off(L, cg, k, PID_CMD) + i] = 0.0;
      } else {  // Pid.cpp:127-187
        const FlexGains g = flexr_gains(row);
        const double e = __dsub_rn(desired, actual);
        const double dt = __dsub_rn(now, sm[(M::kLtime + c) * TPB]);
        double pe = e;
        if (NF > 0 && A.flex_ps > 0) pe = flexr_cascade<TPB, NF, false>(A.flex_pf, sm + (M::kFilt + c * M::FS) * TPB, A.flex_ps, A.pc[k].p_casc, e);
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
          if (fresh >= (unsigned)kFlexLen) {
            derived = fma(A.fir[kFlexLen - 1], e, flexr_fir_older<CPL * TPB>(A.firx0 + (kFlexLen - 1 - head), sm + (M::kRing + c) * TPB));
          } else {
            const double stale = sm[(M::kStale + c) * TPB];
            if (((w >> (kRunBit0 + k)) & 1u) && stale >= 0.0 && stale < 65536.0) {
              if (fresh != gw_fresh || stale != gw_stale) {
                flexr_gap_weights(A.pc[0].degree, fresh, stale, sec, nsec, A.dt_ns, now, gw);
                gw_fresh = fresh; gw_stale = stale;
              }
              const double *ringc = sm + (M::kRing + c) * TPB;
              double d0 = 0.0, d1 = 0.0;
#pragma unroll
              for (int a = 0; a < kFlexLen; ++a) {
                int sl = head - a;
                sl += (sl < 0) ? kFlexLen : 0;
                const double y = ringc[sl * (CPL * TPB)];
                if (a & 1) d1 = fma(gw[a], y, d1); else d0 = fma(gw[a], y, d0);
              }
              derived = __dadd_rn(d0, d1);
            } else
              derived = flex_gap_fit(A, cg, k, hd, now, i);
          }
        }
        double de = derived;
        if (NF > 0 && A.flex_ds > 0) de = flexr_cascade<TPB, NF, false>(A.flex_df, sm + (M::kFilt + c * M::FS + 4 * NF) * TPB, A.flex_ds, A.pc[k].d_casc, derived);
        const FlexrPidOut o = flexr_pid(g, rc.effort_limit_abs, desired, e, dt, pe, de, sm[(M::kIerr + c) * TPB]);
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
    const double eff = clamp_sym(force, rc.effort_limit_abs);  // +inf when Joint::SetForce does not truncate
    if (last) {
      if (valid) publish_joint(A, L.nc, cg, kin.qp, kin.qd, eff, i);
      L.cab[cab_off(L, cg, CAB_EFFORT) + i] = eff;
      L.cab[cab_off(L, cg, CAB_PID_FORCE) + i] = force;
    }
    const double tl = fma(eff, kin.il, __dmul_rn(__dmul_rn(-rc.cdamp, kin.qd), kin.il));  // tension / L, as in the hot body
    W.fx = fma(tl, kin.dx, W.fx); W.fy = fma(tl, kin.dy, W.fy); W.fz = fma(tl, kin.dz, W.fz);
    W.mx = fma(tl, kin.cx, W.mx); W.my = fma(tl, kin.cy, W.my); W.mz = fma(tl, kin.cz, W.mz);
  }
  return W;
}

// HOLD = false: velocityEpsilon < 0, no cable can ever hold, the Pid follows the instance's mode alone
template <int NC, int TPB, int NF, bool HOLD, int LANES, bool ISO>
__global__ void __launch_bounds__(TPB, (LANES == 2 && NF == 0) ? CDPR_FLEXR_MINB0 : 1) k_step_flexr(const __grid_constant__ StepArgs A) {
  constexpr int CPL = NC / LANES;
  // Unroll factor of the hot body's cable loop.  The kernel is bound by INSTRUCTION FETCH as soon as warps alternate between
  // the hot loop and the full path (ncu: 30-60 % of the stall samples `no_instruction`, clustered on 128-byte line starts):
  // the hot loop is 1237 / 759 / 487 instructions at 4 / 2 / 1 cables per iteration and the full path another ~1500.  Measured
  // at NC=8, 2^20 x 1000 steps (ms), unroll 4 / 2 / 1 (first build; the final one in brackets): steady 82 / 88 / 92 [75 / 80 / -];
  // with hold transitions 175 / 116 / 117 [129 / 99 / -]; hold + one P and one D stage 250 / 165 / 156 [195 / 134 / 140].
  // So: everything unrolled when no cable can ever hold, else 2.
  constexpr int kUnr = (CDPR_FLEXR_UNR > 0) ? CDPR_FLEXR_UNR : (HOLD ? 2 : CPL);
  static_assert(CPL * LANES == NC && (LANES == 1 || LANES == 2 || LANES == 4) && NF <= 2, "lanes must divide the cables; at most two biquad slots per filter");
  using M = FlexRSmem<CPL, TPB, NF, LANES>;
  extern __shared__ double smem[];
  const int tid = (int)threadIdx.x;
  const long long gt = (long long)blockIdx.x * TPB + tid;
  const long long i = gt / LANES;
  const int lane = (int)(gt % LANES);
  const int c0 = lane * CPL;
  const bool lead = (c0 == 0);
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const long long np = L.np;
  const bool valid = i < L.n;  // i < np always: the grid covers the padded columns
  double *sm = smem + tid;
  unsigned *sw = reinterpret_cast<unsigned *>(smem + M::kWords * TPB) + tid;
  double *tabw = smem + M::kPerThread * TPB;
  const double *tab = tabw;

  // ---- the block's table: gains and biquad coefficients of the two Pids, robot constants of each lane's cables
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const PidConsts &pc = A.pc[k];
      double *row = tabw + k * M::kRow;
      row[0] = pc.kf; row[1] = pc.kp; row[2] = pc.ki; row[3] = pc.kd; row[4] = pc.i_max; row[5] = pc.i_max_over_ki; row[6] = pc.cmd_max;
      row[7] = fmin(pc.cmd_max, rc.effort_limit_abs);  // a command within it passes every clamp unchanged
    }
    for (int c = 0; c < NC; ++c) {
      double *q = tabw + M::kTabCab + c * 7;
      q[0] = rc.b[c][0]; q[1] = rc.b[c][1]; q[2] = rc.b[c][2]; q[3] = rc.a[c][0]; q[4] = rc.a[c][1]; q[5] = rc.a[c][2]; q[6] = rc.home_len[c];
    }
  }
  __syncthreads();
  const double *cabtab = tab + M::kTabCab + c0 * 7;

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
    sm[(M::kDes + c) * TPB] = 0.0;
    sm[(M::kStale + c) * TPB] = 0.0;
#pragma unroll
    for (int f = 0; f < M::FS; ++f) sm[(M::kFilt + c * M::FS + f) * TPB] = 0.0;
    const unsigned live = fctl_live(w);
    if (live != 0u) {
      const int k = (int)live - 1;
      // the HBM ring is current at a launch boundary whatever `fresh` is: the whole window comes on chip
      flexr_wake<CPL, TPB, NF>(A, sm, w, c0, c, k, head0, i);
      const unsigned fresh = fctl_fresh(w);
      if (fresh < (unsigned)kFlexLen) {  // newest stale sample = logical position 10 - fresh from the oldest slot on
        unsigned sl = gctl_head(w, k) + (unsigned)(kFlexLen - 1) - fresh;
        sl -= (sl >= (unsigned)kFlexLen) ? (unsigned)kFlexLen : 0u;
        sm[(M::kStale + c) * TPB] = L.win_x[win_off(L, c0 + c, k, (int)sl) + i];
      }
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

  // ---- what the hot body carries in registers
  // The controller state of the live Pids (integrals, biquad state) stays in shared memory in the hot body too.  Held in
  // registers it raised the pressure of the unrolled cable loop past 255 registers, and with the shared-memory carve-out at
  // its maximum a spill is an L2 round trip -- measured 25 % slower than the explicit LDS / STS.  What the hot body skips is
  // mLastTime (the same for every live Pid: tprev) and the P filter's x1, x2 (the last two errors, which a hot thread finds
  // in the ring: its live Pids have pushed every one of the last 11 steps); `spill` puts them back for the general body.
  unsigned posmask = 0u, holdmask = 0u;  // per cable: runs the position Pid / holds (position Pid in Velocity mode)
  bool hot = false;  // the previous step ran the hot body
  const bool has_fir = A.pc[0].degree >= 1;

  auto spill = [&]() {  // runs after the clock tick of a step: `head` is the slot this step's sample WILL take
    int h1 = head - 1, h2 = head - 2;
    h1 += (h1 < 0) ? kFlexLen : 0;
    h2 += (h2 < 0) ? kFlexLen : 0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      sm[(M::kLtime + c) * TPB] = tprev;
      if (NF > 0) {
        double *q = sm + (M::kFilt + c * M::FS) * TPB;
        q[0] = sm[(M::kRing + h1 * CPL + c) * TPB]; q[TPB] = sm[(M::kRing + h2 * CPL + c) * TPB];
      }
    }
    hot = false;
  };
  // Which Pid every cable of this lane runs, on which set point, and whether all of them are live, primed and on a window
  // of the last 11 steps.  Depends on the targets, the latched positions and the control words only, so it is re-evaluated
  // after a command or a general step, not every step.
  auto evaluate = [&]() {
    posmask = 0u; holdmask = 0u;
    bool ok = (mode != MODE_FORCE);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const double tg = sm[(M::kTgt + c) * TPB];
      const bool hold = HOLD && (mode == MODE_VELOCITY) && !(fabs(tg) > rc.vel_eps);
      const bool pos = (mode == MODE_POSITION) || hold;
      sm[(M::kDes + c) * TPB] = hold ? sm[(M::kLastp + c) * TPB] : tg;
      posmask |= (pos ? 1u : 0u) << c;
      holdmask |= (hold ? 1u : 0u) << c;
      ok = ok && fctl_steady(sw[c * TPB], pos ? PID_POS : PID_VEL);
    }
    return ok;
  };
  bool steady = false, recheck = true;
  // the lanes of a robot are adjacent threads; the whole warp is at the same place of the loop below (see `full`), so the
  // shuffles name every lane (a computed mask costs a MATCH + WARPSYNC per shuffle)
  auto robot_sum = [&](double v) { return lane_sum<LANES>(v); };
  auto clock_tick = [&](double &now) {
    // World::Step: simTime += dt, then the plugin callback
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    now = time_double(sec, nsec);
    head = (head + 1 == kFlexLen) ? 0 : head + 1;
  };
  auto platform_step = [&](const Rot &R, Wrench6 &W) {
    // the robot's wrench = sum over its lanes; then every lane integrates the same platform step
    W.fx = robot_sum(W.fx); W.fy = robot_sum(W.fy); W.fz = robot_sum(W.fz);
    W.mx = robot_sum(W.mx); W.my = robot_sum(W.my); W.mz = robot_sum(W.mz);
    // ISO: isotropic body inertia (detected at cdpr_create), compiled in so the hot loop holds one rigid-body update
    if (ISO) rigid_body_step<SPEC_DIAG | SPEC_ISO>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    else if (rc.diag_inertia) rigid_body_step<SPEC_DIAG>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    else rigid_body_step<0>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    if (A.cost) {
      const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
      cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
    }
  };

  // The K steps as in step_fast.cuh: every step whose commands, flags or bookkeeping need attention goes through the full
  // path at the top of the loop; a robot that comes out of it steady runs the following steps up to its next event (a
  // command of the sine publisher or the command table, a snapshot, the last step of the launch) in an inner loop that
  // holds the hot body and nothing else -- no call, no flag -- so its registers are not shared with the rare paths.
  int s = 0;
  while (s < A.k_steps) {
    const bool last = (s == A.k_steps - 1);
    double now;
    clock_tick(now);

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
      if (hot) spill();
      mode = pending ? flex_apply_pending<CPL, TPB, NF>(A, sm, sw, c0, mode, vel_pending0, pos_pending0, vel_event, i)
                     : flex_enter_velocity<CPL, TPB, NF>(A, sm, sw, c0, i);
      recheck = true;
    }
    if (vel_event) recheck = true;
    if (recheck) {
      steady = evaluate();
      recheck = false;
    }

    // The threads of a warp stay in step, and they take the same body: as soon as ONE of them needs the full path in this
    // step, all of them run it (it is the general body -- a steady robot gets the same bits out of it).  Alternating
    // between the two bodies step by step, which is what a per-thread choice amounts to while a neighbour is in transition,
    // misses the instruction cache on every switch (both bodies together are > 40 KB): measured 228 ms against 162 ms per
    // 2^20 x 1000 steps with hold transitions.  A robot left to run ahead in its own loop would be worse still.
    const bool full = __any_sync(0xffffffffu, !(steady && !last));
    if (full) {
      // ================= full path: one step with every flag honoured =================
      if (hot) spill();
      const Rot R = make_rot(S);
      Wrench6 W = flexr_general_step<CPL, TPB, NF, LANES>(A, S, sm, sw, tab, c0, lead, valid, mode, now, head, sec, nsec, last, i);
      recheck = true;
      if (last && lead && valid) publish_platform(A, S, i);
      platform_step(R, W);
      tprev = now;
      ++s;
      if (A.snap_every > 0 && ++snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (lead && valid && snap_idx < A.snap_capacity) write_snapshot(A, S, snap_idx * 13 * A.snap_stride + A.snap_offset + i);
        ++snap_idx;
      }
      continue;
    }

    // ================= hot run: this step and the steps up to the next event =================
    int run = A.k_steps - 1 - s;  // the last step of the launch takes the full path (it publishes)
    if (A.sine_on) run = min(run, (sine_ctr == 0) ? 1 : A.sine_period - sine_ctr + 1);
    if (cmd_row) run = min(run, (cmd_ctr == 0) ? 1 : A.steps_per_cmd - cmd_ctr + 1);
    if (A.snap_every > 0) run = (int)min((long long)run, A.snap_every - snap_ctr);
    hot = true;
    {
      double *ring = sm + M::kRing * TPB;
      int r = 0;
#pragma unroll 1
      for (;;) {
        // ---- hot body: every cable of the lane on its live Pid, no flag
        const double dt = __dsub_rn(now, tprev);  // == now - mLastTime of every live Pid
        const Rot R = make_rot(S);
        Wrench6 W;
        W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
        W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
        // ring offsets of this step's slot and of the two samples before it (the P filter's x1, x2); FIR weights by slot
        int o1 = head - 1, o2 = head - 2;
        o1 += (o1 < 0) ? kFlexLen : 0;
        o2 += (o2 < 0) ? kFlexLen : 0;
        const int o0 = head * (CPL * TPB);
        o1 *= CPL * TPB; o2 *= CPL * TPB;
        const double *gw0 = A.firx0 + (kFlexLen - 1 - head);
        // kUnr cables per iteration, spelled out (an unroll pragma on the cable loop lets the compiler peel or clone it)
#pragma unroll 1
        for (int cb = 0; cb < CPL; cb += kUnr) {
#pragma unroll
        for (int cu = 0; cu < kUnr; ++cu) {
          const int c = cb + cu;
          // ---- every load of this cable first: nothing here depends on this step's sample
          double *rc_ = ring + c * TPB;
          double *fq = sm + (M::kFilt + c * M::FS) * TPB;
          const double older = has_fir ? flexr_fir_older<CPL * TPB>(gw0, rc_) : 0.0;
          const double desired = sm[(M::kDes + c) * TPB];
          const double prev_ie = sm[(M::kIerr + c) * TPB];
          double rx1 = 0.0, rx2 = 0.0, fp[NF > 0 ? 4 * NF : 1], fd[NF > 0 ? 4 * NF : 1];
          if (NF > 0) {
            rx1 = rc_[o1]; rx2 = rc_[o2];  // the P filter's x1, x2: the last two errors
#pragma unroll
            for (int f = 0; f < 4 * NF; ++f) { fp[f] = (f < 2) ? 0.0 : fq[f * TPB]; fd[f] = fq[(4 * NF + f) * TPB]; }
          }
          CableKin kin;
          if (LANES == 1 && kUnr >= CPL) {
            kin = cable_kin_v(rc.b[c][0], rc.b[c][1], rc.b[c][2], rc.a[c][0], rc.a[c][1], rc.a[c][2], rc.home_len[c], S, R);
          } else {
            const double *q = cabtab + c * 7;
            kin = cable_kin_v(q[0], q[1], q[2], q[3], q[4], q[5], q[6], S, R);
          }
          const bool pos = ((posmask >> c) & 1u) != 0u;
          const double *row = tab + (pos ? M::kRow : 0);
          const double actual = pos ? kin.qp : kin.qd;
          const FlexGains g = flexr_gains(row);
          const double e = __dsub_rn(desired, actual);
          // (no test for "this filter has no stage at all": a uniform branch here makes the compiler clone the whole cable loop,
          // and the clone costs more instruction fetch than the operations it saves; the select inside returns the input)
          double pe = e;
          if (NF > 0) pe = flexr_cascade_regs<NF, true>(A.flex_pf, fp, A.flex_ps, pos ? A.pc[1].p_casc : A.pc[0].p_casc, e, rx1, rx2);
          double derived = 0.0;
          if (has_fir) derived = fma(A.fir[kFlexLen - 1], e, older);
          double de = derived;
          if (NF > 0) de = flexr_cascade_regs<NF, false>(A.flex_df, fd, A.flex_ds, pos ? A.pc[1].d_casc : A.pc[0].d_casc, derived);
          const FlexrPidOut o = flexr_pid(g, rc.effort_limit_abs, desired, e, dt, pe, de, prev_ie);
          const double ie_new = o.ierr, eff = o.eff;
          // ---- every store of this cable last
          rc_[o0] = e;
          sm[(M::kIerr + c) * TPB] = ie_new;
          if (!((holdmask >> c) & 1u)) sm[(M::kLastp + c) * TPB] = kin.qp;  // mLastPosition follows the joint unless the cable holds (JointForceCalculator.cpp:78,84,88)
          if (NF > 0) {
#pragma unroll
            for (int f = 0; f < 4 * NF; ++f) {
              if (f >= 2) fq[f * TPB] = fp[f];   // (the P filter's first x1, x2 live in the ring)
              fq[(4 * NF + f) * TPB] = fd[f];
            }
          }
          const double tl = fma(eff, kin.il, __dmul_rn(__dmul_rn(-rc.cdamp, kin.qd), kin.il));  // tension / L = (effort - c q') / L, the damping part ahead of the effort
          W.fx = fma(tl, kin.dx, W.fx); W.fy = fma(tl, kin.dy, W.fy); W.fz = fma(tl, kin.dz, W.fz);
          W.mx = fma(tl, kin.cx, W.mx); W.my = fma(tl, kin.cy, W.my); W.mz = fma(tl, kin.cz, W.mz);
        }
        }
        platform_step(R, W);
        tprev = now;
        if (++r == run) break;
        clock_tick(now);
      }
    }
    // the steps of the run after the first saw no command: move the counters over them
    if (A.sine_on) { sine_ctr += run - 1; sine_ctr -= (sine_ctr >= A.sine_period) ? A.sine_period : 0; }
    if (cmd_row) { cmd_ctr += run - 1; cmd_ctr -= (cmd_ctr >= A.steps_per_cmd) ? A.steps_per_cmd : 0; }
    s += run;
    if (A.snap_every > 0) {
      snap_ctr += run;
      if (snap_ctr == A.snap_every) {
        snap_ctr = 0;
        if (lead && valid && snap_idx < A.snap_capacity) write_snapshot(A, S, snap_idx * 13 * A.snap_stride + A.snap_offset + i);
        ++snap_idx;
      }
    }
  }
  // the last step always runs the general body, so the controller state is back in shared memory here

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
