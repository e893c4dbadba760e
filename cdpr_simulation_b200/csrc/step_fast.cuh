// step_fast.cuh -- K2: the persistent K-step kernel (kinematics -> force law -> wrench -> rigid-body
// step), one thread per robot instance, state on chip for the whole launch.
//
// Covers reference rows a1,a3,a4,a5,a7,a8,a9,a11 of SURVEY.md 8(a) for the configurations in which
//   - velocity hold cannot trigger (velocityEpsilon < 0, launch/cdpr_gazebo.launch:18), so exactly one
//     Pid per cable is live (JointForceCalculator.cpp:71-89),
//   - no biquad stage is configured (cascade = 0, launch:29,32; CdprGazeboPlugin.cpp:133),
//   - cmdLimit != 0, iGain >= 0, window length 11,
// which is the reference's launch configuration.  Everything else runs in step_general.cuh.
//
// On-chip residency:  platform state, integral errors, targets, window sums + Kd D, flags -> registers
//                     D-term error windows (Pid::mDbufferY, LEN per cable)            -> shared memory,
//                     a ring [slot][cable][thread]: a warp touches 256 contiguous bytes per access
//
// D-term.  The reference fits a degree-d polynomial through the last LEN (time, error) samples and
// evaluates its derivative at `now` (Pid.cpp:193-247).  With uniform time stamps that is a fixed FIR
// sum_j w_j y_j whose weights are a degree-d polynomial in the sample position.  Two forms:
//   DMOM = false  plain FIR: LEN-1 shared-memory reads + LEN DFMA per cable and step (any degree);
//   DMOM = true   (degree <= 2) sliding moments S_m = sum_j p_j^m y_j, m = 0,1,2, with positions p_j = j + 1
//                 (oldest sample 1, newest LEN): after a slide every kept sample moves to p - 1 and the sample
//                 that leaves sits at 0, so it drops out of S1, S2 by itself:
//                   S0' = S0 - y_old + y_new,  S1' = S1 - S0 + LEN y_new,  S2' = S2 - 2 S1 + S0 + LEN^2 y_new
//                 (old S0, S1 on the right) and D = a S0 + b S1 + c S2.  S2 is only ever needed inside D, so the
//                 kernel carries (S0, S1, Kd D) instead: substituting the slide into D gives
//                   Kd D' = Kd D + Kd (a + LEN b + LEN^2 c) y_new - Kd a y_old + Kd (c - b) S0 - 2 Kd c S1
//                 -- 1 read + 1 write of shared memory and 8 FP64 instructions per cable and step, and the command
//                 is Ki Ierr + (Kp e + Kd D): 2 more.  Every kResync steps (global step index, so results do not
//                 depend on how steps are split into launches) S0, S1, S2 are re-summed from the ring and D rebuilt,
//                 which bounds the rounding drift of the recursion.
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace cdpr {

// Order in which the cable loop visits the cables: ascending, or -- SPEC_PAIR -- pair by pair (0, NC/2, 1, NC/2 + 1, ...), so
// the second cable of a pair follows the first while the shared kinematics are still in registers.  Every body of the
// kernel (hot, clamping, warm-up, last, saturated pass) uses the same order, so the wrench sums are the same bits everywhere.
template <int NC, int SPEC>
__device__ __forceinline__ constexpr int pair_order(int it) { return (SPEC & SPEC_PAIR) ? ((it & 1) ? (it >> 1) + NC / 2 : (it >> 1)) : it; }

// Wrench of one cable (tl = tension / L) into the running sums.  SPEC_PAIR: the two cables of a pair have the same d_x, d_y
// and (g x d)_z, so those three sums take the pair's summed tension once (1 DADD + 3 DFMA instead of 6 DFMA).
template <int SPEC>
__device__ __forceinline__ void wrench_add(int it, double tl, double &tl_first, const CableKin &k, double &fx, double &fy, double &fz, double &mx,
                                           double &my, double &mz) {
  if (SPEC & SPEC_PAIR) {
    fz = fma(tl, k.dz, fz); mx = fma(tl, k.cx, mx); my = fma(tl, k.cy, my);
    if (!(it & 1)) {
      tl_first = tl;
    } else {
      const double ts = tl_first + tl;
      fx = fma(ts, k.dx, fx); fy = fma(ts, k.dy, fy); mz = fma(ts, k.cz, mz);
    }
  } else {
    fx = fma(tl, k.dx, fx); fy = fma(tl, k.dy, fy); fz = fma(tl, k.dz, fz);
    mx = fma(tl, k.cx, mx); my = fma(tl, k.cy, my); mz = fma(tl, k.cz, mz);
  }
}

constexpr int kResync = 64;
constexpr int kSatHold = 32;  // clean steps before a warp that saw a clamp fire returns to the optimistic body
#ifndef CDPR_NC4_BLOCKS
#define CDPR_NC4_BLOCKS 2
#endif
#ifndef CDPR_NC8_BLOCKS
#define CDPR_NC8_BLOCKS 2
#endif
#ifndef CDPR_NC4_TPB
#define CDPR_NC4_TPB 128
#endif
#ifndef CDPR_NC8_TPB
#define CDPR_NC8_TPB 128
#endif
#ifndef CDPR_NC8_LEAN_TPB
#define CDPR_NC8_LEAN_TPB 128
#endif
#ifndef CDPR_NC8_LEAN_BLOCKS
#define CDPR_NC8_LEAN_BLOCKS 2
#endif
// "lean" = SPEC_UTGT | SPEC_NOFF: the one target lives in a register, so no target / feed-forward arrays in shared memory
template <int NC, int SPEC = 0> struct FastCfg {
  static constexpr bool lean = (SPEC & (8 | 16)) == (8 | 16);  // SPEC_NOFF | SPEC_UTGT
  // optimistic saturation handling parks the pre-update integrals in [NC][tpb] doubles of scratch; the non-lean
  // 8-cable layout has no room for it and commits the integrals after the vote instead (one more LDS + DFMA per cable)
  static constexpr bool scratch = (NC <= 4) || lean;
  static constexpr int tpb = (NC <= 4) ? CDPR_NC4_TPB : (lean ? CDPR_NC8_LEAN_TPB : CDPR_NC8_TPB);
  static constexpr int blocks = (NC <= 4) ? CDPR_NC4_BLOCKS : (lean ? CDPR_NC8_LEAN_BLOCKS : CDPR_NC8_BLOCKS);
};

// P + I + D (+ feed-forward `ff`) before any clamp (Pid.cpp:140-172).  DMOM: `d` is Kd * dErr already.
template <int SPEC, bool DMOM>
__device__ __forceinline__ double pid_command(const PidConsts &pc, double e, double ie, double d, double ff) {
  if (DMOM) return fma(pc.ki, ie, fma(pc.kp, e, (SPEC & SPEC_NOFF) ? d : d + ff));
  return (SPEC & SPEC_NOFF) ? fma(pc.ki, ie, fma(pc.kd, d, pc.kp * e)) : fma(pc.ki, ie, fma(pc.kd, d, fma(pc.kp, e, ff)));
}

// The exact clamping chain, from the integrated-but-unclamped integral `ie1`.
template <int SPEC, bool DMOM>
__device__ __forceinline__ void pid_clamped(const StepArgs &A, double dt, double e, double derr, double ff, double ie1, double prev_ierr,
                                            double &force, double &eff, double &ie) {
  const PidConsts &pc = A.live;
  // integral clamp with back-calculation (Pid.cpp:143-150) applied to the integral itself:
  // |Ki * Ierr| > Imax  <=>  |Ierr| > Imax / Ki (Ki >= 0 in this variant), clamped term = Ki * (Imax / Ki)
  ie = (fabs(ie1) > pc.i_max_over_ki) ? copysign(pc.i_max_over_ki, ie1) : ie1;
  const double cmd_raw = pid_command<SPEC, DMOM>(pc, e, ie, derr, ff);
  // clamp + anti-windup (Pid.cpp:175-184): mCmd != cmd  <=>  |cmd| > cmdMax
  const bool csat = fabs(cmd_raw) > pc.cmd_max;
  force = cmd_raw;
  if (csat) {  // rare: saturated command
    force = fma(dt * e, pc.ki, copysign(pc.cmd_max, cmd_raw));
    ie = prev_ierr;
  }
  // Joint::SetForce truncation can only bite on a saturated command when effortLimit >= cmdMax
  eff = force;
  if (csat || !A.effort_ge_cmd) eff = (fabs(force) > A.rc.effort_limit_abs) ? copysign(A.rc.effort_limit_abs, force) : force;
}

// Saturation (Pid.cpp:143-150 integral clamp, :175-184 command clamp + anti-windup, Joint::SetForce truncation) is
// rare, and handling it inline costs ~10 non-FP64 issue slots per cable.  The steady, non-last body (the hot loop)
// therefore runs OPTIMISTICALLY: it integrates and applies the unclamped command, and only accumulates one predicate
// "something in this step would have clamped".  One warp vote per step; if it fires, saturated_pass() recomputes every
// cable's force with the exact clamping chain and rebuilds the wrench from scratch in cable order.  That pass is
// bit-identical to the inline code the other bodies run, so a step's result does not depend on which body ran it.
// The integral before the update, needed by the anti-windup restore, is parked in shared memory (`prv`) where the
// layout has room (FastCfg::scratch); otherwise the integrals are committed after the vote from the ring's newest slot.
// Out of line and by value, so the hot loop's register allocation does not see it.
template <int NC, bool SCR> struct SatIn { FastState S; double ie1[NC], derr[NC]; double prev[SCR ? 1 : NC]; double tgu, dt; };
template <int NC> struct SatOut { double ierr[NC]; double fx, fy, fz, mx, my, mz; };
template <int NC, int MODE, bool DMOM, int SPEC>
__device__ __noinline__ SatOut<NC> saturated_pass(const StepArgs &A, SatIn<NC, FastCfg<NC, SPEC>::scratch> in, const double *tgts, const double *prv) {
  constexpr int kT = FastCfg<NC, SPEC>::tpb;
  const RobotConsts &rc = A.rc;
  const Rot R = make_rot(in.S);
  SatOut<NC> o;
  o.fx = rc.mg[0]; o.fy = rc.mg[1]; o.fz = rc.mg[2];
  o.mx = 0.0; o.my = 0.0; o.mz = 0.0;
  CableKin kpair;  // SPEC_PAIR: the second cable of the pair in flight
  double tl_first = 0.0;
#pragma unroll
  for (int it = 0; it < NC; ++it) {
    const int c = pair_order<NC, SPEC>(it);
    CableKin k;
    if (SPEC & SPEC_PAIR) {
      if (!(it & 1)) cable_kin_pair<SPEC, MODE == MODE_POSITION>(rc, in.S, R, c, c + NC / 2, rc.pair_dz[c], k, kpair);
      else k = kpair;
    } else {
      k = cable_kin<SPEC, MODE == MODE_POSITION>(rc, in.S, R, c);
    }
    const double tg = (SPEC & SPEC_UTGT) ? in.tgu : tgts[c * kT];
    const double ff = (SPEC & SPEC_NOFF) ? 0.0 : tgts[(NC + c) * kT];
    const double e = tg - ((MODE == MODE_VELOCITY) ? k.qd : k.qp);
    double force, eff;
    pid_clamped<SPEC, DMOM>(A, in.dt, e, in.derr[c], ff, in.ie1[c], FastCfg<NC, SPEC>::scratch ? prv[c * kT] : in.prev[c], force, eff, o.ierr[c]);
    const double tl = fma(-rc.cdamp, k.qd, eff) * k.il;
    wrench_add<SPEC>(it, tl, tl_first, k, o.fx, o.fy, o.fz, o.mx, o.my, o.mz);
  }
  return o;
}

// One physics step for one instance.
//   STEADY: every live Pid is primed and its window is full (mWasLastTime && mDbufferMissing == 0)
//   LAST:   last step of the launch: also writes the telemetry columns
//   MODE:   batch-uniform JointForceCalculator::UpdateMode
//   OPT:    optimistic saturation handling (steady, non-last steps only); returns the warp's vote "a clamp fired in
//           this step".  The clamping steady body (OPT = false, VOTE = true) returns the same vote, computed inline.
template <int NC, int LEN, bool STEADY, bool LAST, bool OPT, bool VOTE, int MODE, bool DMOM, int SPEC>
__device__ __forceinline__ bool fast_step(const StepArgs &A, FastState &S, double (&ierr)[NC], const double *__restrict__ tgts, const double tgu,
                                          double (&mom)[NC][3], unsigned &primed, unsigned (&missing)[NC],
                                          double *__restrict__ win, double *__restrict__ prv, int head, double dt, long long i) {
  constexpr int kT = FastCfg<NC, SPEC>::tpb;
  static_assert(!OPT || (STEADY && !LAST && MODE != MODE_FORCE), "optimistic steps are steady, non-last Pid steps");
  constexpr bool SCR = FastCfg<NC, SPEC>::scratch;
  const RobotConsts &rc = A.rc;
  const PidConsts &pc = A.live;
  const Rot R = make_rot(S);

  double fx = rc.mg[0], fy = rc.mg[1], fz = rc.mg[2];
  double mx = 0.0, my = 0.0, mz = 0.0;

  // ring offsets by sample age (0 = this step's slot, which still holds the sample leaving the window)
  int slot[DMOM ? 1 : LEN];
#pragma unroll
  for (int a = 0; a < (DMOM ? 1 : LEN); ++a) {
    int s = head - a;
    s += (s < 0) ? LEN : 0;
    slot[a] = s * (NC * kT);
  }
  // least-squares derivative at `now` from the window state AFTER this step's sample went in (Pid.cpp:193-217);
  // DMOM: times Kd
  auto dterm = [&](int c, double e, const double *w) {
    if (DMOM) return mom[c][2];
    double d0 = A.fir[LEN - 1] * e, d1 = 0.0;
#pragma unroll
    for (int a = 1; a < LEN; ++a) {
      if (a & 1) d1 = fma(A.fir[LEN - 1 - a], w[slot[a]], d1);
      else d0 = fma(A.fir[LEN - 1 - a], w[slot[a]], d0);
    }
    return d0 + d1;
  };

  bool sat = false;
  CableKin kpair;  // SPEC_PAIR: the second cable of the pair in flight
  double tl_first = 0.0;
#pragma unroll
  for (int it = 0; it < NC; ++it) {
    const int c = pair_order<NC, SPEC>(it);
    CableKin k;
    if (SPEC & SPEC_PAIR) {
      if (!(it & 1)) cable_kin_pair<SPEC, MODE == MODE_POSITION || LAST>(rc, S, R, c, c + NC / 2, rc.pair_dz[c], k, kpair);
      else k = kpair;
    } else {
      k = cable_kin<SPEC, MODE == MODE_POSITION || LAST>(rc, S, R, c);
    }

    // ---- force law (a3, a4, a5)
    double force, eff;
    if (MODE == MODE_FORCE) {
      force = tgts[c * kT];  // JointForceCalculator.cpp:67-70
      eff = (fabs(force) > rc.effort_limit_abs) ? copysign(rc.effort_limit_abs, force) : force;
    } else {
      const double tg = (SPEC & SPEC_UTGT) ? tgu : tgts[c * kT];
      const double e = tg - ((MODE == MODE_VELOCITY) ? k.qd : k.qp);
      double *w = win + c * kT;
      if (STEADY || ((primed >> c) & 1u)) {  // Pid.cpp:127-187
        const double prev_ierr = ierr[c];
        const double ie1 = fma(dt, e, prev_ierr);
        // derive(): push the sample
        if (DMOM) {
          const double y_old = w[slot[0]];
          w[slot[0]] = e;
          const double s0 = mom[c][0], s1 = mom[c][1];
          // y_old (a shared-memory read) enters last, so its latency hides behind the rest of the chain
          mom[c][2] = fma(A.dk[3], y_old, fma(A.dk[2], s1, fma(A.dk[1], s0, fma(A.dk[0], e, mom[c][2]))));
          mom[c][1] = fma((double)LEN, e, s1 - s0);
          mom[c][0] = (s0 + e) - y_old;
        } else {
          w[slot[0]] = e;
        }
        double derr = dterm(c, e, w);
        if (!STEADY) {
          missing[c] -= (missing[c] > 0u) ? 1u : 0u;
          if (missing[c] != 0u) derr = 0.0;
        }
        const double ff = (SPEC & SPEC_NOFF) ? 0.0 : tgts[(NC + c) * kT];
        if (OPT) {
          force = pid_command<SPEC, DMOM>(pc, e, ie1, derr, ff);
          eff = force;
          sat = sat || (fabs(ie1) > pc.i_max_over_ki) || (fabs(force) > A.sat_thr);
          if (SCR) { prv[c * kT] = prev_ierr; ierr[c] = ie1; }
        } else {
          double ie;
          pid_clamped<SPEC, DMOM>(A, dt, e, derr, ff, ie1, prev_ierr, force, eff, ie);
          if (VOTE) sat = sat || (ie != ie1) || (fabs(force) > A.sat_thr);  // some clamp changed the integral or the force
          ierr[c] = ie;
        }
        if (LAST) {
          A.L.pid[pid_off(A.L, c, A.live_idx, PID_P_ERR) + i] = e;
          double derr_out = derr;
          if (DMOM) {  // the carried value is Kd * dErr (and Kd may be 0): publish the FIR over the ring instead
            derr_out = 0.0;
            if (STEADY || missing[c] == 0u) {
#pragma unroll
              for (int a = 0; a < LEN; ++a) {
                int sl = head - a;
                sl += (sl < 0) ? LEN : 0;
                derr_out = fma(A.fir[LEN - 1 - a], w[sl * (NC * kT)], derr_out);
              }
            }
          }
          A.L.pid[pid_off(A.L, c, A.live_idx, PID_D_ERR) + i] = derr_out;
          // topic "pid": pTerm, iTerm before its clamp, dTerm, desired (Pid.cpp:140-141,159,167)
          A.L.cab[cab_off(A.L, c, CAB_TERM_P) + i] = pc.kp * e;
          A.L.cab[cab_off(A.L, c, CAB_TERM_I) + i] = pc.ki * ie1;
          A.L.cab[cab_off(A.L, c, CAB_TERM_D) + i] = pc.kd * derr_out;
          A.L.cab[cab_off(A.L, c, CAB_DESIRED) + i] = tg;
        }
      } else {  // first update after a reset: Pid.cpp:123-126
        primed |= 1u << c;
        force = 0.0;
        eff = 0.0;
      }
      if (LAST) A.L.pid[pid_off(A.L, c, A.live_idx, PID_CMD) + i] = force;
    }
    if (LAST) {
      publish_joint(A, NC, c, k.qp, k.qd, eff, i);
      A.L.cab[cab_off(A.L, c, CAB_LAST_POS) + i] = k.qp;
      A.L.cab[cab_off(A.L, c, CAB_EFFORT) + i] = eff;
      A.L.cab[cab_off(A.L, c, CAB_PID_FORCE) + i] = force;
    }
    // ---- explicit joint damping, wrench (a8)
    const double tl = fma(-rc.cdamp, k.qd, eff) * k.il;  // tension / L
    wrench_add<SPEC>(it, tl, tl_first, k, fx, fy, fz, mx, my, mz);
  }

  bool fired = false;
  if (VOTE && !OPT) fired = __any_sync(0xffffffffu, sat);
  if (OPT) {
    // the ring's newest slot holds this step's error of every cable
    fired = __any_sync(0xffffffffu, sat);
    if (fired) {  // rare: redo the force law of this step with the clamps
      SatIn<NC, SCR> in;
      in.S = S; in.tgu = tgu; in.dt = dt;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double e = win[c * kT + slot[0]];
        if (!SCR) in.prev[c] = ierr[c];
        in.ie1[c] = SCR ? ierr[c] : fma(dt, e, ierr[c]);
        in.derr[c] = dterm(c, e, win + c * kT);
      }
      const SatOut<NC> o = saturated_pass<NC, MODE, DMOM, SPEC>(A, in, tgts, prv);
#pragma unroll
      for (int c = 0; c < NC; ++c) ierr[c] = o.ierr[c];
      fx = o.fx; fy = o.fy; fz = o.fz; mx = o.mx; my = o.my; mz = o.mz;
    } else if (!SCR) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ierr[c] = fma(dt, win[c * kT + slot[0]], ierr[c]);
    }
  }

  if (LAST) publish_platform(A, S, i);
  rigid_body_step<SPEC>(rc, S, R, fx, fy, fz, mx, my, mz);
  return fired;
}

// exact re-summation of the window sums from the ring (newest sample in slot `head`), every kResync steps.  Sample-major
// and fully unrolled: one slot address per sample age, NC independent accumulator triples -- a rolled per-cable loop
// is a serial LDS -> FMA chain and cost 6 % of the whole kernel.
template <int NC, int LEN, int SPEC>
__device__ __forceinline__ void resync_moments(const StepArgs &A, double (&mom)[NC][3], const double *__restrict__ win, int head) {
  double s2[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { mom[c][0] = 0.0; mom[c][1] = 0.0; s2[c] = 0.0; }
  int sl = head + 1;  // oldest sample
#pragma unroll
  for (int j = 0; j < LEN; ++j) {
    sl -= (sl >= LEN) ? LEN : 0;
    const double *w = win + sl * (NC * FastCfg<NC, SPEC>::tpb);
    const double p = (double)(j + 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double y = w[c * FastCfg<NC, SPEC>::tpb];
      mom[c][0] += y;
      mom[c][1] = fma(p, y, mom[c][1]);
      s2[c] = fma(p * p, y, s2[c]);
    }
    ++sl;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) mom[c][2] = A.live.kd * fma(A.dmom[0], mom[c][0], fma(A.dmom[1], mom[c][1], A.dmom[2] * s2[c]));
}

// shared memory per block (doubles): ring [LEN][NC][tpb], targets [NC][tpb], feed-forward terms Kf*target [NC][tpb],
// sine parameters [3][tpb], pre-update integrals [NC][tpb] (optimistic variants)
template <int NC, int LEN, int SPEC>
constexpr size_t fast_smem_bytes() { return sizeof(double) * (size_t)FastCfg<NC, SPEC>::tpb * (LEN * NC + (FastCfg<NC, SPEC>::lean ? 0 : 2 * NC) + 3 + (FastCfg<NC, SPEC>::scratch ? NC : 0)); }

template <int NC, int LEN, int MODE, bool DMOM, int SPEC>
__global__ void __launch_bounds__(FastCfg<NC, SPEC>::tpb, FastCfg<NC, SPEC>::blocks) k_step_fast(const __grid_constant__ StepArgs A) {
  extern __shared__ double smem[];
  constexpr int kTpbL = FastCfg<NC, SPEC>::tpb;
  constexpr bool PIDMODE = (MODE != MODE_FORCE);
  const int tid = threadIdx.x;
  const long long gi = (long long)blockIdx.x * kTpbL + tid;
  const bool valid = gi < A.L.n;
  const long long i = valid ? gi : (long long)A.L.n - 1;  // tail threads shadow the last instance, never store
  const long long np = A.L.np;
  const int live = A.live_idx;
  double *mywin = smem + tid;                          // [LEN][NC][tpb]
  double *mytgt = smem + LEN * NC * kTpbL + tid;       // [NC][tpb]
  constexpr bool kLean = FastCfg<NC, SPEC>::lean;       // no target arrays: mytgt is never dereferenced
  double *mysine = mytgt + (kLean ? 0 : 2 * NC * kTpbL);  // [3][tpb]: amp, freq, phase
  double *myprv = mysine + 3 * kTpbL;                      // [NC][tpb], optimistic variants only

  FastState S;
  load_plat(A.L, i, S);
  double ierr[NC], mom[NC][3];
  unsigned primed = 0, missing[NC];
  constexpr int tgt_field = (MODE == MODE_FORCE) ? CAB_FORCE_CMD : (MODE == MODE_POSITION) ? CAB_POS_TARGET : CAB_VEL_TARGET;
  // the ring slot of a sample is (its step index) mod LEN, so the layout does not depend on launch boundaries
  const int head0 = (int)(A.n0 % LEN);  // slot of the newest sample already in the window
  bool steady = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ierr[c] = A.L.pid[pid_off(A.L, c, live, PID_I_ERR) + i];
    if (!kLean) {
      mytgt[c * kTpbL] = A.L.cab[cab_off(A.L, c, tgt_field) + i];
      mytgt[(NC + c) * kTpbL] = A.live.kf * mytgt[c * kTpbL];
    }
    const unsigned ctl = A.L.ctl[(long long)c * np + i];
    primed |= ((ctl >> live) & 1u) << c;
    missing[c] = (ctl >> (8 + 8 * live)) & 0xffu;
    steady = steady && ((ctl >> live) & 1u) && missing[c] == 0u;
#pragma unroll
    for (int m = 0; m < 3; ++m) mom[c][m] = 0.0;
    if (PIDMODE) {
#pragma unroll
      for (int j = 0; j < LEN; ++j) {  // logical j (oldest first) -> slot (head0 + 1 + j) mod LEN
        int sl = head0 + 1 + j;
        sl -= (sl >= LEN) ? LEN : 0;
        mywin[(sl * NC + c) * kTpbL] = A.L.win_y[win_off(A.L, c, live, j) + i];
      }
      if (DMOM) {
#pragma unroll
        for (int m = 0; m < 3; ++m) mom[c][m] = A.L.mom[mom_off(A.L, c, live, m) + i];
      }
    }
  }
  if (A.sine_on) {
#pragma unroll
    for (int m = 0; m < 3; ++m) mysine[m * kTpbL] = A.L.sine[m * np + i];
  }
  // the lean layout is only launched for the sine publisher without a command table or rollout cost (api.cu), so its
  // event code drops those checks at compile time
  const float *cmd_row = nullptr;
  if (!kLean && A.cmd_table) cmd_row = A.cmd_table + (size_t)(i % A.n_seq) * A.n_cmd * NC;
  const bool sine_on = kLean || A.sine_on;
  const bool want_cost = !kLean && A.cost != nullptr;
  double cost = 0.0;
  double tgu = A.L.cab[cab_off(A.L, 0, tgt_field) + i];  // SPEC_UTGT: the one target shared by all cables, kept in a register

  bool warp_steady = __all_sync(0xffffffffu, steady || !PIDMODE);
  int sec = A.sec0, nsec = A.nsec0, head = head0;
  double tprev = A.t0, sine_time = A.sine_time0;
  int sine_ctr = (int)(A.n0 % (A.sine_period > 0 ? A.sine_period : 1));
  int cmd_ctr = 0, cmd_idx = 0;
  int resync_ctr = (int)(A.n0 % kResync);
  long long snap_idx = A.snap_written0;
  long long snap_ctr = A.snap_every > 0 ? (A.n0 % A.snap_every) : 0;
  // The K steps run as short inner loops between EVENTS, so the hot loop body holds nothing but the clock, the force
  // law and the physics.  Events before a step: a new command (sine publisher, command table).  Events after a step:
  // moment re-summation, snapshot.  All event schedules depend on the global step index only.
  auto clock_tick = [&](double &dt) {
    // World::Step: simTime += dt, then the plugin callback
    nsec += A.dt_ns;
    if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; }
    const double t = time_double(sec, nsec);
    dt = __dsub_rn(t, tprev);
    tprev = t;
    head = (head + 1 == LEN) ? 0 : head + 1;
  };
  auto events_before = [&]() {
    if (sine_on && sine_ctr == 0) {  // sinevelocitytest.cpp:35-38,48: float32 axes, accumulated publisher time
      const double amp = mysine[0], freq = mysine[kTpbL], phase = mysine[2 * kTpbL];
      const double arg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(sine_time, freq), 2.0), 3.14159265358979323846), phase);
      const double vel = publisher_value(A.pub_shape, amp, sin(arg));
      if (!kLean) {
#pragma unroll
        for (int c = 0; c < NC; ++c) { mytgt[c * kTpbL] = vel; mytgt[(NC + c) * kTpbL] = A.live.kf * vel; }
      }
      tgu = vel;
      sine_time = __dadd_rn(sine_time, A.sine_pub_dt);
    }
    if (!kLean && cmd_row && cmd_ctr == 0 && cmd_idx < A.n_cmd) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double v = (double)cmd_row[cmd_idx * NC + c];
        mytgt[c * kTpbL] = v; mytgt[(NC + c) * kTpbL] = A.live.kf * v;
      }
      ++cmd_idx;
    }
  };
  // how many steps may run before the next event of any kind (at least 1)
  auto run_length = [&](int remaining) {
    int run = remaining;
    if (sine_on) run = min(run, A.sine_period - sine_ctr);
    if (!kLean && cmd_row) run = min(run, A.steps_per_cmd - cmd_ctr);
    if (DMOM && PIDMODE) run = min(run, kResync - resync_ctr);
    if (A.snap_every > 0) run = (int)min((long long)run, A.snap_every - snap_ctr);
    return run;
  };
  auto advance_counters = [&](int run) {
    if (sine_on) { sine_ctr += run; if (sine_ctr >= A.sine_period) sine_ctr = 0; }
    if (!kLean && cmd_row) { cmd_ctr += run; if (cmd_ctr >= A.steps_per_cmd) cmd_ctr = 0; }
    if (DMOM && PIDMODE) resync_ctr += run;
    if (A.snap_every > 0) snap_ctr += run;
  };
  auto events_after = [&]() {
    if (DMOM && PIDMODE && resync_ctr >= kResync) {
      resync_ctr = 0;
      resync_moments<NC, LEN, SPEC>(A, mom, mywin, head);
    }
    if (A.snap_every > 0 && snap_ctr >= A.snap_every) {
      snap_ctr = 0;
      if (valid && snap_idx < A.snap_capacity) write_snapshot(A, S, snap_idx * 13 * A.snap_stride + A.snap_offset + i);
      ++snap_idx;
    }
  };
  auto add_cost = [&]() {
    const double ex = S.px - A.target[0], ey = S.py - A.target[1], ez = S.pz - A.target[2];
    cost += fma(ex, ex, fma(ey, ey, ez * ez)) + A.lambda * fma(S.wx, S.wx, fma(S.wy, S.wy, S.wz * S.wz));
  };

  int s = 0, sat_hold = 0;
  const int k_body = A.k_steps - 1;  // the last step runs separately (it also publishes telemetry)
  while (s < k_body) {
    events_before();
    if (!warp_steady) {
      // warm-up (at most LEN + 1 steps after a Pid reset): some live Pid of the warp is un-primed or its window is not full
      double dt;
      clock_tick(dt);
      fast_step<NC, LEN, false, false, false, false, MODE, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
      if (want_cost) add_cost();
      bool st = true;
#pragma unroll
      for (int c = 0; c < NC; ++c) st = st && ((primed >> c) & 1u) && missing[c] == 0u;
      warp_steady = __all_sync(0xffffffffu, st);
      if (warp_steady && PIDMODE) {  // the flags are constants from here on: keep them out of the hot loop's registers
        primed = (1u << NC) - 1u;
#pragma unroll
        for (int c = 0; c < NC; ++c) missing[c] = 0u;
      }
      advance_counters(1);
      ++s;
    } else {
      const int run = run_length(k_body - s);
      // Two interchangeable steady bodies (same bits out): the optimistic one, and -- while clamps keep firing -- the
      // one that clamps inline.  sat_hold counts down the clean steps left before the warp goes back to optimistic.
      int r = 0;
      while (r < run) {
        if (!PIDMODE) {
          for (; r < run; ++r) {
            double dt;
            clock_tick(dt);
            fast_step<NC, LEN, true, false, false, false, MODE, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
            if (want_cost) add_cost();
          }
        } else if (sat_hold > 0) {
          for (; r < run && sat_hold > 0; ++r) {
            double dt;
            clock_tick(dt);
            const bool fired = fast_step<NC, LEN, true, false, false, true, PIDMODE ? MODE : MODE_VELOCITY, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
            sat_hold = fired ? kSatHold : sat_hold - 1;
            if (want_cost) add_cost();
          }
        } else if (want_cost) {
          for (; r < run;) {
            double dt;
            clock_tick(dt);
            const bool fired = fast_step<NC, LEN, true, false, true, true, PIDMODE ? MODE : MODE_VELOCITY, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
            add_cost();
            ++r;
            if (fired) { sat_hold = kSatHold; break; }
          }
        } else {
          bool fired;
          do {  // the hot loop
            double dt;
            clock_tick(dt);
            fired = fast_step<NC, LEN, true, false, true, true, PIDMODE ? MODE : MODE_VELOCITY, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
            ++r;
          } while (!fired && r < run);
          if (fired) sat_hold = kSatHold;
        }
      }
      advance_counters(run);
      s += run;
    }
    events_after();
  }
  if (A.k_steps > 0) {  // last step of the launch
    events_before();
    double dt;
    clock_tick(dt);
    if (valid) {
      if (warp_steady) fast_step<NC, LEN, true, true, false, false, MODE, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
      else fast_step<NC, LEN, false, true, false, false, MODE, DMOM, SPEC>(A, S, ierr, mytgt, tgu, mom, primed, missing, mywin, myprv, head, dt, i);
    }
    if (want_cost) add_cost();
    advance_counters(1);
    events_after();
  }

  if (!valid) return;
  store_plat(A.L.plat + i, np, S);
  if (want_cost) A.cost[i] = cost;
  if (!PIDMODE) return;  // Force mode touches no Pid state
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    A.L.pid[pid_off(A.L, c, live, PID_I_ERR) + i] = ierr[c];
    A.L.pid[pid_off(A.L, c, live, PID_LAST_TIME) + i] = tprev;
    if (sine_on || cmd_row) A.L.cab[cab_off(A.L, c, CAB_VEL_TARGET) + i] = kLean ? tgu : mytgt[c * kTpbL];
    unsigned ctl = A.L.ctl[(long long)c * np + i];
    ctl &= ~((1u << live) | (0xffu << (8 + 8 * live)));
    ctl |= (((primed >> c) & 1u) << live) | (missing[c] << (8 + 8 * live));
    A.L.ctl[(long long)c * np + i] = ctl;
    for (int j = 0; j < LEN; ++j) {
      int sl = head + 1 + j;
      sl -= (sl >= LEN) ? LEN : 0;
      A.L.win_y[win_off(A.L, c, live, j) + i] = mywin[(sl * NC + c) * kTpbL];
    }
    if (DMOM) {
#pragma unroll
      for (int m = 0; m < 3; ++m) A.L.mom[mom_off(A.L, c, live, m) + i] = mom[c][m];
    }
  }
}

}  // namespace cdpr
