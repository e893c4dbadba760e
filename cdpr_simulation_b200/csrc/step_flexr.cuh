// step_flexr.cuh -- K2''r: the full-semantics kernel of step_flex.cuh with a REGISTER-resident hot body and on-chip gap fits.
//
// Same semantics, same HBM state and the same rare paths (flush / wake / reset / pending commands: the helpers of
// step_flex.cuh) as k_step_flex -- hold through the position Pid (JointForceCalculator.cpp:72-82), biquad cascades
// (Pid.cpp:27-44, Filter.h:152-165), the exact clamp chain of Pid::update (Pid.cpp:136-187), per-instance modes and command
// latches (CdprGazeboPlugin.cpp:67-83,206-219).  What changes is where the time went (ncu + SASS of k_step_flex, round 2):
//
//   * the hot body was 291 instructions per cable, 87 of them FP64: a rolled cable loop turns every robot constant, gain and
//     filter coefficient into an indexed constant load (48 LDC per cable) and every per-Pid value into a pair of selects
//     (30 FSEL per cable); the biquad state and the integrals went through shared memory (16 LDS/STS per cable).
//     Here the cable loop of the hot body is fully unrolled over the CPL = 4 (or 8) cables of a lane, the integrals and
//     the biquad state of the live Pids stay in REGISTERS while the thread is hot, the per-Pid gains / coefficients come
//     from a two-row table in shared memory indexed by "this cable runs the position Pid" (one broadcast LDS per value),
//     and which Pid a cable runs (hold) and its set point are worked out when a command arrives, not every step.
//   * a Pid that woke up fitted its gap-spanning window out of HBM for 11 steps: 22 dependent-latency loads, normal equations
//     and 4 divisions per cable and step, with 1-2 threads of the warp active -- 40 % of the whole run with hold
//     transitions every few hundred steps.  Here the shared-memory ring always holds the live Pid's whole window (the
//     stale samples sit in the slots the next pushes overwrite); when the stale part is a run of consecutive steps
//     (ctl bits 30/31, set when a Pid goes to sleep on a full window of fresh samples) its time stamps follow from one
//     value, so the fit needs no load at all.  Windows that are stale twice over keep the HBM fit of step_flex.cuh.
//
// Bodies are chosen per thread as in step_flex.cuh; both run the same inlined arithmetic helpers with explicit roundings, so
// which body ran never shows in the bits (GPU tests: bitwise launch-split and checkpoint invariance through hold transitions).
//
// Biquad slots: NF = 0 or 1 stage per filter.  A Pid without a stage where the other Pid has one runs the identity
// biquad (a0 = 1, rest 0; exact: 1 x + 0 = x), so the cable loop has no per-Pid stage count.  More stages, or the leg
// model: k_step_flex.
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
  static constexpr int kRow = 17;             // kf kp ki kd i_max i_max/ki c_max | P a0 a1 a2 b1 b2 | D a0 a1 a2 b1 b2
  static constexpr int kTabCab = 2 * kRow;    // [LANES][CPL][7]: b xyz, a xyz, home length
  static constexpr int kTabDoubles = kTabCab + LANES * CPL * 7;
  static constexpr size_t bytes = sizeof(double) * ((size_t)kPerThread * TPB + kTabDoubles);
};

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

// BiQuad::process (Filter.h:152-165) on four values: a0 x + a1 x1 + a2 x2 - b1 y1 - b2 y2, left to right, then the shift
__device__ __forceinline__ double biquad_step(const double *co, double &x1, double &x2, double &y1, double &y2, double x) {
  double y0 = __dmul_rn(co[0], x);
  y0 = __dadd_rn(y0, __dmul_rn(co[1], x1));
  y0 = __dadd_rn(y0, __dmul_rn(co[2], x2));
  y0 = __dsub_rn(y0, __dmul_rn(co[3], y1));
  y0 = __dsub_rn(y0, __dmul_rn(co[4], y2));
  x2 = x1; x1 = x; y2 = y1; y1 = y0;
  return y0;
}
// the same on the shared-memory columns of one filter (x1 x2 y1 y2)
template <int TPB>
__device__ __forceinline__ double biquad_step_sm(const double *co, double *q, double x) {
  double x1 = q[0], x2 = q[TPB], y1 = q[2 * TPB], y2 = q[3 * TPB];
  const double y0 = biquad_step(co, x1, x2, y1, y2, x);
  q[0] = x1; q[TPB] = x2; q[2 * TPB] = y1; q[3 * TPB] = y2;
  return y0;
}

__device__ __forceinline__ FlexGains flexr_gains(const double *row) {
  FlexGains g;
  g.kf = row[0]; g.kp = row[1]; g.ki = row[2]; g.kd = row[3]; g.i_max = row[4]; g.i_max_over_ki = row[5]; g.c_max = row[6];
  return g;
}

// (sec, nsec) - back * dt_ns as a gazebo time stamp, without 64-bit divisions
__device__ __forceinline__ void stamp_dec(int &sec, int &nsec, int dt_ns) {
  nsec -= dt_ns;
  while (nsec < 0) { nsec += 1000000000; --sec; }
}

// Derivative at `now` of the degree-D least-squares polynomial through 11 (stamp, value) pairs; xs[10] is the oldest stamp.
// The arithmetic of ls_derivative (step_general.cuh): window-relative, span-scaled time, normal equations, elimination.
template <int D>
__device__ __forceinline__ double ls_fit11(const double (&xs)[kFlexLen], const double (&ys)[kFlexLen], double now) {
  constexpr int M = D + 1;
  const double span = now - xs[kFlexLen - 1], inv_span = 1.0 / span;
  double sx[2 * D + 1], sy[M];
#pragma unroll
  for (int p = 0; p <= 2 * D; ++p) sx[p] = 0.0;
#pragma unroll
  for (int p = 0; p < M; ++p) sy[p] = 0.0;
#pragma unroll
  for (int j = 0; j < kFlexLen; ++j) {
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

// Gap fit on chip: ring position (head - a) holds the sample of age a; the newest `fresh` are the last steps, the others
// the run of consecutive steps that ended at `stale` (a gazebo time stamp: sec + nsec 1e-9 with nsec < 1e9, sec < 2^16).
template <int STRIDE>
static __device__ __noinline__ double flexr_gap_fit(int degree, const double *ringc, int head, unsigned fresh, double stale, int sec, int nsec, int dt_ns, double now) {
  double xs[kFlexLen], ys[kFlexLen];
  int s = sec, ns = nsec;
#pragma unroll
  for (int a = 0; a < kFlexLen; ++a) {
    if ((unsigned)a == fresh) {  // from here on the stale run: its newest stamp back to integers (exact below 2^16 s)
      const double fl = floor(stale);
      s = (int)fl;
      ns = (int)__double2ll_rn(__dmul_rn(__dsub_rn(stale, fl), 1e9));
    }
    xs[a] = time_double(s, ns);
    int sl = head - a;
    sl += (sl < 0) ? kFlexLen : 0;
    ys[a] = ringc[sl * STRIDE];
    stamp_dec(s, ns, dt_ns);
  }
  if (degree == 1) return ls_fit11<1>(xs, ys, now);
  if (degree == 2) return ls_fit11<2>(xs, ys, now);
  if (degree == 3) return ls_fit11<3>(xs, ys, now);
  if (degree == 4) return ls_fit11<4>(xs, ys, now);
  return 0.0;
}

// One step of THIS LANE's cables with every flag honoured (the out-of-line body): flex_general_step of step_flex.cuh with
// the table-driven gains, the always-present biquad slots and the on-chip window of a Pid that woke up.
template <int CPL, int TPB, int NF, int LANES>
static __device__ __noinline__ Wrench6 flexr_general_step(const StepArgs &A, FastState S, double *sm, unsigned *sw, const double *tab, int c0, bool lead, bool valid,
                                                          int mode, double now, int head, int sec, int nsec, bool last, long long i) {
  using M = FlexRSmem<CPL, TPB, NF, LANES>;
  const DevLayout &L = A.L;
  const RobotConsts &rc = A.rc;
  const Rot R = make_rot(S);
  Wrench6 W;
  W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
  W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
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
        flex_wake<CPL, TPB, NF>(A, sm, c0, c, (int)run - 1, i);
        sm[(M::kStale + c) * TPB] = sm[(M::kLtime + c) * TPB];  // its newest sample was pushed at its last update
        flex_load_window<CPL, TPB, NF>(A, sm, w, c0, c, (int)run - 1, slot_prev, i);
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
        if (NF > 0 && A.flex_ps > 0) pe = biquad_step_sm<TPB>(row + 7, sm + (M::kFilt + c * M::FS) * TPB, e);
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
            derived = flex_fir<CPL * TPB>(A, sm + (M::kRing + c) * TPB, head, e);
          } else {
            const double stale = sm[(M::kStale + c) * TPB];
            if (((w >> (kRunBit0 + k)) & 1u) && stale >= 0.0 && stale < 65536.0)
              derived = flexr_gap_fit<CPL * TPB>(A.pc[0].degree, sm + (M::kRing + c) * TPB, head, fresh, stale, sec, nsec, A.dt_ns, now);
            else
              derived = flex_gap_fit(A, cg, k, hd, now, i);
          }
        }
        double de = derived;
        if (NF > 0 && A.flex_ds > 0) de = biquad_step_sm<TPB>(row + 12, sm + (M::kFilt + c * M::FS + 4 * NF) * TPB, derived);
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

// HOLD = false: velocityEpsilon < 0, no cable can ever hold, the Pid follows the instance's mode alone
template <int NC, int TPB, int NF, bool HOLD, int LANES>
__global__ void __launch_bounds__(TPB) k_step_flexr(const __grid_constant__ StepArgs A) {
  constexpr int CPL = NC / LANES;
  static_assert(CPL * LANES == NC && (LANES == 1 || LANES == 2) && NF <= 1, "lanes must divide the cables; one biquad slot per filter");
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
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const double ident = (q == 0) ? 1.0 : 0.0;  // a Pid without a stage runs the identity biquad in the slot
        row[7 + q] = (pc.p_casc > 0) ? pc.pf[q] : ident;
        row[12 + q] = (pc.d_casc > 0) ? pc.df[q] : ident;
      }
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
      flex_wake<CPL, TPB, NF>(A, sm, c0, c, k, i);
      // the HBM ring is current at a launch boundary whatever `fresh` is: the whole window comes on chip
      flex_load_window<CPL, TPB, NF>(A, sm, w, c0, c, k, head0, i);
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
  // integrals of the live Pids; their biquad state: P y1 y2, D x1 x2 y1 y2 (the P filter's x1, x2 are the last two errors,
  // which a hot thread finds in the ring: its live Pids have pushed every one of the last 11 steps)
  double ierr[CPL], fst[CPL][NF > 0 ? 6 : 1];
  unsigned posmask = 0u, holdmask = 0u;  // per cable: runs the position Pid / holds (position Pid in Velocity mode)
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    ierr[c] = 0.0;
#pragma unroll
    for (int f = 0; f < (NF > 0 ? 6 : 1); ++f) fst[c][f] = 0.0;
  }
  bool hot = false;  // the previous step ran the hot body: the registers above are the truth, shared memory is stale
  const bool has_p = NF > 0 && A.flex_ps > 0, has_d = NF > 0 && A.flex_ds > 0, has_fir = A.pc[0].degree >= 1;

  auto spill = [&]() {  // runs after the clock tick of a step: `head` is the slot this step's sample WILL take
    int h1 = head - 1, h2 = head - 2;
    h1 += (h1 < 0) ? kFlexLen : 0;
    h2 += (h2 < 0) ? kFlexLen : 0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      sm[(M::kIerr + c) * TPB] = ierr[c];
      sm[(M::kLtime + c) * TPB] = tprev;
      if (NF > 0) {
        double *q = sm + (M::kFilt + c * M::FS) * TPB;
        q[0] = sm[(M::kRing + h1 * CPL + c) * TPB]; q[TPB] = sm[(M::kRing + h2 * CPL + c) * TPB];
        q[2 * TPB] = fst[c][0]; q[3 * TPB] = fst[c][1];
#pragma unroll
        for (int f = 0; f < 4; ++f) q[(4 + f) * TPB] = fst[c][2 + f];
      }
    }
    hot = false;
  };
  auto fill = [&]() {
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      ierr[c] = sm[(M::kIerr + c) * TPB];
      if (NF > 0) {
        const double *q = sm + (M::kFilt + c * M::FS) * TPB;
        fst[c][0] = q[2 * TPB]; fst[c][1] = q[3 * TPB];
#pragma unroll
        for (int f = 0; f < 4; ++f) fst[c][2 + f] = q[(4 + f) * TPB];
      }
    }
    hot = true;
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
        const double vel = (double)(float)__dmul_rn(sm[M::kSine * TPB], sin(arg));
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
    if (recheck) { steady = evaluate(); recheck = false; }

    const Rot R = make_rot(S);
    Wrench6 W;
    if (steady && !last) {
      // ================= hot body: straight-line, every cable of the lane on its live Pid, state in registers =================
      if (!hot) fill();
      const double dt = __dsub_rn(now, tprev);  // == now - mLastTime of every live Pid
      W.fx = lead ? rc.mg[0] : 0.0; W.fy = lead ? rc.mg[1] : 0.0; W.fz = lead ? rc.mg[2] : 0.0;
      W.mx = 0.0; W.my = 0.0; W.mz = 0.0;
      // ring offsets by sample age, once per step for all cables (age 0 = this step's slot)
      int ro[kFlexLen];
#pragma unroll
      for (int a = 0; a < kFlexLen; ++a) {
        int sl = head - a;
        sl += (sl < 0) ? kFlexLen : 0;
        ro[a] = sl * (CPL * TPB);
      }
      double *ring = sm + M::kRing * TPB;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        CableKin kin;
        if (LANES == 1) {
          kin = cable_kin_v(rc.b[c][0], rc.b[c][1], rc.b[c][2], rc.a[c][0], rc.a[c][1], rc.a[c][2], rc.home_len[c], S, R);
        } else {
          const double *q = cabtab + c * 7;
          kin = cable_kin_v(q[0], q[1], q[2], q[3], q[4], q[5], q[6], S, R);
        }
        // mLastPosition follows the joint unless the cable holds (JointForceCalculator.cpp:78,84,88)
        if (!((holdmask >> c) & 1u)) sm[(M::kLastp + c) * TPB] = kin.qp;
        const bool pos = ((posmask >> c) & 1u) != 0u;
        const double *row = tab + (pos ? M::kRow : 0);
        const double desired = sm[(M::kDes + c) * TPB];
        const double actual = pos ? kin.qp : kin.qd;
        const FlexGains g = flexr_gains(row);
        const double e = __dsub_rn(desired, actual);
        double *rc_ = ring + c * TPB;
        double pe = e;
        if (has_p) {
          double x1 = rc_[ro[1]], x2 = rc_[ro[2]];
          pe = biquad_step(row + 7, x1, x2, fst[c][0], fst[c][1], e);
        }
        rc_[ro[0]] = e;
        double derived = 0.0;
        if (has_fir) {  // flex_fir with the offsets above: same order of operations
          double d0 = __dmul_rn(A.fir[kFlexLen - 1], e), d1 = 0.0;
#pragma unroll
          for (int a = 1; a < kFlexLen; ++a) {
            const double y = rc_[ro[a]];
            if (a & 1) d1 = fma(A.fir[kFlexLen - 1 - a], y, d1); else d0 = fma(A.fir[kFlexLen - 1 - a], y, d0);
          }
          derived = __dadd_rn(d0, d1);
        }
        double de = derived;
        if (has_d) de = biquad_step(row + 12, fst[c][2], fst[c][3], fst[c][4], fst[c][5], derived);
        const FlexPidOut o = flex_pid(g, desired, e, dt, pe, de, ierr[c]);
        ierr[c] = o.ierr;
        const double eff = (rc.effort_limit >= 0.0) ? clampd(o.cmd, -rc.effort_limit, rc.effort_limit) : o.cmd;
        const double tl = __dmul_rn(fma(-rc.cdamp, kin.qd, eff), kin.il);  // tension / L
        W.fx = fma(tl, kin.dx, W.fx); W.fy = fma(tl, kin.dy, W.fy); W.fz = fma(tl, kin.dz, W.fz);
        W.mx = fma(tl, kin.cx, W.mx); W.my = fma(tl, kin.cy, W.my); W.mz = fma(tl, kin.cz, W.mz);
      }
    } else {
      if (hot) spill();
      W = flexr_general_step<CPL, TPB, NF, LANES>(A, S, sm, sw, tab, c0, lead, valid, mode, now, head, sec, nsec, last, i);
      recheck = true;
    }
    // ---- the robot's wrench = sum over its lanes; then every lane integrates the same platform step
    W.fx = lane_sum<LANES>(W.fx); W.fy = lane_sum<LANES>(W.fy); W.fz = lane_sum<LANES>(W.fz);
    W.mx = lane_sum<LANES>(W.mx); W.my = lane_sum<LANES>(W.my); W.mz = lane_sum<LANES>(W.mz);
    if (last && lead && valid) publish_platform(A, S, i);
    if (rc.spec & SPEC_ISO) rigid_body_step<SPEC_DIAG | SPEC_ISO>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    else if (rc.diag_inertia) rigid_body_step<SPEC_DIAG>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
    else rigid_body_step<0>(rc, S, R, W.fx, W.fy, W.fz, W.mx, W.my, W.mz);
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
