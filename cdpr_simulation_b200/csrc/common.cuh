// common.cuh -- device-side data layout and constant blocks of the batched CDPR step.
//
// HBM layout (struct-of-arrays; instance index i is the fastest-varying one; Np = N padded to a
// multiple of 128 so every column starts 1 KB aligned):
//   plat  [13][Np]                px py pz | qw qx qy qz | vx vy vz | wx wy wz
//   cab   [NC][CAB_F][Np]         last_pos, force_cmd, pos_target, vel_target, effort, pid_force, pid terms P I D, desired
//   pid   [NC][2][PID_F][Np]      last_time, p_err, i_err, d_err, cmd      (pid 0 = velocity, 1 = position)
//   win_y [NC][2][LEN][Np]        D-term error window, logical order (oldest first) = Pid::mDbufferY
//   mom   [NC][2][3][Np]          window state S0, S1, Kd*D of the fast variant's D-term (see step_fast.cuh)
//   win_x [NC][2][LEN][Np]        D-term time stamps = Pid::mDbufferX      (general variant only; there win_x/win_y are
//                                 rings whose head lives in ctl, see step_general.cuh)
//   filt  [NC][2][2][CASC][4][Np] biquad x1 x2 y1 y2 (P filter, D filter)  (general variant only)
//   ctl   [NC][Np] uint32         bit0 vel.wasLast, bit1 pos.wasLast, bits8-15 vel.missing, 16-23 pos.missing
//   ictl  [Np] uint32             flex variant: bits 0-1 UpdateMode of the instance, bit 2 / 3 velocity / position command pending
//   sine  [3][Np]                 amp, freq, phase of the in-kernel sinevelocitytest generator
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cdpr {

constexpr int kMaxCables = 8;
constexpr int kMaxDbuf = 32;
constexpr int kMaxDegree = 4;
constexpr int kMaxCascade = 4;
constexpr int kTpb = 128;  // threads per block of the step kernels = instances per block

// CAB_TERM_P .. CAB_DESIRED: what the reference publishes on topic "pid" from inside Pid::update (pTerm, pre-clamp iTerm,
// dTerm, desired; Pid.cpp:140-141,159,167) -- written by the last step of a launch, untouched by a priming update
enum CabField { CAB_LAST_POS = 0, CAB_FORCE_CMD, CAB_POS_TARGET, CAB_VEL_TARGET, CAB_EFFORT, CAB_PID_FORCE,
                CAB_TERM_P, CAB_TERM_I, CAB_TERM_D, CAB_DESIRED, CAB_F };
enum PidField { PID_LAST_TIME = 0, PID_P_ERR, PID_I_ERR, PID_D_ERR, PID_CMD, PID_F };
enum { PID_VEL = 0, PID_POS = 1 };
enum { MODE_FORCE = 0, MODE_POSITION = 1, MODE_VELOCITY = 2 };

struct DevLayout {
  double *plat, *cab, *pid, *win_y, *win_x, *filt, *sine, *mom;
  uint32_t *ctl;
  uint32_t *ictl;  // [Np] per-instance word of the flex variant: UpdateMode + pending-command bits (step_flex.cuh)
  long long np;  // padded instance count (column stride)
  int n;         // live instances
  int nc, len, casc;
};

__host__ __device__ inline long long cab_off(const DevLayout &L, int c, int f) { return ((long long)c * CAB_F + f) * L.np; }
__host__ __device__ inline long long pid_off(const DevLayout &L, int c, int k, int f) { return (((long long)c * 2 + k) * PID_F + f) * L.np; }
__host__ __device__ inline long long win_off(const DevLayout &L, int c, int k, int j) { return (((long long)c * 2 + k) * L.len + j) * L.np; }
__host__ __device__ inline long long mom_off(const DevLayout &L, int c, int k, int m) { return (((long long)c * 2 + k) * 3 + m) * L.np; }
__host__ __device__ inline long long filt_off(const DevLayout &L, int c, int k, int pd, int s, int f) {
  return (((((long long)c * 2 + k) * 2 + pd) * L.casc + s) * 4 + f) * L.np;
}

// gains of one gazebo::common::Pid, pre-digested on the host (Pid.cpp:63-77)
struct PidConsts {
  double kf, kp, ki, kd;
  double i_max, i_min, i_max_over_ki, i_min_over_ki;  // iTerm / mIgain of Pid.cpp:145,149
  double cmd_max, cmd_min;
  int degree, len, p_casc, d_casc;
  double pf[5], df[5];  // biquad a0 a1 a2 b1 b2 (Filter.h:130-140), identical for every stage
};

// robot constants (sdf/cube.sdf) in the form the kernels consume; passed as a __grid_constant__
// kernel parameter so compile-time-indexed reads become constant-bank operands of DFMA.
struct RobotConsts {
  double a[kMaxCables][3];  // frame anchors
  double b[kMaxCables][3];  // platform anchors (body frame)
  double home_len[kMaxCables];
  double mg[3];             // mass * gravity
  double h, h_over_m, half_h;
  double ib[6], ib_inv[6];  // xx yy zz xy xz yz
  int diag_inertia;
  int spec;                 // SPEC_* bits (physics.cuh) that hold for this robot
  double pair_dz[kMaxCables / 2];  // SPEC_PAIR: a[c + NC/2][2] - a[c][2]
  double cdamp, effort_limit;  // effort_limit < 0: no truncation
  double effort_limit_abs;     // effort_limit, or +inf when truncation is off
  double vel_eps;
  // leg fidelity (legs.cuh): 0 = massless legs; square roots of link inertia / mass / passive damping (and of twice them), COM
  // offset of the cable link, rev_X axis (frame), rev_Zpf axis in the leg triad, rev_Xpf axis (platform body)
  int leg_model;
  double mass, grav[3];
  double leg_sI, leg_s2I, leg_sm, leg_s2m, leg_sc, leg_lc;
  double leg_x0[kMaxCables][3], leg_alpha[kMaxCables][3], leg_a1[kMaxCables][3];
};

struct StepArgs {
  DevLayout L;
  RobotConsts rc;
  PidConsts pc[2];     // [PID_VEL], [PID_POS]
  PidConsts live;      // copy of pc[live_idx]: the Pid the fast kernel runs
  int live_idx;
  double fir[kMaxDbuf];  // D-term FIR weights of the LIVE pid, logical order, already divided by the window span
  double fir2[2][kMaxDbuf];  // FIR weights of BOTH pids (general variant: used whenever a window's time stamps are uniform)
  double dmom[3];        // the same weights as a quadratic in the centred sample position: w_j = dmom[0] + dmom[1] k + dmom[2] k^2
  double dk[4];          // Kd * D recursion of the fast variant (step_fast.cuh): coefficients of y_new, S0, S1, y_old
  int flex_ps, flex_ds;  // flex variant: biquad stages held on chip for the P input / D output (max over the two Pids)
  // step_flexr.cuh: the ONE coefficient set of the P-input / D-output stages (the Pids that have stages share it, api.cu checks)
  double flex_pf[5], flex_df[5];
  double firx[21];       // step_flexr.cuh: fir[0..10] twice over, so that the weight of ring SLOT s at ring head h is firx[s + 10 - h]
  double firx0[21];      // ... with the newest sample's weight (fir[10]) replaced by 0: the weights of the ten older samples
  int effort_ge_cmd;     // effort limit >= cmdMax of the live pid: truncation can only bite on a saturated command
  double sat_thr;        // min(cmdMax, effort limit): an unclamped command within it passes every clamp unchanged
  int mode;            // batch-uniform JointForceCalculator::UpdateMode
  int k_steps;
  long long n0;        // physics steps done before this launch
  int sec0, nsec0, dt_ns;
  double t0;           // time_double(sec0, nsec0)
  // sine generator (sinevelocitytest.cpp)
  int sine_on, sine_period;
  double sine_time0, sine_pub_dt;
  int pub_shape;       // wave form of the in-kernel publisher: 0 sinevelocitytest, 1 squarevelocitytest (dead band)
  // snapshots
  // snapshots: every peer buffer is [capacity][13][snap_stride]; this handle's instances start at column snap_offset.
  // One local buffer (stride n, offset 0), or the gather buffers of ALL ranks mapped over NVLink (fused all-gather).
  double *snap_peers[8];
  int n_snap_peers;
  int snap_multimem;  // snap_peers[0] is an NVLS multicast address: one multimem.st reaches every rank's buffer
  long long snap_stride, snap_offset;
  long long snap_every, snap_written0, snap_capacity;
  // cdpr_update: what the plugin publishes from inside update() (CdprGazeboPlugin.cpp:248-280), written by the LAST step of the
  // launch: joint position / velocity and platform pose / twist as read at that update (before the body integrates), and the
  // effort applied in it. [N][NC] x 3, [N][7] (x y z qx qy qz qw), [N][6]; null = not wanted. May be mapped host memory.
  double *pub_pos, *pub_vel, *pub_eff, *pub_pose, *pub_twist;
  // rollout mode (cdpr_rollout)
  const float *cmd_table;  // [n_seq][n_cmd][NC]
  int n_seq, n_cmd, steps_per_cmd;
  double *cost;            // [N] or nullptr
  double target[3], lambda;
};

// The value one loop iteration of a reference command driver publishes, from sine = sin(time * freq * 2 pi [+ phase]):
//   shape 0  sinevelocitytest.cpp:36-38    amp * sine
//   shape 1  squarevelocitytest.cpp:21-22  +-amp outside the dead band |sine| >= sqrt(0.5), else 0
// stored into a float32 Joy axis like the driver does.
__device__ __forceinline__ double publisher_value(int shape, double amp, double sine) {
#ifdef __CUDA_ARCH__
  const double v = (shape == 1) ? ((fabs(sine) >= 0.70710678118654757) ? copysign(amp, sine) : 0.0) : __dmul_rn(amp, sine);
#else
  const double v = 0.0;
#endif
  return (double)(float)v;
}

// gazebo::common::Time::Double(): two roundings, never contracted
__host__ __device__ inline double time_double(int sec, int nsec) {
#ifdef __CUDA_ARCH__
  return __dadd_rn((double)sec, __dmul_rn((double)nsec, 1e-9));
#else
  volatile double frac = (double)nsec * 1e-9;
  return (double)sec + frac;
#endif
}

}  // namespace cdpr
