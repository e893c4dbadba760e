/*
 * cdpr_b200.h -- C ABI of the B200-native batched CDPR step (libcdpr_b200.so).
 *
 * One handle = N independent robots ("instances") resident on one B200.  The entry points
 * are what a binding of the reference plugin's hot path would call; each cites the
 * reference interface it replaces (paths relative to /root/reference/src/cdpr_gazebo/).
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Layout conventions
 *   host buffers ("reference layout", array-of-structs, instance-major):
 *     axes      float32 [N][NC]        like sensor_msgs/Joy.axes (CdprGazeboPlugin.cpp:67-83)
 *     pose7     float64 [N][7]         x y z qx qy qz qw        (CdprGazeboPlugin.cpp:262-269)
 *     twist6    float64 [N][6]         lin xyz, ang xyz         (CdprGazeboPlugin.cpp:270-277)
 *     joint     float64 [N][NC]        position / velocity / effort (CdprGazeboPlugin.cpp:248-256)
 *   device buffers (struct-of-arrays, instance index fastest): documented per call.
 *   Cable index c is the numeric suffix of the reference joint name "cable<c>"
 *   (CdprGazeboPlugin.cpp:150-157); instance index i is the caller's, never permuted.
 *
 * Device buffers handed to the library must be ready when the call is made (or the handle must share the producer's
 * stream, cdpr_set_stream): the handle's own stream does not synchronise with the legacy default stream.
 *
 * Every function returns CDPR_OK (0) or a negative error code and never throws.
 * A handle is not thread-safe; distinct handles are independent (no globals).
 * There is NO CPU fallback: cdpr_create fails with CDPR_ERR_NO_DEVICE without a CUDA device.
 */
#ifndef CDPR_B200_H
#define CDPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDPR_MAX_CABLES 8
#define CDPR_MAX_DBUF 32
#define CDPR_MAX_DEGREE 4
#define CDPR_MAX_CASCADE 4

enum {
  CDPR_OK = 0,
  CDPR_ERR_BAD_ARG = -1,
  CDPR_ERR_BAD_CABLE_COUNT = -2, /* gazebo::common::Exception("invalid joint count"), CdprGazeboPlugin.cpp:167-169 */
  CDPR_ERR_BAD_LENGTH = -3,      /* command with axes.size() != NC: dropped, state untouched (CdprGazeboPlugin.cpp:68,77) */
  CDPR_ERR_NO_DEVICE = -4,
  CDPR_ERR_CUDA = -5,
  CDPR_ERR_UNSUPPORTED = -6,
  CDPR_ERR_NOMEM = -7
};

/* JointForceCalculator::UpdateMode, include/cdpr_gazebo/JointForceCalculator.h:33-35 */
enum { CDPR_MODE_FORCE = 0, CDPR_MODE_POSITION = 1, CDPR_MODE_VELOCITY = 2 };

/* gazebo::common::Pid::PidParameters, include/cdpr_gazebo/Pid.h:70-81; the ROS parameters of
 * CdprGazeboPlugin.h:34-54 map 1:1 onto these fields. */
typedef struct cdpr_pid_params {
  double forward_gain, p_gain, i_gain, d_gain;
  int32_t d_degree, d_buffer_length;
  double i_limit, cmd_limit;
  double p_cutoff, p_quality;
  int32_t p_cascade;
  double d_cutoff, d_quality;
  int32_t d_cascade;
} cdpr_pid_params;

/* Everything the plugin reads at Load() (ROS parameters, CdprGazeboPlugin.cpp:98-139) plus the
 * robot constants it gets implicitly from sdf/cube.sdf through Gazebo. All fields explicit. */
typedef struct cdpr_config {
  int32_t n_cables;                            /* cWireCount, CdprGazeboPlugin.h:20 (4); 8 = synthetic extension */
  double frame_anchor[CDPR_MAX_CABLES][3];     /* a_i, frame coordinates (cube.sdf:383,559,735,911) */
  double platform_anchor[CDPR_MAX_CABLES][3];  /* b_i, platform body coordinates (cube.sdf:458,634,810,986) */
  double home_pos[3];                          /* pose where every joint coordinate is 0 (cube.sdf:310) */
  double home_quat[4];                         /* w x y z */
  double mass;                                 /* cube.sdf:340 */
  double inertia[6];                           /* ixx iyy izz ixy ixz iyz, body frame (cube.sdf:331-338) */
  double gravity[3];                           /* Gazebo world default (0,0,-9.8) */
  double cable_damping;                        /* cube.sdf:442 */
  double effort_limit;                         /* cube.sdf:438; <0 disables truncation */
  double dt;                                   /* physics step, s; must be a whole number of ns */
  cdpr_pid_params vel_pid, pos_pid;            /* launch/cdpr_gazebo.launch:19-39 */
  double velocity_epsilon;                     /* launch/cdpr_gazebo.launch:18 */
  /* ---- leg fidelity (the five links of every UPS leg and their passive joints, cube.sdf:344-518). leg_model = 0 (default)
   * keeps the legs massless: the reduced model. leg_model = 1 adds the configuration-dependent mass matrix of the leg links,
   * gravity on them and the viscous damping of the five passive revolute joints of each leg (csrc/legs.cuh); it runs in the
   * flex kernel. Integration semantics are parity-unpinned like the rest of the rigid-body model (no Gazebo/ODE to compare). */
  int32_t leg_model;
  double leg_link_mass, leg_link_inertia;      /* 1e-3 kg, 1e-3 kg m^2 isotropic, each link (cube.sdf:359-369,372-382,401-411,447-457,476-486) */
  double leg_cable_com;                        /* platform anchor -> COM of the cable link along the leg, l/2 = 0.51961524 (cube.sdf:344) */
  double passive_damping;                      /* 0.01 (cube.sdf:396,425,471,500,515) */
  double leg_axis_frame[CDPR_MAX_CABLES][3];   /* rev_X axis, fixed in the frame (cube.sdf:390) */
  double leg_axis_cable[CDPR_MAX_CABLES][3];   /* rev_Zpf axis at the home pose, fixed in the cable link (cube.sdf:506-512) */
  double leg_axis_platform[CDPR_MAX_CABLES][3];/* rev_Xpf axis, platform body frame (cube.sdf:462-468) */
  double slider_lower, slider_upper;           /* cube.sdf:436-437: +-0.51961524 m, unreachable with the platform inside the frame; constants only */
  double slider_velocity_limit;                /* cube.sdf:439: 10 m/s; ODE does not enforce joint velocity limits; constant only */
  double sine_publish_hz;                      /* sinevelocitytest.cpp:7 (100 Hz) */
} cdpr_config;

typedef struct cdpr_batch *cdpr_handle;

/* ---- lifecycle (CdprGazeboPlugin::Load, .cpp:49-65; destructor .h:88-90) ----------------- */
/* Reference constants. n_cables 4 = the reference robot; 8 = synthetic. */
int cdpr_config_default(cdpr_config *cfg, int n_cables);
/* State after Load(): platform at home, every cable in Position mode with target 0, both PIDs
 * un-primed (CdprGazeboPlugin.cpp:153-157), sim time 0. device = CUDA ordinal. */
int cdpr_create(const cdpr_config *cfg, int64_t n_instances, int device, cdpr_handle *out);
int cdpr_destroy(cdpr_handle h);
/* back to the post-Load state (Gazebo world reset + JointForceCalculator::reset, JointForceCalculator.h:69-73);
 * keeps the sine generator parameters and the snapshot buffer, restarts both at 0. */
int cdpr_reset(cdpr_handle h);
/* Last error text of this handle (h may be NULL: text of the last failed cdpr_create). */
const char *cdpr_last_error(cdpr_handle h);
/* Kernels run on this cudaStream_t (default: a stream owned by the handle). */
int cdpr_set_stream(cdpr_handle h, void *cuda_stream);
int cdpr_synchronize(cdpr_handle h);
/* on != 0: calls that take host buffers only ENQUEUE their copies/kernels and return; the caller keeps the (pinned)
 * buffers alive and untouched until cdpr_synchronize. In this mode everything that is not a step kernel (reset,
 * uploads, layout kernels, downloads) runs on a high-priority stream owned by the handle, ordered against the step
 * kernels by events, so the I/O of one handle is not starved behind step kernels of another handle that fill every SM.
 * Default off (every call completes on the handle's stream). */
int cdpr_set_async(cdpr_handle h, int on);

/* ---- options --------------------------------------------------------------------------------------- */
enum {
  /* value != 0: the step kernel evaluates the D-term as the plain FIR over the window instead of the sliding-moment
   * recursion (same law, different rounding: used to separate recursion drift from the dynamics' own error growth) */
  CDPR_OPT_DTERM_FIR = 1,
  /* value == 0: no CUDA event records around the launches (cdpr_last_kernel_ms then returns -1); needed when the calls
   * are captured into a CUDA graph, and saves two driver calls per cdpr_step in plugin-style stepping. Default 1. */
  CDPR_OPT_KERNEL_TIMING = 2,
  /* value != 0: INDEPENDENT ROBOTS. Every instance gets its own UpdateMode and its own command latch, like N separate
   * plugin instances (CdprGazeboPlugin.cpp:67-83,206-219): the *_masked commands address subsets, robots may sit in
   * different modes and receive commands at different steps. Runs the on-chip full-semantics kernel ("flex"); allowed
   * before the first step or right after cdpr_reset (the state is re-initialised to the post-Load state). Handles whose
   * configuration needs hold / biquad stages are independent from the start. */
  CDPR_OPT_INDEPENDENT = 3,
  /* wave form of the in-kernel command publisher that cdpr_set_sine_cmd switches on (per-instance amp, freq, phase; rate
   * cdpr_config.sine_publish_hz): 0 = sinevelocitytest (P/src/sinevelocitytest.cpp:33-49, amp * sin), 1 = squarevelocitytest
   * (P/src/squarevelocitytest.cpp:19-33: +-amp outside the dead band |sin| >= sqrt(0.5), else 0 -- with velocityEpsilon >= 0
   * the cables HOLD in the dead band, JointForceCalculator.cpp:72-82).  Float32 axes and accumulated publisher time as in
   * the drivers.  May change between steps; takes effect at the next publish. */
  CDPR_OPT_PUBLISHER_SHAPE = 4
};
int cdpr_set_option(cdpr_handle h, int option, int64_t value);

/* ---- commands (topics jointVelocities / jointPositions, CdprGazeboPlugin.cpp:67-83,206-219;
 *      JointForceCalculator::setForce, JointForceCalculator.h:92-95) ------------------------
 * n_axes != NC  =>  CDPR_ERR_BAD_LENGTH and nothing changes (the plugin drops the message).
 * A command becomes visible to the NEXT step; velocity is applied before position. */
int cdpr_set_velocity_cmd(cdpr_handle h, const float *axes, int64_t n_instances, int n_axes);
int cdpr_set_position_cmd(cdpr_handle h, const float *axes, int64_t n_instances, int n_axes);
int cdpr_set_effort_cmd(cdpr_handle h, const double *force, int64_t n_instances, int n_axes);
/* The same messages addressed to a SUBSET of the robots: mask[N] (uint8), instance i receives its row of axes iff
 * mask[i] != 0; the others keep their targets, modes and pending commands. mask == NULL addresses all. Needs independent
 * robots (CDPR_OPT_INDEPENDENT), else CDPR_ERR_UNSUPPORTED and nothing changes. */
int cdpr_set_velocity_cmd_masked(cdpr_handle h, const float *axes, const unsigned char *mask, int64_t n_instances, int n_axes);
int cdpr_set_position_cmd_masked(cdpr_handle h, const float *axes, const unsigned char *mask, int64_t n_instances, int n_axes);
int cdpr_set_effort_cmd_masked(cdpr_handle h, const double *force, const unsigned char *mask, int64_t n_instances, int n_axes);
/* JointForceCalculator::mUpdateMode of every instance after the last step (commands still pending are not applied yet):
 * modes[N], CDPR_MODE_* */
int cdpr_get_modes(cdpr_handle h, int32_t *modes);
/* sinevelocitytest.cpp:33-49 run inside the step kernel, per instance: every
 * (1/sine_publish_hz)/dt steps a new command (float)(amp*sin(time*freq*2*M_PI + phase)) goes to
 * all cables; `time` accumulates in double from 0. amp == NULL disables the generator. */
int cdpr_set_sine_cmd(cdpr_handle h, const double *amp, const double *freq, const double *phase, int64_t n_instances);

/* ---- stepping (CdprGazeboPlugin::update .cpp:202-246 followed by the physics step) ------- */
int cdpr_step(cdpr_handle h, int64_t k_steps);
/* ONE plugin update per call -- the literal drop-in for CdprGazeboPlugin::update (.cpp:202-246) at the plugin's own
 * operating point (one call per physics step): latch the messages given (vel_axes / pos_axes float32 [N][NC], NULL = no
 * message this step; velocity is fanned out before position), run one physics step, and publish like the plugin does from
 * inside update(): joint position / velocity [N][NC] and platform pose7 / twist6 as READ AT THIS UPDATE (before the body
 * integrates), effort [N][NC] = the force applied in this step (.cpp:248-280). Any output may be NULL. Synchronous. The
 * step kernel publishes straight into pinned host memory: one kernel launch and one synchronisation per update. */
int cdpr_update(cdpr_handle h, const float *vel_axes, const float *pos_axes, double *position, double *velocity, double *effort,
                double *pose7, double *twist6);
int64_t cdpr_step_count(cdpr_handle h);
double cdpr_sim_time(cdpr_handle h);

/* ---- outputs (publishJointStates .cpp:248-256, publishPlatformState .cpp:258-280) -------- */
/* position/velocity follow from the CURRENT platform state; effort = force applied in the last step. */
int cdpr_get_joint_states(cdpr_handle h, double *position, double *velocity, double *effort);
int cdpr_get_platform_state(cdpr_handle h, double *pose7, double *twist6);
/* overwrite platform pose/twist (Gazebo's SetWorldPose/SetWorldTwist equivalent); either may be NULL */
int cdpr_set_platform_state(cdpr_handle h, const double *pose7, const double *twist6);
/* telemetry of topic "pid" (Pid.cpp:140-167), all cables: [N][NC][6] = pid_force, p_err, i_err, d_err, cmd, mode */
int cdpr_get_pid_state(cdpr_handle h, double *out);
/* topic "pid" as the plugin publishes it every update (CdprGazeboPlugin.cpp:226,233-235), for EVERY cable (the reference
 * publishes cable 0): [N][NC][5] = pTerm, iTerm BEFORE its clamp, dTerm, desired (Pid.cpp:140-141,159,167), applied force
 * (Joint::GetForce after truncation) of the last step. A priming update (first after a Pid reset) leaves the first
 * four untouched, like the reference's message buffer. Step with k = 1 for the per-step stream. */
int cdpr_get_pid_terms(cdpr_handle h, double *out);

/* ---- checkpoint / resume ---------------------------------------------------------------- */
size_t cdpr_state_bytes(cdpr_handle h);
int cdpr_get_state(cdpr_handle h, void *blob, size_t bytes);
int cdpr_set_state(cdpr_handle h, const void *blob, size_t bytes);

/* ---- decimated trajectory snapshots (the publishPeriod gate, .cpp:236-242) --------------- */
/* every `every` steps the step kernel writes the platform state into dev_buf, a device buffer
 * laid out [capacity][13][n_instances] = px py pz qw qx qy qz vx vy vz wx wy wz; the write index
 * restarts at 0 on every call of this function. every == 0 disables. */
int cdpr_set_snapshots(cdpr_handle h, int64_t every, void *dev_buf, int64_t capacity);
/* Fused all-gather: the same snapshots, written by the step kernel into the gather buffer of EVERY rank. peer_bufs[p]
 * (p < n_peers <= 8) is rank p's buffer [capacity][13][total_instances] as mapped into this process (NVLink peer /
 * symmetric memory); this handle's instances occupy columns [instance_offset, instance_offset + N). After the step
 * and a cross-rank barrier every buffer holds the whole trajectory in global instance order -- no collective call. */
int cdpr_set_snapshot_peers(cdpr_handle h, int64_t every, void *const *peer_bufs, int n_peers, int64_t instance_offset,
                            int64_t total_instances, int64_t capacity);
/* Same, through ONE NVLS multicast address covering all ranks' buffers: each snapshot value is a single
 * multimem.st that the NVSwitch replicates to every rank (fast kernel variant only). */
int cdpr_set_snapshot_multicast(cdpr_handle h, int64_t every, void *multicast_buf, int64_t instance_offset,
                                int64_t total_instances, int64_t capacity);
int64_t cdpr_snapshot_count(cdpr_handle h);

/* ---- kinematics only (Joint::Position/GetVelocity read-backs + the wrench Jacobian) ------- */
/* host, reference layout: length/length_rate [N][NC], wmat [N][NC][6] = (u, r x u) per cable */
int cdpr_ik(cdpr_handle h, int64_t n, const double *pose7, const double *twist6, double *length, double *length_rate, double *wmat);
/* device, SoA: state13 [13][n] (order as snapshots); out [NC][8][n] = L, dL/dt, u xyz, (r x u) xyz */
int cdpr_ik_device(cdpr_handle h, int64_t n, const void *dev_state13, void *dev_out);

/* ---- sampled rollouts (north_star config 5) ---------------------------------------------- */
/* Instances are (robot r, sequence s), i = r*n_seq + s, N = n_robots*n_seq. All sequences of a
 * robot start from that robot's entry in pose7/twist6 ([n_robots][..]; NULL = home, at rest),
 * the controller in its post-Load state. cmds float32 [n_seq][n_cmd][NC] are velocity commands,
 * one per `steps_per_cmd` steps. cost[i] = sum over steps of |p - target|^2 + lambda*|w|^2.
 * dev_cost_seq (device, float64 [n_seq]) receives sum over this handle's robots of cost, in
 * robot order (deterministic) -- the vector the caller all-reduces across GPUs. */
int cdpr_rollout(cdpr_handle h, int64_t n_robots, int64_t n_seq, const double *pose7, const double *twist6,
                 const float *cmds, int64_t n_cmd, int64_t steps_per_cmd, const double target_pos[3], double lambda,
                 void *dev_cost_seq, double *host_cost /* [N] or NULL */);

/* ---- multi-GPU for a C++ host: ONE process, G devices of one box ------------------------- */
/* Instances shard by contiguous range, one handle per device (created by the caller with cdpr_create(..., device_r, ...)),
 * no exchange inside a step. The two exchanges of the path run over NVLink peer memory, written by this library's kernels;
 * no NCCL, no torch. (The one-rank-per-GPU form with NVLS multicast is cdpr_simulation_b200/distributed.py.) */
typedef struct cdpr_comm *cdpr_comm_t;
/* enables peer access between all pairs of `devices` (NULL = 0 .. n_devices-1); CDPR_ERR_UNSUPPORTED without P2P */
int cdpr_comm_create(int n_devices, const int *devices, cdpr_comm_t *out);
int cdpr_comm_destroy(cdpr_comm_t c);
int cdpr_comm_size(cdpr_comm_t c);
const char *cdpr_comm_last_error(cdpr_comm_t c);
/* Config 4, fused trajectory gather: allocates on every device the FULL buffer [capacity][13][sum instances] and points
 * handle r (instances[r] robots, on device r of the comm) at all of them: its step kernel stores every snapshot of its
 * shard into its column range of EVERY device's buffer (cdpr_set_snapshot_peers). */
int cdpr_comm_attach_gather(cdpr_comm_t c, cdpr_handle *handles, const int64_t *instances, int64_t every, int64_t capacity);
void *cdpr_comm_gather_buffer(cdpr_comm_t c, int rank); /* device pointer on device `rank` */
/* one pass of k steps on every device (launches enqueued back to back, then all devices synchronised: after it every
 * gather buffer holds the snapshots of ALL shards in global instance order) */
int cdpr_comm_step(cdpr_comm_t c, cdpr_handle *handles, int64_t k_steps);
/* Config 5, cost all-reduce: dev_vectors[r] = float64 [n_elements] on device r (e.g. the dev_cost_seq of cdpr_rollout);
 * afterwards every vector holds the sum over the ranks in ascending rank order -- the same bits on every device. */
int cdpr_comm_allreduce(cdpr_comm_t c, void *const *dev_vectors, int64_t n_elements);

/* ---- D-term constants (host only, no device needed) ---------------------------------------- */
/* With uniform time stamps Pid::derive (Pid.cpp:193-247) is a fixed FIR: derivative = sum_j fir[j] * y[j], j = 0
 * oldest. fir: [d_buffer_length]. quadratic[3] (may be NULL): fir[j] = q0 + q1 p + q2 p^2 with p = j + 1, valid when
 * *is_quadratic = 1 (degree <= 2): the coefficients of the sliding-moment form used by the step kernel. */
int cdpr_dterm_weights(const cdpr_pid_params *pid, double dt, double *fir, double *quadratic, int *is_quadratic);

/* ---- raw device access for zero-copy callers (torch.distributed gathers) ----------------- */
int64_t cdpr_padded_instances(cdpr_handle h);
void *cdpr_device_platform_state(cdpr_handle h); /* [13][padded] float64, same order as snapshots */

/* ---- measurement helpers ----------------------------------------------------------------- */
/* sustained DFMA rate of this GPU (FP64 roofline denominator); returns TFLOP/s, <0 on error */
double cdpr_measure_fp64_tflops(int device, int iters);
/* elapsed ms of the kernels launched by the last cdpr_step / cdpr_ik_device / cdpr_rollout,
 * from CUDA events recorded on the handle's stream around the launch */
float cdpr_last_kernel_ms(cdpr_handle h);
int64_t cdpr_launch_count(cdpr_handle h);
const char *cdpr_kernel_variant(cdpr_handle h); /* "fast", "flex" (full semantics, on chip) or "general" (catch-all, state in HBM) */
/* which instance of the variant runs, for tests and tuning logs: "fast", "general", "flex:lanes=L,nf=F,unroll=U"
 * (k_step_flex) or "flexr:lanes=L,nf=F,hold=H" (k_step_flexr).  The string lives in the handle. */
const char *cdpr_kernel_detail(cdpr_handle h);

#ifdef __cplusplus
}
#endif
#endif
