/*
 * cdpr_oracle.h -- CPU ORACLE (level L1) for the CDPR step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (cdpr_simulation_b200/)
 * never links, imports or executes it and has no CPU fallback.
 *
 * What it is: a scalar, single-robot, plain-C restatement of
 *   (a) the reference plugin's force law -- Pid.cpp / JointForceCalculator.cpp /
 *       Filter.h / CdprGazeboPlugin::update() -- statement by statement, and
 *   (b) the reduced-coordinate rigid-body model that stands in for Gazebo/ODE
 *       (SURVEY.md App. C), which the reference does not contain as code.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - force law P/I/clamp/anti-windup/mode logic: PINNED against the
 *     reference's own Pid.cpp + JointForceCalculator.cpp compiled unmodified
 *     (oracle/_ref, tests/test_oracle_vs_reference.py), bit-exact.
 *   - D-term: the reference fits the window in ABSOLUTE time (pow + normal
 *     equations + Eigen QR, Pid.cpp:219-247) which is ill-conditioned; the
 *     oracle fits in WINDOW-RELATIVE time (same polynomial in exact
 *     arithmetic).  Pinned to the reference within its own conditioning noise.
 *   - geometry: PINNED against the numeric literals in sdf/cube.sdf.
 *   - rigid-body integration (ODE semi-implicit Euler): PARITY UNPINNED --
 *     Gazebo/ODE are absent and the reference has no test that fixes them.
 *
 * All reference citations are relative to /root/reference/src/cdpr_gazebo/.
 */
#ifndef CDPR_ORACLE_H
#define CDPR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_CABLES 8
#define ORC_MAX_DBUF 32
#define ORC_MAX_DEGREE 4
#define ORC_MAX_CASCADE 4

/* Pid::PidParameters, include/cdpr_gazebo/Pid.h:70-81 */
typedef struct {
  double forward_gain, p_gain, i_gain, d_gain;
  int32_t d_degree, d_buffer_length;
  double i_limit, cmd_limit;
  double p_cutoff, p_quality;
  int32_t p_cascade;
  double d_cutoff, d_quality;
  int32_t d_cascade;
} orc_pid_params;

typedef struct {
  int32_t n_cables;
  double frame_anchor[ORC_MAX_CABLES][3];    /* a_i, frame coordinates */
  double platform_anchor[ORC_MAX_CABLES][3]; /* b_i, platform body coordinates */
  double home_pos[3];                        /* pose at which joint coordinate q_i = 0 */
  double home_quat[4];                       /* w x y z */
  double mass;
  double inertia[6];                         /* ixx iyy izz ixy ixz iyz (body) */
  double gravity[3];
  double cable_damping;                      /* N s / m on the prismatic joint */
  double effort_limit;                       /* joint effort truncation, N */
  double dt;                                 /* physics step, s (integer ns) */
  orc_pid_params vel_pid, pos_pid;
  double velocity_epsilon;
  /* ---- leg fidelity: the five links of every UPS leg and their passive joints, P/sdf/cube.sdf:359-518.
   * PARITY UNPINNED like the rest of the rigid-body model (no Gazebo/ODE here); see orc_legs_* in cdpr_oracle.c. */
  int32_t leg_model;                       /* 0: massless legs (reduced model); 1: leg links + passive damping */
  double leg_link_mass, leg_link_inertia;  /* 1e-3 kg, 1e-3 kg m^2 isotropic, each link (cube.sdf:359-369,372-382,401-411,447-457,476-486) */
  double leg_cable_com;                    /* platform anchor -> COM of the cable link, along the leg: l/2 = 0.51961524 (cube.sdf:344) */
  double passive_damping;                  /* 0.01 N m s on the five revolute joints (cube.sdf:396,425,471,500,515) */
  double leg_axis_frame[ORC_MAX_CABLES][3];    /* rev_X axis, fixed in the frame (cube.sdf:390) */
  double leg_axis_cable[ORC_MAX_CABLES][3];    /* rev_Zpf axis at home, fixed in the cable link (cube.sdf:506-512: "0 0 1", model frame) */
  double leg_axis_platform[ORC_MAX_CABLES][3]; /* rev_Xpf axis, platform body frame (cube.sdf:462-468: "1 0 0") */
  double slider_lower, slider_upper;       /* cube.sdf:436-437 (unreachable inside the frame; carried as constants) */
  double slider_velocity_limit;            /* cube.sdf:439 (ODE does not enforce joint velocity limits; carried as a constant) */
  int32_t derive_absolute_time; /* 0: window-relative fit (default); 1: reference-style absolute time */
} orc_config;

/* gazebo::math::BiQuad<double>, include/cdpr_gazebo/Filter.h:102-172 */
typedef struct {
  double a0, a1, a2, b1, b2;
  double x1, x2, y1, y2;
} orc_biquad;

/* gazebo::common::Pid, include/cdpr_gazebo/Pid.h:112-164 */
typedef struct {
  orc_pid_params prm;
  double i_max, i_min, cmd_max, cmd_min;
  int32_t was_last_time;
  double last_time;
  double p_err, i_err, d_err, cmd;
  orc_biquad p_filter[ORC_MAX_CASCADE], d_filter[ORC_MAX_CASCADE];
  int32_t d_missing;
  double d_x[ORC_MAX_DBUF], d_y[ORC_MAX_DBUF];
  int32_t derive_absolute_time;
  /* last-update telemetry (the reference publishes these on topic "pid") */
  double dbg_p, dbg_i, dbg_d;
} orc_pid;

enum { ORC_MODE_FORCE = 0, ORC_MODE_POSITION = 1, ORC_MODE_VELOCITY = 2 };

/* gazebo::physics::JointForceCalculator, include/cdpr_gazebo/JointForceCalculator.h:32-96 */
typedef struct {
  orc_pid pos_pid, vel_pid;
  int32_t mode;
  double velocity_epsilon;
  double last_position, force, position_target, velocity_target;
  int32_t last_sec, last_nsec; /* mLastUpdateTime */
} orc_cable;

typedef struct {
  orc_config cfg;
  double home_len[ORC_MAX_CABLES]; /* L0_i */
  double leg_alpha[ORC_MAX_CABLES][3]; /* rev_Zpf axis in the leg's body triad (e1, e2, u), fixed at the home pose */
  /* platform state in frame coordinates */
  double p[3], q[4] /* w x y z */, v[3], w[3];
  orc_cable cable[ORC_MAX_CABLES];
  int32_t sec, nsec; /* gazebo::common::Time simTime */
  int64_t step_count;
  /* pending commands (CdprGazeboPlugin::m*Command + m*CommandReceived) */
  int32_t vel_cmd_received, pos_cmd_received;
  float vel_cmd[ORC_MAX_CABLES], pos_cmd[ORC_MAX_CABLES];
  /* outputs of the last step */
  double joint_pos[ORC_MAX_CABLES], joint_vel[ORC_MAX_CABLES];
  double pid_force[ORC_MAX_CABLES]; /* returned by JointForceCalculator::update */
  double effort[ORC_MAX_CABLES];    /* after effort truncation (Joint::GetForce) */
  /* command publisher (sinevelocitytest.cpp, squarevelocitytest.cpp) */
  int32_t sine_enabled;
  double sine_amp, sine_freq, sine_phase, sine_time;
  int32_t sine_period_steps; /* physics steps per published command */
  double sine_pub_dt;        /* 1/cPublishFrequency */
  int32_t pub_shape;         /* 0: amp * sin (sinevelocitytest.cpp:36); 1: +-amp outside the dead band, else 0 (squarevelocitytest.cpp:21-22) */
} orc_robot;

typedef struct {
  double len[ORC_MAX_CABLES];     /* L_i */
  double len_rate[ORC_MAX_CABLES];/* dL_i/dt */
  double unit[ORC_MAX_CABLES][3]; /* u_i, platform -> frame */
  double arm[ORC_MAX_CABLES][3];  /* r_i x u_i */
  double joint_pos[ORC_MAX_CABLES], joint_vel[ORC_MAX_CABLES];
} orc_kinematics;

/* external force law hook: lets the reference's own JointForceCalculator (L0)
 * drive the reduced model; returns the PID force for one cable. */
typedef double (*orc_force_fn)(void *ctx, int cable, double sim_time, double joint_pos, double joint_vel);

void orc_config_default(orc_config *cfg, int n_cables);
double orc_time_double(int32_t sec, int32_t nsec);

void orc_pid_init(orc_pid *pid, const orc_pid_params *prm);
void orc_pid_reset(orc_pid *pid);
double orc_pid_update(orc_pid *pid, double desired, double actual, double now);
double orc_pid_derive(orc_pid *pid, double value, double now);
/* reference-faithful absolute-time variant of derive (pow + normal equations) */
double orc_pid_derive_abs(orc_pid *pid, double value, double now);

void orc_cable_init(orc_cable *c, const orc_config *cfg, int32_t sec, int32_t nsec);
void orc_cable_set_position_target(orc_cable *c, double target);
void orc_cable_set_velocity_target(orc_cable *c, double target);
void orc_cable_set_force(orc_cable *c, double force);
double orc_cable_update(orc_cable *c, int32_t sec, int32_t nsec, double joint_pos, double joint_vel);

void orc_robot_init(orc_robot *r, const orc_config *cfg);
void orc_home_lengths(const orc_config *cfg, double *len);
void orc_kinematics_eval(const orc_config *cfg, const double *home_len, const double p[3], const double q[4],
                         const double v[3], const double w[3], orc_kinematics *out);
/* returns 0, or -1 (command dropped, state untouched) when n_axes != n_cables */
int orc_robot_velocity_cmd(orc_robot *r, const float *axes, int n_axes);
int orc_robot_position_cmd(orc_robot *r, const float *axes, int n_axes);
int orc_robot_effort_cmd(orc_robot *r, const double *force, int n_axes);
void orc_robot_sine(orc_robot *r, double amp, double freq, double phase);
void orc_robot_step(orc_robot *r);
void orc_robot_step_ext(orc_robot *r, orc_force_fn fn, void *ctx);
void orc_robot_platform_state(const orc_robot *r, double pose7[7], double twist6[6]);
/* leg model diagnostics (tests): 6x6 generalised mass matrix of platform + legs at the robot's pose, row-major; kinetic energy
 * 1/2 xi^T M xi; potential energy of platform + leg links; the five passive joint rates of leg i for the current twist */
void orc_legs_mass_matrix(const orc_robot *r, double M[36]);
double orc_legs_kinetic_energy(const orc_robot *r);
double orc_legs_potential_energy(const orc_robot *r);
void orc_legs_joint_rates(const orc_robot *r, int leg, double rates[5]);

/* batched helpers used by the tests and the CPU baseline (OpenMP over robots) */
void orc_batch_step(orc_robot *robots, int64_t n, int64_t k_steps, int n_threads);
void orc_batch_ik(const orc_config *cfg, int64_t n, const double *pose7, const double *twist6,
                  double *len, double *len_rate, double *wmat /* [n][nc][6] */, int n_threads);
int orc_sizeof_robot(void);
/* robots[i] = fresh robot (plugin state right after Load); every array argument may be NULL.
 * pose7 = [n][7] x y z qx qy qz qw, twist6 = [n][6]; amp/freq/phase enable the sine publisher. */
void orc_batch_init(orc_robot *robots, int64_t n, const orc_config *cfg, const double *pose7, const double *twist6,
                    const double *amp, const double *freq, const double *phase);
/* wave form and rate of the publisher of every robot: shape as in orc_robot.pub_shape, publish_hz = cPublishFrequency of the driver
 * (100 for sinevelocitytest.cpp:7, 10 for squarevelocitytest.cpp:6) */
void orc_batch_publisher(orc_robot *robots, int64_t n, int shape, double publish_hz);
void orc_batch_velocity_cmd(orc_robot *robots, int64_t n, const float *axes /* [n][nc] */);
void orc_batch_position_cmd(orc_robot *robots, int64_t n, const float *axes);
void orc_batch_effort_cmd(orc_robot *robots, int64_t n, const double *force);
void orc_batch_platform_state(const orc_robot *robots, int64_t n, double *pose7, double *twist6);
/* joint states as the plugin would publish them at its NEXT call: position/velocity from the
 * current platform state, effort = force applied in the last step */
void orc_batch_joint_states(const orc_robot *robots, int64_t n, double *pos, double *vel, double *effort);

void orc_batch_last_outputs(const orc_robot *robots, int64_t n, double *jpos, double *jvel, double *pid_force, double *effort);
void orc_batch_pid_terms(const orc_robot *robots, int64_t n, double *out);

void orc_batch_targets(const orc_robot *robots, int64_t n, double *vel_target, double *pos_target, double *mode);
orc_pid *orc_pid_new(const orc_pid_params *prm, int derive_absolute_time);
void orc_pid_free(orc_pid *p);
void orc_pid_get(const orc_pid *p, double *out);
orc_cable *orc_cable_new(const orc_config *cfg);
void orc_cable_free(orc_cable *c);
int orc_cable_mode(const orc_cable *c);
double orc_cable_last_position(const orc_cable *c);

#ifdef __cplusplus
}
#endif
#endif
