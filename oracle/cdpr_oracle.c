/*
 * cdpr_oracle.c -- CPU ORACLE (level L1).  TEST INFRASTRUCTURE ONLY (see header).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps every a*b+c as two roundings, like the reference's
 * catkin build on plain x86-64 (no -march => no FMA).
 *
 * Citations: P/ = /root/reference/src/cdpr_gazebo/.
 */
#include "cdpr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* constants: P/sdf/cube.sdf, P/launch/cdpr_gazebo.launch   */
/* ------------------------------------------------------------------------- */
void orc_config_default(orc_config *cfg, int n_cables) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->n_cables = n_cables;
  /* frame / platform anchors: P/sdf/cube.sdf:383,559,735,911 and :458,634,810,986 (minus :310) */
  static const double sx[4] = {-1.0, -1.0, 1.0, 1.0};
  static const double sy[4] = {-1.0, 1.0, 1.0, -1.0};
  for (int i = 0; i < n_cables && i < ORC_MAX_CABLES; ++i) {
    int k = i & 3;
    cfg->frame_anchor[i][0] = 0.3 * sx[k];
    cfg->frame_anchor[i][1] = 0.3 * sy[k];
    /* cables 4..7: synthetic 8-cable extension, lower frame face */
    cfg->frame_anchor[i][2] = (i < 4) ? 0.6 : 0.0;
    cfg->platform_anchor[i][0] = 0.03 * sx[k];
    cfg->platform_anchor[i][1] = 0.03 * sy[k];
    cfg->platform_anchor[i][2] = 0.0;
  }
  cfg->home_pos[0] = 0.0; cfg->home_pos[1] = 0.0; cfg->home_pos[2] = 0.3; /* cube.sdf:310 */
  cfg->home_quat[0] = 1.0;
  cfg->mass = 1.0;                                   /* cube.sdf:340 */
  cfg->inertia[0] = cfg->inertia[1] = cfg->inertia[2] = 1.0; /* cube.sdf:331-338 */
  cfg->gravity[2] = -9.8;                            /* Gazebo default world */
  cfg->cable_damping = 1.0;                          /* cube.sdf:442 */
  cfg->effort_limit = 100.0;                         /* cube.sdf:438 */
  cfg->dt = 0.001;                                   /* Gazebo default max_step_size */
  /* P/launch/cdpr_gazebo.launch:17-39 */
  orc_pid_params *v = &cfg->vel_pid, *p = &cfg->pos_pid;
  v->forward_gain = 0.0; v->p_gain = 200.0; v->i_gain = 20.0; v->d_gain = 1.0;
  v->d_degree = 2; v->d_buffer_length = 11; v->i_limit = 100.0; v->cmd_limit = 100.0;
  v->p_cutoff = 0.1; v->p_quality = 0.707; v->p_cascade = 0;
  v->d_cutoff = 0.1; v->d_quality = 0.707; v->d_cascade = 0;
  *p = *v;
  p->forward_gain = 0.0;                             /* CdprGazeboPlugin.cpp:123 */
  p->p_gain = 200.0; p->i_gain = 70.0; p->d_gain = 80.0;
  p->p_cascade = p->d_cascade = 0;                   /* CdprGazeboPlugin.cpp:133 */
  cfg->velocity_epsilon = -0.001;
  /* leg links and passive joints: P/sdf/cube.sdf:359-518; off by default (reduced model) */
  cfg->leg_model = 0;
  cfg->leg_link_mass = 0.001; cfg->leg_link_inertia = 0.001;
  cfg->leg_cable_com = 0.51961524;   /* l/2, l = |(0.6, 0.6, 0.6)| (gen_cdpr.py:104,124-125; cube.sdf:344) */
  cfg->passive_damping = 0.01;
  cfg->slider_lower = -0.51961524; cfg->slider_upper = 0.51961524; cfg->slider_velocity_limit = 10.0;
  for (int i = 0; i < n_cables && i < ORC_MAX_CABLES; ++i) {
    /* gen_cdpr.py:113-125,152: the leg frame is the rotation that takes z onto the frame->platform direction about
     * z x u_fp; rev_X turns about its first column (cube.sdf:390). */
    double ufp[3], n = 0.0;
    for (int k = 0; k < 3; ++k) { ufp[k] = cfg->home_pos[k] + cfg->platform_anchor[i][k] - cfg->frame_anchor[i][k]; n += ufp[k] * ufp[k]; }
    n = sqrt(n);
    for (int k = 0; k < 3; ++k) ufp[k] /= n;
    double ax[3] = {-ufp[1], ufp[0], 0.0}; /* z x u_fp */
    double sn = sqrt(ax[0] * ax[0] + ax[1] * ax[1]), cs = ufp[2];
    if (sn > 0.0) { ax[0] /= sn; ax[1] /= sn; }
    /* Rodrigues, first column: e_x cos + (ax x e_x) sin + ax (ax . e_x)(1 - cos) */
    cfg->leg_axis_frame[i][0] = cs + ax[0] * ax[0] * (1.0 - cs);
    cfg->leg_axis_frame[i][1] = ax[2] * sn + ax[1] * ax[0] * (1.0 - cs);
    cfg->leg_axis_frame[i][2] = -ax[1] * sn + ax[2] * ax[0] * (1.0 - cs);
    cfg->leg_axis_cable[i][2] = 1.0;    /* "0 0 1" in the model frame (SDF 1.4) */
    cfg->leg_axis_platform[i][0] = 1.0; /* "1 0 0" */
  }
}

/* gazebo::common::Time::Double(): sec + nsec * 1e-9 */
double orc_time_double(int32_t sec, int32_t nsec) { return (double)sec + (double)nsec * 1e-9; }

/* ------------------------------------------------------------------------- */
/* BiQuad: P/include/cdpr_gazebo/Filter.h:130-165                              */
/* ------------------------------------------------------------------------- */
static void biquad_set_fc(orc_biquad *f, double fc, double fs, double q) {
  double k = tan(M_PI * fc / fs);
  double den = k * k + k / q + 1.0;
  f->a0 = k * k / den;
  f->a1 = 2 * f->a0;
  f->a2 = f->a0;
  f->b1 = 2 * (k * k - 1.0) / den;
  f->b2 = (k * k - k / q + 1.0) / den;
}
static void biquad_set_value(orc_biquad *f, double v) { f->x1 = f->x2 = f->y1 = f->y2 = v; }
static double biquad_process(orc_biquad *f, double x) {
  double y0 = f->a0 * x + f->a1 * f->x1 + f->a2 * f->x2 - f->b1 * f->y1 - f->b2 * f->y2;
  f->x2 = f->x1; f->x1 = x; f->y2 = f->y1; f->y1 = y0;
  return y0;
}
/* Pid::CascadeFilter::update, P/src/Pid.cpp:38-44 */
static double cascade_update(orc_biquad *f, int n, double x) {
  double out = x;
  for (int i = 0; i < n; ++i) out = biquad_process(&f[i], out);
  return out;
}

/* ------------------------------------------------------------------------- */
/* Pid: P/src/Pid.cpp                                                         */
/* ------------------------------------------------------------------------- */
void orc_pid_reset(orc_pid *pid) { /* Pid.cpp:100-115 (mLastTime is NOT cleared) */
  pid->was_last_time = 0;
  pid->p_err = pid->i_err = pid->d_err = pid->cmd = 0.0;
  for (int i = 0; i < pid->prm.p_cascade; ++i) biquad_set_value(&pid->p_filter[i], 0.0);
  for (int i = 0; i < pid->prm.d_cascade; ++i) biquad_set_value(&pid->d_filter[i], 0.0);
  for (int i = 0; i < ORC_MAX_DBUF; ++i) pid->d_x[i] = pid->d_y[i] = 0.0;
  pid->d_missing = pid->prm.d_buffer_length;
}

void orc_pid_init(orc_pid *pid, const orc_pid_params *prm) { /* Pid.cpp:63-77 */
  memset(pid, 0, sizeof(*pid));
  pid->prm = *prm;
  pid->i_max = fabs(prm->i_limit);  /* abs() resolves to the double overload in the real build */
  pid->i_min = -fabs(prm->i_limit);
  pid->cmd_max = fabs(prm->cmd_limit);
  pid->cmd_min = -fabs(prm->cmd_limit);
  for (int i = 0; i < prm->p_cascade; ++i) { /* Pid.cpp:27-36 */
    biquad_set_value(&pid->p_filter[i], 0.0);
    biquad_set_fc(&pid->p_filter[i], prm->p_cutoff, 1.0, prm->p_quality);
  }
  for (int i = 0; i < prm->d_cascade; ++i) {
    biquad_set_value(&pid->d_filter[i], 0.0);
    biquad_set_fc(&pid->d_filter[i], prm->d_cutoff, 1.0, prm->d_quality);
  }
  pid->last_time = 0.0;
  pid->derive_absolute_time = 0;
  orc_pid_reset(pid);
}

static void dbuf_push(orc_pid *pid, double value, double now) { /* Pid.cpp:194-200 */
  int n = pid->prm.d_buffer_length;
  for (int i = 1; i < n; ++i) {
    pid->d_x[i - 1] = pid->d_x[i];
    pid->d_y[i - 1] = pid->d_y[i];
  }
  pid->d_x[n - 1] = now;
  pid->d_y[n - 1] = value;
  pid->d_missing -= (pid->d_missing > 0 ? 1 : 0);
}

/*
 * Least-squares polynomial derivative at `now`, window-relative time.
 * Same polynomial as Pid.cpp:203-212 + :219-247 in exact arithmetic: shifting
 * and scaling the abscissa does not change the fitted polynomial.  Solved by
 * Householder QR of the Vandermonde matrix (cond ~1e2) instead of the
 * reference's normal equations in absolute time (cond 1e10..1e19).
 */
double orc_pid_derive(orc_pid *pid, double value, double now) {
  dbuf_push(pid, value, now);
  if (pid->d_missing != 0) return 0.0;
  int n = pid->prm.d_buffer_length, m = pid->prm.d_degree + 1;
  if (m < 2) return 0.0; /* degree 0: derivative of a constant */
  double span = now - pid->d_x[0];
  double a[ORC_MAX_DBUF][ORC_MAX_DEGREE + 1], b[ORC_MAX_DBUF];
  for (int j = 0; j < n; ++j) {
    double x = (pid->d_x[j] - now) / span, pw = 1.0;
    for (int k = 0; k < m; ++k) { a[j][k] = pw; pw *= x; }
    b[j] = pid->d_y[j];
  }
  for (int k = 0; k < m; ++k) { /* Householder, no pivoting (columns well scaled) */
    double nrm = 0.0;
    for (int j = k; j < n; ++j) nrm += a[j][k] * a[j][k];
    nrm = sqrt(nrm);
    if (nrm == 0.0) return 0.0;
    double alpha = a[k][k] > 0 ? -nrm : nrm;
    double vv[ORC_MAX_DBUF];
    for (int j = k; j < n; ++j) vv[j] = a[j][k];
    vv[k] -= alpha;
    double vnorm2 = 0.0;
    for (int j = k; j < n; ++j) vnorm2 += vv[j] * vv[j];
    if (vnorm2 == 0.0) continue;
    for (int c = k; c < m; ++c) {
      double s = 0.0;
      for (int j = k; j < n; ++j) s += vv[j] * a[j][c];
      s = 2.0 * s / vnorm2;
      for (int j = k; j < n; ++j) a[j][c] -= s * vv[j];
    }
    double s = 0.0;
    for (int j = k; j < n; ++j) s += vv[j] * b[j];
    s = 2.0 * s / vnorm2;
    for (int j = k; j < n; ++j) b[j] -= s * vv[j];
  }
  double c[ORC_MAX_DEGREE + 1];
  for (int k = m - 1; k >= 0; --k) {
    double s = b[k];
    for (int cc = k + 1; cc < m; ++cc) s -= a[k][cc] * c[cc];
    c[k] = s / a[k][k];
  }
  return c[1] / span;
}

/*
 * Reference-faithful variant: absolute time, pow(), normal equations
 * (Pid.cpp:219-244) then a dense solve.  Eigen is absent here, so the solve is
 * Gaussian elimination with complete pivoting in long double -- i.e. this
 * returns (nearly) the exact solution of the reference's ill-conditioned
 * system and is used only to quantify the reference's D-term noise.
 */
double orc_pid_derive_abs(orc_pid *pid, double value, double now) {
  dbuf_push(pid, value, now);
  if (pid->d_missing != 0) return 0.0;
  int n = pid->prm.d_buffer_length, deg = pid->prm.d_degree, m = deg + 1;
  double fx[2 * ORC_MAX_DEGREE + 1];
  for (int i = 0; i < 2 * deg + 1; ++i) {
    fx[i] = 0.0;
    for (int j = 0; j < n; ++j) fx[i] += pow(pid->d_x[j], i);
  }
  long double A[ORC_MAX_DEGREE + 1][ORC_MAX_DEGREE + 2];
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < m; ++j) A[i][j] = fx[i + j];
    double tmp = 0.0;
    for (int j = 0; j < n; ++j) tmp += pow(pid->d_x[j], i) * pid->d_y[j];
    A[i][m] = tmp;
  }
  int perm[ORC_MAX_DEGREE + 1];
  for (int i = 0; i < m; ++i) perm[i] = i;
  for (int k = 0; k < m; ++k) {
    int pr = k, pc = k; long double best = 0;
    for (int i = k; i < m; ++i) for (int j = k; j < m; ++j)
      if (fabsl(A[i][j]) > best) { best = fabsl(A[i][j]); pr = i; pc = j; }
    if (best == 0) return 0.0;
    for (int j = 0; j <= m; ++j) { long double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; }
    for (int i = 0; i < m; ++i) { long double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; }
    { int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t; }
    for (int i = k + 1; i < m; ++i) {
      long double f = A[i][k] / A[k][k];
      for (int j = k; j <= m; ++j) A[i][j] -= f * A[k][j];
    }
  }
  long double z[ORC_MAX_DEGREE + 1];
  for (int k = m - 1; k >= 0; --k) {
    long double s = A[k][m];
    for (int j = k + 1; j < m; ++j) s -= A[k][j] * z[j];
    z[k] = s / A[k][k];
  }
  double coef[ORC_MAX_DEGREE + 2];
  for (int k = 0; k < m; ++k) coef[perm[k]] = (double)z[k];
  /* Pid.cpp:205-212 */
  for (int i = 1; i <= deg; ++i) coef[i - 1] = i * coef[i];
  coef[deg] = 0.0;
  double derived = 0.0;
  for (int i = deg; i > 0; --i) derived = now * (derived + coef[i]);
  derived += coef[0];
  return derived;
}

static double clampd(double v, double lo, double hi) { /* ignition::math::clamp */
  double t = v > lo ? v : lo;
  return t < hi ? t : hi;
}

double orc_pid_update(orc_pid *pid, double desired, double actual, double now) { /* Pid.cpp:122-191 */
  if (!pid->was_last_time) {
    pid->was_last_time = 1;
    pid->cmd = 0.0;
  } else {
    double f_term = pid->prm.forward_gain * desired;
    double error = desired - actual;
    double dt = now - pid->last_time;
    pid->last_time = now;

    pid->p_err = cascade_update(pid->p_filter, pid->prm.p_cascade, error);
    double p_term = pid->prm.p_gain * pid->p_err;

    double prev_ierr = pid->i_err;
    pid->i_err += dt * error;
    double i_term = pid->prm.i_gain * pid->i_err;
    pid->dbg_p = p_term;
    pid->dbg_i = i_term; /* pre-clamp, like pidMsg.axes[1] (Pid.cpp:141) */
    if (i_term > pid->i_max) {
      i_term = pid->i_max;
      pid->i_err = i_term / pid->prm.i_gain;
    } else if (i_term < pid->i_min) {
      i_term = pid->i_min;
      pid->i_err = i_term / pid->prm.i_gain;
    }
    if (dt > 0.0) {
      double derived = pid->derive_absolute_time ? orc_pid_derive_abs(pid, error, now) : orc_pid_derive(pid, error, now);
      pid->d_err = cascade_update(pid->d_filter, pid->prm.d_cascade, derived);
    }
    double d_term = pid->prm.d_gain * pid->d_err;
    pid->dbg_d = d_term;

    double cmd = f_term + p_term + i_term + d_term;
    if (pid->cmd_max > pid->cmd_min) pid->cmd = clampd(cmd, pid->cmd_min, pid->cmd_max);
    if (pid->cmd != cmd) {
      pid->i_err = prev_ierr;
      pid->cmd += dt * error * pid->prm.i_gain;
    }
  }
  pid->last_time = now;
  return pid->cmd;
}

/* ------------------------------------------------------------------------- */
/* JointForceCalculator: P/src/JointForceCalculator.cpp                       */
/* ------------------------------------------------------------------------- */
static void cable_reset(orc_cable *c) { /* JointForceCalculator.h:69-73 */
  c->force = c->position_target = c->velocity_target = 0.0;
  orc_pid_reset(&c->vel_pid);
  orc_pid_reset(&c->pos_pid);
}

void orc_cable_set_position_target(orc_cable *c, double target) { /* .cpp:99-107 */
  c->position_target = target;
  if (c->mode != ORC_MODE_POSITION) orc_pid_reset(&c->pos_pid);
  c->mode = ORC_MODE_POSITION;
}
void orc_cable_set_velocity_target(orc_cable *c, double target) { /* .cpp:111-119 */
  c->velocity_target = target;
  if (c->mode != ORC_MODE_VELOCITY) orc_pid_reset(&c->vel_pid);
  c->mode = ORC_MODE_VELOCITY;
}
void orc_cable_set_force(orc_cable *c, double force) { /* .h:92-95 */
  c->force = force;
  c->mode = ORC_MODE_FORCE;
}

/* state of one force calculator right after CdprGazeboPlugin::initJointsAndController
 * (CdprGazeboPlugin.cpp:153-157): ctor, setPositionTarget(joint->Position()), then
 * operator= which copies mode/targets and calls reset() => Position mode, target 0,
 * both PIDs un-primed, mLastPosition = 0 (not copied by operator=). */
void orc_cable_init(orc_cable *c, const orc_config *cfg, int32_t sec, int32_t nsec) {
  memset(c, 0, sizeof(*c));
  orc_pid_init(&c->pos_pid, &cfg->pos_pid);
  orc_pid_init(&c->vel_pid, &cfg->vel_pid);
  c->pos_pid.derive_absolute_time = c->vel_pid.derive_absolute_time = cfg->derive_absolute_time;
  c->velocity_epsilon = cfg->velocity_epsilon;
  c->mode = ORC_MODE_FORCE;
  orc_cable_set_position_target(c, 0.0);
  cable_reset(c);
  c->last_position = 0.0;
  c->last_sec = sec;
  c->last_nsec = nsec;
}

double orc_cable_update(orc_cable *c, int32_t sec, int32_t nsec, double joint_pos, double joint_vel) { /* .cpp:59-96 */
  int64_t step_ns = ((int64_t)sec - c->last_sec) * 1000000000LL + ((int64_t)nsec - c->last_nsec);
  c->last_sec = sec;
  c->last_nsec = nsec;
  double now = orc_time_double(sec, nsec);
  double force = 0.0;
  if (step_ns > 0) {
    if (c->mode == ORC_MODE_FORCE) {
      c->last_position = joint_pos;
      force = c->force;
    } else if (c->mode == ORC_MODE_VELOCITY) {
      if (fabs(c->velocity_target) > c->velocity_epsilon) {
        c->last_position = joint_pos;
        force = orc_pid_update(&c->vel_pid, c->velocity_target, joint_vel, now);
      } else {
        force = orc_pid_update(&c->pos_pid, c->last_position, joint_pos, now);
      }
    } else if (c->mode == ORC_MODE_POSITION) {
      c->last_position = joint_pos;
      force = orc_pid_update(&c->pos_pid, c->position_target, joint_pos, now);
    }
  }
  return force;
}

/* ------------------------------------------------------------------------- */
/* reduced model: SURVEY.md App. C                                            */
/* ------------------------------------------------------------------------- */
static void quat_to_rot(const double q[4], double R[3][3]) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = 1.0 - 2.0 * (y * y + z * z); R[0][1] = 2.0 * (x * y - w * z); R[0][2] = 2.0 * (x * z + w * y);
  R[1][0] = 2.0 * (x * y + w * z); R[1][1] = 1.0 - 2.0 * (x * x + z * z); R[1][2] = 2.0 * (y * z - w * x);
  R[2][0] = 2.0 * (x * z - w * y); R[2][1] = 2.0 * (y * z + w * x); R[2][2] = 1.0 - 2.0 * (x * x + y * y);
}
static void mat_vec(const double R[3][3], const double a[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = R[i][0] * a[0] + R[i][1] * a[1] + R[i][2] * a[2];
}
static void matT_vec(const double R[3][3], const double a[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = R[0][i] * a[0] + R[1][i] * a[1] + R[2][i] * a[2];
}
static void cross(const double a[3], const double b[3], double out[3]) {
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

void orc_kinematics_eval(const orc_config *cfg, const double *home_len, const double p[3], const double q[4],
                         const double v[3], const double w[3], orc_kinematics *out) { /* App. C.2/C.3 */
  double R[3][3];
  quat_to_rot(q, R);
  for (int i = 0; i < cfg->n_cables; ++i) {
    double r[3], d[3];
    mat_vec(R, cfg->platform_anchor[i], r);
    for (int k = 0; k < 3; ++k) d[k] = cfg->frame_anchor[i][k] - p[k] - r[k];
    double len = sqrt(dot3(d, d));
    for (int k = 0; k < 3; ++k) out->unit[i][k] = d[k] / len;
    cross(r, out->unit[i], out->arm[i]);
    out->len[i] = len;
    double qd = dot3(out->unit[i], v) + dot3(out->arm[i], w);
    out->joint_vel[i] = qd;
    out->len_rate[i] = -qd;
    out->joint_pos[i] = (home_len ? home_len[i] : 0.0) - len;
  }
}

/* ------------------------------------------------------------------------- */
/* Leg fidelity -- PARITY UNPINNED (no Gazebo/ODE here).   */
/* ------------------------------------------------------------------------- */
/*
 * Every UPS leg of P/sdf/cube.sdf:344-518 is a chain  frame -rev_X- virt_X -rev_Y- virt_Y -cable(prismatic)- cable
 * -rev_Zpf- virt_Ypf -rev_Ypf- virt_Xpf -rev_Xpf- platform  with five links of mass m_l and isotropic inertia I_l.  Its six
 * joint coordinates are functions of the platform pose, so the robot keeps 6 degrees of freedom; the legs add a
 * configuration-dependent term to the generalised mass matrix, gravity on the moving leg links and viscous torques on
 * the five passive revolute joints.
 *
 * Geometry (frame coordinates): A frame anchor, B = p + R b platform anchor, u = (A - B)/L.  virt_X and virt_Y have their
 * centre of mass at A (at rest), virt_Xpf and virt_Ypf at B, the cable link at B + l_c u (cube.sdf:344: a rod centred l/2 from
 * the platform anchor).  Body triad of the leg (fixed in virt_Y / cable): e2 = rev_Y axis = (u x x0)/c, e1 = e2 x u, u, with x0 the
 * frame-fixed rev_X axis, s = u.x0, c = sqrt(1 - s^2).  With v_B = v + w x r the anchor velocity:
 *     rev_Y rate  thy = -(e1 . v_B)/L,   rev_X rate  thx = (e2 . v_B)/(L c),   w_leg = thx x0 + thy e2,
 *     cable COM velocity  v_c = v_B - (l_c/L)(v_B - u (u . v_B)).
 * Gimbal at B: a3 = rev_Zpf axis (fixed in the cable link: constant components in the leg triad), a1 = R a1_body the rev_Xpf
 * axis (fixed in the platform), a2 = (a3 x a1)/|a3 x a1| the rev_Ypf axis.  Closing the loop, w_leg + phi a3 = w + psx a1 + psy a2:
 *     psy = D . a2,  psx = (D . a1 - t D . a3)/(1 - t^2),  phi = psx t - D . a3,   D = w_leg - w,  t = a3 . a1,
 *     w(virt_Ypf) = w_leg + phi a3,   w(virt_Xpf) = w + psx a1.
 * All of these are linear in the platform twist xi = (v, w).  Kinetic energy of the leg = 1/2 |y|^2 with
 *     y = [ sqrt(I_l) thx | sqrt(2 I_l) w_leg | sqrt(I_l) w(virt_Ypf) | sqrt(I_l) w(virt_Xpf) | sqrt(m_l) v_c | sqrt(2 m_l) v_B ]  (16 rows)
 * so M(x) = diag(m, m, m, R I_b R^T) + sum_legs Jy^T Jy.  Passive damping: Rayleigh function 1/2 c_p |z|^2, z = the five joint
 * rates, generalised force -Jz^T z (explicit, like the actuated joint's damping).  Gravity: m_l g . (v_c + 2 v_B) per leg.
 * Velocity-product (Coriolis / centrifugal) terms of the leg links: each link obeys m a = f, I alpha = tau (isotropic inertia: no
 * gyroscopic torque) with a = J xi' + J' xi, so the projected equation gains -Jy^T (J'y xi); J'y xi = d/dt y(x(t), xi held) is
 * a one-sided difference along the motion over LEG_BIAS_DT.
 * Step (same semi-implicit order as App. C.6):  M(x_n) (xi+ - xi)/h = Q_cables + Q_gravity + Q_passive - Jy^T J'y xi - gyro(platform).
 * NEGLECTED, stated: ODE's constraint softness; the slider's position stops (|q| <= 0.5196 m cannot be reached with
 * the platform inside the 0.6 m frame) and velocity limit (ODE does not enforce joint velocity limits).
 */
#define LEG_BIAS_DT 1e-6 /* s: step of the one-sided difference behind the velocity-product terms */
typedef struct {
  double r[3], u[3], e1[3], e2[3], a1[3], a2[3], a3[3];
  double L, c, t;
  const double *x0;
} leg_geom;

static void leg_triad(const double x0[3], const double u[3], double e1[3], double e2[3], double *c_out) {
  double s = dot3(u, x0), c = sqrt(1.0 - s * s);
  for (int k = 0; k < 3; ++k) e1[k] = (x0[k] - s * u[k]) / c;
  cross(u, e1, e2);
  *c_out = c;
}

static void leg_geometry(const orc_config *cfg, const double alpha[3], int i, const double p[3], const double R[3][3], leg_geom *g) {
  double d[3];
  mat_vec(R, cfg->platform_anchor[i], g->r);
  for (int k = 0; k < 3; ++k) d[k] = cfg->frame_anchor[i][k] - p[k] - g->r[k];
  g->L = sqrt(dot3(d, d));
  for (int k = 0; k < 3; ++k) g->u[k] = d[k] / g->L;
  g->x0 = cfg->leg_axis_frame[i];
  leg_triad(g->x0, g->u, g->e1, g->e2, &g->c);
  for (int k = 0; k < 3; ++k) g->a3[k] = alpha[0] * g->e1[k] + alpha[1] * g->e2[k] + alpha[2] * g->u[k];
  mat_vec(R, cfg->leg_axis_platform[i], g->a1);
  g->t = dot3(g->a3, g->a1);
  double n[3];
  cross(g->a3, g->a1, n);
  double nn = sqrt(1.0 - g->t * g->t);
  for (int k = 0; k < 3; ++k) g->a2[k] = n[k] / nn;
}

/* y[16], z[5] (see above) for platform twist (v, w) */
static void leg_rates(const orc_config *cfg, const leg_geom *g, const double v[3], const double w[3], double y[16], double z[5]) {
  double wr[3], vB[3];
  cross(w, g->r, wr);
  for (int k = 0; k < 3; ++k) vB[k] = v[k] + wr[k];
  double thy = -dot3(g->e1, vB) / g->L;
  double thx = dot3(g->e2, vB) / (g->L * g->c);
  double wleg[3], D[3];
  for (int k = 0; k < 3; ++k) { wleg[k] = thx * g->x0[k] + thy * g->e2[k]; D[k] = wleg[k] - w[k]; }
  double psy = dot3(D, g->a2);
  double Da3 = dot3(D, g->a3);
  double psx = (dot3(D, g->a1) - g->t * Da3) / (1.0 - g->t * g->t);
  double phi = psx * g->t - Da3;
  double uv = dot3(g->u, vB), lam = cfg->leg_cable_com / g->L;
  double sI = sqrt(cfg->leg_link_inertia), s2I = sqrt(2.0 * cfg->leg_link_inertia);
  double sm = sqrt(cfg->leg_link_mass), s2m = sqrt(2.0 * cfg->leg_link_mass), sc = sqrt(cfg->passive_damping);
  y[0] = sI * thx;
  for (int k = 0; k < 3; ++k) {
    y[1 + k] = s2I * wleg[k];
    y[4 + k] = sI * (wleg[k] + phi * g->a3[k]);
    y[7 + k] = sI * (w[k] + psx * g->a1[k]);
    y[10 + k] = sm * (vB[k] - lam * (vB[k] - g->u[k] * uv));
    y[13 + k] = s2m * vB[k];
  }
  z[0] = sc * thx; z[1] = sc * thy; z[2] = sc * phi; z[3] = sc * psy; z[4] = sc * psx;
}

static void legs_home_alpha(const orc_config *cfg, double alpha[ORC_MAX_CABLES][3]) {
  double R[3][3];
  quat_to_rot(cfg->home_quat, R);
  for (int i = 0; i < cfg->n_cables; ++i) {
    double r[3], d[3], u[3], e1[3], e2[3], c;
    mat_vec(R, cfg->platform_anchor[i], r);
    for (int k = 0; k < 3; ++k) d[k] = cfg->frame_anchor[i][k] - cfg->home_pos[k] - r[k];
    double L = sqrt(dot3(d, d));
    for (int k = 0; k < 3; ++k) u[k] = d[k] / L;
    leg_triad(cfg->leg_axis_frame[i], u, e1, e2, &c);
    alpha[i][0] = dot3(e1, cfg->leg_axis_cable[i]);
    alpha[i][1] = dot3(e2, cfg->leg_axis_cable[i]);
    alpha[i][2] = dot3(u, cfg->leg_axis_cable[i]);
  }
}

/* M (6x6, symmetric, row-major) of platform + legs, and the extra generalised forces of the legs (gravity on the links,
 * passive joint damping) added to Q[6] */
static void legs_assemble(const orc_robot *r, const double R[3][3], double M[6][6], double Q[6]) {
  const orc_config *cfg = &r->cfg;
  const double *I = cfg->inertia;
  double Ib[3][3] = {{I[0], I[3], I[4]}, {I[3], I[1], I[5]}, {I[4], I[5], I[2]}};
  memset(M, 0, sizeof(double) * 36);
  for (int k = 0; k < 3; ++k) M[k][k] = cfg->mass;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0.0;
      for (int j = 0; j < 3; ++j)
        for (int l = 0; l < 3; ++l) s += R[a][j] * Ib[j][l] * R[b][l];
      M[3 + a][3 + b] = s;
    }
  double sm = sqrt(cfg->leg_link_mass), s2m = sqrt(2.0 * cfg->leg_link_mass);
  /* pose a moment later at the current twist (for the velocity-product terms below) */
  double p1[3], q1[4], R1[3][3];
  {
    const double *w = r->w, *q = r->q;
    double dq[4] = {0.5 * (-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]), 0.5 * (w[0] * q[0] + w[1] * q[3] - w[2] * q[2]),
                    0.5 * (-w[0] * q[3] + w[1] * q[0] + w[2] * q[1]), 0.5 * (w[0] * q[2] - w[1] * q[1] + w[2] * q[0])};
    double n2 = 0.0;
    for (int k = 0; k < 4; ++k) { q1[k] = q[k] + LEG_BIAS_DT * dq[k]; n2 += q1[k] * q1[k]; }
    n2 = sqrt(n2);
    for (int k = 0; k < 4; ++k) q1[k] /= n2;
    for (int k = 0; k < 3; ++k) p1[k] = r->p[k] + LEG_BIAS_DT * r->v[k];
    quat_to_rot(q1, R1);
  }
  for (int i = 0; i < cfg->n_cables; ++i) {
    leg_geom g;
    leg_geometry(cfg, r->leg_alpha[i], i, r->p, R, &g);
    double Jy[16][6], Jz[5][6];
    for (int k = 0; k < 6; ++k) {
      double ev[3] = {0, 0, 0}, ew[3] = {0, 0, 0}, y[16], z[5];
      if (k < 3) ev[k] = 1.0; else ew[k - 3] = 1.0;
      leg_rates(cfg, &g, ev, ew, y, z);
      for (int j = 0; j < 16; ++j) Jy[j][k] = y[j];
      for (int j = 0; j < 5; ++j) Jz[j][k] = z[j];
    }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double s = 0.0;
        for (int j = 0; j < 16; ++j) s += Jy[j][a] * Jy[j][b];
        M[a][b] += s;
      }
    if (Q) {
      double y[16], z[5], y1[16], z1[5];
      leg_rates(cfg, &g, r->v, r->w, y, z);
      /* velocity-product terms of the leg links: a_link = J xi' + J' xi, and J' xi = d/dt y(x(t), xi held) is taken as a
       * one-sided difference along the motion: the pose advanced by LEG_BIAS_DT at the current twist (first order, like
       * the integrator).  Isotropic link inertia => no gyroscopic torque of the links themselves. */
      leg_geom g1;
      leg_geometry(cfg, r->leg_alpha[i], i, p1, R1, &g1);
      leg_rates(cfg, &g1, r->v, r->w, y1, z1);
      for (int k = 0; k < 6; ++k) {
        double damp = 0.0;
        for (int j = 0; j < 5; ++j) damp += Jz[j][k] * z[j];
        double grav = 0.0;
        for (int j = 0; j < 3; ++j) grav += cfg->gravity[j] * (sm * Jy[10 + j][k] + s2m * Jy[13 + j][k]);
        double bias = 0.0;
        for (int j = 0; j < 16; ++j) bias += Jy[j][k] * ((y1[j] - y[j]) / LEG_BIAS_DT);
        Q[k] += grav - damp - bias;
      }
    }
  }
}

/* Cholesky solve of the 6x6 SPD system M x = b */
static void chol6_solve(double M[6][6], const double b[6], double x[6]) {
  double Lc[6][6];
  memset(Lc, 0, sizeof(Lc));
  for (int j = 0; j < 6; ++j) {
    double s = M[j][j];
    for (int k = 0; k < j; ++k) s -= Lc[j][k] * Lc[j][k];
    Lc[j][j] = sqrt(s);
    for (int i = j + 1; i < 6; ++i) {
      double t = M[i][j];
      for (int k = 0; k < j; ++k) t -= Lc[i][k] * Lc[j][k];
      Lc[i][j] = t / Lc[j][j];
    }
  }
  double yv[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= Lc[i][k] * yv[k];
    yv[i] = s / Lc[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = yv[i];
    for (int k = i + 1; k < 6; ++k) s -= Lc[k][i] * x[k];
    x[i] = s / Lc[i][i];
  }
}

void orc_legs_mass_matrix(const orc_robot *r, double Mout[36]) {
  double R[3][3], M[6][6];
  quat_to_rot(r->q, R);
  legs_assemble(r, R, M, NULL);
  memcpy(Mout, M, sizeof(M));
}
double orc_legs_kinetic_energy(const orc_robot *r) {
  double M[36], xi[6] = {r->v[0], r->v[1], r->v[2], r->w[0], r->w[1], r->w[2]}, e = 0.0;
  orc_legs_mass_matrix(r, M);
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) e += 0.5 * xi[a] * M[6 * a + b] * xi[b];
  return e;
}
double orc_legs_potential_energy(const orc_robot *r) {
  const orc_config *cfg = &r->cfg;
  double R[3][3], e = -cfg->mass * dot3(cfg->gravity, r->p);
  quat_to_rot(r->q, R);
  for (int i = 0; i < cfg->n_cables; ++i) {
    leg_geom g;
    leg_geometry(cfg, r->leg_alpha[i], i, r->p, R, &g);
    double B[3], C[3];
    for (int k = 0; k < 3; ++k) { B[k] = r->p[k] + g.r[k]; C[k] = B[k] + cfg->leg_cable_com * g.u[k]; }
    e -= cfg->leg_link_mass * (dot3(cfg->gravity, C) + 2.0 * dot3(cfg->gravity, B));
  }
  return e;
}
void orc_legs_joint_rates(const orc_robot *r, int leg, double rates[5]) {
  double R[3][3], y[16], z[5];
  leg_geom g;
  quat_to_rot(r->q, R);
  leg_geometry(&r->cfg, r->leg_alpha[leg], leg, r->p, R, &g);
  leg_rates(&r->cfg, &g, r->v, r->w, y, z);
  double sc = sqrt(r->cfg.passive_damping);
  for (int k = 0; k < 5; ++k) rates[k] = sc > 0.0 ? z[k] / sc : 0.0;
}

void orc_home_lengths(const orc_config *cfg, double *len) {
  orc_kinematics k;
  double zero[3] = {0, 0, 0};
  orc_kinematics_eval(cfg, NULL, cfg->home_pos, cfg->home_quat, zero, zero, &k);
  for (int i = 0; i < cfg->n_cables; ++i) len[i] = k.len[i];
}

void orc_robot_init(orc_robot *r, const orc_config *cfg) {
  memset(r, 0, sizeof(*r));
  r->cfg = *cfg;
  orc_home_lengths(cfg, r->home_len);
  legs_home_alpha(cfg, r->leg_alpha);
  for (int k = 0; k < 3; ++k) r->p[k] = cfg->home_pos[k];
  for (int k = 0; k < 4; ++k) r->q[k] = cfg->home_quat[k];
  for (int i = 0; i < cfg->n_cables; ++i) orc_cable_init(&r->cable[i], cfg, 0, 0);
  r->sine_period_steps = 10; /* 100 Hz publisher, 1 kHz physics (sinevelocitytest.cpp:7) */
  r->sine_pub_dt = 1.0 / 100.0;
}

/* CdprGazeboPlugin::cable*CommandCallback (.cpp:67-83): wrong length => silently dropped */
int orc_robot_velocity_cmd(orc_robot *r, const float *axes, int n_axes) {
  if (n_axes != r->cfg.n_cables) return -1;
  for (int i = 0; i < n_axes; ++i) r->vel_cmd[i] = axes[i];
  r->vel_cmd_received = 1;
  return 0;
}
int orc_robot_position_cmd(orc_robot *r, const float *axes, int n_axes) {
  if (n_axes != r->cfg.n_cables) return -1;
  for (int i = 0; i < n_axes; ++i) r->pos_cmd[i] = axes[i];
  r->pos_cmd_received = 1;
  return 0;
}
int orc_robot_effort_cmd(orc_robot *r, const double *force, int n_axes) {
  if (n_axes != r->cfg.n_cables) return -1;
  for (int i = 0; i < n_axes; ++i) orc_cable_set_force(&r->cable[i], force[i]);
  return 0;
}
void orc_robot_sine(orc_robot *r, double amp, double freq, double phase) {
  r->sine_enabled = 1;
  r->sine_amp = amp; r->sine_freq = freq; r->sine_phase = phase;
  r->sine_time = 0.0;
}

static void robot_step_impl(orc_robot *r, orc_force_fn fn, void *ctx) {
  const orc_config *cfg = &r->cfg;
  const int nc = cfg->n_cables;
  /* World::Step: simTime += dt, then the WorldUpdateBegin event (App. C.1) */
  int32_t dt_ns = (int32_t)llround(cfg->dt * 1e9);
  r->nsec += dt_ns;
  while (r->nsec >= 1000000000) { r->nsec -= 1000000000; r->sec += 1; }
  r->step_count += 1;

  /* sinevelocitytest.cpp:33-49, headless schedule of SURVEY.md App. A.4 */
  if (r->sine_enabled && ((r->step_count - 1) % r->sine_period_steps) == 0) {
    double sine = sin(r->sine_time * r->sine_freq * 2 * M_PI + r->sine_phase);
    double velocity = r->sine_amp * sine;
    /* squarevelocitytest.cpp:21-22: velocity = abs(sine) >= sqrt(0.5) ? copysign(cVelocityAmplitude, sine) : 0.0 */
    if (r->pub_shape == 1) velocity = fabs(sine) >= sqrt(0.5) ? copysign(r->sine_amp, sine) : 0.0;
    float axes[ORC_MAX_CABLES];
    for (int i = 0; i < nc; ++i) axes[i] = (float)velocity;
    orc_robot_velocity_cmd(r, axes, nc);
    r->sine_time += r->sine_pub_dt;
  }

  /* CdprGazeboPlugin::update, .cpp:206-221: velocity first, then position */
  if (r->vel_cmd_received) {
    for (int i = 0; i < nc; ++i) orc_cable_set_velocity_target(&r->cable[i], (double)r->vel_cmd[i]);
    r->vel_cmd_received = 0;
  }
  if (r->pos_cmd_received) {
    for (int i = 0; i < nc; ++i) orc_cable_set_position_target(&r->cable[i], (double)r->pos_cmd[i]);
    r->pos_cmd_received = 0;
  }

  orc_kinematics kin;
  orc_kinematics_eval(cfg, r->home_len, r->p, r->q, r->v, r->w, &kin);

  double F[3], M[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) F[k] = cfg->mass * cfg->gravity[k];
  double now = orc_time_double(r->sec, r->nsec);
  for (int i = 0; i < nc; ++i) { /* .cpp:222-228, ascending cable index */
    double f = fn ? fn(ctx, i, now, kin.joint_pos[i], kin.joint_vel[i])
                  : orc_cable_update(&r->cable[i], r->sec, r->nsec, kin.joint_pos[i], kin.joint_vel[i]);
    r->pid_force[i] = f;
    double eff = clampd(f, -cfg->effort_limit, cfg->effort_limit); /* Joint::SetForce truncation, cube.sdf:438 */
    r->effort[i] = eff;
    r->joint_pos[i] = kin.joint_pos[i];
    r->joint_vel[i] = kin.joint_vel[i];
    double tau = eff - cfg->cable_damping * kin.joint_vel[i]; /* App. C.4 */
    for (int k = 0; k < 3; ++k) {
      F[k] += tau * kin.unit[i][k];
      M[k] += tau * kin.arm[i][k];
    }
  }

  /* App. C.6: ODE-order semi-implicit Euler */
  double R[3][3];
  quat_to_rot(r->q, R);
  const double *I = cfg->inertia;
  double Ib[3][3] = {{I[0], I[3], I[4]}, {I[3], I[1], I[5]}, {I[4], I[5], I[2]}};
  double det = Ib[0][0] * (Ib[1][1] * Ib[2][2] - Ib[1][2] * Ib[2][1]) - Ib[0][1] * (Ib[1][0] * Ib[2][2] - Ib[1][2] * Ib[2][0]) +
               Ib[0][2] * (Ib[1][0] * Ib[2][1] - Ib[1][1] * Ib[2][0]);
  double Iinv[3][3];
  Iinv[0][0] = (Ib[1][1] * Ib[2][2] - Ib[1][2] * Ib[2][1]) / det;
  Iinv[0][1] = (Ib[0][2] * Ib[2][1] - Ib[0][1] * Ib[2][2]) / det;
  Iinv[0][2] = (Ib[0][1] * Ib[1][2] - Ib[0][2] * Ib[1][1]) / det;
  Iinv[1][0] = Iinv[0][1];
  Iinv[1][1] = (Ib[0][0] * Ib[2][2] - Ib[0][2] * Ib[2][0]) / det;
  Iinv[1][2] = (Ib[0][2] * Ib[1][0] - Ib[0][0] * Ib[1][2]) / det;
  Iinv[2][0] = Iinv[0][2];
  Iinv[2][1] = Iinv[1][2];
  Iinv[2][2] = (Ib[0][0] * Ib[1][1] - Ib[0][1] * Ib[1][0]) / det;

  double wb[3], Lb[3], Lw[3], gyro[3], Mb[3], ab[3], alpha[3];
  matT_vec(R, r->w, wb);
  mat_vec(Ib, wb, Lb);
  mat_vec(R, Lb, Lw);
  cross(r->w, Lw, gyro);
  for (int k = 0; k < 3; ++k) M[k] -= gyro[k];
  matT_vec(R, M, Mb);
  mat_vec(Iinv, Mb, ab);
  mat_vec(R, ab, alpha);

  const double h = cfg->dt;
  if (cfg->leg_model) { /* N2: platform + legs, M(x) (xi+ - xi)/h = Q */
    double Mm[6][6], Q[6] = {F[0], F[1], F[2], M[0], M[1], M[2]}, acc[6]; /* M[] already carries -gyro */
    legs_assemble(r, R, Mm, Q);
    chol6_solve(Mm, Q, acc);
    for (int k = 0; k < 3; ++k) {
      r->v[k] += h * acc[k];
      r->w[k] += h * acc[3 + k];
    }
  } else
  for (int k = 0; k < 3; ++k) {
    r->v[k] += h * (F[k] / cfg->mass);
    r->w[k] += h * alpha[k];
  }
  for (int k = 0; k < 3; ++k) r->p[k] += h * r->v[k];
  double qw = r->q[0], qx = r->q[1], qy = r->q[2], qz = r->q[3];
  double wx = r->w[0], wy = r->w[1], wz = r->w[2];
  double dq[4];
  dq[0] = 0.5 * (-wx * qx - wy * qy - wz * qz);
  dq[1] = 0.5 * (wx * qw + wy * qz - wz * qy);
  dq[2] = 0.5 * (-wx * qz + wy * qw + wz * qx);
  dq[3] = 0.5 * (wx * qy - wy * qx + wz * qw);
  double nq[4], n2 = 0.0;
  for (int k = 0; k < 4; ++k) { nq[k] = r->q[k] + h * dq[k]; n2 += nq[k] * nq[k]; }
  double nrm = sqrt(n2);
  for (int k = 0; k < 4; ++k) r->q[k] = nq[k] / nrm;
}

void orc_robot_step(orc_robot *r) { robot_step_impl(r, NULL, NULL); }
void orc_robot_step_ext(orc_robot *r, orc_force_fn fn, void *ctx) { robot_step_impl(r, fn, ctx); }

/* publishPlatformState, CdprGazeboPlugin.cpp:258-280: pos xyz, quat x y z w, lin, ang */
void orc_robot_platform_state(const orc_robot *r, double pose7[7], double twist6[6]) {
  pose7[0] = r->p[0]; pose7[1] = r->p[1]; pose7[2] = r->p[2];
  pose7[3] = r->q[1]; pose7[4] = r->q[2]; pose7[5] = r->q[3]; pose7[6] = r->q[0];
  for (int k = 0; k < 3; ++k) { twist6[k] = r->v[k]; twist6[3 + k] = r->w[k]; }
}

void orc_batch_step(orc_robot *robots, int64_t n, int64_t k_steps, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    for (int64_t s = 0; s < k_steps; ++s) orc_robot_step(&robots[i]);
}

void orc_batch_ik(const orc_config *cfg, int64_t n, const double *pose7, const double *twist6, double *len,
                  double *len_rate, double *wmat, int n_threads) {
  double home[ORC_MAX_CABLES];
  orc_home_lengths(cfg, home);
  const int nc = cfg->n_cables;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double *ps = pose7 + 7 * i, *tw = twist6 + 6 * i;
    double q[4] = {ps[6], ps[3], ps[4], ps[5]};
    orc_kinematics k;
    orc_kinematics_eval(cfg, home, ps, q, tw, tw + 3, &k);
    for (int c = 0; c < nc; ++c) {
      len[i * nc + c] = k.len[c];
      len_rate[i * nc + c] = k.len_rate[c];
      for (int d = 0; d < 3; ++d) {
        wmat[(i * nc + c) * 6 + d] = k.unit[c][d];
        wmat[(i * nc + c) * 6 + 3 + d] = k.arm[c][d];
      }
    }
  }
}

int orc_sizeof_robot(void) { return (int)sizeof(orc_robot); }

void orc_batch_init(orc_robot *robots, int64_t n, const orc_config *cfg, const double *pose7, const double *twist6,
                    const double *amp, const double *freq, const double *phase) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    orc_robot *r = &robots[i];
    orc_robot_init(r, cfg);
    if (pose7) {
      const double *ps = pose7 + 7 * i;
      r->p[0] = ps[0]; r->p[1] = ps[1]; r->p[2] = ps[2];
      r->q[0] = ps[6]; r->q[1] = ps[3]; r->q[2] = ps[4]; r->q[3] = ps[5];
    }
    if (twist6) {
      for (int k = 0; k < 3; ++k) { r->v[k] = twist6[6 * i + k]; r->w[k] = twist6[6 * i + 3 + k]; }
    }
    if (amp) orc_robot_sine(r, amp[i], freq ? freq[i] : 0.1, phase ? phase[i] : 0.0);
  }
}
void orc_batch_publisher(orc_robot *robots, int64_t n, int shape, double publish_hz) {
  for (int64_t i = 0; i < n; ++i) {
    orc_robot *r = &robots[i];
    r->pub_shape = shape;
    r->sine_pub_dt = 1.0 / publish_hz;                                   /* time += 1.0 / cPublishFrequency */
    r->sine_period_steps = (int32_t)llround(r->sine_pub_dt / r->cfg.dt); /* headless schedule, SURVEY.md App. A.4 */
    if (r->sine_period_steps < 1) r->sine_period_steps = 1;
  }
}
void orc_batch_velocity_cmd(orc_robot *robots, int64_t n, const float *axes) {
  for (int64_t i = 0; i < n; ++i) orc_robot_velocity_cmd(&robots[i], axes + i * robots[i].cfg.n_cables, robots[i].cfg.n_cables);
}
void orc_batch_position_cmd(orc_robot *robots, int64_t n, const float *axes) {
  for (int64_t i = 0; i < n; ++i) orc_robot_position_cmd(&robots[i], axes + i * robots[i].cfg.n_cables, robots[i].cfg.n_cables);
}
void orc_batch_effort_cmd(orc_robot *robots, int64_t n, const double *force) {
  for (int64_t i = 0; i < n; ++i) orc_robot_effort_cmd(&robots[i], force + i * robots[i].cfg.n_cables, robots[i].cfg.n_cables);
}
void orc_batch_platform_state(const orc_robot *robots, int64_t n, double *pose7, double *twist6) {
  for (int64_t i = 0; i < n; ++i) orc_robot_platform_state(&robots[i], pose7 + 7 * i, twist6 + 6 * i);
}
void orc_batch_joint_states(const orc_robot *robots, int64_t n, double *pos, double *vel, double *effort) {
  for (int64_t i = 0; i < n; ++i) {
    const orc_robot *r = &robots[i];
    orc_kinematics k;
    orc_kinematics_eval(&r->cfg, r->home_len, r->p, r->q, r->v, r->w, &k);
    for (int c = 0; c < r->cfg.n_cables; ++c) {
      pos[i * r->cfg.n_cables + c] = k.joint_pos[c];
      vel[i * r->cfg.n_cables + c] = k.joint_vel[c];
      effort[i * r->cfg.n_cables + c] = r->effort[c];
    }
  }
}

void orc_batch_last_outputs(const orc_robot *robots, int64_t n, double *jpos, double *jvel, double *pid_force, double *effort) {
  for (int64_t i = 0; i < n; ++i) {
    const orc_robot *r = &robots[i];
    for (int c = 0; c < r->cfg.n_cables; ++c) {
      int64_t o = i * r->cfg.n_cables + c;
      jpos[o] = r->joint_pos[c]; jvel[o] = r->joint_vel[c]; pid_force[o] = r->pid_force[c]; effort[o] = r->effort[c];
    }
  }
}
/* out[n][nc][2 (vel,pos)][6]: p_term, i_term (pre-clamp), d_term, i_err, cmd, d_err */
void orc_batch_pid_terms(const orc_robot *robots, int64_t n, double *out) {
  for (int64_t i = 0; i < n; ++i) {
    const orc_robot *r = &robots[i];
    for (int c = 0; c < r->cfg.n_cables; ++c)
      for (int k = 0; k < 2; ++k) {
        const orc_pid *p = k == 0 ? &r->cable[c].vel_pid : &r->cable[c].pos_pid;
        double *o = out + ((i * r->cfg.n_cables + c) * 2 + k) * 6;
        o[0] = p->dbg_p; o[1] = p->dbg_i; o[2] = p->dbg_d; o[3] = p->i_err; o[4] = p->cmd; o[5] = p->d_err;
      }
  }
}

/* ---- opaque single-object helpers for the ctypes tests --------------------------------------- */
orc_pid *orc_pid_new(const orc_pid_params *prm, int derive_absolute_time) {
  orc_pid *p = (orc_pid *)malloc(sizeof(orc_pid));
  orc_pid_init(p, prm);
  p->derive_absolute_time = derive_absolute_time;
  return p;
}
void orc_pid_free(orc_pid *p) { free(p); }
/* out[8]: dbg_p, dbg_i, dbg_d, p_err, i_err, d_err, cmd, last_time */
void orc_pid_get(const orc_pid *p, double *out) {
  out[0] = p->dbg_p; out[1] = p->dbg_i; out[2] = p->dbg_d; out[3] = p->p_err; out[4] = p->i_err; out[5] = p->d_err;
  out[6] = p->cmd; out[7] = p->last_time;
}
orc_cable *orc_cable_new(const orc_config *cfg) {
  orc_cable *c = (orc_cable *)malloc(sizeof(orc_cable));
  orc_cable_init(c, cfg, 0, 0);
  return c;
}
void orc_cable_free(orc_cable *c) { free(c); }
int orc_cable_mode(const orc_cable *c) { return c->mode; }
double orc_cable_last_position(const orc_cable *c) { return c->last_position; }

/* latched targets and mode per cable: vel_target[n][nc], pos_target[n][nc], mode[n][nc] (as double) */
void orc_batch_targets(const orc_robot *robots, int64_t n, double *vel_target, double *pos_target, double *mode) {
  for (int64_t i = 0; i < n; ++i) {
    const orc_robot *r = &robots[i];
    for (int c = 0; c < r->cfg.n_cables; ++c) {
      int64_t o = i * r->cfg.n_cables + c;
      vel_target[o] = r->cable[c].velocity_target; pos_target[o] = r->cable[c].position_target; mode[o] = r->cable[c].mode;
    }
  }
}
