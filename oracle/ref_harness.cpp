// ref_harness.cpp -- oracle level L0: drives the REFERENCE's own force-law sources
// (src/Pid.cpp, src/JointForceCalculator.cpp, include/cdpr_gazebo/{Pid,JointForceCalculator,
// Filter}.h), compiled unmodified from /root/reference against oracle/ref_shim/*.
// TEST INFRASTRUCTURE ONLY; built into the git-ignored oracle/_ref/libcdpr_ref.so.
//
// The harness restates only the glue the reference keeps in CdprGazeboPlugin.cpp, which
// cannot compile here (roscpp/Gazebo runtime): initJointsAndController (:98-172) and the
// command fan-out + force loop of update() (:202-228).
#include <memory>
#include <vector>
#include <cstring>
#include "cdpr_gazebo/JointForceCalculator.h"
#include <gazebo/physics/World.hh>
#include <sensor_msgs/Joy.h>

// process-wide globals the reference defines in CdprGazeboPlugin.cpp:174,200
bool theZeroest = false;
sensor_msgs::Joy pidMsg;

extern "C" {

typedef struct {
  double forward_gain, p_gain, i_gain, d_gain;
  int32_t d_degree, d_buffer_length;
  double i_limit, cmd_limit;
  double p_cutoff, p_quality;
  int32_t p_cascade;
  double d_cutoff, d_quality;
  int32_t d_cascade;
} ref_pid_params;

struct ref_plugin {
  gazebo::physics::WorldPtr world;
  gazebo::physics::ModelPtr model;
  std::vector<gazebo::physics::JointPtr> joints;
  std::vector<gazebo::physics::JointForceCalculator> calc;
  int n;
};

static gazebo::common::Pid::PidParameters convert(const ref_pid_params *p) {
  gazebo::common::Pid::PidParameters q;
  q.forwardGain = p->forward_gain; q.pGain = p->p_gain; q.iGain = p->i_gain; q.dGain = p->d_gain;
  q.dDegree = p->d_degree; q.dBufferLength = p->d_buffer_length;
  q.iLimit = p->i_limit; q.cmdLimit = p->cmd_limit;
  q.pFilter.relCutoff = p->p_cutoff; q.pFilter.quality = p->p_quality; q.pFilter.cascade = p->p_cascade;
  q.dFilter.relCutoff = p->d_cutoff; q.dFilter.quality = p->d_quality; q.dFilter.cascade = p->d_cascade;
  return q;
}

// CdprGazeboPlugin::initJointsAndController, :98-172 (sim time 0, joint position 0)
ref_plugin *ref_plugin_create(int n_cables, const ref_pid_params *vel, const ref_pid_params *pos, double velocity_epsilon) {
  if (pidMsg.axes.size() < 9) pidMsg.axes.resize(9);
  ref_plugin *h = new ref_plugin;
  h->n = n_cables;
  h->world = std::make_shared<gazebo::physics::World>();
  h->model = std::make_shared<gazebo::physics::Model>();
  h->model->mWorld = h->world;
  gazebo::common::Pid velocityPid(convert(vel));
  gazebo::common::Pid positionPid(convert(pos));
  h->joints.resize(n_cables);
  h->calc.resize(n_cables);
  for (int i = 0; i < n_cables; ++i) {
    h->joints[i] = std::make_shared<gazebo::physics::Joint>();
    gazebo::physics::JointForceCalculator fc(h->model, h->joints[i], positionPid, velocityPid, velocity_epsilon);
    fc.setPositionTarget(h->joints[i]->Position());
    h->calc[i] = fc;
  }
  return h;
}
void ref_plugin_destroy(ref_plugin *h) { delete h; }

void ref_plugin_set_time(ref_plugin *h, int32_t sec, int32_t nsec) { h->world->mTime = gazebo::common::Time(sec, nsec); }
void ref_plugin_set_joint(ref_plugin *h, int cable, double position, double velocity) {
  h->joints[cable]->mPosition = position;
  h->joints[cable]->mVelocity = velocity;
}
// update(), :206-219
void ref_plugin_velocity_cmd(ref_plugin *h, const float *axes) {
  for (int i = 0; i < h->n; ++i) h->calc[i].setVelocityTarget(axes[i]);
}
void ref_plugin_position_cmd(ref_plugin *h, const float *axes) {
  for (int i = 0; i < h->n; ++i) h->calc[i].setPositionTarget(axes[i]);
}
void ref_plugin_effort_cmd(ref_plugin *h, const double *force) {
  for (int i = 0; i < h->n; ++i) h->calc[i].setForce(force[i]);
}
// update(), :222-228 for one cable; also returns the P/I/D terms the reference publishes on "pid"
double ref_plugin_update_cable(ref_plugin *h, int cable, double *terms3) {
  theZeroest = true;
  double f = h->calc[cable].update();
  if (terms3) { terms3[0] = pidMsg.axes[0]; terms3[1] = pidMsg.axes[1]; terms3[2] = pidMsg.axes[2]; }
  theZeroest = false;
  return f;
}

// force-law hook with the signature of orc_force_fn (oracle/cdpr_oracle.h); ctx = ref_plugin*.
// The caller must have set the sim time for this step.
double ref_force_fn(void *ctx, int cable, double sim_time, double joint_pos, double joint_vel) {
  (void)sim_time;
  ref_plugin *h = static_cast<ref_plugin *>(ctx);
  h->joints[cable]->mPosition = joint_pos;
  h->joints[cable]->mVelocity = joint_vel;
  return h->calc[cable].update();
}

// direct access to one reference Pid (derive/fitPolynomial are public, Pid.h:108-109)
struct ref_pid { gazebo::common::Pid pid; };
ref_pid *ref_pid_create(const ref_pid_params *p) {
  if (pidMsg.axes.size() < 9) pidMsg.axes.resize(9);
  ref_pid *r = new ref_pid;
  r->pid = gazebo::common::Pid(convert(p));
  return r;
}
void ref_pid_destroy(ref_pid *r) { delete r; }
void ref_pid_reset(ref_pid *r) { r->pid.reset(); }
double ref_pid_update(ref_pid *r, double desired, double actual, double now) { return r->pid.update(desired, actual, now); }
double ref_pid_derive(ref_pid *r, double value, double now) { return r->pid.derive(value, now); }
double ref_pid_update_terms(ref_pid *r, double desired, double actual, double now, double *terms3) {
  theZeroest = true;
  double c = r->pid.update(desired, actual, now);
  terms3[0] = pidMsg.axes[0]; terms3[1] = pidMsg.axes[1]; terms3[2] = pidMsg.axes[2];
  theZeroest = false;
  return c;
}
}

// ---------------------------------------------------------------------------------------
// Reference arm of bench.py: the reduced model (oracle L1 kinematics + integration) with the
// REFERENCE's JointForceCalculator/Pid as the force law, OpenMP over robots.
#include "cdpr_oracle.h"
#include <omp.h>
extern "C" {
static void to_ref(const orc_pid_params *p, ref_pid_params *q) {
  q->forward_gain = p->forward_gain; q->p_gain = p->p_gain; q->i_gain = p->i_gain; q->d_gain = p->d_gain;
  q->d_degree = p->d_degree; q->d_buffer_length = p->d_buffer_length; q->i_limit = p->i_limit; q->cmd_limit = p->cmd_limit;
  q->p_cutoff = p->p_cutoff; q->p_quality = p->p_quality; q->p_cascade = p->p_cascade;
  q->d_cutoff = p->d_cutoff; q->d_quality = p->d_quality; q->d_cascade = p->d_cascade;
}
// robots: n oracle robots already initialised (orc_batch_init). Creates one reference plugin per
// robot, runs k_steps with the reference force law, leaves the final state in robots[].
// theZeroest/pidMsg are process-wide in the reference, so telemetry is off (theZeroest=false)
// and the loop is thread-safe.
void ref_batch_step(orc_robot *robots, int64_t n, int64_t k_steps, int n_threads) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  if (pidMsg.axes.size() < 9) pidMsg.axes.resize(9);
  theZeroest = false;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    orc_robot *r = &robots[i];
    ref_pid_params vel, pos;
    to_ref(&r->cfg.vel_pid, &vel); to_ref(&r->cfg.pos_pid, &pos);
    ref_plugin *h = ref_plugin_create(r->cfg.n_cables, &vel, &pos, r->cfg.velocity_epsilon);
    for (int64_t s = 0; s < k_steps; ++s) {
      // the oracle advances sim time inside the step; mirror it into the shim World first
      int32_t dt_ns = (int32_t)llround(r->cfg.dt * 1e9);
      int32_t sec = r->sec, nsec = r->nsec + dt_ns;
      while (nsec >= 1000000000) { nsec -= 1000000000; sec += 1; }
      ref_plugin_set_time(h, sec, nsec);
      // command fan-out happens inside the oracle step on ITS cables; replay it on the reference's
      bool sinePublish = r->sine_enabled && ((r->step_count % r->sine_period_steps) == 0);
      if (sinePublish) {
        double velocity = r->sine_amp * sin(r->sine_time * r->sine_freq * 2 * M_PI + r->sine_phase);
        float axes[ORC_MAX_CABLES];
        for (int c = 0; c < r->cfg.n_cables; ++c) axes[c] = (float)velocity;
        ref_plugin_velocity_cmd(h, axes);
      }
      else if (r->vel_cmd_received) ref_plugin_velocity_cmd(h, r->vel_cmd);
      if (r->pos_cmd_received) ref_plugin_position_cmd(h, r->pos_cmd);
      orc_robot_step_ext(r, ref_force_fn, h);
    }
    ref_plugin_destroy(h);
  }
}
}
