// CdprBatchPlugin.cpp -- see CdprBatchPlugin.h.  Links against libcdpr_b200.so only.
#include "CdprBatchPlugin.h"

namespace cdpr_host {

CdprBatchPlugin::~CdprBatchPlugin() {
  if (mHandle) cdpr_destroy(mHandle);
}

void CdprBatchPlugin::check(int rc, const char *what) const {
  if (rc != CDPR_OK) throw std::runtime_error(std::string(what) + ": " + cdpr_last_error(mHandle));
}

void CdprBatchPlugin::Load(const cdpr_config &cfg, int64_t nInstances, int device) {
  int rc = cdpr_create(&cfg, nInstances, device, &mHandle);
  if (rc == CDPR_ERR_BAD_CABLE_COUNT) throw std::runtime_error("invalid joint count");  // CdprGazeboPlugin.cpp:168
  if (rc != CDPR_OK) throw std::runtime_error(std::string("cdpr_create: ") + cdpr_last_error(nullptr));
  mInstances = nInstances;
  mWireCount = cfg.n_cables;
  mJointStates.name.resize(mWireCount);
  for (int i = 0; i < mWireCount; ++i) mJointStates.name[i] = "cable" + std::to_string(i);
  const size_t nj = (size_t)nInstances * mWireCount;
  mJointStates.position.resize(nj);
  mJointStates.velocity.resize(nj);
  mJointStates.effort.resize(nj);
  mPlatformState.pose.resize((size_t)nInstances * 7);
  mPlatformState.twist.resize((size_t)nInstances * 6);
  mPreviousProcessingTime = 0.0;
  mStep = cfg.dt;
}

void CdprBatchPlugin::cableVelocityCommandCallback(const Joy &msg) {
  if (msg.axes.size() == (size_t)mInstances * mWireCount) {
    mVelocityCommand = msg;
    mVelocityCommandReceived = true;
  }
}

void CdprBatchPlugin::cablePositionCommandCallback(const Joy &msg) {
  if (msg.axes.size() == (size_t)mInstances * mWireCount) {
    mPositionCommand = msg;
    mPositionCommandReceived = true;
  }
}

void CdprBatchPlugin::update() {
  // fan-out order of CdprGazeboPlugin.cpp:206-219: velocity, then position (the library applies them in that order
  // at the next step; the mode switch and Pid reset of the setters happen there)
  if (mVelocityCommandReceived) {
    check(cdpr_set_velocity_cmd(mHandle, mVelocityCommand.axes.data(), mInstances, mWireCount), "cdpr_set_velocity_cmd");
    mVelocityCommandReceived = false;
  }
  if (mPositionCommandReceived) {
    check(cdpr_set_position_cmd(mHandle, mPositionCommand.axes.data(), mInstances, mWireCount), "cdpr_set_position_cmd");
    mPositionCommandReceived = false;
  }
  // the plugin publishes from inside update(): the state BEFORE this step together with the force set in this step.
  // Position/velocity are read first, the step runs, then the effort of this step is read.
  const double now = simTime() + mStep;  // World::Step advances sim time before the callback
  const bool publish = (now - mPreviousProcessingTime) > mPublishPeriod;  // CdprGazeboPlugin.cpp:237
  if (publish) {
    check(cdpr_get_joint_states(mHandle, mJointStates.position.data(), mJointStates.velocity.data(), nullptr), "cdpr_get_joint_states");
    publishPlatformState();
  }
  check(cdpr_step(mHandle, 1), "cdpr_step");
  if (publish) {
    mPreviousProcessingTime = now;
    publishJointStates();
  }
}

void CdprBatchPlugin::publishJointStates() {  // effort = Joint::GetForce(0) of the step just taken (.cpp:253)
  check(cdpr_get_joint_states(mHandle, nullptr, nullptr, mJointStates.effort.data()), "cdpr_get_joint_states");
}

void CdprBatchPlugin::publishPlatformState() {  // .cpp:258-280
  check(cdpr_get_platform_state(mHandle, mPlatformState.pose.data(), mPlatformState.twist.data()), "cdpr_get_platform_state");
}

double CdprBatchPlugin::simTime() const { return cdpr_sim_time(mHandle); }

}  // namespace cdpr_host
