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
  // One call does what CdprGazeboPlugin::update does (.cpp:202-246): the latched messages are fanned out velocity first,
  // then position (.cpp:206-219; the mode switch and Pid reset of the setters happen inside), every cable's force is set,
  // the step runs, and the publish pairs the state READ AT THIS UPDATE with the force of this step (.cpp:248-280).
  const double now = simTime() + mStep;  // World::Step advances sim time before the callback
  const bool publish = (now - mPreviousProcessingTime) > mPublishPeriod;  // CdprGazeboPlugin.cpp:237
  check(cdpr_update(mHandle, mVelocityCommandReceived ? mVelocityCommand.axes.data() : nullptr,
                    mPositionCommandReceived ? mPositionCommand.axes.data() : nullptr,
                    publish ? mJointStates.position.data() : nullptr, publish ? mJointStates.velocity.data() : nullptr,
                    publish ? mJointStates.effort.data() : nullptr, publish ? mPlatformState.pose.data() : nullptr,
                    publish ? mPlatformState.twist.data() : nullptr),
        "cdpr_update");
  mVelocityCommandReceived = mPositionCommandReceived = false;
  if (publish) mPreviousProcessingTime = now;
}

double CdprBatchPlugin::simTime() const { return cdpr_sim_time(mHandle); }

}  // namespace cdpr_host
