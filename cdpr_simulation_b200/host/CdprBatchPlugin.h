// CdprBatchPlugin.h -- C++ host shim shaped like the reference plugin.
//
// Mirrors gazebo::CdprGazeboPlugin (include/cdpr_gazebo/CdprGazeboPlugin.h:18-103, src/CdprGazeboPlugin.cpp)
// for a BATCH of N robots with the ROS / Gazebo types replaced by plain structs: the same members, callbacks and
// per-step order (drain command queues -> velocity fan-out -> position fan-out -> forces -> publish), but every
// numeric step is one call into the C ABI (include/cdpr_b200.h).  A maintainer of the reference would keep their
// Load()/callbacks and swap the body of update() for these calls.
#ifndef CDPR_BATCH_PLUGIN_H
#define CDPR_BATCH_PLUGIN_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cdpr_b200.h"

namespace cdpr_host {

// sensor_msgs/Joy: float32 axes[]; here the axes of all N robots, robot-major ([N][NC])
struct Joy {
  std::vector<float> axes;
};
// sensor_msgs/JointState for the batch: [N][NC] each
struct JointState {
  std::vector<std::string> name;  // "cable0".."cable<NC-1>" (CdprGazeboPlugin.cpp:150-157)
  std::vector<double> position, velocity, effort;
};
// cdpr_gazebo/PlatformState (msg/PlatformState.msg) for the batch: pose [N][7] = x y z qx qy qz qw, twist [N][6]
struct PlatformState {
  std::vector<double> pose, twist;
};

class CdprBatchPlugin {
public:
  CdprBatchPlugin() = default;
  ~CdprBatchPlugin();
  CdprBatchPlugin(const CdprBatchPlugin &) = delete;
  CdprBatchPlugin &operator=(const CdprBatchPlugin &) = delete;

  // Load(): reads the "ROS parameters" (cfg), creates N robots on `device`; throws like
  // gazebo::common::Exception("invalid joint count") when the cable count is unusable (.cpp:167-169)
  void Load(const cdpr_config &cfg, int64_t nInstances, int device);

  // subscriber callbacks: a message whose axes.size() != N * cWireCount is dropped silently (.cpp:67-83)
  void cableVelocityCommandCallback(const Joy &msg);
  void cablePositionCommandCallback(const Joy &msg);

  // one WorldUpdateBegin callback + physics step (.cpp:202-246); publishes every publishPeriod seconds of sim time
  void update();

  const JointState &jointStates() const { return mJointStates; }
  const PlatformState &platformState() const { return mPlatformState; }
  double simTime() const;
  int64_t instances() const { return mInstances; }
  int wireCount() const { return mWireCount; }
  void setPublishPeriod(double seconds) { mPublishPeriod = seconds; }
  cdpr_handle handle() const { return mHandle; }

private:
  void check(int rc, const char *what) const;

  cdpr_handle mHandle = nullptr;
  int64_t mInstances = 0;
  int mWireCount = 0;
  double mPublishPeriod = 0.0, mPreviousProcessingTime = 0.0, mStep = 0.001;
  Joy mVelocityCommand, mPositionCommand;
  bool mVelocityCommandReceived = false, mPositionCommandReceived = false;
  JointState mJointStates;
  PlatformState mPlatformState;
};

}  // namespace cdpr_host
#endif
