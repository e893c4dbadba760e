// Stand-in for gazebo/physics/{World,Model,Joint}.hh -- TEST INFRASTRUCTURE ONLY.
// The harness sets sim time and the joint read-backs; the reference code reads them.
#ifndef CDPR_SHIM_GZ_PHYSICS
#define CDPR_SHIM_GZ_PHYSICS
#include <memory>
#include "gazebo/common/Time.hh"
#include "gazebo/gazebo.hh"   // the real physics headers pull in common/Console.hh (gzdbg)
namespace gazebo { namespace physics {
class World { public: common::Time SimTime() const { return mTime; } common::Time mTime; };
typedef std::shared_ptr<World> WorldPtr;
class Model { public: WorldPtr GetWorld() const { return mWorld; } WorldPtr mWorld; };
typedef std::shared_ptr<Model> ModelPtr;
class Joint {
public:
  double Position(unsigned = 0) const { return mPosition; }
  double GetVelocity(unsigned) const { return mVelocity; }
  double mPosition = 0.0, mVelocity = 0.0;
};
typedef std::shared_ptr<Joint> JointPtr;
}}
#endif
