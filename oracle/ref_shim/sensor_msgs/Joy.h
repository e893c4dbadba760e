// Stand-in for sensor_msgs/Joy.h -- TEST INFRASTRUCTURE ONLY (float32 axes[] only).
#ifndef CDPR_SHIM_JOY
#define CDPR_SHIM_JOY
#include <vector>
namespace sensor_msgs { struct Joy { std::vector<float> axes; }; }
#endif
