// Stand-in for ignition/math/Helpers.hh -- TEST INFRASTRUCTURE ONLY (clamp only).
#ifndef CDPR_SHIM_IGN_HELPERS
#define CDPR_SHIM_IGN_HELPERS
#include <algorithm>
namespace ignition { namespace math {
template <typename T> inline T clamp(T _v, T _min, T _max) { return std::max(std::min(_v, _max), _min); }
}}
#endif
