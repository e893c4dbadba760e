// Stand-in for gazebo/gazebo.hh -- TEST INFRASTRUCTURE ONLY: gzdbg swallows everything.
#ifndef CDPR_SHIM_GZ_GAZEBO
#define CDPR_SHIM_GZ_GAZEBO
#include <ostream>
namespace gazebo { namespace shim {
struct NullStream {
  template <class T> NullStream &operator<<(T const &) { return *this; }
  NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
inline NullStream &nullStream() { static NullStream s; return s; }
}}
#define gzdbg (gazebo::shim::nullStream())
#endif
