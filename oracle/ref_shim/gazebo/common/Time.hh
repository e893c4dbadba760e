// Stand-in for gazebo/common/Time.hh (Gazebo 9) -- TEST INFRASTRUCTURE ONLY.
// Integer sec/nsec like gazebo::common::Time; Double() = sec + nsec*1e-9.
#ifndef CDPR_SHIM_GZ_TIME
#define CDPR_SHIM_GZ_TIME
#include <cstdint>
#include <cmath>
namespace gazebo { namespace common {
class Time {
public:
  Time() : sec(0), nsec(0) {}
  Time(int32_t s, int32_t ns) : sec(s), nsec(ns) { Correct(); }
  Time(double t) { sec = (int32_t)std::floor(t); nsec = (int32_t)std::round((t - sec) * 1e9); Correct(); }
  double Double() const { return (double)sec + (double)nsec * 1e-9; }
  Time operator-(Time const &o) const { return Time(sec - o.sec, nsec - o.nsec); }
  Time operator+(Time const &o) const { return Time(sec + o.sec, nsec + o.nsec); }
  bool operator>(double t) const { Time o(t); return sec > o.sec || (sec == o.sec && nsec > o.nsec); }
  bool operator>(int t) const { return *this > (double)t; }
  int32_t sec, nsec;
private:
  void Correct() {
    if (sec > 0 && nsec < 0) { int32_t n = std::abs(nsec / 1000000000) + 1; sec -= n; nsec += n * 1000000000; }
    if (sec < 0 && nsec > 0) { int32_t n = std::abs(nsec / 1000000000) + 1; sec += n; nsec -= n * 1000000000; }
    sec += nsec / 1000000000; nsec = nsec % 1000000000;
  }
};
}}
#endif
