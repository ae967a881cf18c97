// stand-in for <ros/ros.h> (see README.md in this directory)
#pragma once
#include <cstdio>
#include <memory>
#include <string>
namespace ros {
struct Time {
  double sec_ = 0;
  double toSec() const { return sec_; }
  Time &fromSec(double s) { sec_ = s; return *this; }
  static Time now() { return Time(); }
};
struct Publisher {
  template <class M> void publish(const M &) const {}
  unsigned getNumSubscribers() const { return 0; }
};
struct Subscriber {};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string &, unsigned) { return Publisher(); }
  template <class M, class T> Subscriber subscribe(const std::string &, unsigned, void (T::*)(const std::shared_ptr<const M> &), T *) { return Subscriber(); }
  template <class V> bool param(const std::string &, V &v, const V &d) const { v = d; return false; }
};
}  // namespace ros
#define ROS_FATAL(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_WARN(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_INFO(...) std::fprintf(stderr, __VA_ARGS__)
