// Stand-in for roscpp (+ the bits of boost it drags in): enough for /root/reference's nodelets to compile UNMODIFIED and be driven
// synchronously by oracle/refbuild/ref_*.cpp.  TEST INFRASTRUCTURE (oracle/_ref build only).
// Publishers do not transport anything: publish() records the message per topic in alego_ref::bus() and calls an optional hook while
// the publishing callback is still on the stack (that is how the driver reads ImageProjection's private images before pcCB clears them).
#ifndef ALEGO_REF_SHIM_ROS_H
#define ALEGO_REF_SHIM_ROS_H
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <chrono>
#include <thread>

namespace boost {
using std::shared_ptr;
template <typename F, typename... A>
auto bind(F &&f, A &&... a) -> decltype(std::bind(std::forward<F>(f), std::forward<A>(a)...)) {
  return std::bind(std::forward<F>(f), std::forward<A>(a)...);
}
}  // namespace boost
using std::placeholders::_1;  // boost/bind.hpp puts its placeholders in the global namespace
using std::placeholders::_2;

namespace alego_ref {
// Per-THREAD state of the stand-in middleware: a driver call and everything the reference code does underneath it run on the
// caller's thread, so independent nodes can be driven from independent threads (bench.py's parity check does).
struct Bus {
  std::map<std::string, std::shared_ptr<const void>> last;  // most recent message per topic
  std::map<std::string, int> count;
  std::function<void(const std::string &)> hook;            // called inside publish()
  std::function<bool()> ok_fn;                              // if set, decides ros::ok() on this thread
  bool capture_log = false;                                 // keep the reference's own timing log lines (see log_printf)
  std::vector<std::string> log;
};
inline Bus &bus() { static thread_local Bus b; return b; }
// Per-LIBRARY state, seen by every thread — including the worker threads LaserMapping::onInit spawns.
struct Globals {
  std::atomic<bool> ok{true};             // what ros::ok() returns on a thread without ok_fn
  std::atomic<bool> park_sleepers{false}; // ros::Rate::sleep() never returns (parks stray worker threads)
  std::atomic<int> subscribers{1};        // what Publisher::getNumSubscribers() returns
};
inline Globals &globals() { static Globals g; return g; }
// The reference reports its stage timings (TicToc, utility.h:99-120) through NODELET_INFO.  Only the lines whose format string
// starts with one of these prefixes are formatted and kept (several other log calls in the reference pass fewer arguments than
// their format string names, so nothing else is ever handed to vsnprintf).
inline void log_printf(const char *fmt, ...) {
  Bus &b = bus();
  if (!b.capture_log) return;
  static const char *const keep[] = {"mapping data assosiation time", "mapping solver time", "build kdtree time", "downsampleCurrentScan:",
                                     "mapping whole time"};
  bool hit = false;
  for (const char *k : keep) hit = hit || std::strncmp(fmt, k, std::strlen(k)) == 0;
  if (!hit) return;
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  b.log.push_back(buf);
}
}  // namespace alego_ref

namespace ros {

struct Time {
  double sec_ = 0;
  Time() {}
  explicit Time(double s) : sec_(s) {}
  double toSec() const { return sec_; }
  Time &fromSec(double s) { sec_ = s; return *this; }
  static Time now() { return Time(); }
};
struct Duration {
  double d_;
  explicit Duration(double d = 0) : d_(d) {}
  bool sleep() const { return true; }
};
inline bool ok() { alego_ref::Bus &b = alego_ref::bus(); return b.ok_fn ? b.ok_fn() : alego_ref::globals().ok.load(); }
inline void spinOnce() {}
struct Rate {
  explicit Rate(double) {}
  bool sleep() {
    while (alego_ref::globals().park_sleepers.load()) std::this_thread::sleep_for(std::chrono::hours(1));
    return true;
  }
};

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string &topic) : topic_(topic) {}
  int getNumSubscribers() const { return alego_ref::globals().subscribers.load(); }
  template <typename M>
  void publish(const std::shared_ptr<M> &msg) const {
    alego_ref::Bus &b = alego_ref::bus();
    b.last[topic_] = std::static_pointer_cast<const void>(std::shared_ptr<const M>(msg));
    ++b.count[topic_];
    if (b.hook) b.hook(topic_);
  }
  const std::string &getTopic() const { return topic_; }

 private:
  std::string topic_;
};
struct Subscriber {};
struct ServiceServer {};

class NodeHandle {
 public:
  template <typename M>
  Publisher advertise(const std::string &topic, int) { return Publisher(topic); }
  template <typename M, typename F>
  Subscriber subscribe(const std::string &, int, F) { return Subscriber(); }
  template <typename... A>
  ServiceServer advertiseService(const std::string &, A...) { return ServiceServer(); }
};

}  // namespace ros

namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs

#define ALEGO_REF_NOLOG(...) do { } while (0)
#define ROS_INFO(...) ALEGO_REF_NOLOG()
#define ROS_WARN(...) ALEGO_REF_NOLOG()
#define ROS_ERROR(...) ALEGO_REF_NOLOG()
#define ROS_INFO_STREAM(x) ALEGO_REF_NOLOG()
#define ROS_WARN_STREAM(x) ALEGO_REF_NOLOG()
#define NODELET_INFO(...) alego_ref::log_printf(__VA_ARGS__)
#define NODELET_WARN(...) ALEGO_REF_NOLOG()
#define NODELET_ERROR(...) ALEGO_REF_NOLOG()
#define NODELET_INFO_STREAM(x) ALEGO_REF_NOLOG()
#define NODELET_WARN_STREAM(x) ALEGO_REF_NOLOG()
#define NODELET_WARN_COND(c, ...) ALEGO_REF_NOLOG()
#define NODELET_WARN_ONCE(...) ALEGO_REF_NOLOG()
#endif
