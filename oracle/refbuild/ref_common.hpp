// Shared helpers of the oracle/_ref drivers (TEST INFRASTRUCTURE): named result blobs read back through one C entry point.
#ifndef ALEGO_REF_COMMON_HPP
#define ALEGO_REF_COMMON_HPP
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace alego_ref {
struct Blobs {
  std::map<std::string, std::vector<uint8_t>> m;
  template <typename T>
  void put(const std::string &k, const T *p, std::size_t n) {
    std::vector<uint8_t> &b = m[k];
    b.resize(n * sizeof(T));
    if (n) std::memcpy(b.data(), p, n * sizeof(T));
  }
  template <typename T>
  void put(const std::string &k, const std::vector<T> &v) { put(k, v.data(), v.size()); }
  template <typename T>
  void put1(const std::string &k, T v) { put(k, &v, 1); }
  int64_t get(const char *k, void *dst, std::size_t cap) const {
    auto it = m.find(k);
    if (it == m.end()) return -1;
    if (dst) {
      if (it->second.size() > cap) return -2;
      if (!it->second.empty()) std::memcpy(dst, it->second.data(), it->second.size());
    }
    return static_cast<int64_t>(it->second.size());
  }
};

// the reference prints progress with std::cout (e.g. laserMapping.cpp:477); keep the host process' stdout clean while it runs
// Counted per library: calls may run on several threads at once (one chain per thread), and std::cout is process-wide.
struct MuteCout {
  static std::mutex &mu() { static std::mutex m; return m; }
  static int &depth() { static int d = 0; return d; }
  static std::streambuf *&saved() { static std::streambuf *s = nullptr; return s; }
  MuteCout() {
    std::lock_guard<std::mutex> g(mu());
    if (depth()++ == 0) saved() = std::cout.rdbuf(nullptr);
  }
  ~MuteCout() {
    std::lock_guard<std::mutex> g(mu());
    if (--depth() == 0) { std::cout.rdbuf(saved()); std::cout.clear(); }
  }
};

template <typename Cloud>
inline std::vector<float> cloud_xyzi(const Cloud &c) {
  std::vector<float> v(c.points.size() * 4);
  for (std::size_t i = 0; i < c.points.size(); ++i) {
    v[4 * i] = c.points[i].x; v[4 * i + 1] = c.points[i].y; v[4 * i + 2] = c.points[i].z; v[4 * i + 3] = c.points[i].intensity;
  }
  return v;
}
}  // namespace alego_ref
#endif
