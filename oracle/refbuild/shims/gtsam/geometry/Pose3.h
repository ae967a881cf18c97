// Stand-in for the slice of GTSAM 4 that LaserMapping touches (TEST INFRASTRUCTURE, oracle/_ref build only).  The pose graph is off
// the pinned hot path (iSAM2 out of scope).  ISAM2 here is NOT an optimiser: calculateEstimate() returns the initial
// values that were inserted, which is the exact optimum of the prior + odometry chain LaserMapping::saveKeyFramesAndFactor builds as
// long as no loop-closure factor was added (every between-factor is consistent with the inserted values by construction,
// laserMapping.cpp:512-515).  Rot3 / Pose3 algebra restates gtsam/geometry/Rot3M.cpp (RzRyRx, Quaternion -> matrix, xyz() via RQ).
#ifndef ALEGO_REF_SHIM_GTSAM_H
#define ALEGO_REF_SHIM_GTSAM_H
#include <cmath>
#include <map>
#include <memory>
#include <vector>
#include <Eigen/Geometry>

namespace gtsam {
typedef Eigen::VectorXd Vector;
typedef std::size_t Key;

struct Point3 {
  double v[3];
  Point3() : v{0, 0, 0} {}
  Point3(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};

class Rot3 {
 public:
  Rot3() { R_.setIdentity(); }
  explicit Rot3(const Eigen::Matrix3d &R) : R_(R) {}
  static Rot3 Quaternion(double w, double x, double y, double z) { return Rot3(Eigen::Quaterniond(w, x, y, z).toRotationMatrix()); }
  static Rot3 Rx(double t) { Eigen::Matrix3d m; const double c = std::cos(t), s = std::sin(t); m.setIdentity(); m(1, 1) = c; m(1, 2) = -s; m(2, 1) = s; m(2, 2) = c; return Rot3(m); }
  static Rot3 Ry(double t) { Eigen::Matrix3d m; const double c = std::cos(t), s = std::sin(t); m.setIdentity(); m(0, 0) = c; m(0, 2) = s; m(2, 0) = -s; m(2, 2) = c; return Rot3(m); }
  static Rot3 Rz(double t) { Eigen::Matrix3d m; const double c = std::cos(t), s = std::sin(t); m.setIdentity(); m(0, 0) = c; m(0, 1) = -s; m(1, 0) = s; m(1, 1) = c; return Rot3(m); }
  static Rot3 RzRyRx(double x, double y, double z) {  // Rot3M.cpp
    const double cx = std::cos(x), sx = std::sin(x), cy = std::cos(y), sy = std::sin(y), cz = std::cos(z), sz = std::sin(z);
    const double ss_ = sx * sy, cs_ = cx * sy, sc_ = sx * cy, cc_ = cx * cy, c_s = cx * sz, s_s = sx * sz, _cs = cy * sz, _cc = cy * cz;
    const double s_c = sx * cz, c_c = cx * cz, ssc = ss_ * cz, csc = cs_ * cz, sss = ss_ * sz, css = cs_ * sz;
    Eigen::Matrix3d m;
    m(0, 0) = _cc; m(0, 1) = -c_s + ssc; m(0, 2) = s_s + csc;
    m(1, 0) = _cs; m(1, 1) = c_c + sss; m(1, 2) = -s_c + css;
    m(2, 0) = -sy; m(2, 1) = sc_; m(2, 2) = cc_;
    return Rot3(m);
  }
  const Eigen::Matrix3d &matrix() const { return R_; }
  Rot3 operator*(const Rot3 &o) const { return Rot3(R_ * o.R_); }
  Rot3 inverse() const { return Rot3(R_.transpose()); }
  void xyz(double *q) const {  // RQ(matrix())
    const Eigen::Matrix3d &A = R_;
    const double x = -std::atan2(-A(2, 1), A(2, 2));
    const Eigen::Matrix3d B = A * Rx(-x).matrix();
    const double y = -std::atan2(B(2, 0), B(2, 2));
    const Eigen::Matrix3d C = B * Ry(-y).matrix();
    const double z = -std::atan2(-C(1, 0), C(1, 1));
    q[0] = x; q[1] = y; q[2] = z;
  }
  double roll() const { double q[3]; xyz(q); return q[0]; }
  double pitch() const { double q[3]; xyz(q); return q[1]; }
  double yaw() const { double q[3]; xyz(q); return q[2]; }

 private:
  Eigen::Matrix3d R_;
};

class Pose3 {
 public:
  Pose3() {}
  Pose3(const Rot3 &R, const Point3 &t) : R_(R), t_(t) {}
  const Rot3 &rotation() const { return R_; }
  const Point3 &translation() const { return t_; }
  Pose3 inverse() const {
    const Rot3 Rt = R_.inverse();
    const Eigen::Vector3d p = Rt.matrix() * Eigen::Vector3d(t_.x(), t_.y(), t_.z());
    return Pose3(Rt, Point3(-p.x(), -p.y(), -p.z()));
  }
  Pose3 operator*(const Pose3 &o) const {
    const Eigen::Vector3d p = R_.matrix() * Eigen::Vector3d(o.t_.x(), o.t_.y(), o.t_.z());
    return Pose3(R_ * o.R_, Point3(p.x() + t_.x(), p.y() + t_.y(), p.z() + t_.z()));
  }
  Pose3 between(const Pose3 &o) const { return inverse() * o; }

 private:
  Rot3 R_;
  Point3 t_;
};

namespace noiseModel {
struct Diagonal {
  typedef std::shared_ptr<Diagonal> shared_ptr;
  static shared_ptr Variances(const Vector &) { return std::make_shared<Diagonal>(); }
};
}  // namespace noiseModel

struct NonlinearFactor { virtual ~NonlinearFactor() {} };
template <typename T>
struct PriorFactor : NonlinearFactor {
  PriorFactor(Key, const T &, const noiseModel::Diagonal::shared_ptr &) {}
};
template <typename T>
struct BetweenFactor : NonlinearFactor {
  BetweenFactor(Key, Key, const T &, const noiseModel::Diagonal::shared_ptr &) {}
};

class NonlinearFactorGraph {
 public:
  template <typename F>
  void add(const F &) { ++n_; }
  void resize(std::size_t n) { n_ = n; }
  std::size_t size() const { return n_; }

 private:
  std::size_t n_ = 0;
};

class Values {
 public:
  void insert(Key k, const Pose3 &p) { v_[k] = p; }
  void clear() { v_.clear(); }
  std::size_t size() const { return v_.size(); }
  template <typename T>
  const T &at(Key k) const { return v_.at(k); }
  const std::map<Key, Pose3> &all() const { return v_; }

 private:
  std::map<Key, Pose3> v_;
};

struct ISAM2Params {
  double relinearizeThreshold = 0.1;
  int relinearizeSkip = 10;
};
class ISAM2 {
 public:
  explicit ISAM2(const ISAM2Params &) {}
  void update(const NonlinearFactorGraph &, const Values &init) { for (const auto &kv : init.all()) est_.insert(kv.first, kv.second); }
  void update(const NonlinearFactorGraph &) {}
  void update() {}
  Values calculateEstimate() const { return est_; }

 private:
  Values est_;
};
}  // namespace gtsam
#endif
