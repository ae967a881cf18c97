// Stand-in for tf (TEST INFRASTRUCTURE, oracle/_ref build only).  Matrix3x3(Quaternion) and getRPY restate tf's LinearMath
// (Matrix3x3::setRotation, getEulerYPR solution 1).
#ifndef ALEGO_REF_SHIM_TF_H
#define ALEGO_REF_SHIM_TF_H
#include <cmath>
#include <string>
#include <geometry_msgs/PoseWithCovarianceStamped.h>
namespace tf {
struct Vector3 {
  double v[3];
  Vector3() : v{0, 0, 0} {}
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
struct Quaternion {
  double x_, y_, z_, w_;
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double w() const { return w_; }
};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
class Matrix3x3 {
 public:
  explicit Matrix3x3(const Quaternion &q) {
    const double d = q.x() * q.x() + q.y() * q.y() + q.z() * q.z() + q.w() * q.w();
    const double s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
    m[1][0] = xy + wz; m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
    m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1.0 - (xx + yy);
  }
  void getRPY(double &roll, double &pitch, double &yaw) const {
    if (std::fabs(m[2][0]) >= 1) {
      yaw = 0;
      const double delta = std::atan2(m[2][1], m[2][2]);
      if (m[2][0] < 0) { pitch = M_PI / 2.0; roll = delta; } else { pitch = -M_PI / 2.0; roll = delta; }
    } else {
      pitch = -std::asin(m[2][0]);
      roll = std::atan2(m[2][1] / std::cos(pitch), m[2][2] / std::cos(pitch));
      yaw = std::atan2(m[1][0] / std::cos(pitch), m[0][0] / std::cos(pitch));
    }
  }

 private:
  double m[3][3];
};
struct Transform {
  Vector3 origin;
  Quaternion rotation;
  void setOrigin(const Vector3 &o) { origin = o; }
  void setRotation(const Quaternion &q) { rotation = q; }
};
inline void poseMsgToTF(const geometry_msgs::Pose &p, Transform &t) {
  t.origin = Vector3(p.position.x, p.position.y, p.position.z);
  t.rotation = Quaternion(p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w);
}
struct StampedTransform : Transform {
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
  StampedTransform(const Transform &t, const ros::Time &s, const std::string &f, const std::string &c)
      : Transform(t), stamp_(s), frame_id_(f), child_frame_id_(c) {}
};
}  // namespace tf
#endif
