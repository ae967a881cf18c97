// Stand-in nav_msgs/Odometry (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_NAV_ODOMETRY_H
#define ALEGO_REF_SHIM_NAV_ODOMETRY_H
#include <geometry_msgs/PoseWithCovarianceStamped.h>
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
  geometry_msgs::TwistWithCovariance twist;
};
typedef std::shared_ptr<Odometry> OdometryPtr;
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}  // namespace nav_msgs
#endif
