#pragma once
#include <memory>
#include <sensor_msgs/Imu.h>
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
};
typedef std::shared_ptr<Odometry> OdometryPtr;
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}  // namespace nav_msgs
