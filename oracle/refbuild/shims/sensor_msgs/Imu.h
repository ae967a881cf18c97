// Stand-in sensor_msgs/Imu (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_SENSOR_IMU_H
#define ALEGO_REF_SHIM_SENSOR_IMU_H
#include <geometry_msgs/PoseWithCovarianceStamped.h>
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Quaternion orientation;
  geometry_msgs::Vector3 angular_velocity;
  geometry_msgs::Vector3 linear_acceleration;
};
typedef std::shared_ptr<Imu> ImuPtr;
typedef std::shared_ptr<const Imu> ImuConstPtr;
}  // namespace sensor_msgs
#endif
