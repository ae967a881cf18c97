// Stand-in geometry_msgs (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_GEOMETRY_MSGS_H
#define ALEGO_REF_SHIM_GEOMETRY_MSGS_H
#include <ros/ros.h>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36] = {0}; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseWithCovarianceStamped { std_msgs::Header header; PoseWithCovariance pose; };
}  // namespace geometry_msgs
#endif
