// stand-in for <nodelet/nodelet.h>
#pragma once
#include <ros/ros.h>
namespace nodelet {
class Nodelet {
 public:
  virtual ~Nodelet() {}
  virtual void onInit() = 0;
  ros::NodeHandle &getNodeHandle() { return nh_; }
  ros::NodeHandle &getPrivateNodeHandle() { return pnh_; }
 private:
  ros::NodeHandle nh_, pnh_;
};
}  // namespace nodelet
#define NODELET_FATAL(...) ROS_FATAL(__VA_ARGS__)
#define NODELET_WARN(...) ROS_WARN(__VA_ARGS__)
#define NODELET_INFO(...) ROS_INFO(__VA_ARGS__)
