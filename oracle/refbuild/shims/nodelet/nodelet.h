// Stand-in for nodelet::Nodelet (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_NODELET_H
#define ALEGO_REF_SHIM_NODELET_H
#include <ros/ros.h>
namespace nodelet {
class Nodelet {
 public:
  virtual ~Nodelet() {}
  virtual void onInit() = 0;

 protected:
  ros::NodeHandle &getMTNodeHandle() { return nh_; }
  ros::NodeHandle &getMTPrivateNodeHandle() { return nh_; }
  ros::NodeHandle &getNodeHandle() { return nh_; }
  ros::NodeHandle &getPrivateNodeHandle() { return nh_; }

 private:
  ros::NodeHandle nh_;
};
}  // namespace nodelet
#endif
