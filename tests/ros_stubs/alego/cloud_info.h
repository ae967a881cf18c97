// stand-in for the header catkin generates from a-lego-loam_b200/ros/msg/cloud_info.msg
#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace alego {
struct cloud_info {
  std_msgs::Header header;
  std::vector<int32_t> startRingIndex, endRingIndex;
  float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
  std::vector<uint8_t> segmentedCloudGroundFlag;
  std::vector<int32_t> segmentedCloudColInd;
  std::vector<float> segmentedCloudRange;
};
typedef std::shared_ptr<cloud_info> cloud_infoPtr;
typedef std::shared_ptr<const cloud_info> cloud_infoConstPtr;
}  // namespace alego
