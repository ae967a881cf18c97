// What genmsg would generate from /root/reference/msg/cloud_info.msg:1-12 (field names and element types), without serialisation.
// TEST INFRASTRUCTURE (oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_CLOUD_INFO_H
#define ALEGO_REF_SHIM_CLOUD_INFO_H
#include <vector>
#include <ros/ros.h>
namespace alego {
struct cloud_info {
  std_msgs::Header header;
  std::vector<int32_t> startRingIndex;
  std::vector<int32_t> endRingIndex;
  float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
  std::vector<uint8_t> segmentedCloudGroundFlag;  // bool[] is uint8 in roscpp
  std::vector<int32_t> segmentedCloudColInd;
  std::vector<float> segmentedCloudRange;
};
typedef std::shared_ptr<cloud_info> cloud_infoPtr;
typedef std::shared_ptr<const cloud_info> cloud_infoConstPtr;
}  // namespace alego
#endif
