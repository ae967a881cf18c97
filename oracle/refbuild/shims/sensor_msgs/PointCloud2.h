// Stand-in sensor_msgs/PointCloud2 (TEST INFRASTRUCTURE, oracle/_ref build only): the payload is kept as decoded x, y, z, intensity
// floats (4 per point) instead of the serialised byte blob; pcl::fromROSMsg / toROSMsg in pcl_conversions copy it verbatim.
#ifndef ALEGO_REF_SHIM_SENSOR_POINTCLOUD2_H
#define ALEGO_REF_SHIM_SENSOR_POINTCLOUD2_H
#include <vector>
#include <ros/ros.h>
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 1, width = 0;
  bool is_dense = true;
  std::vector<float> xyzi;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}  // namespace sensor_msgs
#endif
