// Stand-in for pcl_conversions (TEST INFRASTRUCTURE, oracle/_ref build only): the shim PointCloud2 carries decoded x, y, z,
// intensity floats, so fromROSMsg / toROSMsg are plain copies (what the real (de)serialisation amounts to for PointXYZI).
#ifndef ALEGO_REF_SHIM_PCL_CONVERSIONS_H
#define ALEGO_REF_SHIM_PCL_CONVERSIONS_H
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <typename PointT>
void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<PointT> &cloud) {
  const std::size_t n = msg.xyzi.size() / 4;
  cloud.points.resize(n);
  for (std::size_t i = 0; i < n; ++i) {
    PointT p;
    p.x = msg.xyzi[4 * i]; p.y = msg.xyzi[4 * i + 1]; p.z = msg.xyzi[4 * i + 2]; p.intensity = msg.xyzi[4 * i + 3];
    cloud.points[i] = p;
  }
  cloud.width = static_cast<uint32_t>(n);
  cloud.height = 1;
  cloud.is_dense = msg.is_dense;
}
template <typename PointT>
void toROSMsg(const PointCloud<PointT> &cloud, sensor_msgs::PointCloud2 &msg) {
  const std::size_t n = cloud.points.size();
  msg.xyzi.resize(4 * n);
  for (std::size_t i = 0; i < n; ++i) {
    msg.xyzi[4 * i] = cloud.points[i].x; msg.xyzi[4 * i + 1] = cloud.points[i].y;
    msg.xyzi[4 * i + 2] = cloud.points[i].z; msg.xyzi[4 * i + 3] = cloud.points[i].intensity;
  }
  msg.width = static_cast<uint32_t>(n);
  msg.height = 1;
  msg.is_dense = cloud.is_dense;
}
}  // namespace pcl
#endif
