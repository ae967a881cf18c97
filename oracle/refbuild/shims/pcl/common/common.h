// Stand-in for pcl/common (TEST INFRASTRUCTURE, oracle/_ref build only): copyPointCloud, removeNaNFromPointCloud and
// transformPointCloud(Matrix4f) restated from PCL 1.8 (common/impl/transforms.hpp: x' = m00*x + m01*y + m02*z + m03, left to right,
// in float; filters/impl/filter.hpp: a cloud flagged is_dense is passed through unfiltered).
#ifndef ALEGO_REF_SHIM_PCL_COMMON_H
#define ALEGO_REF_SHIM_PCL_COMMON_H
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
void copyPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out) { out = in; }

template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, std::vector<int> &index) {
  if (&in != &out) { out.header = in.header; out.points.resize(in.points.size()); }
  index.resize(in.points.size());
  std::size_t j = 0;
  if (in.is_dense) {
    if (&in != &out) out = in;
    for (j = 0; j < out.points.size(); ++j) index[j] = static_cast<int>(j);
  } else {
    for (std::size_t i = 0; i < in.points.size(); ++i) {
      if (!std::isfinite(in.points[i].x) || !std::isfinite(in.points[i].y) || !std::isfinite(in.points[i].z)) continue;
      out.points[j] = in.points[i];
      index[j] = static_cast<int>(i);
      ++j;
    }
    if (j != in.points.size()) { out.points.resize(j); index.resize(j); }
    out.height = 1;
    out.width = static_cast<uint32_t>(j);
    out.is_dense = true;
  }
}

template <typename PointT>
void transformPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, const Eigen::Matrix4f &m) {
  if (&in != &out) out = in;
  for (std::size_t i = 0; i < out.points.size(); ++i) {
    const PointT &s = in.points[i];
    if (!in.is_dense && (!std::isfinite(s.x) || !std::isfinite(s.y) || !std::isfinite(s.z))) continue;
    const float x = s.x, y = s.y, z = s.z;
    out.points[i].x = static_cast<float>(m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3));
    out.points[i].y = static_cast<float>(m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3));
    out.points[i].z = static_cast<float>(m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3));
  }
}
}  // namespace pcl
#endif
