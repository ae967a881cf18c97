// Stand-in for pcl::PointCloud (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_PCL_POINT_CLOUD_H
#define ALEGO_REF_SHIM_PCL_POINT_CLOUD_H
#include <cstdint>
#include <memory>
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; height = 0; }
  void push_back(const PointT &p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  PointT &operator[](std::size_t i) { return points[i]; }
  const PointT &operator[](std::size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointCloud &operator+=(const PointCloud &rhs) {  // PCL: appends, is_dense &= rhs.is_dense
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
};
}  // namespace pcl
#endif
