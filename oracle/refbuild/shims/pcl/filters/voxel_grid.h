// Stand-in for pcl::VoxelGrid (TEST INFRASTRUCTURE, oracle/_ref build only).  THIRD-PARTY SEMANTICS RESTATED from PCL 1.8-1.10
// voxel_grid.hpp (VoxelGrid<PointT>::applyFilter, downsample_all_data_ = true, min_points_per_voxel_ = 0, no filter field):
// bounding box -> voxel index ijk0 + ijk1*dx + ijk2*dx*dy -> std::sort by index only (so the order INSIDE a voxel is libstdc++'s
// introsort order, as in PCL built against the same libstdc++) -> float centroid of x, y, z, intensity per voxel, in index order.
#ifndef ALEGO_REF_SHIM_PCL_VOXEL_GRID_H
#define ALEGO_REF_SHIM_PCL_VOXEL_GRID_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
class VoxelGrid {
 public:
  void setLeafSize(float lx, float ly, float lz) {
    leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
    for (int a = 0; a < 3; ++a) inv_[a] = 1.0f / leaf_[a];  // inverse_leaf_size_ = Array4f::Ones () / leaf_size_.array ()
  }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { in_ = c; }
  void filter(PointCloud<PointT> &out) {
    out.header = in_->header;
    out.points.clear();
    out.height = 1;
    out.is_dense = true;
    const std::vector<PointT> &P = in_->points;
    if (P.empty()) { out.width = 0; return; }
    const bool dense = in_->is_dense;
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::max(); mx[a] = -std::numeric_limits<float>::max(); }
    for (const PointT &p : P) {  // getMinMax3D
      if (!dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv_[0]) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv_[1]) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv_[2]) + 1;
    if (dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {  // "Leaf size is too small"
      out = *in_;
      return;
    }
    int minb[3], maxb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
      minb[a] = static_cast<int>(std::floor(mn[a] * inv_[a]));
      maxb[a] = static_cast<int>(std::floor(mx[a] * inv_[a]));
      divb[a] = maxb[a] - minb[a] + 1;
    }
    const int mul[3] = {1, divb[0], divb[0] * divb[1]};
    std::vector<Cell> cells;
    cells.reserve(P.size());
    for (unsigned k = 0; k < P.size(); ++k) {
      const PointT &p = P[k];
      if (!dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      const int i0 = static_cast<int>(std::floor(p.x * inv_[0]) - static_cast<float>(minb[0]));
      const int i1 = static_cast<int>(std::floor(p.y * inv_[1]) - static_cast<float>(minb[1]));
      const int i2 = static_cast<int>(std::floor(p.z * inv_[2]) - static_cast<float>(minb[2]));
      cells.push_back(Cell(static_cast<unsigned>(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), k));
    }
    std::sort(cells.begin(), cells.end(), std::less<Cell>());
    std::size_t a = 0;
    while (a < cells.size()) {
      std::size_t b = a + 1;
      while (b < cells.size() && cells[b].idx == cells[a].idx) ++b;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;  // CentroidPoint: AccumulatorXYZ + AccumulatorIntensity
      for (std::size_t t = a; t < b; ++t) {
        const PointT &p = P[cells[t].pt];
        sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
      }
      const float n = static_cast<float>(b - a);
      PointT c;
      c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
      out.points.push_back(c);
      a = b;
    }
    out.width = static_cast<uint32_t>(out.points.size());
  }

 private:
  struct Cell {  // pcl::cloud_point_index_idx
    unsigned idx, pt;
    Cell(unsigned i, unsigned p) : idx(i), pt(p) {}
    bool operator<(const Cell &o) const { return idx < o.idx; }
  };
  float leaf_[3] = {0, 0, 0}, inv_[3] = {0, 0, 0};
  typename PointCloud<PointT>::ConstPtr in_;
};
}  // namespace pcl
#endif
