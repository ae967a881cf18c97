// Stand-in for pcl::KdTreeFLANN (TEST INFRASTRUCTURE, oracle/_ref build only).  THIRD-PARTY SEMANTICS RESTATED: exact k-NN /
// radius search, squared L2 accumulated in float in x, y, z order like ::flann::L2_Simple<float>, results ascending by
// (distance, index) — PCL/FLANN's order among exactly equal distances depends on its tree traversal and is not reproduced.
// Tree: median-split bounding kd-tree, leaves of <= 15 points (KDTreeSingleIndexParams(15)); an implementation detail, the
// search is exact either way.
#ifndef ALEGO_REF_SHIM_PCL_KDTREE_FLANN_H
#define ALEGO_REF_SHIM_PCL_KDTREE_FLANN_H
#include <algorithm>
#include <limits>
#include <memory>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
  void setInputCloud(const CloudConstPtr &cloud) {
    cloud_ = cloud;
    const int n = static_cast<int>(cloud->points.size());
    perm_.resize(n);
    for (int i = 0; i < n; ++i) perm_[i] = i;
    nodes_.clear();
    if (n) build(0, n);
  }
  int nearestKSearch(const PointT &q, int k, std::vector<int> &idx, std::vector<float> &dist) const {
    std::vector<Hit> best(k, Hit{std::numeric_limits<float>::infinity(), std::numeric_limits<int>::max()});
    if (!nodes_.empty()) descend(0, q, best.data(), k, std::numeric_limits<float>::infinity());
    idx.resize(k);
    dist.resize(k);
    int found = 0;
    for (int t = 0; t < k; ++t) {
      idx[t] = best[t].i == std::numeric_limits<int>::max() ? -1 : best[t].i;
      dist[t] = best[t].d;
      found += idx[t] >= 0;
    }
    if (found < k) { idx.resize(found); dist.resize(found); }
    return found;
  }
  int radiusSearch(const PointT &q, double radius, std::vector<int> &idx, std::vector<float> &dist, unsigned = 0) const {
    std::vector<Hit> hits;
    const float r2 = static_cast<float>(radius * radius);
    if (cloud_)
      for (int i = 0; i < static_cast<int>(cloud_->points.size()); ++i) {
        const float d = l2(q, cloud_->points[i]);
        if (d < r2) hits.push_back(Hit{d, i});
      }
    std::sort(hits.begin(), hits.end(), before);
    idx.resize(hits.size());
    dist.resize(hits.size());
    for (std::size_t t = 0; t < hits.size(); ++t) { idx[t] = hits[t].i; dist[t] = hits[t].d; }
    return static_cast<int>(hits.size());
  }

 private:
  struct Hit { float d; int i; };
  struct Node { int lo, hi, left, right, dim; float lmax, rmin; };
  static bool before(const Hit &a, const Hit &b) { return a.d != b.d ? a.d < b.d : a.i < b.i; }
  static float l2(const PointT &a, const PointT &b) {
    float r = 0.f, t;
    t = a.x - b.x; r += t * t;
    t = a.y - b.y; r += t * t;
    t = a.z - b.z; r += t * t;
    return r;
  }
  static float at(const PointT &p, int d) { return d == 0 ? p.x : d == 1 ? p.y : p.z; }
  int build(int lo, int hi) {
    const int id = static_cast<int>(nodes_.size());
    nodes_.push_back(Node{lo, hi, -1, -1, -1, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<float>::max(); mx[d] = -mn[d]; }
    for (int t = lo; t < hi; ++t)
      for (int d = 0; d < 3; ++d) {
        const float c = at(cloud_->points[perm_[t]], d);
        if (c < mn[d]) mn[d] = c;
        if (c > mx[d]) mx[d] = c;
      }
    int dim = 0;
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (!(mx[dim] > mn[dim])) return id;
    const int mid = lo + (hi - lo) / 2;
    std::nth_element(perm_.begin() + lo, perm_.begin() + mid, perm_.begin() + hi,
                     [&](int a, int b) { return at(cloud_->points[a], dim) < at(cloud_->points[b], dim); });
    float lmax = -std::numeric_limits<float>::max(), rmin = std::numeric_limits<float>::max();
    for (int t = lo; t < mid; ++t) lmax = std::max(lmax, at(cloud_->points[perm_[t]], dim));
    for (int t = mid; t < hi; ++t) rmin = std::min(rmin, at(cloud_->points[perm_[t]], dim));
    const int l = build(lo, mid), r = build(mid, hi);
    nodes_[id].left = l; nodes_[id].right = r; nodes_[id].dim = dim; nodes_[id].lmax = lmax; nodes_[id].rmin = rmin;
    return id;
  }
  void descend(int node, const PointT &q, Hit *best, int k, float) const {
    const Node &n = nodes_[node];
    if (n.left < 0) {
      for (int t = n.lo; t < n.hi; ++t) {
        const Hit h{l2(q, cloud_->points[perm_[t]]), perm_[t]};
        if (before(h, best[k - 1])) {
          int pos = k - 1;
          while (pos > 0 && before(h, best[pos - 1])) { best[pos] = best[pos - 1]; --pos; }
          best[pos] = h;
        }
      }
      return;
    }
    const float c = at(q, n.dim);
    const double gl = c > n.lmax ? static_cast<double>(c) - n.lmax : 0.0;  // gap to the left / right child's slab
    const double gr = c < n.rmin ? static_cast<double>(n.rmin) - c : 0.0;
    const bool left_first = gl <= gr;
    descend(left_first ? n.left : n.right, q, best, k, 0.f);
    const double g = left_first ? gr : gl;
    if (g * g * (1.0 - 1e-6) <= static_cast<double>(best[k - 1].d)) descend(left_first ? n.right : n.left, q, best, k, 0.f);
  }
  CloudConstPtr cloud_;
  std::vector<int> perm_;
  std::vector<Node> nodes_;
};
}  // namespace pcl
#endif
