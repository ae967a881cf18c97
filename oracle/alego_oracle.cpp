// =====================================================================================================
// alego_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++17, zero dependencies) of A-LeGO-LOAM's per-scan numeric hot path, used only
// as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Nothing under a-lego-loam_b200/ may include, link or call this file.
//
// PARITY UNPINNED: the reference has no tests, fixtures or golden vectors and cannot
// be compiled here (ROS / PCL / FLANN / Eigen / Ceres / GTSAM are absent).  This file is
// therefore a "port" oracle: each function follows the cited reference lines of the NODELET twin
// (src/imageProjection.cpp, src/laserOdometry.cpp, src/laserMapping.cpp, include/alego/utility.h) and
// restates the published algorithms of the absent third-party pieces (PCL VoxelGrid, FLANN exact k-NN with
// L2_Simple<float>, Ceres trust-region Levenberg-Marquardt with DENSE_QR + HuberLoss, Eigen 3x3 symmetric
// eigen-solve and column-pivoted Householder QR).  Independent cross-checks (scipy cKDTree, numpy eigh /
// lstsq, scipy least_squares, finite differences) live in tests/test_oracle_*.py.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off (no -march=native, no -ffast-math) — mirrors the reference's
// Release / -std=c++11 / baseline x86-64 flags (CMakeLists.txt:4-5), i.e. no FMA contraction.
// =====================================================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../include/alego_b200.h"

namespace {

using std::vector;

struct P4 {
  float x, y, z, i;
};

static inline double rad2deg(double x) { return x * 180.0 / M_PI; }  // RAD2ANGLE utility.h:47
static inline double deg2rad(double x) { return x / 180.0 * M_PI; }  // ANGLE2RAD utility.h:48

struct Stopwatch {  // TicToc utility.h:99-120 (steady clock instead of system clock)
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// -----------------------------------------------------------------------------------------------------
// pcl::VoxelGrid<pcl::PointXYZI>::applyFilter restated (PCL 1.8-1.10 voxel_grid.hpp; call sites
// laserOdometry.cpp:288-293, laserMapping.cpp:37-41,325-342).  downsample_all_data_=true (intensity is
// averaged too), min_points_per_voxel_=0, no filter field.  `stable_ties`=false sorts exactly like PCL
// (std::sort, key-only operator<, so the order of points INSIDE a voxel — and therefore the last bits of
// the float centroid — is whatever libstdc++'s introsort leaves); true uses (key, input index) order.
// -----------------------------------------------------------------------------------------------------
struct VoxKey {
  unsigned idx;
  unsigned pt;
  bool operator<(const VoxKey &o) const { return idx < o.idx; }
};

static void voxel_grid(const vector<P4> &in, float leaf, vector<P4> &out, bool stable_ties, vector<unsigned> *keys_out = nullptr) {
  out.clear();
  if (keys_out) keys_out->clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (const P4 &p : in) {  // getMinMax3D
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    out = in;  // "Leaf size is too small for the input dataset" → output = input
    return;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  vector<VoxKey> iv;
  iv.reserve(in.size());
  for (unsigned k = 0; k < in.size(); ++k) {
    const P4 &p = in[k];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int i0 = static_cast<int>(std::floor(p.x * inv) - static_cast<float>(min_b[0]));
    int i1 = static_cast<int>(std::floor(p.y * inv) - static_cast<float>(min_b[1]));
    int i2 = static_cast<int>(std::floor(p.z * inv) - static_cast<float>(min_b[2]));
    int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
    iv.push_back({static_cast<unsigned>(idx), k});
  }
  if (stable_ties)
    std::sort(iv.begin(), iv.end(), [](const VoxKey &a, const VoxKey &b) { return a.idx != b.idx ? a.idx < b.idx : a.pt < b.pt; });
  else
    std::sort(iv.begin(), iv.end(), std::less<VoxKey>());
  size_t a = 0;
  while (a < iv.size()) {
    size_t b = a + 1;
    while (b < iv.size() && iv[b].idx == iv[a].idx) ++b;
    // CentroidPoint<PointXYZI>: float accumulators for xyz and intensity, divided by the count
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t k = a; k < b; ++k) {
      const P4 &p = in[iv[k].pt];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
    }
    const float n = static_cast<float>(b - a);
    out.push_back({sx / n, sy / n, sz / n, si / n});
    if (keys_out) keys_out->push_back(iv[a].idx);
    a = b;
  }
}

// -----------------------------------------------------------------------------------------------------
// Exact k-NN, squared L2 accumulated in float like ::flann::L2_Simple<float> (diff=a-b; result+=diff*diff
// over x,y,z), which is what pcl::KdTreeFLANN<PointXYZI>::nearestKSearch returns (call sites
// laserOdometry.cpp:341,431 k=1; laserMapping.cpp:375,423 k=5).  A bounding-box kd-tree (leaf 15, like
// PCL's KDTreeSingleIndexParams(15)) with exact search; result order = ascending (dist, index).
// -----------------------------------------------------------------------------------------------------
static inline float l2f(const P4 &q, const P4 &p) {
  float r = 0.f, d;
  d = q.x - p.x; r += d * d;
  d = q.y - p.y; r += d * d;
  d = q.z - p.z; r += d * d;
  return r;
}

struct KdTree {
  struct Node {
    int lo, hi, left, right, dim;
    float split_lo, split_hi;  // max of left side / min of right side along dim
  };
  const vector<P4> *pts = nullptr;
  vector<int> order;
  vector<Node> nodes;
  void build(const vector<P4> &p) {
    pts = &p;
    order.resize(p.size());
    for (size_t k = 0; k < p.size(); ++k) order[k] = (int)k;
    nodes.clear();
    nodes.reserve(p.size() / 4 + 4);
    if (!p.empty()) build_rec(0, (int)p.size());
  }
  static float coord(const P4 &p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }
  int build_rec(int lo, int hi) {
    int id = (int)nodes.size();
    nodes.push_back({lo, hi, -1, -1, -1, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int k = lo; k < hi; ++k) {
      const P4 &p = (*pts)[order[k]];
      for (int d = 0; d < 3; ++d) {
        float c = coord(p, d);
        mn[d] = std::min(mn[d], c);
        mx[d] = std::max(mx[d], c);
      }
    }
    int dim = 0;
    for (int d = 1; d < 3; ++d)
      if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (!(mx[dim] > mn[dim])) return id;  // all points identical → leaf
    int mid = (lo + hi) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                     [&](int a, int b) { return coord((*pts)[a], dim) < coord((*pts)[b], dim); });
    float slo = -1e30f, shi = 1e30f;
    for (int k = lo; k < mid; ++k) slo = std::max(slo, coord((*pts)[order[k]], dim));
    for (int k = mid; k < hi; ++k) shi = std::min(shi, coord((*pts)[order[k]], dim));
    int l = build_rec(lo, mid);
    int r = build_rec(mid, hi);
    nodes[id].left = l; nodes[id].right = r; nodes[id].dim = dim; nodes[id].split_lo = slo; nodes[id].split_hi = shi;
    return id;
  }
  struct Cand {
    float d;
    int idx;
  };
  static bool closer(const Cand &a, const Cand &b) { return a.d != b.d ? a.d < b.d : a.idx < b.idx; }
  // best[] kept sorted ascending, size k (filled with +inf / INT_MAX)
  void search(int node, const P4 &q, Cand *best, int k) const {
    const Node &n = nodes[node];
    if (n.left < 0) {
      for (int t = n.lo; t < n.hi; ++t) {
        Cand c{l2f(q, (*pts)[order[t]]), order[t]};
        if (closer(c, best[k - 1])) {
          int pos = k - 1;
          while (pos > 0 && closer(c, best[pos - 1])) { best[pos] = best[pos - 1]; --pos; }
          best[pos] = c;
        }
      }
      return;
    }
    float qc = coord(q, n.dim);
    // distance (double, conservative) from q to each child's slab along the split dimension
    double dl = qc > n.split_lo ? (double)qc - (double)n.split_lo : 0.0;
    double dr = qc < n.split_hi ? (double)n.split_hi - (double)qc : 0.0;
    int first = dl <= dr ? n.left : n.right, second = dl <= dr ? n.right : n.left;
    double dsecond = dl <= dr ? dr : dl;
    search(first, q, best, k);
    // prune only when the slab alone is already strictly farther than the current k-th best (margin for float rounding)
    if (dsecond * dsecond * (1.0 - 1e-6) <= (double)best[k - 1].d) search(second, q, best, k);
  }
  // returns number found (<k if fewer points)
  int knn(const P4 &q, int k, int *idx, float *dist) const {
    Cand best[8];
    for (int t = 0; t < k; ++t) best[t] = {std::numeric_limits<float>::infinity(), std::numeric_limits<int>::max()};
    if (!nodes.empty()) search(0, q, best, k);
    int found = 0;
    for (int t = 0; t < k; ++t) {
      idx[t] = best[t].idx == std::numeric_limits<int>::max() ? -1 : best[t].idx;
      dist[t] = best[t].d;
      if (idx[t] >= 0) ++found;
    }
    return found;
  }
};

// -----------------------------------------------------------------------------------------------------
// Cost functions (utility.h:122-349) as one tagged residual.  kind: 0 CornerCostFunction (:122-179),
// 1 SurfCostFunction (:181-240), 2 LidarEdgeCostFunction (:242-299), 3 LidarPlaneCostFunction (:301-349).
// -----------------------------------------------------------------------------------------------------
struct Resid {
  int kind;
  double cp[3];  // current point (sensor frame)
  double a[3];   // lpj  | plane unit normal
  double b[3];   // lpl
  double c[3];   // lpm (surf only)
  double d;      // negative_OA_dot_norm (plane only)
};

struct PoseTrig {
  double sr, cr, sp, cp, sy, cy;
  double R[9];
  explicit PoseTrig(const double *x) {
    sr = std::sin(x[3]); cr = std::cos(x[3]);
    sp = std::sin(x[4]); cp = std::cos(x[4]);
    sy = std::sin(x[5]); cy = std::cos(x[5]);
    // Rz(yaw) * Ry(pitch) * Rx(roll)   (AngleAxisd products, utility.h:128 etc.)
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
  }
};

// residual r and (optionally) the 6-vector Jacobian exactly as the reference's Evaluate() bodies fill them
static void eval_resid(const Resid &f, const double *x, const PoseTrig &T, double *r, double *J) {
  const double px = f.cp[0], py = f.cp[1], pz = f.cp[2];
  const double lx = T.R[0] * px + T.R[1] * py + T.R[2] * pz + x[0];
  const double ly = T.R[3] * px + T.R[4] * py + T.R[5] * pz + x[1];
  const double lz = T.R[6] * px + T.R[7] * py + T.R[8] * pz + x[2];
  // analytic d(lp)/d(roll,pitch,yaw) (utility.h:148-158 and its three clones).  D[row=x,y,z][col=r,p,y].
  // NOTE the reference's dy_dp ends in cr*sr*cp*cp_.z (utility.h:153,217,273,325); the true term is
  // sy*cp*cr*cp_.z.  Replicated on purpose.
  double D[3][3];
  const double sr = T.sr, cr = T.cr, sp = T.sp, cp = T.cp, sy = T.sy, cy = T.cy;
  if (J) {
    D[0][0] = (cy * sp * cr + sr * sy) * py + (sy * cr - cy * sr * sp) * pz;
    D[1][0] = (-cy * sr + sy * sp * cr) * py + (-sr * sy * sp - cy * cr) * pz;
    D[2][0] = cp * cr * py - cp * sr * pz;
    D[0][1] = -cy * sp * px + cy * cp * sr * py + cy * cr * cp * pz;
    D[1][1] = -sp * sy * px + sy * cp * sr * py + cr * sr * cp * pz;  // <- reference quirk
    D[2][1] = -cp * px - sp * sr * py - sp * cr * pz;
    D[0][2] = -sy * cp * px - (sy * sp * sr + cr * cy) * py + (cy * sr - sy * cr * sp) * pz;
    D[1][2] = cp * cy * px + (-sy * cr + cy * sp * sr) * py + (cy * cr * sp + sy * sr) * pz;
    D[2][2] = 0.;
  }
  if (f.kind == 0 || f.kind == 2) {
    const double *j = f.a, *l = f.b;
    const double k = std::sqrt(std::pow(j[0] - l[0], 2) + std::pow(j[1] - l[1], 2) + std::pow(j[2] - l[2], 2));
    const double a = (ly - j[1]) * (lz - l[2]) - (lz - j[2]) * (ly - l[1]);
    const double b = (lz - j[2]) * (lx - l[0]) - (lx - j[0]) * (lz - l[2]);
    const double c = (lx - j[0]) * (ly - l[1]) - (ly - j[1]) * (lx - l[0]);
    const double m = std::sqrt(a * a + b * b + c * c);
    *r = m / k;
    if (J) {
      const double gx = (b * (l[2] - j[2]) + c * (j[1] - l[1])) / m;
      const double gy = (a * (j[2] - l[2]) - c * (j[0] - l[0])) / m;
      const double gz = (-a * (j[1] - l[1]) + b * (j[0] - l[0])) / m;
      if (f.kind == 0) {  // x, y, yaw only (utility.h:162-167)
        J[0] = gx / k; J[1] = gy / k; J[2] = 0.; J[3] = 0.; J[4] = 0.;
        J[5] = (gx * D[0][2] + gy * D[1][2] + gz * D[2][2]) / k;
      } else {  // full (utility.h:282-287)
        J[0] = gx / k; J[1] = gy / k; J[2] = gz / k;
        J[3] = (gx * D[0][0] + gy * D[1][0] + gz * D[2][0]) / k;
        J[4] = (gx * D[0][1] + gy * D[1][1] + gz * D[2][1]) / k;
        J[5] = (gx * D[0][2] + gy * D[1][2] + gz * D[2][2]) / k;
      }
    }
  } else if (f.kind == 1) {
    const double *j = f.a, *l = f.b, *mm = f.c;
    double a = (j[1] - l[1]) * (j[2] - mm[2]) - (j[2] - l[2]) * (j[1] - mm[1]);
    double b = (j[2] - l[2]) * (j[0] - mm[0]) - (j[0] - l[0]) * (j[2] - mm[2]);
    double c = (j[0] - l[0]) * (j[1] - mm[1]) - (j[1] - l[1]) * (j[0] - mm[0]);
    a *= a; b *= b; c *= c;  // component-wise squares (utility.h:191-193)
    const double m = std::sqrt(std::pow(lx - j[0], 2) * a + std::pow(ly - j[1], 2) * b + std::pow(lz - j[2], 2) * c);
    const double k = std::sqrt(a + b + c);
    *r = m / k;
    if (J) {
      const double tmp = m * k;
      const double gz = ((lz - j[2]) * c) / tmp;  // note the extra 1/k (utility.h:199-203)
      J[0] = 0.; J[1] = 0.; J[2] = gz / k; J[3] = 0.; J[4] = 0.; J[5] = 0.;  // z only (utility.h:226-231)
    }
  } else {
    const double *n = f.a;
    *r = n[0] * lx + n[1] * ly + n[2] * lz + f.d;
    if (J) {
      J[0] = n[0]; J[1] = n[1]; J[2] = n[2];
      J[3] = n[0] * D[0][0] + n[1] * D[1][0] + n[2] * D[2][0];
      J[4] = n[0] * D[0][1] + n[1] * D[1][1] + n[2] * D[2][1];
      J[5] = n[0] * D[0][2] + n[1] * D[1][2] + n[2] * D[2][2];
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// ceres::Solve restated for one 6-parameter block, trust-region Levenberg-Marquardt, DENSE_QR, HuberLoss
// (call sites laserOdometry.cpp:413-418,487-492; laserMapping.cpp:468-475).  Defaults of Ceres 1.13/1.14:
// initial radius 1e4, max 1e16, min 1e-32, min/max LM diagonal 1e-6/1e32, min_relative_decrease 1e-3,
// function/gradient/parameter tolerance 1e-6/1e-10/1e-8, jacobi scaling from the initial Jacobian,
// monotonic steps, max 5 consecutive invalid steps (trust_region_minimizer.cc,
// levenberg_marquardt_strategy.cc, corrector.cc, loss_function.cc — recalled, see header caveat).
// -----------------------------------------------------------------------------------------------------
struct SolveSummary {
  int iterations = 0;  // accepted + rejected step attempts
  int successful = 0;
  double initial_cost = 0, final_cost = 0;
  int termination = 0;  // 0 max iterations, 1 parameter tol, 2 function tol, 3 gradient tol, 4 failure
};

struct Evaluated {
  double cost = 0;
  vector<double> r;  // corrected residuals
  vector<double> J;  // corrected Jacobian, row-major n x 6 (unscaled)
  double g[6];       // gradient J^T r
};

static double evaluate(const vector<Resid> &rs, const double *x, double huber_a, Evaluated *out) {
  const PoseTrig T(x);
  const size_t n = rs.size();
  double cost = 0;
  if (out) {
    out->r.resize(n);
    out->J.resize(n * 6);
    for (double &g : out->g) g = 0;
  }
  const double b = huber_a * huber_a;
  for (size_t k = 0; k < n; ++k) {
    double r, J[6];
    eval_resid(rs[k], x, T, &r, out ? J : nullptr);
    const double s = r * r;
    double rho0, rho1;
    if (s > b) {  // HuberLoss::Evaluate outlier region
      const double rr = std::sqrt(s);
      rho0 = 2.0 * huber_a * rr - b;
      rho1 = std::max(std::numeric_limits<double>::min(), huber_a / rr);
    } else {
      rho0 = s;
      rho1 = 1.0;
    }
    cost += 0.5 * rho0;
    if (out) {  // Corrector with rho'' <= 0: scale residual and Jacobian by sqrt(rho')
      const double w = std::sqrt(rho1);
      out->r[k] = w * r;
      for (int c = 0; c < 6; ++c) {
        out->J[k * 6 + c] = w * J[c];
        out->g[c] += out->J[k * 6 + c] * out->r[k];
      }
    }
  }
  if (out) out->cost = cost;
  return cost;
}

// min || [A; diag(D)] s - [b; 0] ||  by Householder QR of the stacked (n+6) x 6 matrix (DenseQRSolver)
static bool dense_qr_solve(const vector<double> &A, const vector<double> &bvec, const double *D, size_t n, double *s) {
  const size_t m = n + 6;
  vector<double> M(m * 6), rhs(m, 0.0);
  for (size_t i = 0; i < n; ++i) {
    for (int c = 0; c < 6; ++c) M[i * 6 + c] = A[i * 6 + c];
    rhs[i] = bvec[i];
  }
  for (int c = 0; c < 6; ++c) {
    for (int c2 = 0; c2 < 6; ++c2) M[(n + c) * 6 + c2] = 0.0;
    M[(n + c) * 6 + c] = D[c];
  }
  for (int c = 0; c < 6; ++c) {
    double nrm = 0;
    for (size_t i = c; i < m; ++i) nrm += M[i * 6 + c] * M[i * 6 + c];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = M[c * 6 + c] > 0 ? -nrm : nrm;
    vector<double> v(m - c);
    for (size_t i = c; i < m; ++i) v[i - c] = M[i * 6 + c];
    v[0] -= alpha;
    double vn = 0;
    for (double t : v) vn += t * t;
    if (vn == 0.0) continue;
    for (int c2 = c; c2 < 6; ++c2) {
      double dot = 0;
      for (size_t i = c; i < m; ++i) dot += v[i - c] * M[i * 6 + c2];
      const double f = 2.0 * dot / vn;
      for (size_t i = c; i < m; ++i) M[i * 6 + c2] -= f * v[i - c];
    }
    double dot = 0;
    for (size_t i = c; i < m; ++i) dot += v[i - c] * rhs[i];
    const double f = 2.0 * dot / vn;
    for (size_t i = c; i < m; ++i) rhs[i] -= f * v[i - c];
  }
  for (int c = 5; c >= 0; --c) {
    double acc = rhs[c];
    for (int c2 = c + 1; c2 < 6; ++c2) acc -= M[c * 6 + c2] * s[c2];
    if (M[c * 6 + c] == 0.0) return false;
    s[c] = acc / M[c * 6 + c];
  }
  for (int c = 0; c < 6; ++c)
    if (!std::isfinite(s[c])) return false;
  return true;
}

static void ceres_like_solve(const vector<Resid> &rs, double *x, int max_iters, double huber_a, SolveSummary *sum,
                             vector<double> *trace /* per attempt: cost, x[6] */) {
  const size_t n = rs.size();
  Evaluated E;
  evaluate(rs, x, huber_a, &E);
  double cost = E.cost;
  sum->initial_cost = cost;
  double scale[6];
  for (int c = 0; c < 6; ++c) {
    double q = 0;
    for (size_t i = 0; i < n; ++i) q += E.J[i * 6 + c] * E.J[i * 6 + c];
    scale[c] = 1.0 / (1.0 + std::sqrt(q));
  }
  auto scale_cols = [&](Evaluated &e) {
    for (size_t i = 0; i < n; ++i)
      for (int c = 0; c < 6; ++c) e.J[i * 6 + c] *= scale[c];
  };
  scale_cols(E);
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double diag[6];
  int iter = 0, invalid_run = 0;
  double x_norm = 0;
  for (int c = 0; c < 6; ++c) x_norm += x[c] * x[c];
  x_norm = std::sqrt(x_norm);
  sum->termination = 0;
  if (trace) { trace->push_back(cost); for (int c = 0; c < 6; ++c) trace->push_back(x[c]); }
  while (true) {
    if (iter >= max_iters) { sum->termination = 0; break; }
    if (radius <= 1e-32) { sum->termination = 4; break; }
    ++iter;
    if (!reuse_diagonal) {
      for (int c = 0; c < 6; ++c) {
        double q = 0;
        for (size_t i = 0; i < n; ++i) q += E.J[i * 6 + c] * E.J[i * 6 + c];
        diag[c] = std::min(std::max(q, 1e-6), 1e32);
      }
    }
    double lm[6], step[6];
    for (int c = 0; c < 6; ++c) lm[c] = std::sqrt(diag[c] / radius);
    bool ok = dense_qr_solve(E.J, E.r, lm, n, step);
    reuse_diagonal = true;
    double model_change = 0;
    if (ok) {
      for (int c = 0; c < 6; ++c) step[c] = -step[c];
      for (size_t i = 0; i < n; ++i) {
        double mr = 0;
        for (int c = 0; c < 6; ++c) mr += E.J[i * 6 + c] * step[c];
        model_change -= mr * (E.r[i] + mr / 2.0);
      }
    }
    if (!ok || !(model_change > 0.0)) {  // invalid step
      if (++invalid_run >= 5) { sum->termination = 4; break; }
      radius *= 0.5;
      if (trace) { trace->push_back(cost); for (int c = 0; c < 6; ++c) trace->push_back(x[c]); }
      continue;
    }
    invalid_run = 0;
    double delta[6], xc[6], step_norm = 0;
    for (int c = 0; c < 6; ++c) {
      delta[c] = step[c] * scale[c];
      xc[c] = x[c] + delta[c];
      step_norm += (x[c] - xc[c]) * (x[c] - xc[c]);
    }
    step_norm = std::sqrt(step_norm);
    const double cand_cost = evaluate(rs, xc, huber_a, nullptr);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { sum->termination = 1; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= 1e-6 * cost) { sum->termination = 2; break; }
    const double rho = cost_change / model_change;
    if (rho > 1e-3) {
      for (int c = 0; c < 6; ++c) x[c] = xc[c];
      x_norm = 0;
      for (int c = 0; c < 6; ++c) x_norm += x[c] * x[c];
      x_norm = std::sqrt(x_norm);
      evaluate(rs, x, huber_a, &E);
      cost = E.cost;
      scale_cols(E);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      ++sum->successful;
      if (trace) { trace->push_back(cost); for (int c = 0; c < 6; ++c) trace->push_back(x[c]); }
      double gmax = 0;
      for (int c = 0; c < 6; ++c) gmax = std::max(gmax, std::fabs(E.g[c]));
      if (gmax <= 1e-10) { sum->termination = 3; break; }
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      if (trace) { trace->push_back(cost); for (int c = 0; c < 6; ++c) trace->push_back(x[c]); }
    }
  }
  sum->iterations = iter;
  sum->final_cost = cost;
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi); eigenvalues ascending like Eigen's
// SelfAdjointEigenSolver (laserMapping.cpp:397-403).  V columns = eigenvectors.
static void eig3_sym(const double A[9], double w[3], double V[9]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double dsum = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-40 * dsum || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
  for (int k = 0; k < 3; ++k) {
    w[k] = a[ord[k]][ord[k]];
    for (int r = 0; r < 3; ++r) V[r * 3 + k] = v[r][ord[k]];
  }
}

// least-squares solve of the 5x3 system A n = b by column-pivoted Householder QR
// (matA0.colPivHouseholderQr().solve(matB0), laserMapping.cpp:435)
static void lstsq_5x3(const double A_in[15], const double b_in[5], double n[3]) {
  double A[5][3], b[5];
  for (int i = 0; i < 5; ++i) { for (int c = 0; c < 3; ++c) A[i][c] = A_in[i * 3 + c]; b[i] = b_in[i]; }
  int perm[3] = {0, 1, 2};
  int rank = 3;
  for (int c = 0; c < 3; ++c) {
    int best = c; double bn = -1;
    for (int c2 = c; c2 < 3; ++c2) {
      double q = 0;
      for (int i = c; i < 5; ++i) q += A[i][c2] * A[i][c2];
      if (q > bn) { bn = q; best = c2; }
    }
    if (best != c) { for (int i = 0; i < 5; ++i) std::swap(A[i][c], A[i][best]); std::swap(perm[c], perm[best]); }
    double nrm = std::sqrt(bn);
    if (nrm < 1e-300) { rank = c; break; }
    double alpha = A[c][c] > 0 ? -nrm : nrm;
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = c; i < 5; ++i) v[i] = A[i][c];
    v[c] -= alpha;
    double vn = 0;
    for (int i = c; i < 5; ++i) vn += v[i] * v[i];
    if (vn > 0) {
      for (int c2 = c; c2 < 3; ++c2) {
        double dot = 0;
        for (int i = c; i < 5; ++i) dot += v[i] * A[i][c2];
        double f = 2.0 * dot / vn;
        for (int i = c; i < 5; ++i) A[i][c2] -= f * v[i];
      }
      double dot = 0;
      for (int i = c; i < 5; ++i) dot += v[i] * b[i];
      double f = 2.0 * dot / vn;
      for (int i = c; i < 5; ++i) b[i] -= f * v[i];
    }
  }
  double y[3] = {0, 0, 0};
  for (int c = rank - 1; c >= 0; --c) {
    double acc = b[c];
    for (int c2 = c + 1; c2 < rank; ++c2) acc -= A[c][c2] * y[c2];
    y[c] = acc / A[c][c];
  }
  for (int c = 0; c < 3; ++c) n[perm[c]] = y[c];
}

static inline void mat3_mul(const double *A, const double *B, double *C) {
  double t[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
  std::memcpy(C, t, sizeof t);
}
static inline void mat3_vec(const double *A, const double *v, double *o) {
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = A[r * 3] * v[0] + A[r * 3 + 1] * v[1] + A[r * 3 + 2] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}

// =====================================================================================================
struct Oracle {
  AlegoParams P;
  int R, C;
  double seg_alpha_x, seg_alpha_y;

  // ---- ImageProjection state (imageProjection.h:19-27)
  vector<P4> full_cloud;
  vector<double> range_mat;  // row-major R x C (reference: Eigen col-major; layout is not observable)
  vector<int32_t> label_mat;
  vector<uint8_t> ground_mat;
  vector<int32_t> startRing, endRing;
  float startOri = 0, endOri = 0, oriDiff = 0;
  vector<uint8_t> segGround;
  vector<int32_t> segCol;
  vector<float> segRange;
  vector<P4> seg_cloud, outlier_cloud;
  double min_margin_row = 0, min_margin_col = 0;  // audit: distance (in cells) of any point to a rounding boundary
  int n_dup_cells = 0;

  // ---- LaserOdometry feature state (laserOdometry.h:55-58)
  vector<double> curvature;
  vector<uint8_t> picked;
  vector<int32_t> cloud_label, sort_idx;
  vector<int32_t> sharp_idx, less_sharp_idx, flat_idx;
  vector<P4> sharp, less_sharp, flat, less_flat;
  vector<P4> less_flat_stable;  // same but with (key,index)-ordered voxel sums
  vector<int32_t> less_flat_scan_idx;  // indices pushed to less_flat_scan, all rings
  int n_tie_segments = 0;        // segments in which two points share a curvature value
  int tie_sensitive = 0;         // 1 if a (curv,index) tie-break order would change any feature list

  // ---- LaserOdometry scan-to-scan state (laserOdometry.h:60-80)
  bool lo_init = false;
  vector<P4> surf_last, corner_last;
  KdTree kd_surf_last, kd_corner_last;
  double lo_params[6] = {0, 0, 0, 0, 0, 0};
  double t_w[3] = {0, 0, 0};
  double r_w[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  vector<Resid> lo_resids;
  vector<int32_t> lo_surf_corr, lo_corner_corr;  // (j, closest, idx2, idx3) / (j, closest, idx2)
  AlegoSolveReport lo_report{};
  vector<double> lo_trace;

  // ---- LaserMapping state (laserMapping.h:140-177)
  vector<P4> map_corner, map_surf;
  KdTree kd_map_corner, kd_map_surf;
  vector<P4> lm_corner, lm_surf, lm_outlier;  // laser_corner_, laser_surf_, laser_outlier_
  vector<P4> lm_corner_ds, lm_surf_ds, lm_outlier_ds, lm_surf_total, lm_surf_total_ds;
  double lm_params[6] = {0, 0, 0, 0, 0, 0};
  double t_m2o[3] = {0, 0, 0}, r_m2o[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double t_o2l[3] = {0, 0, 0}, r_o2l[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double t_m2l[3] = {0, 0, 0}, r_m2l[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  vector<Resid> lm_resids;
  vector<int32_t> lm_corner_sel, lm_surf_sel;  // query indices that produced a residual
  AlegoSolveReport lm_report{};
  vector<double> lm_trace;
  bool stable_voxel = false;

  int scan_count = 0;
  int lm_every = 1;
  double t_ip = 0, t_feat = 0, t_s2s = 0, t_lm = 0;  // ms of the last call of each stage

  explicit Oracle(const AlegoParams &p) : P(p), R(p.n_scan), C(p.horizon_scan) {
    seg_alpha_x = deg2rad(P.ang_res_x);  // utility.h:60-61
    seg_alpha_y = deg2rad(P.ang_res_y);
    const size_t n = (size_t)R * C;
    P4 nan_p{0, 0, 0, -1};  // imageProjection.cpp:24-27 (x,y,z of a default PointXYZI are 0)
    full_cloud.assign(n, nan_p);
    range_mat.assign(n, std::numeric_limits<double>::max());
    label_mat.assign(n, 0);
    ground_mat.assign(n, 0);
    startRing.assign(R, 0);
    endRing.assign(R, 0);
    curvature.assign(n, 0);
    picked.assign(n, 0);
    cloud_label.assign(n, 0);
    sort_idx.assign(n, 0);
  }

  // -------------------------------------------------------------------------------------------------
  // ImageProjection::pcCB (imageProjection.cpp:49-208)
  // -------------------------------------------------------------------------------------------------
  int ip(const P4 *pts_in, int n_in) {
    Stopwatch sw;
    vector<P4> in;
    in.reserve(n_in);
    for (int k = 0; k < n_in; ++k)  // pcl::removeNaNFromPointCloud (:59)
      if (std::isfinite(pts_in[k].x) && std::isfinite(pts_in[k].y) && std::isfinite(pts_in[k].z)) in.push_back(pts_in[k]);
    const int n = (int)in.size();
    if (n == 0) return ALEGO_BAD_ARG;
    // reset (the reference resets at the END of pcCB, :197-205; equivalent)
    const P4 nan_p{0, 0, 0, -1};
    std::fill(full_cloud.begin(), full_cloud.end(), nan_p);
    std::fill(range_mat.begin(), range_mat.end(), std::numeric_limits<double>::max());
    std::fill(label_mat.begin(), label_mat.end(), 0);
    std::fill(ground_mat.begin(), ground_mat.end(), 0);
    seg_cloud.clear();
    outlier_cloud.clear();
    segGround.clear(); segCol.clear(); segRange.clear();

    // orientation (:62-72).  atan2(float,float) is the float overload; message fields are float32.
    startOri = -std::atan2(in[0].y, in[0].x);
    endOri = (float)(-std::atan2(in[n - 1].y, in[n - 1].x) + 2 * M_PI);
    if (endOri - startOri > 3 * M_PI) endOri = (float)(endOri - 2 * M_PI);
    else if (endOri - startOri < M_PI) endOri = (float)(endOri + 2 * M_PI);
    oriDiff = endOri - startOri;

    // projection (:74-104)
    min_margin_row = min_margin_col = 0.5;
    n_dup_cells = 0;
    for (int k = 0; k < n; ++k) {
      P4 p = in[k];
      const double vertical_ang = rad2deg(std::atan2(p.z, std::hypot(p.x, p.y)));  // float overloads (Appendix A.1)
      const double row_f = (vertical_ang + P.ang_bottom) / P.ang_res_y + 0.5;
      const int row_id = (int)row_f;
      if (row_id < 0 || row_id >= R) continue;
      const double horizon_ang = rad2deg(-std::atan2(p.y, p.x) + 2 * M_PI);
      const double col_f = horizon_ang / P.ang_res_x;
      int col_id = (int)col_f;
      if (col_id >= C) col_id -= C;
      if (col_id < 0 || col_id >= C) continue;
      min_margin_row = std::min(min_margin_row, std::fabs(row_f - std::floor(row_f + 0.5)));
      min_margin_col = std::min(min_margin_col, std::fabs(col_f - std::floor(col_f + 0.5)));
      const size_t index = (size_t)col_id + (size_t)row_id * C;
      if (full_cloud[index].i != -1) ++n_dup_cells;
      range_mat[index] = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);  // float sqrt of a float sum (:99)
      p.i = (float)(row_id + col_id / 10000.0);                          // (:101)
      full_cloud[index] = p;
    }

    // groundRemoval (:106-132)
    for (int j = 0; j < C; ++j)
      for (int i = 0; i < P.ground_scan_id; ++i) {
        if (i + 1 >= R) break;  // guard for ground_scan_id >= n_scan (reference would read out of bounds)
        const size_t lo = (size_t)j + (size_t)i * C, up = (size_t)j + (size_t)(i + 1) * C;
        if (full_cloud[lo].i == -1 || full_cloud[up].i == -1) continue;
        const double dx = full_cloud[up].x - full_cloud[lo].x;  // float subtraction, widened
        const double dy = full_cloud[up].y - full_cloud[lo].y;
        const double dz = full_cloud[up].z - full_cloud[lo].z;
        const double angle = rad2deg(std::atan2(dz, std::hypot(dx, dy)));
        if (std::abs(angle - P.sensor_mount_ang) < 10.) ground_mat[lo] = ground_mat[up] = 1;
      }
    // label init (:134-143)
    for (size_t c = 0; c < full_cloud.size(); ++c)
      if (ground_mat[c] == 1 || range_mat[c] == std::numeric_limits<double>::max()) label_mat[c] = -1;
    // cloudSegmentation (:147-156)
    int label_cnt = 1;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j)
        if (label_mat[(size_t)i * C + j] == 0) label_components(i, j, label_cnt);

    // compaction (:158-191)
    int line_size = 0;
    for (int i = 0; i < R; ++i) {
      startRing[i] = line_size + 5;
      for (int j = 0; j < C; ++j) {
        const size_t c = (size_t)i * C + j;
        if (label_mat[c] > 0 || ground_mat[c] == 1) {
          if (label_mat[c] == 999999) {
            if (i > P.ground_scan_id && j % 5 == 0) outlier_cloud.push_back(full_cloud[c]);
            continue;
          } else if (ground_mat[c] == 1) {
            if (j % 5 != 0 && j > 4 && j < C - 5) continue;
          }
          segGround.push_back(ground_mat[c] == 1);
          segCol.push_back(j);
          segRange.push_back((float)range_mat[c]);
          seg_cloud.push_back(full_cloud[c]);
          ++line_size;
        }
      }
      endRing[i] = line_size - 1 - 5;
    }
    t_ip = sw.ms();
    return ALEGO_OK;
  }

  // ImageProjection::labelComponents (imageProjection.cpp:210-316).  FIFO order and the neighbour order
  // (-1,0),(1,0),(0,-1),(0,1) (:37-40) are kept although the result does not depend on them.
  void label_components(int row, int col, int &label_cnt) {
    vector<std::pair<int, int>> all;  // doubles as the BFS queue (head index)
    vector<uint8_t> line_flag(R, 0);
    all.emplace_back(row, col);
    line_flag[row] = 1;
    size_t head = 0;
    const double sx = std::sin(seg_alpha_x), cx = std::cos(seg_alpha_x);
    const double sy = std::sin(seg_alpha_y), cy = std::cos(seg_alpha_y);
    static const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
    // the queue and the "all pushed" list hold the same cells in the same order in the reference
    label_mat[(size_t)row * C + col] = label_cnt;
    while (head < all.size()) {
      const int fi = all[head].first, fj = all[head].second;
      ++head;
      label_mat[(size_t)fi * C + fj] = label_cnt;
      line_flag[fi] = 1;
      for (int t = 0; t < 4; ++t) {
        const int ti = fi + di[t];
        int tj = fj + dj[t];
        if (ti < 0 || ti >= R) continue;
        if (tj < 0) tj = C - 1;
        else if (tj >= C) tj = 0;
        if (label_mat[(size_t)ti * C + tj]) continue;
        const double ra = range_mat[(size_t)fi * C + fj], rb = range_mat[(size_t)ti * C + tj];
        const double d1 = std::max(ra, rb), d2 = std::min(ra, rb);
        const bool horiz = di[t] == 0;
        const double angle = std::atan2(d2 * (horiz ? sx : sy), d1 - d2 * (horiz ? cx : cy));
        if (angle > P.seg_theta) {
          label_mat[(size_t)ti * C + tj] = label_cnt;
          line_flag[ti] = 1;
          all.emplace_back(ti, tj);
        }
      }
    }
    bool feasible = false;
    if ((int)all.size() >= P.seg_min_cluster) feasible = true;
    else if ((int)all.size() >= P.seg_valid_point_num) {
      int lines = 0;
      for (int i = 0; i < R; ++i) lines += line_flag[i];
      if (lines >= P.seg_valid_line_num) feasible = true;
    }
    if (feasible) ++label_cnt;
    else
      for (auto &rc : all) label_mat[(size_t)rc.first * C + rc.second] = 999999;
  }

  // -------------------------------------------------------------------------------------------------
  // LaserOdometry::mainLoop steps 2-4 (laserOdometry.cpp:118-297)
  // -------------------------------------------------------------------------------------------------
  // One pass of the per-segment sort + greedy picks.  `stable`=false is the reference (std::sort with a
  // curvature-only comparator, :185); true orders ties by index (used only to detect tie sensitivity).
  void select_features(bool stable, vector<uint8_t> &pk, vector<int32_t> &lab, vector<int32_t> &sidx, vector<int32_t> &o_sharp,
                       vector<int32_t> &o_less_sharp, vector<int32_t> &o_flat, vector<vector<int32_t>> &o_less_flat_scan,
                       int *tie_segments) {
    o_sharp.clear(); o_less_sharp.clear(); o_flat.clear();
    o_less_flat_scan.assign(R, {});
    if (tie_segments) *tie_segments = 0;
    for (int i = 0; i < R; ++i) {
      for (int j = 0; j < 6; ++j) {
        const int sp = (startRing[i] * (6 - j) + endRing[i] * j) / 6;
        const int ep = (startRing[i] * (5 - j) + endRing[i] * (j + 1)) / 6 - 1;
        if (sp >= ep) continue;
        if (stable)
          std::sort(sidx.begin() + sp, sidx.begin() + ep + 1,
                    [this](int a, int b) { return curvature[a] != curvature[b] ? curvature[a] < curvature[b] : a < b; });
        else
          std::sort(sidx.begin() + sp, sidx.begin() + ep + 1, [this](int a, int b) { return curvature[a] < curvature[b]; });
        if (tie_segments) {
          for (int k = sp; k < ep; ++k)
            if (curvature[sidx[k]] == curvature[sidx[k + 1]]) { ++*tie_segments; break; }
        }
        int picked_num = 0;
        for (int k = ep; k >= sp; --k) {
          const int idx = sidx[k];
          if (pk[idx] == 0 && curvature[idx] > 0.1 && segGround[idx] == 0) {
            ++picked_num;
            pk[idx] = 1;
            if (picked_num <= 2) { lab[idx] = 2; o_sharp.push_back(idx); o_less_sharp.push_back(idx); }
            else if (picked_num <= 20) { lab[idx] = 1; o_less_sharp.push_back(idx); }
            else break;
            for (int l = 1; l <= 5; ++l) {
              if (std::abs(segCol[idx + l] - segCol[idx + l - 1]) > 10) break;
              pk[idx + l] = 1;
            }
            for (int l = -1; l >= -5; --l) {
              if (std::abs(segCol[idx + l] - segCol[idx + l + 1]) > 10) break;
              pk[idx + l] = 1;
            }
          }
        }
        picked_num = 0;
        for (int k = sp; k <= ep; ++k) {
          const int idx = sidx[k];
          if (pk[idx] == 0 && curvature[idx] < 0.1 && segGround[idx] == 1) {
            lab[idx] = -1;
            o_flat.push_back(idx);
            ++picked_num;
            pk[idx] = 1;
            if (picked_num >= 4) break;
            for (int l = 1; l <= 5; ++l) {
              if (std::abs(segCol[idx + l] - segCol[idx + l - 1]) > 10) break;
              pk[idx + l] = 1;
            }
            for (int l = -1; l >= -5; --l) {
              if (std::abs(segCol[idx + l] - segCol[idx + l + 1]) > 10) break;
              pk[idx + l] = 1;
            }
          }
        }
        for (int k = sp; k <= ep; ++k)
          if (lab[k] <= 0) o_less_flat_scan[i].push_back(k);
      }
    }
  }

  int lo_features() {
    Stopwatch sw;
    const int M = (int)seg_cloud.size();
    // calculateSmoothness (:122-129): float sum, left to right, r[i]*10 a float product
    for (int i = 5; i < M - 5; ++i) {
      const float *r = segRange.data();
      double diff_range = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 + r[i + 1] + r[i + 2] + r[i + 3] + r[i + 4] + r[i + 5];
      curvature[i] = diff_range * diff_range;
      picked[i] = 0;
      cloud_label[i] = 0;
      sort_idx[i] = i;
    }
    // markOccludedPoints (:131-159)
    for (int i = 5; i < M - 5; ++i) {
      const double depth1 = segRange[i], depth2 = segRange[i + 1];
      const int col_diff = std::abs(segCol[i] - segCol[i + 1]);
      if (col_diff < 10) {
        if (depth1 - depth2 > 0.5) {
          for (int l = 0; l <= 5; ++l) picked[i - l] = 1;
          continue;
        } else if (depth2 - depth1 > 0.5) {
          for (int l = 1; l <= 5; ++l) picked[i + l] = 1;
        }
      }
      const double diff1 = std::abs(segRange[i - 1] - depth1);  // float - double → double
      const double diff2 = std::abs(depth2 - depth1);
      if (diff1 > 0.02 * segRange[i] && diff2 > 0.02 * segRange[i]) picked[i] = 1;
    }
    // tie-sensitivity audit on copies
    {
      vector<uint8_t> pk2 = picked;
      vector<int32_t> lab2 = cloud_label, sidx2 = sort_idx, s2, ls2, f2;
      vector<vector<int32_t>> lf2;
      select_features(true, pk2, lab2, sidx2, s2, ls2, f2, lf2, nullptr);
      vector<vector<int32_t>> lfs;
      select_features(false, picked, cloud_label, sort_idx, sharp_idx, less_sharp_idx, flat_idx, lfs, &n_tie_segments);
      tie_sensitive = !(s2 == sharp_idx && ls2 == less_sharp_idx && f2 == flat_idx && lf2 == lfs);
      // extractFeatures clouds + per-ring VoxelGrid (:288-293)
      sharp.clear(); less_sharp.clear(); flat.clear(); less_flat.clear(); less_flat_stable.clear();
      less_flat_scan_idx.clear();
      for (int k : sharp_idx) sharp.push_back(seg_cloud[k]);
      for (int k : less_sharp_idx) less_sharp.push_back(seg_cloud[k]);
      for (int k : flat_idx) flat.push_back(seg_cloud[k]);
      for (int i = 0; i < R; ++i) {
        vector<P4> scan, ds;
        for (int k : lfs[i]) { scan.push_back(seg_cloud[k]); less_flat_scan_idx.push_back(k); }
        voxel_grid(scan, (float)P.less_flat_leaf, ds, false);
        less_flat.insert(less_flat.end(), ds.begin(), ds.end());
        voxel_grid(scan, (float)P.less_flat_leaf, ds, true);
        less_flat_stable.insert(less_flat_stable.end(), ds.begin(), ds.end());
      }
    }
    t_feat = sw.ms();
    return ALEGO_OK;
  }

  // LaserOdometry::transformToStart (laserOdometry.cpp:728-740), s = 1
  void transform_to_start(const P4 &pi, P4 &po) const {
    const PoseTrig T(lo_params);
    const double x = pi.x, y = pi.y, z = pi.z;
    po.x = (float)(T.R[0] * x + T.R[1] * y + T.R[2] * z + lo_params[0]);
    po.y = (float)(T.R[3] * x + T.R[4] * y + T.R[5] * z + lo_params[1]);
    po.z = (float)(T.R[6] * x + T.R[7] * y + T.R[8] * z + lo_params[2]);
    po.i = pi.i;
  }
  static double sqdist_d(const P4 &a, const P4 &q) {  // pow(float-float,2) summed in double (:354 etc.)
    return std::pow(a.x - q.x, 2) + std::pow(a.y - q.y, 2) + std::pow(a.z - q.z, 2);
  }

  // -------------------------------------------------------------------------------------------------
  // scan-to-scan (laserOdometry.cpp:316-535)
  // -------------------------------------------------------------------------------------------------
  int lo_scan2scan() {
    Stopwatch sw;
    const vector<P4> &lf = stable_voxel ? less_flat_stable : less_flat;
    lo_report = AlegoSolveReport{};
    lo_resids.clear(); lo_surf_corr.clear(); lo_corner_corr.clear(); lo_trace.clear();
    if (!lo_init) {  // :316-324
      lo_init = true;
      surf_last = lf;
      corner_last = less_sharp;
      kd_surf_last.build(surf_last);
      kd_corner_last.build(corner_last);
      lo_report.status = ALEGO_OK;
      t_s2s = sw.ms();
      return ALEGO_OK;
    }
    int status = ALEGO_OK;
    const double gate = P.nearest_feature_dist;
    int idx1[1]; float d1[1];
    // surf association (:337-407)
    for (size_t j = 0; j < flat.size(); ++j) {
      P4 sel;
      transform_to_start(flat[j], sel);
      if (kd_surf_last.knn(sel, 1, idx1, d1) < 1) continue;
      int closest = -1, min_idx2 = -1, min_idx3 = -1;
      if (d1[0] < gate) {
        closest = idx1[0];
        double min_dist2 = gate, min_dist3 = gate;
        const int closest_scan = (int)surf_last[closest].i;
        for (int k = closest + 1; k < (int)surf_last.size(); ++k) {
          if ((int)surf_last[k].i > closest_scan + 2.5) break;
          const double pd = sqdist_d(surf_last[k], sel);
          if ((int)surf_last[k].i == closest_scan) { if (pd < min_dist2) { min_dist2 = pd; min_idx2 = k; } }
          else if (pd < min_dist3) { min_dist3 = pd; min_idx3 = k; }
        }
        for (int k = closest - 1; k >= 0; --k) {
          if ((int)surf_last[k].i < closest_scan - 2.5) break;
          const double pd = sqdist_d(surf_last[k], sel);
          if ((int)surf_last[k].i == closest_scan) { if (pd < min_dist2) { min_dist2 = pd; min_idx2 = k; } }
          else if (pd < min_dist3) { min_dist3 = pd; min_idx3 = k; }
        }
        if (min_idx2 >= 0 && min_idx3 >= 0) {
          Resid f{};
          f.kind = 1;
          f.cp[0] = flat[j].x; f.cp[1] = flat[j].y; f.cp[2] = flat[j].z;
          const P4 &pj = surf_last[closest], &pl = surf_last[min_idx2], &pm = surf_last[min_idx3];
          f.a[0] = pj.x; f.a[1] = pj.y; f.a[2] = pj.z;
          f.b[0] = pl.x; f.b[1] = pl.y; f.b[2] = pl.z;
          f.c[0] = pm.x; f.c[1] = pm.y; f.c[2] = pm.z;
          lo_resids.push_back(f);
          lo_surf_corr.insert(lo_surf_corr.end(), {(int)j, closest, min_idx2, min_idx3});
        }
      }
    }
    const int n_surf = (int)lo_resids.size();
    lo_report.n_surf = n_surf;
    bool first_solve = true;
    if (n_surf >= 10) {  // :410-421
      SolveSummary s;
      ceres_like_solve(lo_resids, lo_params, P.lo_surf_iters, P.huber_delta, &s, &lo_trace);
      lo_report.iterations += s.iterations;
      lo_report.initial_cost = s.initial_cost;
      lo_report.final_cost = s.final_cost;
      first_solve = false;
    } else status = ALEGO_FEW_FEATURES;
    // corner association (:427-481) with the params_ updated by the first solve
    int n_corner = 0;
    for (size_t j = 0; j < sharp.size(); ++j) {
      P4 sel;
      transform_to_start(sharp[j], sel);
      if (kd_corner_last.knn(sel, 1, idx1, d1) < 1) continue;
      int closest = -1, min_idx2 = -1;
      if (d1[0] < gate) {
        closest = idx1[0];
        const int closest_scan = (int)corner_last[closest].i;
        double min_dist2 = gate;
        for (int k = closest + 1; k < (int)corner_last.size(); ++k) {
          if ((int)corner_last[k].i > closest_scan + 2) break;
          const double pd = sqdist_d(corner_last[k], sel);
          if ((int)corner_last[k].i > closest_scan && pd < min_dist2) { min_dist2 = pd; min_idx2 = k; }
        }
        for (int k = closest - 1; k >= 0; --k) {
          if ((int)corner_last[k].i < closest_scan - 2) break;
          const double pd = sqdist_d(corner_last[k], sel);
          if ((int)corner_last[k].i < closest_scan && pd < min_dist2) { min_dist2 = pd; min_idx2 = k; }
        }
      }
      if (min_idx2 >= 0) {
        Resid f{};
        f.kind = 0;
        f.cp[0] = sharp[j].x; f.cp[1] = sharp[j].y; f.cp[2] = sharp[j].z;
        const P4 &pj = corner_last[closest], &pl = corner_last[min_idx2];
        f.a[0] = pj.x; f.a[1] = pj.y; f.a[2] = pj.z;
        f.b[0] = pl.x; f.b[1] = pl.y; f.b[2] = pl.z;
        lo_resids.push_back(f);
        lo_corner_corr.insert(lo_corner_corr.end(), {(int)j, closest, min_idx2});
        ++n_corner;
      }
    }
    lo_report.n_corner = n_corner;
    if (n_corner >= 10) {  // :484-495 — same Problem: surf + corner blocks
      SolveSummary s;
      ceres_like_solve(lo_resids, lo_params, P.lo_corner_iters, P.huber_delta, &s, &lo_trace);
      lo_report.iterations += s.iterations;
      if (first_solve) lo_report.initial_cost = s.initial_cost;
      lo_report.final_cost = s.final_cost;
    } else status = ALEGO_FEW_FEATURES;
    // pose integration (:504-508): translation + yaw only
    {
      const double t_lc[3] = {lo_params[0], lo_params[1], lo_params[2]};
      const double cy = std::cos(lo_params[5]), sy = std::sin(lo_params[5]);
      const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
      double rt[3];
      mat3_vec(r_w, t_lc, rt);
      for (int c = 0; c < 3; ++c) t_w[c] += rt[c];
      mat3_mul(r_w, Rz, r_w);
    }
    surf_last = lf;  // :531-534
    corner_last = less_sharp;
    kd_surf_last.build(surf_last);
    kd_corner_last.build(corner_last);
    lo_report.status = status;
    t_s2s = sw.ms();
    return status;
  }

  // -------------------------------------------------------------------------------------------------
  // LaserMapping (laserMapping.cpp:188-192, 325-489; laserMapping.h:187-194)
  // -------------------------------------------------------------------------------------------------
  void transform_associate_to_map() {  // :188-192
    double rt[3];
    mat3_vec(r_m2o, t_o2l, rt);
    for (int c = 0; c < 3; ++c) t_m2l[c] = rt[c] + t_m2o[c];
    mat3_mul(r_m2o, r_o2l, r_m2l);
  }
  void point_associate_to_map(const P4 &pi, P4 &po) const {  // laserMapping.h:187-194
    const double x = pi.x, y = pi.y, z = pi.z;
    po.x = (float)(r_m2l[0] * x + r_m2l[1] * y + r_m2l[2] * z + t_m2l[0]);
    po.y = (float)(r_m2l[3] * x + r_m2l[4] * y + r_m2l[5] * z + t_m2l[1]);
    po.z = (float)(r_m2l[6] * x + r_m2l[7] * y + r_m2l[8] * z + t_m2l[2]);
    po.i = pi.i;
  }
  void downsample_current_scan() {  // :325-346
    voxel_grid(lm_corner, (float)P.lm_corner_leaf, lm_corner_ds, stable_voxel);
    voxel_grid(lm_surf, (float)P.lm_surf_leaf, lm_surf_ds, stable_voxel);
    voxel_grid(lm_outlier, (float)P.lm_outlier_leaf, lm_outlier_ds, stable_voxel);
    lm_surf_total = lm_surf_ds;
    lm_surf_total.insert(lm_surf_total.end(), lm_outlier_ds.begin(), lm_outlier_ds.end());
    voxel_grid(lm_surf_total, (float)P.lm_surf_leaf, lm_surf_total_ds, stable_voxel);
  }
  void transform_update() {  // :481-489
    const PoseTrig T(lm_params);
    std::memcpy(r_m2l, T.R, sizeof r_m2l);
    for (int c = 0; c < 3; ++c) t_m2l[c] = lm_params[c];
    double inv[9] = {r_o2l[0], r_o2l[3], r_o2l[6], r_o2l[1], r_o2l[4], r_o2l[7], r_o2l[2], r_o2l[5], r_o2l[8]};
    mat3_mul(r_m2l, inv, r_m2o);
    double rt[3];
    mat3_vec(r_m2o, t_o2l, rt);
    for (int c = 0; c < 3; ++c) t_m2o[c] = t_m2l[c] - rt[c];
  }

  int lm_scan2map() {
    Stopwatch sw;
    lm_report = AlegoSolveReport{};
    lm_resids.clear(); lm_corner_sel.clear(); lm_surf_sel.clear(); lm_trace.clear();
    transform_associate_to_map();
    downsample_current_scan();
    int status = ALEGO_OK;
    if (lm_corner_ds.size() < 10 || lm_surf_total.size() < 100 || map_corner.size() < 10) {  // :350-354
      status = ALEGO_FEW_FEATURES;
    } else {
      kd_map_corner.build(map_corner);  // :356-357, every mapped frame
      kd_map_surf.build(map_surf);
      for (int outer = 0; outer < P.lm_outer_iters; ++outer) {  // :360
        lm_resids.clear(); lm_corner_sel.clear(); lm_surf_sel.clear();
        int idx[5]; float dist[5];
        for (size_t i = 0; i < lm_corner_ds.size(); ++i) {  // :371-417
          P4 sel;
          point_associate_to_map(lm_corner_ds[i], sel);
          if (kd_map_corner.knn(sel, 5, idx, dist) < 5) continue;
          if (dist[4] < 1.0) {
            double near[5][3], center[3] = {0, 0, 0};
            for (int j = 0; j < 5; ++j) {
              near[j][0] = map_corner[idx[j]].x; near[j][1] = map_corner[idx[j]].y; near[j][2] = map_corner[idx[j]].z;
              for (int c = 0; c < 3; ++c) center[c] = center[c] + near[j][c];
            }
            for (int c = 0; c < 3; ++c) center[c] = center[c] / 5.0;
            double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < 5; ++j) {
              double zm[3] = {near[j][0] - center[0], near[j][1] - center[1], near[j][2] - center[2]};
              for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) cov[r * 3 + c] = cov[r * 3 + c] + zm[r] * zm[c];
            }
            double w[3], V[9];
            eig3_sym(cov, w, V);
            if (w[2] > 3 * w[1]) {
              Resid f{};
              f.kind = 2;
              f.cp[0] = lm_corner_ds[i].x; f.cp[1] = lm_corner_ds[i].y; f.cp[2] = lm_corner_ds[i].z;
              for (int c = 0; c < 3; ++c) {
                const double u = V[c * 3 + 2];
                f.a[c] = 0.1 * u + center[c];
                f.b[c] = -0.1 * u + center[c];
              }
              lm_resids.push_back(f);
              lm_corner_sel.push_back((int)i);
            }
          }
        }
        const int n_corner = (int)lm_resids.size();
        for (size_t i = 0; i < lm_surf_total_ds.size(); ++i) {  // :419-462
          P4 sel;
          point_associate_to_map(lm_surf_total_ds[i], sel);
          if (kd_map_surf.knn(sel, 5, idx, dist) < 5) continue;
          if (dist[4] < 1.0) {
            double A[15], B[5] = {-1, -1, -1, -1, -1}, nrm[3];
            for (int j = 0; j < 5; ++j) { A[j * 3] = map_surf[idx[j]].x; A[j * 3 + 1] = map_surf[idx[j]].y; A[j * 3 + 2] = map_surf[idx[j]].z; }
            lstsq_5x3(A, B, nrm);
            const double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            const double d = 1 / nn;
            for (int c = 0; c < 3; ++c) nrm[c] /= nn;
            bool valid = true;
            for (int j = 0; j < 5; ++j)
              if (std::fabs(nrm[0] * map_surf[idx[j]].x + nrm[1] * map_surf[idx[j]].y + nrm[2] * map_surf[idx[j]].z + d) > 0.2) { valid = false; break; }
            if (valid) {
              Resid f{};
              f.kind = 3;
              f.cp[0] = lm_surf_total_ds[i].x; f.cp[1] = lm_surf_total_ds[i].y; f.cp[2] = lm_surf_total_ds[i].z;
              f.a[0] = nrm[0]; f.a[1] = nrm[1]; f.a[2] = nrm[2];
              f.d = d;
              lm_resids.push_back(f);
              lm_surf_sel.push_back((int)i);
            }
          }
        }
        lm_report.n_corner = n_corner;
        lm_report.n_surf = (int)lm_resids.size() - n_corner;
        SolveSummary s;  // :467-477 — solved whatever the correspondence count
        if (!lm_resids.empty()) {
          ceres_like_solve(lm_resids, lm_params, P.lm_max_iters, P.huber_delta, &s, &lm_trace);
          lm_report.iterations += s.iterations;
          if (outer == 0) lm_report.initial_cost = s.initial_cost;
          lm_report.final_cost = s.final_cost;
        }
      }
    }
    transform_update();
    lm_report.status = status;
    t_lm = sw.ms();
    return status;
  }

  int pipeline_step(const P4 *pts, int n) {
    int rc = ip(pts, n);
    if (rc < 0) return rc;
    lo_features();
    int s1 = lo_scan2scan();
    int s2 = ALEGO_OK;
    if (lm_every > 0 && scan_count % lm_every == 0) {
      lm_corner = corner_last;  // /corner_last, /surf_last, /outlier (laserOdometry.cpp:537-546; laserMapping.cpp:133-153)
      lm_surf = surf_last;
      lm_outlier = outlier_cloud;
      std::memcpy(t_o2l, t_w, sizeof t_w);  // laserOdomHandler (:154-164)
      std::memcpy(r_o2l, r_w, sizeof r_w);
      s2 = lm_scan2map();
    }
    ++scan_count;
    return (s1 == ALEGO_OK && s2 == ALEGO_OK) ? ALEGO_OK : ALEGO_FEW_FEATURES;
  }
};

template <typename T>
static int64_t put(const vector<T> &v, void *dst, size_t cap) {
  const size_t bytes = v.size() * sizeof(T);
  if (dst) {
    if (bytes > cap) return -1;
    if (bytes) std::memcpy(dst, v.data(), bytes);
  }
  return (int64_t)bytes;
}
static int64_t putd(const double *p, int n, void *dst, size_t cap) {
  if (dst) {
    if (sizeof(double) * n > cap) return -1;
    std::memcpy(dst, p, sizeof(double) * n);
  }
  return (int64_t)sizeof(double) * n;
}

}  // namespace

// =====================================================================================================
// C ABI for ctypes (tests / bench cpu_baseline only)
// =====================================================================================================
extern "C" {

int oracle_default_params(AlegoParams *p, int preset) {
  if (!p) return ALEGO_BAD_ARG;
  std::memset(p, 0, sizeof *p);
  p->seg_valid_point_num = 5;
  p->seg_valid_line_num = 3;
  p->seg_min_cluster = 30;
  p->lo_surf_iters = 5;
  p->lo_corner_iters = 5;
  p->lm_outer_iters = 2;
  p->lm_max_iters = 20;
  p->sensor_mount_ang = 0.;
  p->seg_theta = 1.047;
  p->nearest_feature_dist = 25.;
  p->huber_delta = 0.1;
  p->less_flat_leaf = 0.4;
  p->lm_corner_leaf = 0.4;
  p->lm_surf_leaf = 0.8;
  p->lm_outlier_leaf = 1.0;
  switch (preset) {
    case ALEGO_PRESET_VLP16_1800:
      p->n_scan = 16; p->ang_res_x = 0.2; p->ang_res_y = 2.0; p->ang_bottom = 15.0; p->ground_scan_id = 7; break;
    case ALEGO_PRESET_HDL64_1800:
      p->n_scan = 64; p->ang_res_x = 0.2; p->ang_res_y = 0.427; p->ang_bottom = 24.9; p->ground_scan_id = 50; break;
    case ALEGO_PRESET_HDL64_2048:
      p->n_scan = 64; p->ang_res_x = 360.0 / 2048.0; p->ang_res_y = 0.427; p->ang_bottom = 24.9; p->ground_scan_id = 50; break;
    case ALEGO_PRESET_REFERENCE:
      p->n_scan = 16; p->ang_res_x = 0.09; p->ang_res_y = 2.0; p->ang_bottom = 15.0; p->ground_scan_id = 10; break;
    default: return ALEGO_BAD_ARG;
  }
  p->horizon_scan = (int)(360.0 / p->ang_res_x + 0.5);  // utility.h:55
  return ALEGO_OK;
}

void *oracle_create(const AlegoParams *p) {
  if (!p || p->n_scan <= 0 || p->horizon_scan <= 0) return nullptr;
  return new Oracle(*p);
}
void oracle_destroy(void *h) { delete static_cast<Oracle *>(h); }

int oracle_config(void *h, int lm_every, int stable_voxel) {
  Oracle *o = static_cast<Oracle *>(h);
  o->lm_every = lm_every;
  o->stable_voxel = stable_voxel != 0;
  return ALEGO_OK;
}
int oracle_ip(void *h, const float *xyzi, int n) { return static_cast<Oracle *>(h)->ip(reinterpret_cast<const P4 *>(xyzi), n); }
int oracle_lo_features(void *h) { return static_cast<Oracle *>(h)->lo_features(); }
int oracle_lo_scan2scan(void *h) { return static_cast<Oracle *>(h)->lo_scan2scan(); }
int oracle_lm_set_map(void *h, const float *corner, int nc, const float *surf, int ns) {
  Oracle *o = static_cast<Oracle *>(h);
  o->map_corner.assign(reinterpret_cast<const P4 *>(corner), reinterpret_cast<const P4 *>(corner) + nc);
  o->map_surf.assign(reinterpret_cast<const P4 *>(surf), reinterpret_cast<const P4 *>(surf) + ns);
  return ALEGO_OK;
}
int oracle_lm_set_scan(void *h, const float *corner, int nc, const float *surf, int ns, const float *outl, int no) {
  Oracle *o = static_cast<Oracle *>(h);
  o->lm_corner.assign(reinterpret_cast<const P4 *>(corner), reinterpret_cast<const P4 *>(corner) + nc);
  o->lm_surf.assign(reinterpret_cast<const P4 *>(surf), reinterpret_cast<const P4 *>(surf) + ns);
  o->lm_outlier.assign(reinterpret_cast<const P4 *>(outl), reinterpret_cast<const P4 *>(outl) + no);
  return ALEGO_OK;
}
int oracle_lm_set_odom(void *h, const double *t, const double *r) {
  Oracle *o = static_cast<Oracle *>(h);
  std::memcpy(o->t_o2l, t, sizeof o->t_o2l);
  std::memcpy(o->r_o2l, r, sizeof o->r_o2l);
  return ALEGO_OK;
}
int oracle_lm_set_params(void *h, const double *p) { std::memcpy(static_cast<Oracle *>(h)->lm_params, p, 6 * sizeof(double)); return ALEGO_OK; }
int oracle_lo_set_params(void *h, const double *p) { std::memcpy(static_cast<Oracle *>(h)->lo_params, p, 6 * sizeof(double)); return ALEGO_OK; }
int oracle_lm_scan2map(void *h) { return static_cast<Oracle *>(h)->lm_scan2map(); }
int oracle_pipeline_step(void *h, const float *xyzi, int n) { return static_cast<Oracle *>(h)->pipeline_step(reinterpret_cast<const P4 *>(xyzi), n); }

// Named getters: returns bytes (copied when dst != NULL), -1 capacity too small, -2 unknown name.
int64_t oracle_get(void *h, const char *name, void *dst, size_t cap) {
  Oracle *o = static_cast<Oracle *>(h);
  const std::string s(name);
  auto putf = [&](float v) { if (dst) { if (cap < 4) return (int64_t)-1; std::memcpy(dst, &v, 4); } return (int64_t)4; };
  auto puti = [&](int32_t v) { if (dst) { if (cap < 4) return (int64_t)-1; std::memcpy(dst, &v, 4); } return (int64_t)4; };
  if (s == "range_mat") return put(o->range_mat, dst, cap);
  if (s == "full_cloud") return put(o->full_cloud, dst, cap);
  if (s == "ground_mat") return put(o->ground_mat, dst, cap);
  if (s == "label_mat") return put(o->label_mat, dst, cap);
  if (s == "startRingIndex") return put(o->startRing, dst, cap);
  if (s == "endRingIndex") return put(o->endRing, dst, cap);
  if (s == "segmentedCloudGroundFlag") return put(o->segGround, dst, cap);
  if (s == "segmentedCloudColInd") return put(o->segCol, dst, cap);
  if (s == "segmentedCloudRange") return put(o->segRange, dst, cap);
  if (s == "segmented_cloud") return put(o->seg_cloud, dst, cap);
  if (s == "outlier_cloud") return put(o->outlier_cloud, dst, cap);
  if (s == "startOrientation") return putf(o->startOri);
  if (s == "endOrientation") return putf(o->endOri);
  if (s == "orientationDiff") return putf(o->oriDiff);
  if (s == "min_margin_row") return putd(&o->min_margin_row, 1, dst, cap);
  if (s == "min_margin_col") return putd(&o->min_margin_col, 1, dst, cap);
  if (s == "n_dup_cells") return puti(o->n_dup_cells);
  if (s == "cloud_curvature") { vector<double> v(o->curvature.begin(), o->curvature.begin() + o->seg_cloud.size()); return put(v, dst, cap); }
  if (s == "cloud_neighbor_picked") { vector<uint8_t> v(o->picked.begin(), o->picked.begin() + o->seg_cloud.size()); return put(v, dst, cap); }
  if (s == "cloud_label") { vector<int32_t> v(o->cloud_label.begin(), o->cloud_label.begin() + o->seg_cloud.size()); return put(v, dst, cap); }
  if (s == "cloud_sort_idx") { vector<int32_t> v(o->sort_idx.begin(), o->sort_idx.begin() + o->seg_cloud.size()); return put(v, dst, cap); }
  if (s == "sharp_idx") return put(o->sharp_idx, dst, cap);
  if (s == "less_sharp_idx") return put(o->less_sharp_idx, dst, cap);
  if (s == "flat_idx") return put(o->flat_idx, dst, cap);
  if (s == "less_flat_scan_idx") return put(o->less_flat_scan_idx, dst, cap);
  if (s == "sharp") return put(o->sharp, dst, cap);
  if (s == "less_sharp") return put(o->less_sharp, dst, cap);
  if (s == "flat") return put(o->flat, dst, cap);
  if (s == "less_flat") return put(o->less_flat, dst, cap);
  if (s == "less_flat_stable") return put(o->less_flat_stable, dst, cap);
  if (s == "n_tie_segments") return puti(o->n_tie_segments);
  if (s == "tie_sensitive") return puti(o->tie_sensitive);
  if (s == "surf_last") return put(o->surf_last, dst, cap);
  if (s == "corner_last") return put(o->corner_last, dst, cap);
  if (s == "lo_params") return putd(o->lo_params, 6, dst, cap);
  if (s == "t_w_cur") return putd(o->t_w, 3, dst, cap);
  if (s == "r_w_cur") return putd(o->r_w, 9, dst, cap);
  if (s == "lo_surf_corr") return put(o->lo_surf_corr, dst, cap);
  if (s == "lo_corner_corr") return put(o->lo_corner_corr, dst, cap);
  if (s == "lo_trace") return put(o->lo_trace, dst, cap);
  if (s == "lm_trace") return put(o->lm_trace, dst, cap);
  if (s == "lm_params") return putd(o->lm_params, 6, dst, cap);
  if (s == "t_map2laser") return putd(o->t_m2l, 3, dst, cap);
  if (s == "r_map2laser") return putd(o->r_m2l, 9, dst, cap);
  if (s == "t_map2odom") return putd(o->t_m2o, 3, dst, cap);
  if (s == "r_map2odom") return putd(o->r_m2o, 9, dst, cap);
  if (s == "lm_corner_ds") return put(o->lm_corner_ds, dst, cap);
  if (s == "lm_surf_ds") return put(o->lm_surf_ds, dst, cap);
  if (s == "lm_outlier_ds") return put(o->lm_outlier_ds, dst, cap);
  if (s == "lm_surf_total_ds") return put(o->lm_surf_total_ds, dst, cap);
  if (s == "lm_corner_sel") return put(o->lm_corner_sel, dst, cap);
  if (s == "lm_surf_sel") return put(o->lm_surf_sel, dst, cap);
  if (s == "lm_resids" || s == "lo_resids") {
    const vector<Resid> &rs = s == "lm_resids" ? o->lm_resids : o->lo_resids;
    vector<double> v;
    for (const Resid &f : rs) {
      v.push_back(f.kind);
      for (int c = 0; c < 3; ++c) v.push_back(f.cp[c]);
      for (int c = 0; c < 3; ++c) v.push_back(f.a[c]);
      for (int c = 0; c < 3; ++c) v.push_back(f.b[c]);
      for (int c = 0; c < 3; ++c) v.push_back(f.c[c]);
      v.push_back(f.d);
    }
    return put(v, dst, cap);
  }
  if (s == "timings_ms") { double t[4] = {o->t_ip, o->t_feat, o->t_s2s, o->t_lm}; return putd(t, 4, dst, cap); }
  return -2;
}
int oracle_get_report(void *h, int which, AlegoSolveReport *out) {
  Oracle *o = static_cast<Oracle *>(h);
  *out = which == 0 ? o->lo_report : o->lm_report;
  return ALEGO_OK;
}

// ---- stand-alone pieces for unit cross-checks ----------------------------------------------------
int oracle_voxel_grid(const float *xyzi, int n, float leaf, int stable, float *out, int *n_out, uint32_t *keys) {
  vector<P4> in(reinterpret_cast<const P4 *>(xyzi), reinterpret_cast<const P4 *>(xyzi) + n), o;
  vector<unsigned> k;
  voxel_grid(in, leaf, o, stable != 0, &k);
  *n_out = (int)o.size();
  if (out && !o.empty()) std::memcpy(out, o.data(), o.size() * sizeof(P4));
  if (keys && !k.empty()) std::memcpy(keys, k.data(), k.size() * sizeof(unsigned));
  return ALEGO_OK;
}
// ---- N1: local-map assembly, the cloud side of extractSurroundingKeyFrames (src/laserMapping.cpp:194-323) ----------
// transformPointCloud (include/alego/laserMapping.h:163-177): Matrix4f from
// (AngleAxisf(yaw, Z) * AngleAxisf(pitch, Y) * AngleAxisf(roll, X)).toRotationMatrix() — Eigen (unpinned, not in the
// container; restated from Eigen 3.3 Geometry/AngleAxis.h, Quaternion.h) converts every AngleAxis to a quaternion
// (w = cos(a/2), vec = sin(a/2) * axis), multiplies the quaternions (generic quat_product) and expands toRotationMatrix();
// then pcl::transformPointCloud (PCL 1.8 common/impl/transforms.hpp): x' = m00*x + m01*y + m02*z + m03 in float, left to
// right, intensity copied.  corner_from_map_ += corner keyframes; surf_from_map_ += surf keyframe, then outlier keyframe
// (:239-243); ds_corner_ (0.4) / ds_surf_ (0.8) VoxelGrid (:316-319).
static void keyframe_matrix(const float pose6[6], float M[12]) {
  auto quat = [](float angle, int axis, float q[4]) {
    const float ha = 0.5f * angle;
    q[0] = std::cos(ha); q[1] = q[2] = q[3] = 0.f;
    q[1 + axis] = std::sin(ha);
  };
  auto mul = [](const float a[4], const float b[4], float o[4]) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  };
  float qz[4], qy[4], qx[4], qzy[4], q[4];
  quat(pose6[5], 2, qz); quat(pose6[4], 1, qy); quat(pose6[3], 0, qx);
  mul(qz, qy, qzy);
  mul(qzy, qx, q);
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M[0] = 1.f - (tyy + tzz); M[1] = txy - twz;         M[2] = txz + twy;          M[3] = pose6[0];
  M[4] = txy + twz;         M[5] = 1.f - (txx + tzz); M[6] = tyz - twx;          M[7] = pose6[1];
  M[8] = txz - twy;         M[9] = tyz + twx;         M[10] = 1.f - (txx + tyy); M[11] = pose6[2];
}
static void append_transformed(const float *xyzi, int n, const float M[12], vector<P4> &out) {
  const P4 *p = reinterpret_cast<const P4 *>(xyzi);
  for (int i = 0; i < n; ++i) {
    P4 o;
    o.x = M[0] * p[i].x + M[1] * p[i].y + M[2] * p[i].z + M[3];
    o.y = M[4] * p[i].x + M[5] * p[i].y + M[6] * p[i].z + M[7];
    o.z = M[8] * p[i].x + M[9] * p[i].y + M[10] * p[i].z + M[11];
    o.i = p[i].i;
    out.push_back(o);
  }
}
int oracle_lm_assemble_map(int n_kf, const float *const *corner, const int *nc, const float *const *surf, const int *ns,
                           const float *const *outl, const int *no, const float *poses6, float leaf_c, float leaf_s, int stable,
                           float *corner_out, int *n_corner_out, float *surf_out, int *n_surf_out, float *matrices_out) {
  vector<P4> cm, sm, cds, sds;
  for (int k = 0; k < n_kf; ++k) {
    float M[12];
    keyframe_matrix(poses6 + (size_t)k * 6, M);
    if (matrices_out) std::memcpy(matrices_out + (size_t)k * 12, M, sizeof M);
    append_transformed(corner[k], nc[k], M, cm);
    append_transformed(surf[k], ns[k], M, sm);
    append_transformed(outl[k], no[k], M, sm);
  }
  voxel_grid(cm, leaf_c, cds, stable != 0);
  voxel_grid(sm, leaf_s, sds, stable != 0);
  *n_corner_out = (int)cds.size();
  *n_surf_out = (int)sds.size();
  if (corner_out && !cds.empty()) std::memcpy(corner_out, cds.data(), cds.size() * sizeof(P4));
  if (surf_out && !sds.empty()) std::memcpy(surf_out, sds.data(), sds.size() * sizeof(P4));
  return ALEGO_OK;
}

// ---- N2: LaserOdometry::adjustDistortion, IMU branch (src/laserOdometry.cpp:557-657) --------------------------------------
// Literal sequential restatement: the forward-only walk of imu_ptr_front_ from imu_ptr_last_iter_ (:587-595), the early
// `return` on unsynchronised stamps (:596-600), the interpolation in double stored to Eigen float vectors (:602-629), the
// start pose taken from point 0 (:633-639) and adjusted_p = r_s_i * (r_c * p + shift_from_start) (:640-655).
// Eigen (unpinned, not in the container; restated from Eigen 3.3): AngleAxisf products through float quaternions
// (keyframe_matrix above), Matrix3f::inverse() = cofactor inverse of LU/InverseImpl.h (determinant from the cofactors of
// column 0, summed by the unrolled redux a0 + (a1 + a2)), `Vector3f * double` converts the scalar to float first
// (promote_scalar_arg), fixed 3x3 products accumulate left to right.  The use_odom branch (:660-714) is dead code
// (use_imu = true, utility.h:68-69) and is not restated.
// q: [10][len] doubles — time, roll, pitch, yaw, shift x y z, velocity x y z.  Returns the number of points visited.
int oracle_adjust_distortion(float *xyzi, int n, const int *col, float start_orientation, float end_orientation, int horizon_scan,
                             double scan_period, double scan_time, const double *q, int len, int ptr_last, int *ptr_last_iter) {
  P4 *pts = reinterpret_cast<P4 *>(xyzi);
  const double *imu_time = q;
  int start_ori = (start_orientation + 2 * M_PI) / horizon_scan;
  int end_ori = (end_orientation + 2 * M_PI) / horizon_scan;
  if (start_ori >= horizon_scan) start_ori -= horizon_scan;
  if (end_ori >= horizon_scan) end_ori -= horizon_scan;
  int ori_diff = end_ori - start_ori;
  if (ori_diff <= 0) ori_diff = horizon_scan;
  float rpy_start[3] = {0, 0, 0}, shift_start[3] = {0, 0, 0}, velo_start[3] = {0, 0, 0}, rpy_cur[3], shift_cur[3], velo_cur[3];
  float r_s_i[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  (void)rpy_start;
  int last_iter = *ptr_last_iter, visited = 0;
  for (int i = 0; i < n; ++i) {
    P4 &p = pts[i];
    const double rel_time = (col[i] - start_ori) * scan_period / ori_diff;
    const double cur_time = scan_time + rel_time;
    if (ptr_last > 0) {
      int front = last_iter;
      while (front != ptr_last) {
        if (cur_time < imu_time[front]) break;
        front = (front + 1) % len;
      }
      if (std::abs(cur_time - imu_time[front]) > scan_period) break;  // "unsync imu and pc msg": return
      if (cur_time > imu_time[front]) {
        for (int k = 0; k < 3; ++k) {
          rpy_cur[k] = q[(1 + k) * len + front];
          shift_cur[k] = q[(4 + k) * len + front];
          velo_cur[k] = q[(7 + k) * len + front];
        }
      } else {
        const int back = (front - 1 + len) % len;
        const double ratio_front = (cur_time - imu_time[back]) / (imu_time[front] - imu_time[back]);
        const double ratio_back = 1. - ratio_front;
        for (int k = 0; k < 3; ++k) {
          rpy_cur[k] = q[(1 + k) * len + front] * ratio_front + q[(1 + k) * len + back] * ratio_back;
          shift_cur[k] = q[(4 + k) * len + front] * ratio_front + q[(4 + k) * len + back] * ratio_back;
          velo_cur[k] = q[(7 + k) * len + front] * ratio_front + q[(7 + k) * len + back] * ratio_back;
        }
      }
      const float pose6[6] = {0.f, 0.f, 0.f, rpy_cur[0], rpy_cur[1], rpy_cur[2]};
      float M[12];
      keyframe_matrix(pose6, M);
      const float r_c[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
      if (i == 0) {
        for (int k = 0; k < 3; ++k) { rpy_start[k] = rpy_cur[k]; shift_start[k] = shift_cur[k]; velo_start[k] = velo_cur[k]; }
        auto cof = [&](int a, int b) {
          const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
          return r_c[a1 * 3 + b1] * r_c[a2 * 3 + b2] - r_c[a1 * 3 + b2] * r_c[a2 * 3 + b1];
        };
        const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        const float det = c0 * r_c[0] + (c1 * r_c[3] + c2 * r_c[6]);
        const float invdet = 1.f / det;
        r_s_i[0] = c0 * invdet; r_s_i[1] = c1 * invdet; r_s_i[2] = c2 * invdet;
        r_s_i[3] = cof(0, 1) * invdet; r_s_i[4] = cof(1, 1) * invdet; r_s_i[5] = cof(2, 1) * invdet;
        r_s_i[6] = cof(0, 2) * invdet; r_s_i[7] = cof(1, 2) * invdet; r_s_i[8] = cof(2, 2) * invdet;
      } else {
        const float relf = (float)rel_time;
        float sh[3], a[3];
        for (int k = 0; k < 3; ++k) sh[k] = (shift_cur[k] - shift_start[k]) - velo_start[k] * relf;
        for (int k = 0; k < 3; ++k) a[k] = ((r_c[k * 3] * p.x + r_c[k * 3 + 1] * p.y) + r_c[k * 3 + 2] * p.z) + sh[k];
        p.x = (r_s_i[0] * a[0] + r_s_i[1] * a[1]) + r_s_i[2] * a[2];
        p.y = (r_s_i[3] * a[0] + r_s_i[4] * a[1]) + r_s_i[5] * a[2];
        p.z = (r_s_i[6] * a[0] + r_s_i[7] * a[1]) + r_s_i[8] * a[2];
      }
      last_iter = front;
      ++visited;
    }
  }
  *ptr_last_iter = last_iter;
  return visited;
}

// ---- N4: loop-closure ICP, LaserMapping::performLoopClosure (src/laserMapping.cpp:652-711) -------------------------------
// pcl::IterativeClosestPoint<PointXYZI, PointXYZI> with setMaxCorrespondenceDistance(100), setMaximumIterations(100),
// setTransformationEpsilon(1e-6), setEuclideanFitnessEpsilon(1e-6), setRANSACIterations(0) (no rejector is installed, so the
// last one has no effect), align() with the identity guess (the computed initial_guess is never passed, :684), then
// hasConverged(), getFitnessScore(), getFinalTransformation() (:686-688).  PCL is not in the container (unpinned; restated
// from PCL 1.8 registration/impl/icp.hpp, correspondence_estimation.hpp, transformation_estimation_svd.hpp,
// default_convergence_criteria.hpp and Eigen 3.3 Geometry/Umeyama.h):
//   loop: correspondences = 1-NN of every (current) source point in the target, kept when d^2 <= max_dist^2; fewer than 3 ->
//   not converged, stop; transformation_ = umeyama(src, tgt, no scaling) in float; the source cloud is transformed in place
//   (float, x' = m00 x + m01 y + m02 z + m03); final_transformation_ = transformation_ * final_transformation_; ++iterations;
//   DefaultConvergenceCriteria: (1) iterations >= max -> converged (failure_after_max_iter_ = false); (2) cos_angle =
//   0.5 (trace R - 1) >= 1 - transformation_epsilon and |t|^2 <= transformation_epsilon -> converged
//   (max_iterations_similar_transforms_ = 0); (3) mse = mean of the correspondence distances (squared, float, summed in
//   double); |mse - prev| < 1e-12 -> converged; |mse - prev| / prev < euclidean_fitness_epsilon -> converged; prev = mse.
//   getFitnessScore(): mean squared 1-NN distance of the ORIGINAL source transformed by final_transformation_.
// exact_sums = 0: umeyama's reductions are accumulated in float, sequentially (the literal reading; Eigen's real order inside
// its GEMM kernel is not knowable here).  exact_sums = 1: the same quantities accumulated in double and rounded to float where
// Eigen holds floats — the variant the device reproduces (like stable_voxel for VoxelGrid); tests compare both ways.
// trace: per iteration 14 doubles — correspondences, mse, incremental R (9, row-major), t (3).
namespace {
void svd_rotation(const float sigma[9], float R[9]) {  // U * diag(1, 1, det(U) det(V)) * V^T of sigma = U S V^T
  double A[9], AtA[9], w[3], V[9];
  for (int k = 0; k < 9; ++k) A[k] = sigma[k];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) AtA[i * 3 + j] = A[0 * 3 + i] * A[0 * 3 + j] + A[1 * 3 + i] * A[1 * 3 + j] + A[2 * 3 + i] * A[2 * 3 + j];
  eig3_sym(AtA, w, V);  // ascending: column 2 = largest singular value
  double v[3][3], u[3][3];  // [k] = k-th singular vector, descending
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 3; ++r) v[k][r] = V[r * 3 + (2 - k)];
  for (int k = 0; k < 2; ++k) {
    double n2 = 0;
    for (int r = 0; r < 3; ++r) { u[k][r] = A[r * 3] * v[k][0] + A[r * 3 + 1] * v[k][1] + A[r * 3 + 2] * v[k][2]; n2 += u[k][r] * u[k][r]; }
    const double inv = n2 > 0 ? 1.0 / std::sqrt(n2) : 0.0;
    for (int r = 0; r < 3; ++r) u[k][r] *= inv;
  }
  {  // u1 re-orthogonalised against u0 (ill-conditioned sigma), u2 = u0 x u1: det(U) = +1
    double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2], n2 = 0;
    for (int r = 0; r < 3; ++r) { u[1][r] -= d * u[0][r]; n2 += u[1][r] * u[1][r]; }
    const double inv = n2 > 0 ? 1.0 / std::sqrt(n2) : 0.0;
    for (int r = 0; r < 3; ++r) u[1][r] *= inv;
  }
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  const double detV = v[0][0] * (v[1][1] * v[2][2] - v[1][2] * v[2][1]) - v[0][1] * (v[1][0] * v[2][2] - v[1][2] * v[2][0]) +
                      v[0][2] * (v[1][0] * v[2][1] - v[1][1] * v[2][0]);
  const double s2 = detV < 0 ? -1.0 : 1.0;  // det(U) = +1 by construction
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = (float)(u[0][r] * v[0][c] + u[1][r] * v[1][c] + s2 * u[2][r] * v[2][c]);
}
}  // namespace

// the rotation step alone (tests: against numpy's SVD, reflections and rank-deficient covariances included)
int oracle_umeyama_rotation(const float *sigma9, float *R9) { svd_rotation(sigma9, R9); return ALEGO_OK; }

int oracle_icp(const float *src_xyzi, int n_src, const float *tgt_xyzi, int n_tgt, double max_corr_dist, int max_iterations,
               double transformation_epsilon, double fitness_epsilon, int exact_sums, float *T16, double *fitness, int *converged,
               int *state_out, double *trace) {
  vector<P4> src(reinterpret_cast<const P4 *>(src_xyzi), reinterpret_cast<const P4 *>(src_xyzi) + n_src);
  const vector<P4> src0 = src;
  vector<P4> tgt(reinterpret_cast<const P4 *>(tgt_xyzi), reinterpret_cast<const P4 *>(tgt_xyzi) + n_tgt);
  KdTree tree;
  tree.build(tgt);
  float Tfin[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const double max_d2 = max_corr_dist * max_corr_dist;
  double prev_mse = std::numeric_limits<double>::max();
  int iterations = 0, state = 0;  // 0 not converged, 1 iterations, 2 transform, 3 abs mse, 4 rel mse, 5 no correspondences
  vector<int> cq, cm;
  vector<float> cd;
  while (state == 0) {
    cq.clear(); cm.clear(); cd.clear();
    for (int i = 0; i < n_src; ++i) {
      int idx; float d;
      if (tree.knn(src[i], 1, &idx, &d) < 1) continue;
      if ((double)d > max_d2) continue;
      cq.push_back(i); cm.push_back(idx); cd.push_back(d);
    }
    const int n = (int)cq.size();
    if (n < 3) { state = 5; break; }
    // Eigen::umeyama(src, dst, false), float
    const float one_over_n = 1.f / (float)n;
    float sm[3], dm[3], sigma[9];
    if (!exact_sums) {
      float ss[3] = {0, 0, 0}, ds[3] = {0, 0, 0};
      for (int k = 0; k < n; ++k) {
        const P4 &a = src[cq[k]], &b = tgt[cm[k]];
        ss[0] += a.x; ss[1] += a.y; ss[2] += a.z;
        ds[0] += b.x; ds[1] += b.y; ds[2] += b.z;
      }
      for (int c = 0; c < 3; ++c) { sm[c] = ss[c] * one_over_n; dm[c] = ds[c] * one_over_n; }
      float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < n; ++k) {
        const P4 &a = src[cq[k]], &b = tgt[cm[k]];
        const float sd[3] = {a.x - sm[0], a.y - sm[1], a.z - sm[2]}, dd[3] = {b.x - dm[0], b.y - dm[1], b.z - dm[2]};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) acc[r * 3 + c] += dd[r] * sd[c];
      }
      for (int k = 0; k < 9; ++k) sigma[k] = one_over_n * acc[k];
    } else {
      double ss[3] = {0, 0, 0}, ds[3] = {0, 0, 0}, cr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < n; ++k) {
        const P4 &a = src[cq[k]], &b = tgt[cm[k]];
        const double s3[3] = {a.x, a.y, a.z}, d3[3] = {b.x, b.y, b.z};
        for (int r = 0; r < 3; ++r) {
          ss[r] += s3[r]; ds[r] += d3[r];
          for (int c = 0; c < 3; ++c) cr[r * 3 + c] += d3[r] * s3[c];
        }
      }
      for (int c = 0; c < 3; ++c) { sm[c] = (float)ss[c] * one_over_n; dm[c] = (float)ds[c] * one_over_n; }
      // sum (d - dm)(s - sm)^T with the float means, exactly
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          sigma[r * 3 + c] = one_over_n * (float)(cr[r * 3 + c] - (double)dm[r] * ss[c] - ds[r] * (double)sm[c] + (double)n * (double)dm[r] * (double)sm[c]);
    }
    float R[9], T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    svd_rotation(sigma, R);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
      T[r * 4 + 3] = dm[r] - ((R[r * 3] * sm[0] + R[r * 3 + 1] * sm[1]) + R[r * 3 + 2] * sm[2]);
    }
    // transformCloud (in place) and final_transformation_ = transformation_ * final_transformation_
    for (int i = 0; i < n_src; ++i) {
      const P4 p = src[i];
      src[i].x = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3];
      src[i].y = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7];
      src[i].z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11];
    }
    float Tn[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) Tn[r * 4 + c] = ((T[r * 4] * Tfin[c] + T[r * 4 + 1] * Tfin[4 + c]) + T[r * 4 + 2] * Tfin[8 + c]) + T[r * 4 + 3] * Tfin[12 + c];
    std::memcpy(Tfin, Tn, sizeof Tn);
    double mse = 0;
    for (int k = 0; k < n; ++k) mse += cd[k];
    mse /= (double)n;
    if (trace) {
      double *tr = trace + (size_t)iterations * 14;
      tr[0] = n; tr[1] = mse;
      for (int k = 0; k < 9; ++k) tr[2 + k] = R[k];
      for (int k = 0; k < 3; ++k) tr[11 + k] = T[k * 4 + 3];
    }
    ++iterations;
    // DefaultConvergenceCriteria::hasConverged
    if (iterations >= max_iterations) { state = 1; break; }
    const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1);
    const double translation_sqr = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
    if (cos_angle >= 1.0 - transformation_epsilon && translation_sqr <= transformation_epsilon) { state = 2; break; }
    if (std::fabs(mse - prev_mse) < 1e-12) { state = 3; break; }
    if (std::fabs(mse - prev_mse) / prev_mse < fitness_epsilon) { state = 4; break; }
    prev_mse = mse;
  }
  // getFitnessScore (max_range = DBL_MAX)
  double fs = 0;
  int nr = 0;
  for (int i = 0; i < n_src; ++i) {
    const P4 p = src0[i];
    P4 q = p;
    q.x = ((Tfin[0] * p.x + Tfin[1] * p.y) + Tfin[2] * p.z) + Tfin[3];
    q.y = ((Tfin[4] * p.x + Tfin[5] * p.y) + Tfin[6] * p.z) + Tfin[7];
    q.z = ((Tfin[8] * p.x + Tfin[9] * p.y) + Tfin[10] * p.z) + Tfin[11];
    int idx; float d;
    if (tree.knn(q, 1, &idx, &d) < 1) continue;
    fs += d; ++nr;
  }
  if (fitness) *fitness = nr > 0 ? fs / nr : std::numeric_limits<double>::max();
  if (converged) *converged = (state >= 1 && state <= 4) ? 1 : 0;
  if (state_out) *state_out = state;
  if (T16) std::memcpy(T16, Tfin, sizeof Tfin);
  return iterations;
}

// k-NN of nq queries against n points; idx [nq][k], dist [nq][k]; brute!=0 uses the O(n) scan
int oracle_knn(const float *pts, int n, const float *q, int nq, int k, int brute, int32_t *idx, float *dist) {
  if (k < 1 || k > 8) return ALEGO_BAD_ARG;
  vector<P4> P(reinterpret_cast<const P4 *>(pts), reinterpret_cast<const P4 *>(pts) + n);
  const P4 *Q = reinterpret_cast<const P4 *>(q);
  if (!brute) {
    KdTree t;
    t.build(P);
    for (int a = 0; a < nq; ++a) t.knn(Q[a], k, idx + (size_t)a * k, dist + (size_t)a * k);
  } else {
    for (int a = 0; a < nq; ++a) {
      vector<KdTree::Cand> c(n);
      for (int b = 0; b < n; ++b) c[b] = {l2f(Q[a], P[b]), b};
      std::partial_sort(c.begin(), c.begin() + std::min(k, n), c.end(), KdTree::closer);
      for (int t = 0; t < k; ++t) {
        idx[(size_t)a * k + t] = t < n ? c[t].idx : -1;
        dist[(size_t)a * k + t] = t < n ? c[t].d : std::numeric_limits<float>::infinity();
      }
    }
  }
  return ALEGO_OK;
}
// residual + Jacobian of one cost function; f: 14 doubles (kind, cp, a, b, c, d)
int oracle_eval_residual(const double *f14, const double *x, double *r, double *J6) {
  Resid f{};
  f.kind = (int)f14[0];
  for (int c = 0; c < 3; ++c) { f.cp[c] = f14[1 + c]; f.a[c] = f14[4 + c]; f.b[c] = f14[7 + c]; f.c[c] = f14[10 + c]; }
  f.d = f14[13];
  PoseTrig T(x);
  eval_resid(f, x, T, r, J6);
  return ALEGO_OK;
}
// solve a problem given as n x 14 doubles; x in/out; returns iterations; summary: initial, final cost
int oracle_solve(const double *f14, int n, double *x, int max_iters, double huber, double *summary4) {
  vector<Resid> rs(n);
  for (int k = 0; k < n; ++k) {
    const double *f = f14 + (size_t)k * 14;
    rs[k].kind = (int)f[0];
    for (int c = 0; c < 3; ++c) { rs[k].cp[c] = f[1 + c]; rs[k].a[c] = f[4 + c]; rs[k].b[c] = f[7 + c]; rs[k].c[c] = f[10 + c]; }
    rs[k].d = f[13];
  }
  SolveSummary s;
  ceres_like_solve(rs, x, max_iters, huber, &s, nullptr);
  if (summary4) { summary4[0] = s.initial_cost; summary4[1] = s.final_cost; summary4[2] = s.successful; summary4[3] = s.termination; }
  return s.iterations;
}
int oracle_eig3(const double *A9, double *w3, double *V9) { eig3_sym(A9, w3, V9); return ALEGO_OK; }
int oracle_lstsq5x3(const double *A15, const double *b5, double *n3) { lstsq_5x3(A15, b5, n3); return ALEGO_OK; }
// std::sort with a key-only comparator on index ranges, exactly the call at laserOdometry.cpp:185
int oracle_std_sort_by_key(const double *key, int32_t *idx, int n) {
  std::sort(idx, idx + n, [key](int a, int b) { return key[a] < key[b]; });
  return ALEGO_OK;
}
}  // extern "C"
