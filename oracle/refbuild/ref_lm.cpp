// oracle/_ref driver for LaserMapping — TEST INFRASTRUCTURE.
// Compiles /root/reference/src/laserMapping.cpp UNMODIFIED (included from where it lies) against the stand-in headers in shims/.
// Two ways in:
//   ref_lm_scan2map  sets the local map (corner_from_map_ds_ / surf_from_map_ds_, what extractSurroundingKeyFrames leaves,
//                    laserMapping.cpp:316-319) and the sweep clouds directly, feeds the odometry pose through laserOdomHandler
//                    (:154-187) and calls the reference's downsampleCurrentScan (:325-346), scan2MapOptimization (:348-479) and
//                    transformUpdate (:481-489) in mainLoop's order (:116-123) — SURVEY.md §8 rows a8, a15-a21 on a given map.
//   ref_lm_frame     pushes one synchronised input set through the reference's four handlers (:133-187) and runs mainLoop()
//                    (:102-131) until it has consumed it: transformAssociateToMap, extractSurroundingKeyFrames,
//                    downsampleCurrentScan, scan2MapOptimization, saveKeyFramesAndFactor, correctPoses, transformUpdate, publish
//                    on every 2nd call (static frame_cnt, :111-126), local map from the reference's own keyframes (row N1).
// onInit() IS called (it carries the leaf sizes, keyframe distance and window length, :37-49): ros::ok() is false at that moment so
// the mainLoop / loopClosure threads it spawns return at once; the visualisation thread (whose `while (ros::ok)` tests a function
// pointer, :599) parks inside the shim ros::Rate::sleep() and never touches the node again.  The node is leaked on purpose
// (std::thread members).  KdTreeFLANN / VoxelGrid / Ceres / Eigen's eigen-solver + QR / GTSAM are restated stand-ins (see shims/;
// the ISAM2 stand-in returns the inserted initial values, exact for a chain without loop closures).
#include "src/laserMapping.cpp"

#include "ref_common.hpp"

namespace {
struct RefLm {
  loam::LaserMapping *node = nullptr;
  alego_ref::Blobs out;
};

PointCloudT::Ptr make_cloud(const float *xyzi, int n) {
  PointCloudT::Ptr c(new PointCloudT);
  c->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointT p;
    p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
    c->points[i] = p;
  }
  c->width = n; c->height = 1;
  return c;
}
sensor_msgs::PointCloud2Ptr make_msg(const float *xyzi, int n, double stamp) {
  sensor_msgs::PointCloud2Ptr m(new sensor_msgs::PointCloud2);
  m->xyzi.assign(xyzi, xyzi + static_cast<std::size_t>(n) * 4);
  m->width = n;
  m->header.stamp.fromSec(stamp);
  return m;
}
nav_msgs::OdometryPtr make_odom(const double *t, const double *q_wxyz, double stamp) {
  nav_msgs::OdometryPtr od(new nav_msgs::Odometry);
  od->header.stamp.fromSec(stamp);
  od->pose.pose.position.x = t[0]; od->pose.pose.position.y = t[1]; od->pose.pose.position.z = t[2];
  od->pose.pose.orientation.w = q_wxyz[0]; od->pose.pose.orientation.x = q_wxyz[1];
  od->pose.pose.orientation.y = q_wxyz[2]; od->pose.pose.orientation.z = q_wxyz[3];
  return od;
}

void capture(RefLm *h) {
  loam::LaserMapping &n = *h->node;
  alego_ref::Blobs &out = h->out;
  out.put("lm_params", n.params_, 6);
  out.put("lm_corner_ds", alego_ref::cloud_xyzi(*n.laser_corner_ds_));
  out.put("lm_surf_ds", alego_ref::cloud_xyzi(*n.laser_surf_ds_));
  out.put("lm_outlier_ds", alego_ref::cloud_xyzi(*n.laser_outlier_ds_));
  out.put("lm_surf_total_ds", alego_ref::cloud_xyzi(*n.laser_surf_total_ds_));
  out.put("corner_from_map_ds", alego_ref::cloud_xyzi(*n.corner_from_map_ds_));
  out.put("surf_from_map_ds", alego_ref::cloud_xyzi(*n.surf_from_map_ds_));
  out.put("t_map2laser", n.t_map2laser_.data(), 3);
  out.put("t_map2odom", n.t_map2odom_.data(), 3);
  const double q1[4] = {n.q_map2laser_.w(), n.q_map2laser_.x(), n.q_map2laser_.y(), n.q_map2laser_.z()};
  const double q2[4] = {n.q_map2odom_.w(), n.q_map2odom_.x(), n.q_map2odom_.y(), n.q_map2odom_.z()};
  out.put("q_map2laser", q1, 4);
  out.put("q_map2odom", q2, 4);
  std::vector<double> kp;
  for (const PointTypePose &p : n.cloud_keyposes_6d_->points) {
    const double v[7] = {p.x, p.y, p.z, p.roll, p.pitch, p.yaw, p.time};
    kp.insert(kp.end(), v, v + 7);
  }
  out.put("keyposes_6d", kp);
  out.put1("n_keyframes", static_cast<int32_t>(n.cloud_keyposes_3d_->points.size()));
  std::vector<double> trace;
  std::vector<int32_t> iters, blocks;
  for (const ceres::Solver::Summary &s : ceres::solve_log()) {
    iters.push_back(s.num_iterations);
    for (std::size_t k = 0; k < s.trace_cost.size(); ++k) {
      trace.push_back(s.trace_cost[k]);
      trace.insert(trace.end(), s.trace_x[k].begin(), s.trace_x[k].end());
    }
  }
  out.put("lm_trace", trace);
  out.put("lm_solve_iterations", iters);
  // the reference's own TicToc figures (laserMapping.cpp:344, 358, 464, 476): per scan2MapOptimization call
  // [downsample ms, kd-tree build ms, association ms (sum over outer iterations), solver ms (sum)]
  double tm[4] = {0, 0, 0, 0};
  for (const std::string &l : alego_ref::bus().log) {
    double v = 0;
    if (std::sscanf(l.c_str(), "downsampleCurrentScan: %lf", &v) == 1) tm[0] += v;
    else if (std::sscanf(l.c_str(), "build kdtree time: %lf", &v) == 1) tm[1] += v;
    else if (std::sscanf(l.c_str(), "mapping data assosiation time %lf", &v) == 1) tm[2] += v;
    else if (std::sscanf(l.c_str(), "mapping solver time %lf", &v) == 1) tm[3] += v;
  }
  out.put("lm_timing_ms", tm, 4);
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int ref_lm_constants(double *out9) {
  out9[0] = N_SCAN; out9[1] = Horizon_SCAN; out9[2] = ground_scan_id; out9[3] = ang_res_x; out9[4] = ang_res_y;
  out9[5] = ang_bottom; out9[6] = 0; out9[7] = 0; out9[8] = 0;
  return 0;
}

void *ref_lm_create() {
  RefLm *h = new RefLm;
  alego_ref::Globals &gl = alego_ref::globals();
  // library-wide and for good (nodes may be created from several threads at once): on a thread without ok_fn — i.e. on the worker
  // threads onInit spawns — ros::ok() is false, so mainLoop / loopClosureThread return immediately; the visualisation thread,
  // which tests the function pointer, parks in ros::Rate::sleep(); nothing is "subscribed" (the parked thread never publishes).
  // The driver's own calls into mainLoop set ok_fn on their thread.
  gl.ok = false;
  gl.park_sleepers = true;
  gl.subscribers = 0;
  h->node = new loam::LaserMapping;  // leaked on purpose, see header
  h->node->onInit();
  h->node->main_thread_.join();
  h->node->loop_thread_.join();
  // the reference leaves these uninitialised until the first message (laserMapping.h:93-100; onInit sets new_laser_corner_ twice
  // and never new_laser_odom_, :36)
  h->node->new_laser_odom_ = false;
  h->node->time_laser_corner_ = h->node->time_laser_surf_ = h->node->time_laser_outlier_ = h->node->time_laser_odom_ = 0;
  return h;
}
void ref_lm_destroy(void *h) { delete static_cast<RefLm *>(h); }

void ref_lm_set_params(void *hv, const double *p6) { std::memcpy(static_cast<RefLm *>(hv)->node->params_, p6, 6 * sizeof(double)); }

int ref_lm_scan2map(void *hv, const float *map_corner, int nmc, const float *map_surf, int nms, const float *corner, int nc,
                    const float *surf, int ns, const float *outlier, int no, const double *t_odom, const double *q_odom_wxyz,
                    const double *params_or_null) {
  RefLm *h = static_cast<RefLm *>(hv);
  loam::LaserMapping &n = *h->node;
  if (params_or_null) std::memcpy(n.params_, params_or_null, 6 * sizeof(double));
  n.corner_from_map_ds_ = make_cloud(map_corner, nmc);
  n.surf_from_map_ds_ = make_cloud(map_surf, nms);
  n.laser_corner_ = make_cloud(corner, nc);
  n.laser_surf_ = make_cloud(surf, ns);
  n.laser_outlier_ = make_cloud(outlier, no);
  n.laserOdomHandler(make_odom(t_odom, q_odom_wxyz, 0.0));
  ceres::solve_log().clear();
  alego_ref::bus().log.clear();
  alego_ref::bus().capture_log = true;
  alego_ref::MuteCout mute;
  n.transformAssociateToMap();
  n.downsampleCurrentScan();
  n.scan2MapOptimization();
  n.transformUpdate();
  h->out.m.clear();
  capture(h);
  return 0;
}

int ref_lm_frame(void *hv, const float *corner, int nc, const float *surf, int ns, const float *outlier, int no, const double *t_odom,
                 const double *q_odom_wxyz, double stamp) {
  RefLm *h = static_cast<RefLm *>(hv);
  loam::LaserMapping &n = *h->node;
  n.surfLastHandler(make_msg(surf, ns, stamp));
  n.cornerLastHandler(make_msg(corner, nc, stamp));
  n.outlierLastHandler(make_msg(outlier, no, stamp));
  n.laserOdomHandler(make_odom(t_odom, q_odom_wxyz, stamp));
  ceres::solve_log().clear();
  alego_ref::bus().log.clear();
  alego_ref::bus().capture_log = true;
  alego_ref::MuteCout mute;
  alego_ref::Bus &bus = alego_ref::bus();
  bus.ok_fn = [&n]() { return n.new_laser_surf_; };  // cleared by mainLoop once the set is consumed (:110)
  n.mainLoop();
  bus.ok_fn = nullptr;
  h->out.m.clear();
  capture(h);
  return 0;
}

// the reference's four cost functions (include/alego/utility.h:122-349) evaluated directly.  f14 = kind (0 CornerCostFunction,
// 1 SurfCostFunction, 2 LidarEdgeCostFunction, 3 LidarPlaneCostFunction), cp[3], a[3] (lpj | unit normal), b[3] (lpl), c[3] (lpm),
// d (negative_OA_dot_norm) — the layout of oracle_eval_residual.
int ref_lm_eval_cost(const double *f14, const double *x6, double *r, double *J6) {
  const Eigen::Vector3d cp(f14[1], f14[2], f14[3]), a(f14[4], f14[5], f14[6]), b(f14[7], f14[8], f14[9]), c(f14[10], f14[11], f14[12]);
  ceres::CostFunction *f = nullptr;
  switch (static_cast<int>(f14[0])) {
    case 0: f = new CornerCostFunction(cp, a, b); break;
    case 1: f = new SurfCostFunction(cp, a, b, c); break;
    case 2: f = new LidarEdgeCostFunction(cp, a, b); break;
    case 3: f = new LidarPlaneCostFunction(cp, a, f14[13]); break;
    default: return -1;
  }
  const double *params[1] = {x6};
  double *jac[1] = {J6};
  const bool ok = f->Evaluate(params, r, J6 ? jac : nullptr);
  delete f;
  return ok ? 0 : -2;
}

int64_t ref_lm_get(void *h, const char *name, void *dst, size_t cap) { return static_cast<RefLm *>(h)->out.get(name, dst, cap); }

#pragma GCC visibility pop
}  // extern "C"
