// oracle/_ref driver for LaserOdometry — TEST INFRASTRUCTURE.
// Compiles /root/reference/src/laserOdometry.cpp UNMODIFIED (included from where it lies) against the stand-in headers in shims/ and
// runs loam::LaserOdometry::mainLoop() (laserOdometry.cpp:79-555: calculateSmoothness, markOccludedPoints, extractFeatures,
// per-ring VoxelGrid, scan-to-scan association + the two Ceres solves, pose integration) one sweep at a time:
//   * the sweep enters through the reference's own handlers segCloudHandler / segInfoHandler / outlierHandler (:742-759);
//   * the shim ros::ok() is true exactly while those buffers are non-empty, so each mainLoop() call processes one sweep and returns;
//   * results are read from the private members (-fno-access-control) and from the messages the loop publishes.
// onInit() (:6-77) is NOT called: besides ROS wiring and the member initialisation repeated in ref_lo_create() it spawns mainLoop on a
// function-local static std::thread, which would race with this driver and abort at exit.  KdTreeFLANN / VoxelGrid / Ceres are the
// restated third-party stand-ins (see shims/); everything else that executes is the reference's code.
#include "src/laserOdometry.cpp"

#include "ref_common.hpp"

namespace {
struct RefLo {
  loam::LaserOdometry node;
  alego_ref::Blobs out;
  double stamp = 100.0;
  int frames = 0;
};

template <typename M>
std::shared_ptr<const M> last_msg(const char *topic) {
  auto &b = alego_ref::bus().last;
  auto it = b.find(topic);
  if (it == b.end()) return nullptr;
  return std::static_pointer_cast<const M>(it->second);
}
void put_cloud_msg(alego_ref::Blobs &out, const char *key, const char *topic) {
  auto m = last_msg<sensor_msgs::PointCloud2>(topic);
  if (m) out.put(key, m->xyzi); else out.put<float>(key, nullptr, 0);
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int ref_lo_constants(double *out9) {
  out9[0] = N_SCAN; out9[1] = Horizon_SCAN; out9[2] = ground_scan_id; out9[3] = ang_res_x; out9[4] = ang_res_y;
  out9[5] = ang_bottom; out9[6] = nearest_feature_dist; out9[7] = scan_period; out9[8] = imu_queue_length;
  return 0;
}

void *ref_lo_create() {
  RefLo *h = new RefLo;
  loam::LaserOdometry &n = h->node;
  // member initialisation of LaserOdometry::onInit (laserOdometry.cpp:17-47)
  n.imu_ptr_front_ = n.odom_ptr_front_ = 0;
  n.imu_ptr_last_ = n.odom_ptr_last_ = -1;
  n.imu_ptr_last_iter_ = n.odom_ptr_last_iter_ = 0;
  n.imu_time_.fill(0); n.imu_roll_.fill(0); n.imu_pitch_.fill(0); n.imu_yaw_.fill(0);
  n.imu_velo_x_.fill(0); n.imu_velo_y_.fill(0); n.imu_velo_z_.fill(0);
  n.imu_shift_x_.fill(0); n.imu_shift_y_.fill(0); n.imu_shift_z_.fill(0);
  n.cloud_curvature_.fill(0); n.cloud_neighbor_picked_.fill(false); n.cloud_label_.fill(0); n.cloud_sort_idx_.fill(0);
  n.system_initialized_ = false;
  n.surf_last_.reset(new PointCloudT);
  n.corner_last_.reset(new PointCloudT);
  n.outlier_last_.reset(new PointCloudT);
  n.kd_surf_last_.reset(new pcl::KdTreeFLANN<PointT>);
  n.kd_corner_last_.reset(new pcl::KdTreeFLANN<PointT>);
  for (int i = 0; i < 6; ++i) n.params_[i] = 0.;
  n.t_w_cur_.setZero();
  n.r_w_cur_.setIdentity();
  // publishers as advertised in onInit (:52-60)
  n.pub_corner_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/corner", 10);
  n.pub_corner_less_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/corner_less", 10);
  n.pub_surf_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/surf", 10);
  n.pub_surf_less_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/surf_less", 10);
  n.pub_undistorted_pc_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/undistorted", 10);
  n.pub_odom_ = n.nh_.advertise<nav_msgs::Odometry>("/odom/lidar", 10);
  n.pub_surf_last_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/surf_last", 10);
  n.pub_corner_last_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/corner_last", 10);
  n.pub_outlier_last_ = n.nh_.advertise<sensor_msgs::PointCloud2>("/outlier_last", 10);
  return h;
}
void ref_lo_destroy(void *h) { delete static_cast<RefLo *>(h); }

void ref_lo_set_params(void *hv, const double *p6) { std::memcpy(static_cast<RefLo *>(hv)->node.params_, p6, 6 * sizeof(double)); }

// One sweep = what ImageProjection publishes: the segmented cloud (M points), cloud_info (ring indices [N_SCAN], the three
// per-point arrays [M]) and the outlier cloud.  Returns 0, or -1 on bad sizes.
int ref_lo_process(void *hv, const float *seg_xyzi, int M, const int32_t *start_ring, const int32_t *end_ring, const uint8_t *ground,
                   const int32_t *col, const float *range, float start_ori, float end_ori, float ori_diff, const float *outlier_xyzi,
                   int n_outlier) {
  RefLo *h = static_cast<RefLo *>(hv);
  loam::LaserOdometry &n = h->node;
  if (M < 0 || M > N_SCAN * Horizon_SCAN) return -1;
  h->stamp += 0.1;
  sensor_msgs::PointCloud2Ptr seg(new sensor_msgs::PointCloud2), outl(new sensor_msgs::PointCloud2);
  seg->xyzi.assign(seg_xyzi, seg_xyzi + static_cast<std::size_t>(M) * 4);
  seg->width = M;
  outl->xyzi.assign(outlier_xyzi, outlier_xyzi + static_cast<std::size_t>(n_outlier) * 4);
  outl->width = n_outlier;
  alego::cloud_infoPtr info(new alego::cloud_info);
  info->startRingIndex.assign(start_ring, start_ring + N_SCAN);
  info->endRingIndex.assign(end_ring, end_ring + N_SCAN);
  // fixed-size arrays like ImageProjection::onInit sizes them (imageProjection.cpp:18-20); only the first M entries are meaningful
  info->segmentedCloudGroundFlag.assign(static_cast<std::size_t>(N_SCAN) * Horizon_SCAN, 0);
  info->segmentedCloudColInd.assign(static_cast<std::size_t>(N_SCAN) * Horizon_SCAN, 0);
  info->segmentedCloudRange.assign(static_cast<std::size_t>(N_SCAN) * Horizon_SCAN, 0);
  std::copy(ground, ground + M, info->segmentedCloudGroundFlag.begin());
  std::copy(col, col + M, info->segmentedCloudColInd.begin());
  std::copy(range, range + M, info->segmentedCloudRange.begin());
  info->startOrientation = start_ori; info->endOrientation = end_ori; info->orientationDiff = ori_diff;
  seg->header.stamp.fromSec(h->stamp);
  outl->header.stamp.fromSec(h->stamp);
  info->header.stamp.fromSec(h->stamp);
  n.segCloudHandler(seg);
  n.segInfoHandler(info);
  n.outlierHandler(outl);

  alego_ref::Bus &bus = alego_ref::bus();
  bus.last.clear();
  ceres::solve_log().clear();
  alego_ref::MuteCout mute;
  bus.ok_fn = [&n]() { return !n.seg_cloud_buf_.empty(); };
  n.mainLoop();
  bus.ok_fn = nullptr;
  ++h->frames;

  alego_ref::Blobs &out = h->out;
  out.m.clear();
  std::vector<uint8_t> picked(M);
  for (int i = 0; i < M; ++i) picked[i] = n.cloud_neighbor_picked_[i] ? 1 : 0;
  out.put("cloud_curvature", n.cloud_curvature_.data(), M);
  out.put("cloud_neighbor_picked", picked);
  out.put("cloud_label", n.cloud_label_.data(), M);
  out.put("cloud_sort_idx", n.cloud_sort_idx_.data(), M);
  put_cloud_msg(out, "sharp", "/corner");
  put_cloud_msg(out, "less_sharp", "/corner_less");
  put_cloud_msg(out, "flat", "/surf");
  put_cloud_msg(out, "less_flat", "/surf_less");
  put_cloud_msg(out, "surf_last", "/surf_last");
  put_cloud_msg(out, "corner_last", "/corner_last");
  out.put("lo_params", n.params_, 6);
  out.put("t_w_cur", n.t_w_cur_.data(), 3);
  double r[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[3 * i + j] = n.r_w_cur_(i, j);
  out.put("r_w_cur", r, 9);
  if (auto od = last_msg<nav_msgs::Odometry>("/odom/lidar")) {
    const double v[7] = {od->pose.pose.position.x, od->pose.pose.position.y, od->pose.pose.position.z, od->pose.pose.orientation.w,
                         od->pose.pose.orientation.x, od->pose.pose.orientation.y, od->pose.pose.orientation.z};
    out.put("odom_lidar", v, 7);
  }
  // iteration traces of the Ceres stand-in: per solve, rows of (cost, x[6]) — record 0 = the starting point
  std::vector<double> trace;
  std::vector<int32_t> iters;
  for (const ceres::Solver::Summary &s : ceres::solve_log()) {
    iters.push_back(s.num_iterations);
    for (std::size_t k = 0; k < s.trace_cost.size(); ++k) {
      trace.push_back(s.trace_cost[k]);
      trace.insert(trace.end(), s.trace_x[k].begin(), s.trace_x[k].end());
    }
  }
  out.put("lo_trace", trace);
  out.put("lo_solve_iterations", iters);
  return 0;
}

// LaserOdometry::imuHandler (:761-804) — feeds the IMU ring buffer used by adjustDistortion
void ref_lo_imu(void *hv, double stamp, const double *quat_xyzw, const double *ang_vel, const double *lin_acc) {
  sensor_msgs::ImuPtr m(new sensor_msgs::Imu);
  m->header.stamp.fromSec(stamp);
  m->orientation.x = quat_xyzw[0]; m->orientation.y = quat_xyzw[1]; m->orientation.z = quat_xyzw[2]; m->orientation.w = quat_xyzw[3];
  m->angular_velocity.x = ang_vel[0]; m->angular_velocity.y = ang_vel[1]; m->angular_velocity.z = ang_vel[2];
  m->linear_acceleration.x = lin_acc[0]; m->linear_acceleration.y = lin_acc[1]; m->linear_acceleration.z = lin_acc[2];
  static_cast<RefLo *>(hv)->node.imuHandler(m);
}

// the IMU ring buffer as imuHandler left it: 10 rows (time, roll, pitch, yaw, shift xyz, velocity xyz) x imu_queue_length, then
// imu_ptr_front_, imu_ptr_last_, imu_ptr_last_iter_
void ref_lo_imu_state(void *hv, double *queue10xN, int32_t *ptrs3) {
  loam::LaserOdometry &n = static_cast<RefLo *>(hv)->node;
  const std::array<double, imu_queue_length> *rows[10] = {&n.imu_time_, &n.imu_roll_, &n.imu_pitch_, &n.imu_yaw_, &n.imu_shift_x_,
                                                          &n.imu_shift_y_, &n.imu_shift_z_, &n.imu_velo_x_, &n.imu_velo_y_, &n.imu_velo_z_};
  for (int r = 0; r < 10; ++r) std::memcpy(queue10xN + static_cast<std::size_t>(r) * imu_queue_length, rows[r]->data(), sizeof(double) * imu_queue_length);
  ptrs3[0] = n.imu_ptr_front_; ptrs3[1] = n.imu_ptr_last_; ptrs3[2] = n.imu_ptr_last_iter_;
}

// LaserOdometry::adjustDistortion (:557-726; the call in mainLoop is commented out, :115) on a cloud + cloud_info pushed the same way
int ref_lo_adjust_distortion(void *hv, float *xyzi, int M, const int32_t *col, float start_ori, float end_ori, double scan_time) {
  loam::LaserOdometry &n = static_cast<RefLo *>(hv)->node;
  alego::cloud_infoPtr info(new alego::cloud_info);
  info->segmentedCloudColInd.assign(static_cast<std::size_t>(N_SCAN) * Horizon_SCAN, 0);
  std::copy(col, col + M, info->segmentedCloudColInd.begin());
  info->startOrientation = start_ori; info->endOrientation = end_ori;
  while (!n.seg_info_buf_.empty()) n.seg_info_buf_.pop();
  n.seg_info_buf_.push(info);
  PointCloudT::Ptr cloud(new PointCloudT);
  cloud->points.resize(M);
  for (int i = 0; i < M; ++i) { PointT p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; cloud->points[i] = p; }
  n.adjustDistortion(cloud, scan_time);
  for (int i = 0; i < M; ++i) { xyzi[4 * i] = cloud->points[i].x; xyzi[4 * i + 1] = cloud->points[i].y; xyzi[4 * i + 2] = cloud->points[i].z; }
  n.seg_info_buf_.pop();
  return n.imu_ptr_last_iter_;
}

int64_t ref_lo_get(void *h, const char *name, void *dst, size_t cap) { return static_cast<RefLo *>(h)->out.get(name, dst, cap); }

#pragma GCC visibility pop
}  // extern "C"
