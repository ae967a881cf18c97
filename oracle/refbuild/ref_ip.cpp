// oracle/_ref driver for ImageProjection — TEST INFRASTRUCTURE.
// Compiles /root/reference/src/imageProjection.cpp UNMODIFIED (included below from where it lies; never copied into this repo)
// against the stand-in headers in shims/, and drives loam::ImageProjection::onInit() / pcCB() (imageProjection.cpp:6-208, 210-316)
// synchronously.  pcCB clears its images before it returns (:197-205), so everything is captured from inside the shim
// ros::Publisher::publish() of "/seg_info" (:320-323), i.e. with the reference's own state, before the reset.
// Built with -fno-access-control so that the private members can be read; no reference source line is altered.
#include "src/imageProjection.cpp"

#include "ref_common.hpp"

namespace {
struct RefIp {
  loam::ImageProjection node;
  alego_ref::Blobs out;
};

void capture(RefIp *h) {
  loam::ImageProjection &n = h->node;
  const int R = N_SCAN, C = Horizon_SCAN;
  std::vector<double> range(static_cast<std::size_t>(R) * C);
  std::vector<int32_t> label(range.size());
  std::vector<uint8_t> ground(range.size());
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) {
      range[static_cast<std::size_t>(i) * C + j] = n.range_mat_(i, j);
      label[static_cast<std::size_t>(i) * C + j] = n.label_mat_(i, j);
      ground[static_cast<std::size_t>(i) * C + j] = n.ground_mat_(i, j) ? 1 : 0;
    }
  h->out.put("range_mat", range);
  h->out.put("label_mat", label);
  h->out.put("ground_mat", ground);
  h->out.put("full_cloud", alego_ref::cloud_xyzi(*n.full_cloud_));
  h->out.put("segmented_cloud", alego_ref::cloud_xyzi(*n.segmented_cloud_));
  h->out.put("outlier_cloud", alego_ref::cloud_xyzi(*n.outlier_cloud_));
  const alego::cloud_info &s = *n.seg_info_msg_;
  const std::size_t M = n.segmented_cloud_->points.size();
  h->out.put("startRingIndex", s.startRingIndex);
  h->out.put("endRingIndex", s.endRingIndex);
  h->out.put("segmentedCloudGroundFlag", s.segmentedCloudGroundFlag.data(), M);
  h->out.put("segmentedCloudColInd", s.segmentedCloudColInd.data(), M);
  h->out.put("segmentedCloudRange", s.segmentedCloudRange.data(), M);
  h->out.put1("startOrientation", s.startOrientation);
  h->out.put1("endOrientation", s.endOrientation);
  h->out.put1("orientationDiff", s.orientationDiff);
  h->out.put1("label_cnt", static_cast<int32_t>(n.label_cnt_));
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

// the reference's compile-time sensor constants (include/alego/utility.h:50-65) as this library was built with them
int ref_ip_constants(double *out9) {
  out9[0] = N_SCAN; out9[1] = Horizon_SCAN; out9[2] = ground_scan_id; out9[3] = ang_res_x; out9[4] = ang_res_y;
  out9[5] = ang_bottom; out9[6] = sensor_mount_ang; out9[7] = seg_theta; out9[8] = scan_period;
  return 0;
}

void *ref_ip_create() {
  RefIp *h = new RefIp;
  h->node.onInit();
  return h;
}
void ref_ip_destroy(void *h) { delete static_cast<RefIp *>(h); }

// one sweep: n points x, y, z, intensity (NaNs allowed: the message is flagged not dense, so removeNaNFromPointCloud filters them)
int ref_ip_process(void *hv, const float *xyzi, int n) {
  RefIp *h = static_cast<RefIp *>(hv);
  if (n <= 0) return -1;
  sensor_msgs::PointCloud2Ptr msg(new sensor_msgs::PointCloud2);
  msg->xyzi.assign(xyzi, xyzi + static_cast<std::size_t>(n) * 4);
  msg->width = n;
  msg->is_dense = false;
  bool all_nan = true;
  for (int k = 0; k < n && all_nan; ++k)
    all_nan = !(std::isfinite(xyzi[4 * k]) && std::isfinite(xyzi[4 * k + 1]) && std::isfinite(xyzi[4 * k + 2]));
  if (all_nan) return -1;  // pcCB would index points[0] of an empty cloud (:62)
  h->out.m.clear();
  alego_ref::bus().hook = [h](const std::string &topic) { if (topic == "/seg_info") capture(h); };
  h->node.pcCB(msg);
  alego_ref::bus().hook = nullptr;
  return h->out.m.empty() ? -2 : 0;
}

int64_t ref_ip_get(void *h, const char *name, void *dst, size_t cap) { return static_cast<RefIp *>(h)->out.get(name, dst, cap); }

#pragma GCC visibility pop
}  // extern "C"
