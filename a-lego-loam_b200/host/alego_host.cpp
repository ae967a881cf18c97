// alego_host.cpp — see alego_host.h.  Thin: argument marshalling only, every numeric step is a C-ABI call.
#include "alego_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace alego {

AlegoContext::~AlegoContext() {
  if (h_) alego_destroy(h_);
}

int AlegoContext::init(const AlegoParams &p, int device, int n_seq, int max_points_per_scan) {
  if (h_) return ALEGO_BAD_ARG;
  p_ = p;
  n_seq_ = n_seq;
  max_pts_ = max_points_per_scan > 0 ? max_points_per_scan : p.n_scan * p.horizon_scan;
  return alego_create(&p_, device, n_seq_, max_pts_, &h_);
}

std::string AlegoContext::last_error() const { return alego_last_error(h_); }

static inline float load_f32(const uint8_t *p, bool swap) {
  uint8_t b[4];
  if (swap) { b[0] = p[3]; b[1] = p[2]; b[2] = p[1]; b[3] = p[0]; }
  else std::memcpy(b, p, 4);
  float v;
  std::memcpy(&v, b, 4);
  return v;
}

static inline bool host_is_bigendian() {
  const uint16_t one = 1;
  return *reinterpret_cast<const uint8_t *>(&one) == 0;
}

long decode_pointcloud2(const PointCloud2View &m, float *out, int stride, size_t capacity_points) {
  if (!out || (stride != 3 && stride != 4)) return -1;
  const size_t n = (size_t)m.width * m.height;
  if (n == 0) return 0;
  // all geometry in 64-bit: width * point_step and offset + 4 wrap in uint32 for crafted headers
  const uint64_t row_bytes = (uint64_t)m.width * m.point_step;
  const uint64_t row_step = m.row_step ? m.row_step : row_bytes;
  const uint64_t offs[4] = {m.off_x, m.off_y, m.off_z, (uint64_t)std::max(m.off_intensity, 0)};
  if (!m.data || n > capacity_points || row_step < row_bytes) return -1;
  for (uint64_t o : offs)
    if (o + 4 > m.point_step) return -1;
  if ((uint64_t)(m.height - 1) * row_step + row_bytes > (uint64_t)m.data_size) return -1;  // truncated / malformed message
  const bool swap = m.is_bigendian != host_is_bigendian();
  float *o = out;
  for (uint32_t r = 0; r < m.height; ++r) {
    const uint8_t *p = m.data + (size_t)(r * row_step);
    for (uint32_t c = 0; c < m.width; ++c, p += m.point_step, o += stride) {
      o[0] = load_f32(p + m.off_x, swap);
      o[1] = load_f32(p + m.off_y, swap);
      o[2] = load_f32(p + m.off_z, swap);
      if (stride == 4) o[3] = m.off_intensity >= 0 ? load_f32(p + m.off_intensity, swap) : 0.f;
    }
  }
  return (long)n;
}

size_t encode_pointcloud2_xyzi(const float *xyzi, size_t n, uint8_t *data) {
  if (!xyzi || !data) return 0;
  std::memset(data, 0, n * 32);
  for (size_t i = 0; i < n; ++i) {
    std::memcpy(data + i * 32, xyzi + i * 4, 12);
    std::memcpy(data + i * 32 + 16, xyzi + i * 4 + 3, 4);
  }
  return n * 32;
}

ImageProjection::~ImageProjection() {
  if (pinned_) alego_host_free(pinned_);
}

int ImageProjection::onInit() {
  if (!ctx_.handle()) return ALEGO_NOT_READY;
  if (pinned_) return ALEGO_OK;
  pinned_ = static_cast<float *>(alego_host_alloc(sizeof(float) * 4 * (size_t)ctx_.n_seq() * ctx_.max_points()));
  n_points_.assign(ctx_.n_seq(), 0);
  return pinned_ ? ALEGO_OK : ALEGO_CUDA_ERROR;
}

int ImageProjection::process(const std::vector<PointCloud> &clouds) {
  if (!pinned_ || (int)clouds.size() != ctx_.n_seq()) return ALEGO_BAD_ARG;
  for (int b = 0; b < ctx_.n_seq(); ++b) {
    const size_t n = clouds[b].size();
    if (n > (size_t)ctx_.max_points()) return ALEGO_BAD_ARG;
    n_points_[b] = (int32_t)n;
    if (n) std::memcpy(pinned_ + (size_t)b * ctx_.max_points() * 4, clouds[b].data(), n * sizeof(PointXYZI));
  }
  return alego_ip_process(ctx_.handle(), pinned_, n_points_.data());
}

int ImageProjection::process(const std::vector<PointCloud2View> &msgs) {
  if (!pinned_ || (int)msgs.size() != ctx_.n_seq()) return ALEGO_BAD_ARG;
  for (int b = 0; b < ctx_.n_seq(); ++b) {
    const long n = decode_pointcloud2(msgs[b], pinned_ + (size_t)b * ctx_.max_points() * 4, 4, (size_t)ctx_.max_points());
    if (n < 0) return ALEGO_BAD_ARG;
    n_points_[b] = (int32_t)n;
  }
  return alego_ip_process(ctx_.handle(), pinned_, n_points_.data());
}

int ImageProjection::results(int seq, CloudInfo *info, PointCloud *segmented, PointCloud *outlier) {
  const AlegoParams &p = ctx_.params();
  const size_t rc = (size_t)p.n_scan * p.horizon_scan;
  CloudInfo tmp;
  CloudInfo *ci = info ? info : &tmp;
  ci->startRingIndex.assign(p.n_scan, 0);
  ci->endRingIndex.assign(p.n_scan, 0);
  ci->segmentedCloudGroundFlag.assign(rc, 0);
  ci->segmentedCloudColInd.assign(rc, 0);
  ci->segmentedCloudRange.assign(rc, 0.f);
  AlegoCloudInfo raw{};
  raw.startRingIndex = ci->startRingIndex.data();
  raw.endRingIndex = ci->endRingIndex.data();
  raw.segmentedCloudGroundFlag = ci->segmentedCloudGroundFlag.data();
  raw.segmentedCloudColInd = ci->segmentedCloudColInd.data();
  raw.segmentedCloudRange = ci->segmentedCloudRange.data();
  if (segmented) segmented->assign(rc, PointXYZI{0, 0, 0, 0});
  if (outlier) outlier->assign(rc, PointXYZI{0, 0, 0, 0});
  int32_t n_out = 0;
  const int rcode = alego_ip_get(ctx_.handle(), seq, &raw, segmented ? &(*segmented)[0].x : nullptr,
                                 outlier ? &(*outlier)[0].x : nullptr, &n_out, nullptr);
  if (rcode != ALEGO_OK) return rcode;
  ci->size = raw.size;
  ci->startOrientation = raw.startOrientation;
  ci->endOrientation = raw.endOrientation;
  ci->orientationDiff = raw.orientationDiff;
  if (segmented) segmented->resize(raw.size);
  if (outlier) outlier->resize(n_out);
  return ALEGO_OK;
}

ImuQueue::ImuQueue(int length) : len_(std::max(length, 1)), a_((size_t)10 * std::max(length, 1), 0.0) {}

AlegoImuQueue ImuQueue::view() const {
  AlegoImuQueue v{};
  v.length = len_;
  v.ptr_last = ptr_last;
  v.ptr_last_iter = ptr_last_iter;
  const double *b = a_.data();
  v.time = b; v.roll = b + len_; v.pitch = b + 2 * len_; v.yaw = b + 3 * len_;
  v.shift_x = b + 4 * len_; v.shift_y = b + 5 * len_; v.shift_z = b + 6 * len_;
  v.velo_x = b + 7 * len_; v.velo_y = b + 8 * len_; v.velo_z = b + 9 * len_;
  return v;
}

// tf::Matrix3x3(tf::Quaternion).getRPY (tf/LinearMath/Matrix3x3.h: setRotation + getEulerYPR, first solution), in double
static void quaternion_to_rpy(double x, double y, double z, double w, double &roll, double &pitch, double &yaw) {
  const double d = x * x + y * y + z * z + w * w;
  const double s = 2.0 / d;
  const double xs = x * s, ys = y * s, zs = z * s;
  const double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
  const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  if (std::fabs(m20) >= 1) {  // gimbal lock
    yaw = 0;
    roll = std::atan2(m21, m22);
    pitch = m20 < 0 ? M_PI / 2.0 : -M_PI / 2.0;
  } else {
    pitch = -std::asin(m20);
    roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
    yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
  }
}

void ImuQueue::push(const ImuMsg &m) {
  double *T = a_.data(), *RO = T + len_, *PI_ = T + 2 * len_, *YA = T + 3 * len_;
  double *S[3] = {T + 4 * len_, T + 5 * len_, T + 6 * len_}, *V[3] = {T + 7 * len_, T + 8 * len_, T + 9 * len_};
  double roll, pitch, yaw;
  quaternion_to_rpy(m.qx, m.qy, m.qz, m.qw, roll, pitch, yaw);
  const double acc_x = m.ax + 9.81 * std::sin(pitch);                    // :768-770
  const double acc_y = m.ay - 9.81 * std::cos(pitch) * std::sin(roll);
  const double acc_z = m.az - 9.81 * std::cos(pitch) * std::cos(roll);
  ptr_last = (ptr_last + 1) % len_;                                      // :772-776
  if ((ptr_last + 1) % len_ == ptr_front) ptr_front = (ptr_front + 1) % len_;
  T[ptr_last] = m.stamp;
  RO[ptr_last] = roll; PI_[ptr_last] = pitch; YA[ptr_last] = yaw;
  // Eigen::Quaternionf(w, x, y, z).toRotationMatrix() * Vector3f(acc) (:784-785), float
  const float w = (float)m.qw, x = (float)m.qx, y = (float)m.qy, z = (float)m.qz;
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const float Rm[9] = {1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1.f - (txx + tyy)};
  const float af[3] = {(float)acc_x, (float)acc_y, (float)acc_z};
  float acc[3];
  for (int k = 0; k < 3; ++k) acc[k] = (Rm[k * 3] * af[0] + Rm[k * 3 + 1] * af[1]) + Rm[k * 3 + 2] * af[2];
  const int back = (ptr_last - 1 + len_) % len_;                         // :790-803
  const double time_diff = T[ptr_last] - T[back];
  if (time_diff < 1.) {
    for (int k = 0; k < 3; ++k) {
      S[k][ptr_last] = S[k][back] + V[k][back] * time_diff + acc[k] * time_diff * time_diff * 0.5;
      V[k][ptr_last] = V[k][back] + acc[k] * time_diff;
    }
  }
}

int LaserOdometry::onInit() {
  if (!ctx_.handle()) return ALEGO_NOT_READY;
  imu_.assign(ctx_.n_seq(), ImuQueue(200));
  return ALEGO_OK;
}

int LaserOdometry::imuHandler(int seq, const ImuMsg &msg) {
  if (seq < 0 || seq >= (int)imu_.size()) return ALEGO_BAD_ARG;
  imu_[seq].push(msg);
  return ALEGO_OK;
}

int LaserOdometry::adjustDistortion(const double *scan_time, int32_t *n_adjusted, double scan_period) {
  if (!scan_time || imu_.empty()) return ALEGO_BAD_ARG;
  std::vector<AlegoImuQueue> v(imu_.size());
  for (size_t b = 0; b < imu_.size(); ++b) v[b] = imu_[b].view();
  const int rc = alego_lo_adjust_distortion(ctx_.handle(), scan_time, v.data(), scan_period, n_adjusted);
  if (rc != ALEGO_OK) return rc;
  for (size_t b = 0; b < imu_.size(); ++b) imu_[b].ptr_last_iter = v[b].ptr_last_iter;  // :656
  return ALEGO_OK;
}

int LaserOdometry::process(AlegoSolveReport *reports) {
  const int rc = alego_lo_extract(ctx_.handle());
  if (rc != ALEGO_OK) return rc;
  return alego_lo_scan2scan(ctx_.handle(), reports);
}

int LaserOdometry::odometry(int seq, double params[6], double t_w_cur[3], double r_w_cur[9]) {
  return alego_lo_get_state(ctx_.handle(), seq, params, t_w_cur, r_w_cur);
}

int LaserOdometry::features(int seq, std::vector<int32_t> *sharp_idx, std::vector<int32_t> *less_sharp_idx,
                            std::vector<int32_t> *flat_idx, PointCloud *less_flat) {
  const AlegoParams &p = ctx_.params();
  const int R = p.n_scan;
  const size_t rc = (size_t)R * p.horizon_scan;
  std::vector<int32_t> s(R * 12), ls(R * 120), f(R * 24);
  PointCloud lf(less_flat ? rc : 0);
  int32_t ns = 0, nls = 0, nf = 0, nlf = 0;
  const int rcode = alego_lo_get_features(ctx_.handle(), seq, s.data(), &ns, ls.data(), &nls, f.data(), &nf,
                                          less_flat ? &lf[0].x : nullptr, &nlf, nullptr);
  if (rcode != ALEGO_OK) return rcode;
  s.resize(ns); ls.resize(nls); f.resize(nf);
  if (sharp_idx) sharp_idx->swap(s);
  if (less_sharp_idx) less_sharp_idx->swap(ls);
  if (flat_idx) flat_idx->swap(f);
  if (less_flat) { lf.resize(nlf); less_flat->swap(lf); }
  return ALEGO_OK;
}

int LaserMapping::setLocalMap(int seq, const PointCloud &corner, const PointCloud &surf) {
  return alego_lm_set_map(ctx_.handle(), seq, corner.empty() ? nullptr : &corner[0].x, (int32_t)corner.size(),
                          surf.empty() ? nullptr : &surf[0].x, (int32_t)surf.size());
}

int LaserMapping::process(AlegoSolveReport *reports) { return alego_lm_scan2map(ctx_.handle(), reports); }

int LaserMapping::saveKeyFrame(int seq, const float pose6[6]) {
  if (seq < 0 || !pose6) return ALEGO_BAD_ARG;
  if ((int)store_.size() <= seq) store_.resize(seq + 1);
  const AlegoParams &p = ctx_.params();
  const size_t cap = (size_t)p.n_scan * p.horizon_scan;
  KeyFrame kf;
  kf.corner.resize((size_t)p.n_scan * 120);
  kf.surf.resize(cap);
  kf.outlier.resize(cap);
  int32_t nc = 0, ns = 0, no = 0;
  const int rc = alego_lm_get_downsampled(ctx_.handle(), seq, &kf.corner[0].x, &nc, &kf.surf[0].x, &ns, &kf.outlier[0].x, &no, nullptr, nullptr);
  if (rc != ALEGO_OK) return rc;
  kf.corner.resize(nc); kf.surf.resize(ns); kf.outlier.resize(no);
  kf.corner.shrink_to_fit(); kf.surf.shrink_to_fit(); kf.outlier.shrink_to_fit();
  for (int q = 0; q < 6; ++q) kf.pose6[q] = pose6[q];
  std::deque<KeyFrame> &dq = store_[seq];
  if (dq.size() >= kRecentKeyframes) dq.pop_front();  // (:230-232)
  dq.push_back(std::move(kf));
  return ALEGO_OK;
}

int LaserMapping::setKeyFramePose(int seq, size_t k, const float pose6[6]) {
  if (seq < 0 || seq >= (int)store_.size() || k >= store_[seq].size() || !pose6) return ALEGO_BAD_ARG;
  for (int q = 0; q < 6; ++q) store_[seq][k].pose6[q] = pose6[q];
  return ALEGO_OK;
}

int LaserMapping::extractSurroundingKeyFrames(int seq) {
  if (seq < 0) return ALEGO_BAD_ARG;
  if ((int)store_.size() <= seq) store_.resize(seq + 1);
  const std::deque<KeyFrame> &dq = store_[seq];
  const size_t K = dq.size();
  std::vector<const float *> cp(K), sp(K), op(K);
  std::vector<int32_t> cn(K), sn(K), on(K);
  std::vector<float> poses(K * 6);
  for (size_t k = 0; k < K; ++k) {
    cp[k] = dq[k].corner.empty() ? nullptr : &dq[k].corner[0].x; cn[k] = (int32_t)dq[k].corner.size();
    sp[k] = dq[k].surf.empty() ? nullptr : &dq[k].surf[0].x; sn[k] = (int32_t)dq[k].surf.size();
    op[k] = dq[k].outlier.empty() ? nullptr : &dq[k].outlier[0].x; on[k] = (int32_t)dq[k].outlier.size();
    for (int q = 0; q < 6; ++q) poses[k * 6 + q] = dq[k].pose6[q];
  }
  return alego_lm_assemble_map(ctx_.handle(), seq, (int)K, cp.data(), cn.data(), sp.data(), sn.data(), op.data(), on.data(), poses.data());
}

int LaserMapping::pose(int seq, double params[6], double t_map2laser[3], double r_map2laser[9], double t_map2odom[3],
                       double r_map2odom[9]) {
  return alego_lm_get_state(ctx_.handle(), seq, params, t_map2laser, r_map2laser, t_map2odom, r_map2odom);
}

}  // namespace alego

extern "C" {
long alego_host_decode_pointcloud2(const uint8_t *data, size_t data_size, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                                   uint32_t off_x, uint32_t off_y, uint32_t off_z, int32_t off_intensity, int is_bigendian, float *out,
                                   int stride, size_t capacity_points) {
  alego::PointCloud2View m;
  m.data = data; m.data_size = data_size; m.width = width; m.height = height; m.point_step = point_step; m.row_step = row_step;
  m.off_x = off_x; m.off_y = off_y; m.off_z = off_z; m.off_intensity = off_intensity; m.is_bigendian = is_bigendian != 0;
  return alego::decode_pointcloud2(m, out, stride, capacity_points);
}
size_t alego_host_encode_pointcloud2_xyzi(const float *xyzi, size_t n, uint8_t *data) { return alego::encode_pointcloud2_xyzi(xyzi, n, data); }
void *alego_host_imu_create(int length) { return new alego::ImuQueue(length); }
void alego_host_imu_destroy(void *q) { delete static_cast<alego::ImuQueue *>(q); }
void alego_host_imu_push(void *q, const double msg[8]) {
  static_cast<alego::ImuQueue *>(q)->push(alego::ImuMsg{msg[0], msg[1], msg[2], msg[3], msg[4], msg[5], msg[6], msg[7]});
}
void alego_host_imu_get(const void *q, double *arrays, int32_t ptrs[3]) {
  const alego::ImuQueue *Q = static_cast<const alego::ImuQueue *>(q);
  const AlegoImuQueue v = Q->view();
  if (arrays) std::memcpy(arrays, v.time, sizeof(double) * 10 * (size_t)v.length);
  if (ptrs) { ptrs[0] = Q->ptr_front; ptrs[1] = Q->ptr_last; ptrs[2] = Q->ptr_last_iter; }
}
}
