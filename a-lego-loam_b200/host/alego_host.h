// alego_host.h — ROS-free host cores of the three A-LeGO-LOAM stages above the C ABI (include/alego_b200.h).
//
// The reference implements each stage as a nodelet class whose numerics are inlined in its ROS callbacks
// (loam::ImageProjection::pcCB, imageProjection.cpp:49-208; loam::LaserOdometry::mainLoop,
// laserOdometry.cpp:79-555; loam::LaserMapping::mainLoop / scan2MapOptimization, laserMapping.cpp:102-127,
// 348-479).  These classes keep the reference's names, entry point (onInit) and the cloud_info hand-off
// (msg/cloud_info.msg:1-12) but hold no numerics: every stage call goes to libalego_b200.so.  A catkin shell only
// has to convert sensor_msgs::PointCloud2 <-> float[n][4] and forward to process(); ROS is not part of this image.
//
// Error model: like the reference's callbacks (log-and-return), process() returns the ALEGO_* status; it never
// throws.  One AlegoContext = one robot (n_seq = 1) or a batch of independent robots.
#pragma once
#include <cstdint>
#include <string>
#include <deque>
#include <vector>

#include "../../include/alego_b200.h"

namespace alego {

// cloud_info.msg as an owning host struct (the message's arrays are pre-sized N_SCAN*Horizon_SCAN,
// imageProjection.cpp:16-20; only the first `size` entries are meaningful)
struct CloudInfo {
  std::vector<int32_t> startRingIndex, endRingIndex;
  float startOrientation = 0.f, endOrientation = 0.f, orientationDiff = 0.f;
  std::vector<uint8_t> segmentedCloudGroundFlag;
  std::vector<int32_t> segmentedCloudColInd;
  std::vector<float> segmentedCloudRange;
  int32_t size = 0;
};

struct PointXYZI {  // pcl::PointXYZI payload (utility.h:44), 16 bytes
  float x, y, z, intensity;
};
typedef std::vector<PointXYZI> PointCloud;

// ---- wire format on either side of the path (the part that needs no ROS) ---------------------------
// A sensor_msgs/PointCloud2 payload as the subscriber receives it (/lslidar_point_cloud, imageProjection.cpp:42,49-59 —
// the reference hands it to pcl::fromROSMsg): `width * height` records of `point_step` bytes (rows `row_step` bytes apart)
// with FLOAT32 fields x / y / z / intensity at the given byte offsets (off_intensity < 0: no such field).
struct PointCloud2View {
  const uint8_t *data = nullptr;
  size_t data_size = 0;  // bytes behind `data` (msg->data.size()): a view whose geometry reaches past it is refused
  uint32_t width = 0, height = 1, point_step = 0, row_step = 0;
  uint32_t off_x = 0, off_y = 4, off_z = 8;
  int32_t off_intensity = -1;
  bool is_bigendian = false;
};
// pcl::fromROSMsg for this path: decode into `stride` floats per point (3: x, y, z — what the kernels read; 4: + intensity,
// 0 when the message has none).  NaN / inf stay in place (removeNaNFromPointCloud happens on the device).  Returns the number
// of points written, or -1 when the view is inconsistent (a field does not fit in point_step, a row does not fit in row_step, the
// last row ends past data_size — all checked in 64-bit arithmetic) or `capacity_points` is too small; nothing is written then.
long decode_pointcloud2(const PointCloud2View &msg, float *out, int stride, size_t capacity_points);
// pcl::toROSMsg layout of a PointXYZI cloud (/segmented_cloud, /outlier, imageProjection.cpp:318-336): x, y, z at 0, 4, 8,
// intensity at 16, point_step 32, little endian.  `data` needs 32 * n bytes; returns the bytes written.
size_t encode_pointcloud2_xyzi(const float *xyzi, size_t n, uint8_t *data);

// shared by the three stage objects of one process (nodelet mode: one manager process, zero-copy hand-offs —
// here: the hand-offs stay in HBM)
class AlegoContext {
 public:
  AlegoContext() = default;
  ~AlegoContext();
  AlegoContext(const AlegoContext &) = delete;
  AlegoContext &operator=(const AlegoContext &) = delete;
  int init(const AlegoParams &p, int device = 0, int n_seq = 1, int max_points_per_scan = 0);
  AlegoHandle *handle() const { return h_; }
  const AlegoParams &params() const { return p_; }
  int n_seq() const { return n_seq_; }
  int max_points() const { return max_pts_; }
  std::string last_error() const;

 private:
  AlegoHandle *h_ = nullptr;
  AlegoParams p_{};
  int n_seq_ = 0, max_pts_ = 0;
};

class ImageProjection {
 public:
  explicit ImageProjection(AlegoContext &ctx) : ctx_(ctx) {}
  ~ImageProjection();
  int onInit();  // imageProjection.cpp:6-47 minus the ROS plumbing: pinned staging for the decoded sweeps
  // pcCB for a batch: clouds[b] is the decoded /lslidar_point_cloud of sequence b (NaNs allowed, dropped on the device)
  int process(const std::vector<PointCloud> &clouds);
  // the same straight from the subscriber's messages: decoded into the pinned staging buffer, no intermediate cloud
  int process(const std::vector<PointCloud2View> &msgs);
  // what publish() sends for sequence `seq` (:318-336): /seg_info, /segmented_cloud, /outlier
  int results(int seq, CloudInfo *info, PointCloud *segmented, PointCloud *outlier);

 private:
  AlegoContext &ctx_;
  float *pinned_ = nullptr;
  std::vector<int32_t> n_points_;
};

// sensor_msgs/Imu as imuHandler reads it (laserOdometry.cpp:761-771): stamp, orientation quaternion, linear acceleration
struct ImuMsg {
  double stamp;
  double qx, qy, qz, qw;
  double ax, ay, az;
};

// The IMU ring buffers of LaserOdometry (laserOdometry.h:36-46) and imuHandler's dead-reckoning (laserOdometry.cpp:761-804):
// orientation -> roll / pitch / yaw (tf::Matrix3x3::getRPY), gravity removed from the acceleration in the sensor frame, rotated
// to the IMU's world frame (float, Eigen::Quaternionf), velocity and shift integrated while consecutive stamps are < 1 s apart.
class ImuQueue {
 public:
  explicit ImuQueue(int length = 200);  // imu_queue_length, utility.h:70
  void push(const ImuMsg &m);
  AlegoImuQueue view() const;           // borrowed pointers, valid until the next push
  int ptr_front = 0, ptr_last = -1, ptr_last_iter = 0;  // laserOdometry.cpp:17-19
  int length() const { return len_; }

 private:
  int len_;
  std::vector<double> a_;  // [10][len]: time, roll, pitch, yaw, shift xyz, velocity xyz
};

class LaserOdometry {
 public:
  explicit LaserOdometry(AlegoContext &ctx) : ctx_(ctx) {}
  // laserOdometry.cpp:6-77: solver state lives in the handle (params_, t_w_cur_, r_w_cur_); the IMU queues start empty (:17-29)
  int onInit();
  // /imu/data subscriber of sequence `seq` (:71, :761-804)
  int imuHandler(int seq, const ImuMsg &msg);
  // step 1 of mainLoop (:111-116, body :557-657) — optional, the reference has the call commented out: motion-compensates the
  // segmented clouds ImageProjection left in HBM with the IMU queues; scan_time[n_seq] = stamp of the sweeps (t1, :96)
  int adjustDistortion(const double *scan_time, int32_t *n_adjusted = nullptr, double scan_period = 0.2);
  const ImuQueue *imuQueue(int seq) const { return seq >= 0 && seq < (int)imu_.size() ? &imu_[seq] : nullptr; }
  // one pass of mainLoop after message sync (:118-535): features + scan-to-scan; reports[n_seq] may be null
  int process(AlegoSolveReport *reports = nullptr);
  int odometry(int seq, double params[6], double t_w_cur[3], double r_w_cur[9]);  // /odom/lidar (:513-529)
  // /corner, /corner_less (indices), /surf, /surf_less (:298-314)
  int features(int seq, std::vector<int32_t> *sharp_idx, std::vector<int32_t> *less_sharp_idx, std::vector<int32_t> *flat_idx,
               PointCloud *less_flat);

 private:
  AlegoContext &ctx_;
  std::vector<ImuQueue> imu_;  // per sequence
};

class LaserMapping {
 public:
  explicit LaserMapping(AlegoContext &ctx) : ctx_(ctx) {}
  int onInit() { return ALEGO_OK; }  // laserMapping.cpp:5-100: leaf sizes / iteration counts are AlegoParams fields
  // corner_from_map_ds_ / surf_from_map_ds_ of sequence `seq` (output of extractSurroundingKeyFrames, :194-323)
  int setLocalMap(int seq, const PointCloud &corner, const PointCloud &surf);
  // downsampleCurrentScan + scan2MapOptimization + transformUpdate (:325-489) on the clouds LaserOdometry left in HBM
  int process(AlegoSolveReport *reports = nullptr);
  int pose(int seq, double params[6], double t_map2laser[3], double r_map2laser[9], double t_map2odom[3], double r_map2odom[9]);

  // ---- keyframe store + local-map assembly (the loop_closure_enabled_ branch of extractSurroundingKeyFrames, :206-243)
  // saveKeyFramesAndFactor's cloud side (:491-545): keep the current sweep's downsampled clouds (laser_corner_ds_,
  // laser_surf_ds_, laser_outlier_ds_) of sequence `seq` with the pose estimate of that keyframe; the deque holds the
  // recent_keyframe_search_num_ (50, :48) most recent keyframes.  pose6 = x, y, z, roll, pitch, yaw (cloud_keyposes_6d_).
  int saveKeyFrame(int seq, const float pose6[6]);
  // correctPoses (:547-584): overwrite the pose of the k-th stored keyframe (0 = oldest) after a pose-graph update
  int setKeyFramePose(int seq, size_t k, const float pose6[6]);
  // transform + concatenate + VoxelGrid on the device -> local map of `seq` (alego_lm_assemble_map)
  int extractSurroundingKeyFrames(int seq);
  size_t keyFrameCount(int seq) const { return seq < (int)store_.size() ? store_[seq].size() : 0; }
  static const size_t kRecentKeyframes = 50;

 private:
  struct KeyFrame {
    PointCloud corner, surf, outlier;
    float pose6[6];
  };
  AlegoContext &ctx_;
  std::vector<std::deque<KeyFrame>> store_;  // per sequence
};

}  // namespace alego

// C entry points of the two wire-format helpers (ctypes tests, other-language shells)
extern "C" {
long alego_host_decode_pointcloud2(const uint8_t *data, size_t data_size, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                                   uint32_t off_x, uint32_t off_y, uint32_t off_z, int32_t off_intensity, int is_bigendian, float *out,
                                   int stride, size_t capacity_points);
size_t alego_host_encode_pointcloud2_xyzi(const float *xyzi, size_t n, uint8_t *data);
// alego::ImuQueue for ctypes: msg = stamp, qx, qy, qz, qw, ax, ay, az; get copies the [10][length] arrays and the three pointers
void *alego_host_imu_create(int length);
void alego_host_imu_destroy(void *q);
void alego_host_imu_push(void *q, const double msg[8]);
void alego_host_imu_get(const void *q, double *arrays, int32_t ptrs[3] /* front, last, last_iter */);
}
