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
  // what publish() sends for sequence `seq` (:318-336): /seg_info, /segmented_cloud, /outlier
  int results(int seq, CloudInfo *info, PointCloud *segmented, PointCloud *outlier);

 private:
  AlegoContext &ctx_;
  float *pinned_ = nullptr;
  std::vector<int32_t> n_points_;
};

class LaserOdometry {
 public:
  explicit LaserOdometry(AlegoContext &ctx) : ctx_(ctx) {}
  int onInit() { return ALEGO_OK; }  // laserOdometry.cpp:6-77: state lives in the handle (params_, t_w_cur_, r_w_cur_)
  // one pass of mainLoop after message sync (:118-535): features + scan-to-scan; reports[n_seq] may be null
  int process(AlegoSolveReport *reports = nullptr);
  int odometry(int seq, double params[6], double t_w_cur[3], double r_w_cur[9]);  // /odom/lidar (:513-529)
  // /corner, /corner_less (indices), /surf, /surf_less (:298-314)
  int features(int seq, std::vector<int32_t> *sharp_idx, std::vector<int32_t> *less_sharp_idx, std::vector<int32_t> *flat_idx,
               PointCloud *less_flat);

 private:
  AlegoContext &ctx_;
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
