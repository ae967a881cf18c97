// alego_nodelets.cpp — the three pluginlib entry points of A-LeGO-LOAM (nodelet_plugins.xml: loam/ImageProjection,
// loam/LaserOdometry, loam/LaserMapping) as thin ROS shells over the host cores of alego_host.h.
//
// Topics and message types are the reference's (imageProjection.cpp:42-45, laserOdometry.cpp:52-72, laserMapping.cpp:82-93):
//   ImageProjection  sub /lslidar_point_cloud          pub /segmented_cloud /seg_info /outlier
//   LaserOdometry    sub /seg_info /imu/data            pub /odom/lidar /corner_last /surf_last /outlier_last
//   LaserMapping     sub /odom/lidar                    pub /odom_aft_mapped
// What differs from the reference, by design: in nodelet mode (one manager process, launch/test.launch) the three plugins share
// ONE AlegoContext, so the segmented cloud, cloud_info arrays and feature clouds stay in HBM between the stages — /seg_info and
// /odom/lidar act as the "stage finished" tokens, the cloud topics are published only for other consumers (rviz, bags).  All
// three use the single-threaded node handle of the manager, so the callbacks are serialised and run in arrival order; a stage
// that finds a token whose stamp is not the sweep the previous stage processed last skips it, like the reference's "unsync msg"
// branch (laserOdometry.cpp:97-110).  Keyframes / iSAM2 / loop-closure bookkeeping (GTSAM) are outside this shell.
//
// ROS is not part of this repository's build image: this file is compiled where catkin finds roscpp (CMakeLists.txt next to it)
// and is syntax-checked against stand-in headers by tests/test_ros_shells.py.
#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <nav_msgs/Odometry.h>
#include <nodelet/nodelet.h>
#include <pluginlib/class_list_macros.h>
#include <ros/ros.h>
#include <sensor_msgs/Imu.h>
#include <sensor_msgs/PointCloud2.h>

#include <alego/cloud_info.h>

#include "alego_host.h"

namespace loam {
namespace {

// one context per process: the nodelets of a manager share the device buffers
struct Shared {
  std::mutex mtx;
  std::unique_ptr<alego::AlegoContext> ctx;
  double ip_stamp = -1., lo_stamp = -1.;  // sweep each stage processed last
  int frame_cnt = 0;
};
Shared &shared() {
  static Shared s;
  return s;
}

alego::AlegoContext *context(ros::NodeHandle &pnh) {
  Shared &s = shared();
  std::lock_guard<std::mutex> lock(s.mtx);
  if (!s.ctx) {
    int preset = ALEGO_PRESET_REFERENCE, device = 0;  // the constants of utility.h:50-65 unless ~preset says otherwise
    pnh.param("preset", preset, preset);
    pnh.param("device", device, device);
    AlegoParams p;
    if (alego_default_params(&p, preset) != ALEGO_OK) return nullptr;
    s.ctx.reset(new alego::AlegoContext());
    if (s.ctx->init(p, device, 1) != ALEGO_OK) {
      ROS_FATAL("alego_create: %s", s.ctx->last_error().c_str());
      s.ctx.reset();
    }
  }
  return s.ctx.get();
}

int field_offset(const sensor_msgs::PointCloud2 &m, const std::string &name) {
  for (const auto &f : m.fields)
    if (f.name == name && f.datatype == sensor_msgs::PointField::FLOAT32) return (int)f.offset;
  return -1;
}

// pcl::toROSMsg of a PointXYZI cloud: x y z @ 0 4 8, intensity @ 16, 32-byte records (what the reference's subscribers expect)
void to_msg(const alego::PointCloud &cloud, const std_msgs::Header &header, sensor_msgs::PointCloud2 &m) {
  m.header = header;
  m.height = 1;
  m.width = (uint32_t)cloud.size();
  m.is_bigendian = false;
  m.is_dense = true;
  m.point_step = 32;
  m.row_step = m.point_step * m.width;
  m.fields.resize(4);
  const char *names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int k = 0; k < 4; ++k) {
    m.fields[k].name = names[k];
    m.fields[k].offset = offs[k];
    m.fields[k].datatype = sensor_msgs::PointField::FLOAT32;
    m.fields[k].count = 1;
  }
  m.data.resize((size_t)m.row_step);
  if (!cloud.empty()) alego::encode_pointcloud2_xyzi(&cloud[0].x, cloud.size(), m.data.data());
}

// rotation matrix (row-major) -> quaternion x, y, z, w (what Eigen::Quaterniond(r_w_cur_) gives, laserOdometry.cpp:513)
void to_quaternion(const double R[9], double q[4]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s;
  }
}

void fill_odometry(nav_msgs::Odometry &o, const ros::Time &stamp, const char *frame, const double t[3], const double R[9]) {
  double q[4];
  to_quaternion(R, q);
  o.header.stamp = stamp;
  o.header.frame_id = frame;
  o.child_frame_id = "/laser";
  o.pose.pose.position.x = t[0]; o.pose.pose.position.y = t[1]; o.pose.pose.position.z = t[2];
  o.pose.pose.orientation.x = q[0]; o.pose.pose.orientation.y = q[1]; o.pose.pose.orientation.z = q[2]; o.pose.pose.orientation.w = q[3];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
class ImageProjection : public nodelet::Nodelet {
 public:
  void onInit() override {  // imageProjection.cpp:6-47
    nh_ = getNodeHandle();
    pnh_ = getPrivateNodeHandle();
    alego::AlegoContext *ctx = context(pnh_);
    if (!ctx) return;
    core_.reset(new alego::ImageProjection(*ctx));
    if (core_->onInit() != ALEGO_OK) { NODELET_FATAL("ImageProjection::onInit failed"); return; }
    pub_segmented_cloud_ = nh_.advertise<sensor_msgs::PointCloud2>("/segmented_cloud", 10);
    pub_seg_info_ = nh_.advertise<alego::cloud_info>("/seg_info", 10);
    pub_outlier_ = nh_.advertise<sensor_msgs::PointCloud2>("/outlier", 10);
    sub_pc_ = nh_.subscribe<sensor_msgs::PointCloud2>("/lslidar_point_cloud", 10, &ImageProjection::pcCB, this);
  }

 private:
  void pcCB(const sensor_msgs::PointCloud2ConstPtr &msg) {  // imageProjection.cpp:49-208
    alego::PointCloud2View v;
    v.data = msg->data.data();
    v.data_size = msg->data.size();
    v.width = msg->width; v.height = msg->height; v.point_step = msg->point_step; v.row_step = msg->row_step;
    const int ox = field_offset(*msg, "x"), oy = field_offset(*msg, "y"), oz = field_offset(*msg, "z");
    if (ox < 0 || oy < 0 || oz < 0) { NODELET_WARN("point cloud without float32 x / y / z fields"); return; }
    v.off_x = ox; v.off_y = oy; v.off_z = oz;
    v.off_intensity = field_offset(*msg, "intensity");
    v.is_bigendian = msg->is_bigendian;
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mtx);
    const int rc = core_->process(std::vector<alego::PointCloud2View>(1, v));
    if (rc != ALEGO_OK) { NODELET_WARN("alego_ip_process rc=%d", rc); return; }
    s.ip_stamp = msg->header.stamp.toSec();
    alego::CloudInfo info;
    alego::PointCloud seg, outlier;
    const bool want_clouds = pub_segmented_cloud_.getNumSubscribers() > 0 || pub_outlier_.getNumSubscribers() > 0;
    if (core_->results(0, &info, want_clouds ? &seg : nullptr, want_clouds ? &outlier : nullptr) != ALEGO_OK) return;
    alego::cloud_infoPtr m(new alego::cloud_info);  // msg/cloud_info.msg, field for field (:16-20, :183-190)
    m->header = msg->header;
    m->startRingIndex.assign(info.startRingIndex.begin(), info.startRingIndex.end());
    m->endRingIndex.assign(info.endRingIndex.begin(), info.endRingIndex.end());
    m->startOrientation = info.startOrientation;
    m->endOrientation = info.endOrientation;
    m->orientationDiff = info.orientationDiff;
    m->segmentedCloudGroundFlag.assign(info.segmentedCloudGroundFlag.begin(), info.segmentedCloudGroundFlag.end());
    m->segmentedCloudColInd.assign(info.segmentedCloudColInd.begin(), info.segmentedCloudColInd.end());
    m->segmentedCloudRange.assign(info.segmentedCloudRange.begin(), info.segmentedCloudRange.end());
    pub_seg_info_.publish(m);  // :318-336
    if (pub_segmented_cloud_.getNumSubscribers() > 0) {
      sensor_msgs::PointCloud2Ptr c(new sensor_msgs::PointCloud2);
      to_msg(seg, msg->header, *c);
      pub_segmented_cloud_.publish(c);
    }
    if (pub_outlier_.getNumSubscribers() > 0) {
      sensor_msgs::PointCloud2Ptr c(new sensor_msgs::PointCloud2);
      to_msg(outlier, msg->header, *c);
      pub_outlier_.publish(c);
    }
  }

  ros::NodeHandle nh_, pnh_;
  ros::Subscriber sub_pc_;
  ros::Publisher pub_segmented_cloud_, pub_seg_info_, pub_outlier_;
  std::unique_ptr<alego::ImageProjection> core_;
};

// ---------------------------------------------------------------------------------------------------------------------------
class LaserOdometry : public nodelet::Nodelet {
 public:
  void onInit() override {  // laserOdometry.cpp:6-77
    nh_ = getNodeHandle();
    pnh_ = getPrivateNodeHandle();
    alego::AlegoContext *ctx = context(pnh_);
    if (!ctx) return;
    core_.reset(new alego::LaserOdometry(*ctx));
    if (core_->onInit() != ALEGO_OK) { NODELET_FATAL("LaserOdometry::onInit failed"); return; }
    pnh_.param("adjust_distortion", adjust_distortion_, false);  // the reference has the call commented out (:115)
    pub_odom_ = nh_.advertise<nav_msgs::Odometry>("/odom/lidar", 10);
    pub_surf_last_ = nh_.advertise<sensor_msgs::PointCloud2>("/surf_last", 10);
    pub_corner_last_ = nh_.advertise<sensor_msgs::PointCloud2>("/corner_last", 10);
    pub_outlier_last_ = nh_.advertise<sensor_msgs::PointCloud2>("/outlier_last", 10);
    ip_view_.reset(new alego::ImageProjection(*ctx));  // read access to the sweep's segmented / outlier clouds on the shared handle
    sub_segmented_info_ = nh_.subscribe<alego::cloud_info>("/seg_info", 10, &LaserOdometry::segInfoHandler, this);
    sub_imu_ = nh_.subscribe<sensor_msgs::Imu>("/imu/data", 100, &LaserOdometry::imuHandler, this);
  }

 private:
  void imuHandler(const sensor_msgs::ImuConstPtr &msg) {  // :761-804
    alego::ImuMsg m;
    m.stamp = msg->header.stamp.toSec();
    m.qx = msg->orientation.x; m.qy = msg->orientation.y; m.qz = msg->orientation.z; m.qw = msg->orientation.w;
    m.ax = msg->linear_acceleration.x; m.ay = msg->linear_acceleration.y; m.az = msg->linear_acceleration.z;
    std::lock_guard<std::mutex> lock(shared().mtx);
    core_->imuHandler(0, m);
  }

  void segInfoHandler(const alego::cloud_infoConstPtr &msg) {  // one pass of mainLoop (:79-555)
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mtx);
    const double t1 = msg->header.stamp.toSec();
    if (std::abs(t1 - s.ip_stamp) > 1e-9) { NODELET_WARN("unsync msg"); return; }  // the device holds another sweep (:97-110)
    if (adjust_distortion_) {
      int32_t visited = 0;
      if (core_->adjustDistortion(&t1, &visited) != ALEGO_OK) NODELET_WARN("adjustDistortion failed");
    }
    AlegoSolveReport rep;
    const int rc = core_->process(&rep);
    if (rc == ALEGO_FEW_FEATURES) NODELET_WARN("few correspondences (surf %d, corner %d)", rep.n_surf, rep.n_corner);  // :424, :498
    else if (rc != ALEGO_OK) { NODELET_WARN("LaserOdometry rc=%d", rc); return; }
    s.lo_stamp = t1;
    double params[6], t_w[3], r_w[9];
    if (core_->odometry(0, params, t_w, r_w) != ALEGO_OK) return;
    nav_msgs::OdometryPtr odom(new nav_msgs::Odometry);  // :513-525
    fill_odometry(*odom, msg->header.stamp, "/odom", t_w, r_w);
    pub_odom_.publish(odom);
    // corner_last_ / surf_last_ / outlier of the sweep for LaserMapping and any recorder (:531-552)
    if (pub_surf_last_.getNumSubscribers() > 0 || pub_corner_last_.getNumSubscribers() > 0 || pub_outlier_last_.getNumSubscribers() > 0) {
      alego::PointCloud less_flat, seg, outlier, less_sharp_cloud;
      std::vector<int32_t> sharp, less_sharp, flat;
      if (core_->features(0, &sharp, &less_sharp, &flat, &less_flat) == ALEGO_OK &&
          ip_view_->results(0, nullptr, &seg, &outlier) == ALEGO_OK) {
        less_sharp_cloud.reserve(less_sharp.size());
        for (int32_t i : less_sharp)
          if (i >= 0 && (size_t)i < seg.size()) less_sharp_cloud.push_back(seg[(size_t)i]);  // corner_less_sharp_ (:209, :223)
        std_msgs::Header h = msg->header;
        h.frame_id = "/laser";
        const alego::PointCloud *clouds[3] = {&less_sharp_cloud, &less_flat, &outlier};
        ros::Publisher *pubs[3] = {&pub_corner_last_, &pub_surf_last_, &pub_outlier_last_};
        for (int k = 0; k < 3; ++k) {
          sensor_msgs::PointCloud2Ptr c(new sensor_msgs::PointCloud2);
          to_msg(*clouds[k], h, *c);
          pubs[k]->publish(c);
        }
      }
    }
  }

  ros::NodeHandle nh_, pnh_;
  ros::Subscriber sub_segmented_info_, sub_imu_;
  ros::Publisher pub_odom_, pub_surf_last_, pub_corner_last_, pub_outlier_last_;
  std::unique_ptr<alego::LaserOdometry> core_;
  std::unique_ptr<alego::ImageProjection> ip_view_;
  bool adjust_distortion_ = false;
};

// ---------------------------------------------------------------------------------------------------------------------------
class LaserMapping : public nodelet::Nodelet {
 public:
  void onInit() override {  // laserMapping.cpp:5-100
    nh_ = getNodeHandle();
    pnh_ = getPrivateNodeHandle();
    alego::AlegoContext *ctx = context(pnh_);
    if (!ctx) return;
    core_.reset(new alego::LaserMapping(*ctx));
    if (core_->onInit() != ALEGO_OK) { NODELET_FATAL("LaserMapping::onInit failed"); return; }
    // no keyframe yet: empty map clouds, the guard of scan2MapOptimization skips the solve until the first keyframe (:196-199, :350)
    if (core_->setLocalMap(0, alego::PointCloud(), alego::PointCloud()) != ALEGO_OK) { NODELET_FATAL("LaserMapping: no local map"); return; }
    pub_odom_aft_mapped_ = nh_.advertise<nav_msgs::Odometry>("/odom_aft_mapped", 10);
    sub_laser_odom_ = nh_.subscribe<nav_msgs::Odometry>("/odom/lidar", 10, &LaserMapping::laserOdomHandler, this);
  }

 private:
  void laserOdomHandler(const nav_msgs::OdometryConstPtr &msg) {  // laserOdomHandler + one pass of mainLoop (:102-127, :154-185)
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mtx);
    if (std::abs(msg->header.stamp.toSec() - s.lo_stamp) > 0.005) return;  // the four inputs must belong to one sweep (:107-108)
    if (s.frame_cnt++ % 2 == 0) {  // every second frame (:112)
      // the first keyframes build the local map; afterwards it is re-assembled from the deque of recent keyframes (:194-323)
      if (core_->keyFrameCount(0) > 0 && core_->extractSurroundingKeyFrames(0) != ALEGO_OK) NODELET_WARN("local map assembly failed");
      AlegoSolveReport rep;
      const int rc = core_->process(&rep);
      if (rc != ALEGO_OK && rc != ALEGO_FEW_FEATURES) { NODELET_WARN("LaserMapping rc=%d", rc); return; }
      double params[6], t_m2l[3], r_m2l[9], t_m2o[3], r_m2o[9];
      if (core_->pose(0, params, t_m2l, r_m2l, t_m2o, r_m2o) != ALEGO_OK) return;
      // saveKeyFramesAndFactor's cloud side (:491-545): the first frame, then whenever the squared distance to the previous
      // keyframe reaches min_keyframe_dist_ = 1.0 (:43, :501-508).  Without the pose graph the keyframe pose is the estimate itself.
      const double dx = t_m2l[0] - last_kf_[0], dy = t_m2l[1] - last_kf_[1], dz = t_m2l[2] - last_kf_[2];
      if (core_->keyFrameCount(0) == 0 || dx * dx + dy * dy + dz * dz >= 1.0) {
        float pose6[6];
        for (int k = 0; k < 6; ++k) pose6[k] = (float)params[k];  // PointTypePose fields are float (utility.h:83-97)
        if (core_->saveKeyFrame(0, pose6) == ALEGO_OK)
          for (int k = 0; k < 3; ++k) last_kf_[k] = t_m2l[k];
      }
      if (pub_odom_aft_mapped_.getNumSubscribers() > 0) {
        nav_msgs::OdometryPtr odom(new nav_msgs::Odometry);
        fill_odometry(*odom, msg->header.stamp, "map", t_m2l, r_m2l);
        pub_odom_aft_mapped_.publish(odom);
      }
    }
  }

  ros::NodeHandle nh_, pnh_;
  ros::Subscriber sub_laser_odom_;
  ros::Publisher pub_odom_aft_mapped_;
  std::unique_ptr<alego::LaserMapping> core_;
  double last_kf_[3] = {0., 0., 0.};
};

}  // namespace loam

PLUGINLIB_EXPORT_CLASS(loam::ImageProjection, nodelet::Nodelet)
PLUGINLIB_EXPORT_CLASS(loam::LaserOdometry, nodelet::Nodelet)
PLUGINLIB_EXPORT_CLASS(loam::LaserMapping, nodelet::Nodelet)
