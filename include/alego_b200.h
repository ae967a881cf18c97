/*
 * alego_b200.h — C ABI of the B200-native A-LeGO-LOAM per-scan hot path.
 *
 * The reference (jyakaranda/A-LeGO-LOAM) has NO library / FFI surface for its numerics: the
 * hot path is inlined in ROS callbacks (src/imageProjection.cpp:49-316, src/laserOdometry.cpp:79-555,
 * src/laserMapping.cpp:325-489).  This header therefore DEFINES the boundary a maintainer would bind
 * from those callbacks; every entry point cites the reference code it replaces.  Plain pointers and
 * sizes only, no torch / CUDA types in any signature, never throws, every call returns an int status.
 *
 * Batch model: one handle owns `n_seq` INDEPENDENT scan sequences that advance in lock-step (one
 * sweep per sequence per call).  n_seq = 1 is the reference's single-robot case; n_seq > 1 is how a
 * B200 leaves the launch-latency regime (a 64x1800 sweep is 1.8 MB).  Sequences never exchange data.
 *
 * Threading: a handle is not thread-safe; distinct handles are fully independent (own stream).
 */
#ifndef ALEGO_B200_H_
#define ALEGO_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference behaviour: log-and-return, laserOdometry.cpp:91-109,424,498;
 *      laserMapping.cpp:350-354) ---- */
#define ALEGO_OK            0
#define ALEGO_FEW_FEATURES  1  /* a solve was skipped (<10 correspondences / LM guard); state still valid */
#define ALEGO_BAD_ARG      -1
#define ALEGO_CUDA_ERROR   -2
#define ALEGO_NOT_READY    -3  /* stage called before its producer stage */

/* ---- presets for alego_default_params ---- */
#define ALEGO_PRESET_VLP16_1800   0  /* 16 x 1800  (BASELINE cfg1/cfg2)                       */
#define ALEGO_PRESET_HDL64_1800   1  /* 64 x 1800  (headline metric)                          */
#define ALEGO_PRESET_HDL64_2048   2  /* 64 x 2048  (BASELINE cfg4/cfg5)                       */
#define ALEGO_PRESET_REFERENCE    3  /* 16 x 4000, the literal constants of utility.h:50-65   */

/* Runtime form of the compile-time constants in include/alego/utility.h:50-73 and of the literals in
 * laserOdometry.cpp:290,415,489 / laserMapping.cpp:37-41,360,470. */
typedef struct AlegoParams {
  int32_t n_scan;               /* N_SCAN            utility.h:50 */
  int32_t horizon_scan;         /* Horizon_SCAN      utility.h:55 = int(360/ang_res_x + 0.5) */
  int32_t ground_scan_id;       /* utility.h:57 */
  int32_t seg_valid_point_num;  /* utility.h:64 (5)  */
  int32_t seg_valid_line_num;   /* utility.h:65 (3)  */
  int32_t seg_min_cluster;      /* imageProjection.cpp:283 (30) */
  int32_t lo_surf_iters;        /* laserOdometry.cpp:415 (5) */
  int32_t lo_corner_iters;      /* laserOdometry.cpp:489 (5; README says 10) */
  int32_t lm_outer_iters;       /* laserMapping.cpp:360 (2) */
  int32_t lm_max_iters;         /* laserMapping.cpp:470 (20) */
  int32_t reserved_i[6];
  double ang_res_x;             /* utility.h:51 (deg) */
  double ang_res_y;             /* utility.h:52 (deg) */
  double ang_bottom;            /* utility.h:56 (deg) */
  double sensor_mount_ang;      /* utility.h:58 (deg) */
  double seg_theta;             /* utility.h:63 (rad) */
  double nearest_feature_dist;  /* utility.h:73 (m^2) */
  double huber_delta;           /* HuberLoss(0.1) laserOdometry.cpp:331, laserMapping.cpp:363 */
  double less_flat_leaf;        /* laserOdometry.cpp:290 (0.4) */
  double lm_corner_leaf;        /* laserMapping.cpp:37 (0.4) */
  double lm_surf_leaf;          /* laserMapping.cpp:38 (0.8) */
  double lm_outlier_leaf;       /* laserMapping.cpp:39 (1.0) */
  double reserved_d[5];
} AlegoParams;

/* Field-for-field mirror of msg/cloud_info.msg:1-12 (Header omitted).  Arrays are caller-owned:
 * ring arrays hold n_scan entries, the per-point arrays n_scan*horizon_scan entries of which the first
 * `size` are meaningful (imageProjection.cpp:16-20,183-190). */
typedef struct AlegoCloudInfo {
  int32_t *startRingIndex;            /* [n_scan] */
  int32_t *endRingIndex;              /* [n_scan] */
  float startOrientation;
  float endOrientation;
  float orientationDiff;
  int32_t size;                       /* M = number of segmented points */
  uint8_t *segmentedCloudGroundFlag;  /* [M] bool */
  int32_t *segmentedCloudColInd;      /* [M] */
  float *segmentedCloudRange;         /* [M] */
} AlegoCloudInfo;

/* Per-sequence result of one Levenberg-Marquardt solve chain (what summary.BriefReport() and the
 * correspondence counters log at laserOdometry.cpp:408,420,482,494 and laserMapping.cpp:465,477). */
typedef struct AlegoSolveReport {
  int32_t status;          /* ALEGO_OK / ALEGO_FEW_FEATURES */
  int32_t n_corner;        /* corner / edge correspondences   */
  int32_t n_surf;          /* surf / plane correspondences    */
  int32_t iterations;      /* LM iterations over all solves (successful + unsuccessful steps) */
  double initial_cost;     /* of the first executed solve */
  double final_cost;       /* of the last executed solve  */
} AlegoSolveReport;

typedef struct AlegoHandle AlegoHandle;

/* ---- lifecycle ------------------------------------------------------------------------------- */
int alego_default_params(AlegoParams *p, int preset);
/* Replaces the three onInit() bodies (imageProjection.cpp:6-47, laserOdometry.cpp:6-77,
 * laserMapping.cpp:5-100): allocates every device buffer for n_seq sequences on `device`. */
int alego_create(const AlegoParams *p, int device, int n_seq, int max_points_per_scan, AlegoHandle **out);
void alego_destroy(AlegoHandle *h);
const char *alego_last_error(const AlegoHandle *h);   /* never NULL */
int alego_synchronize(AlegoHandle *h);
int alego_get_params(const AlegoHandle *h, AlegoParams *out);
int alego_n_seq(const AlegoHandle *h);

/* Layout of the sweep buffers handed to alego_ip_* / alego_stage_upload / alego_pipeline_*: 4 floats per point
 * (x, y, z, intensity — the default, a decoded pcl::PointXYZI) or 3 (packed x, y, z).  The reference never reads the
 * sensor intensity on this path (pcCB overwrites it with row + col/10000, imageProjection.cpp:101), so the packed form
 * gives identical results with 25 % fewer bytes over PCIe.  Applies to every later upload. */
int alego_set_point_stride(AlegoHandle *h, int floats_per_point);

/* Pinned host memory for the sweep buffers (cudaMallocHost) so that the H2D copies are asynchronous. */
void *alego_host_alloc(size_t bytes);
void alego_host_free(void *p);

/* ---- ImageProjection: replaces ImageProjection::pcCB + labelComponents
 *      (src/imageProjection.cpp:49-208, 210-316) -------------------------------------------------- */
/* xyzi_host: [n_seq][max_points_per_scan][stride] float32 (stride 4: x,y,z,intensity; 3: x,y,z) — the decoded
 * sensor_msgs::PointCloud2 of /lslidar_point_cloud; n_points[n_seq].  Copies H2D (async on the
 * handle's stream; pass pinned memory for true overlap) and runs the IP kernels. */
int alego_ip_process(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points);
/* Same, but only the H2D copy / only the kernels (inputs already resident in HBM). */
int alego_ip_upload(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points);
int alego_ip_run(AlegoHandle *h);
/* Pre-stage sweeps in HBM (slot k = one sweep per sequence) and select a slot as the input of the next
 * alego_ip_run / alego_pipeline_step(h, NULL, NULL, ...) — no copy at selection time.  Used to measure the
 * path with inputs already resident in device memory. */
int alego_stage_upload(AlegoHandle *h, int slot, const float *xyzi_host, const int32_t *n_points);
int alego_stage_select(AlegoHandle *h, int slot);
/* Fetch what pcCB publishes for sequence `seq`: /seg_info, /segmented_cloud, /outlier
 * (imageProjection.cpp:318-336).  Any output pointer may be NULL.  label_image is the R x C
 * label_mat_ (-1 ground/empty, 1..K clusters, 999999 rejected), row-major. */
int alego_ip_get(AlegoHandle *h, int seq, AlegoCloudInfo *info, float *segmented_xyzi, float *outlier_xyzi,
                 int32_t *n_outlier, int32_t *label_image);

/* ---- LaserOdometry, motion-distortion correction (optional): LaserOdometry::adjustDistortion
 *      (src/laserOdometry.cpp:557-726, IMU branch :581-657).  The reference has the call commented out (:115), so the
 *      pipeline entry points never run it; a caller that wants it calls it between alego_ip_* and alego_lo_extract. ---- */
/* The IMU ring buffers imuHandler fills (laserOdometry.h:36-46, laserOdometry.cpp:761-804), caller-owned. */
typedef struct AlegoImuQueue {
  int32_t length;         /* imu_queue_length (utility.h:70, 200); at most 2048 */
  int32_t ptr_last;       /* imu_ptr_last_: newest entry; <= 0 means "not enough IMU data", nothing is adjusted (:583) */
  int32_t ptr_last_iter;  /* imu_ptr_last_iter_: in = where the previous sweep stopped; out = where this one stopped (:656) */
  int32_t reserved;
  const double *time, *roll, *pitch, *yaw;        /* [length] */
  const double *shift_x, *shift_y, *shift_z;      /* [length] */
  const double *velo_x, *velo_y, *velo_z;         /* [length] */
} AlegoImuQueue;
/* In place on the segmented cloud of every sequence (what alego_ip_get returns and alego_lo_extract reads).
 * scan_time[n_seq]: stamp of the sweep (t1, :96); queues[n_seq]; scan_period: utility.h:53 (0.2).  n_adjusted[n_seq]
 * (may be NULL): points visited before the "unsync imu and pc msg" return (:596-600) — the cloud size when that never
 * happened, 0 when the queue holds fewer than two messages.  The IMU stamps must not run backwards over the entries
 * ptr_last_iter .. ptr_last (ALEGO_BAD_ARG otherwise).  The use_odom branch (:660-714) is compiled out in the reference
 * (use_imu = true, use_odom = false, utility.h:68-69) and is not provided. */
int alego_lo_adjust_distortion(AlegoHandle *h, const double *scan_time, AlegoImuQueue *queues, double scan_period,
                               int32_t *n_adjusted);

/* ---- LaserOdometry, features: replaces steps 2-4 of LaserOdometry::mainLoop
 *      (src/laserOdometry.cpp:118-297) ------------------------------------------------------------ */
int alego_lo_extract(AlegoHandle *h);
/* Index lists are positions in the segmented cloud, in the reference's push_back order.
 * less_flat_xyzi is the per-ring VoxelGrid(0.4) output concatenated over rings (:288-293).
 * cloud_label is cloud_label_[0..M) (2 sharp, 1 less sharp, -1 flat, 0 rest). */
int alego_lo_get_features(AlegoHandle *h, int seq, int32_t *sharp_idx, int32_t *n_sharp, int32_t *less_sharp_idx,
                          int32_t *n_less_sharp, int32_t *flat_idx, int32_t *n_flat, float *less_flat_xyzi,
                          int32_t *n_less_flat, int32_t *cloud_label);

/* ---- LaserOdometry, scan-to-scan: replaces src/laserOdometry.cpp:316-535 + transformToStart
 *      (:728-740) + CornerCostFunction / SurfCostFunction (utility.h:122-240) ---------------------- */
/* First call per sequence only initialises the targets (:316-324).  reports may be NULL. */
int alego_lo_scan2scan(AlegoHandle *h, AlegoSolveReport *reports /*[n_seq]*/);
/* params_[6] (current->last), t_w_cur_[3], r_w_cur_[9] row-major (laserOdometry.h:76-80). */
int alego_lo_get_state(AlegoHandle *h, int seq, double params[6], double t_w[3], double r_w[9]);
int alego_lo_set_params(AlegoHandle *h, int seq, const double params[6]);

/* ---- LaserMapping, scan-to-map: replaces src/laserMapping.cpp:325-489 + pointAssociateToMap
 *      (laserMapping.h:187-194) + LidarEdge/LidarPlaneCostFunction (utility.h:242-349) ------------- */
/* Local map of sequence `seq` (corner_from_map_ds_, surf_from_map_ds_): uploads and builds the
 * device search grid — replaces the two KdTreeFLANN::setInputCloud calls at laserMapping.cpp:356-357.
 * alego_lm_scan2map rebuilds the grid every call when params rebuild flag is set (reference rebuilds
 * its kd-trees every mapped frame). */
int alego_lm_set_map(AlegoHandle *h, int seq, const float *corner_xyzi, int32_t n_corner, const float *surf_xyzi,
                     int32_t n_surf);
/* Local-map assembly: the cloud side of LaserMapping::extractSurroundingKeyFrames (laserMapping.cpp:194-323) and
 * transformPointCloud (laserMapping.h:163-177).  The caller selects the keyframes (the deque of the 50 most recent ones,
 * :206-243, or the radius search, :246-311) and passes their stored clouds (corner_frames_ / surf_frames_ /
 * outlier_frames_, host memory, xyzi) with their poses (cloud_keyposes_6d_: x, y, z, roll, pitch, yaw, float); the device
 * transforms every cloud by its pose, concatenates them in keyframe order — corner_from_map_ += corner;
 * surf_from_map_ += surf, then outlier (:239-243) — and applies ds_corner_ (lm_corner_leaf) / ds_surf_ (lm_surf_leaf)
 * (:316-319).  The result becomes the local map of sequence `seq` exactly as after alego_lm_set_map, without leaving
 * the device.  alego_lm_get_map reads corner_from_map_ds_ / surf_from_map_ds_ back (buffers sized by the caller: at most
 * the total number of input points; any pointer may be NULL). */
int alego_lm_assemble_map(AlegoHandle *h, int seq, int n_keyframes, const float *const *corner_xyzi, const int32_t *n_corner,
                          const float *const *surf_xyzi, const int32_t *n_surf, const float *const *outlier_xyzi,
                          const int32_t *n_outlier, const float *poses6 /*[n_keyframes][6]*/);
int alego_lm_get_map(AlegoHandle *h, int seq, float *corner_xyzi, int32_t *n_corner, float *surf_xyzi, int32_t *n_surf);
/* ---- LaserMapping, loop-closure ICP: the pcl::IterativeClosestPoint<PointT, PointT> object of
 *      LaserMapping::performLoopClosure (src/laserMapping.cpp:667-688) — point-to-point, SVD (Umeyama) estimation,
 *      DefaultConvergenceCriteria.  Loop detection, the keyframe store and the iSAM2 update stay with the caller. -------- */
typedef struct AlegoIcpResult {
  float final_transformation[16]; /* icp.getFinalTransformation(), row-major 4x4 (correction_frame, :688) */
  double fitness_score;           /* icp.getFitnessScore() (:687): mean squared 1-NN distance of the aligned source */
  int32_t has_converged;          /* icp.hasConverged() (:686) */
  int32_t iterations;             /* nr_iterations_ */
  int32_t convergence_state;      /* 0 not converged, 1 iterations, 2 transform, 3 abs MSE, 4 rel MSE, 5 no correspondences */
  int32_t n_correspondences;      /* of the last iteration */
} AlegoIcpResult;
/* source = latest_keyframe_, target = near_history_keyframes_ (host memory, xyzi, finite coordinates), both already in the
 * map frame (detectLoopClosure, :764-822).  The reference's settings: max_correspondence_distance 100, max_iterations 100,
 * transformation_epsilon 1e-6, euclidean_fitness_epsilon 1e-6 (:668-671); align() runs from the identity guess (:684).
 * trace (may be NULL): [max_iterations][14] doubles per executed iteration — correspondences, their mean squared distance,
 * the incremental rotation (9, row-major) and translation (3).  Returns ALEGO_FEW_FEATURES for an empty cloud. */
int alego_lc_icp(AlegoHandle *h, const float *source_xyzi, int32_t n_source, const float *target_xyzi, int32_t n_target,
                 double max_correspondence_distance, int32_t max_iterations, double transformation_epsilon,
                 double euclidean_fitness_epsilon, AlegoIcpResult *out, double *trace);

/* Stand-alone inputs for sequence `seq` (the /corner_last, /surf_last, /outlier clouds of
 * laserMapping.cpp:133-153) and the odometry prediction odom2laser (:154-164). When not called, the
 * clouds produced on the device by the LO stage of the same handle are used. */
int alego_lm_set_scan(AlegoHandle *h, int seq, const float *corner_xyzi, int32_t n_corner, const float *surf_xyzi,
                      int32_t n_surf, const float *outlier_xyzi, int32_t n_outlier);
int alego_lm_set_odom(AlegoHandle *h, int seq, const double t_odom2laser[3], const double r_odom2laser[9]);
/* downsampleCurrentScan + scan2MapOptimization + transformUpdate (:325-346, 348-479, 481-489). */
int alego_lm_scan2map(AlegoHandle *h, AlegoSolveReport *reports /*[n_seq]*/);
/* params_[6], map2laser (t[3], R[9] row-major), map2odom (t[3], R[9]) (laserMapping.h:171-177). */
int alego_lm_get_state(AlegoHandle *h, int seq, double params[6], double t_map2laser[3], double r_map2laser[9],
                       double t_map2odom[3], double r_map2odom[9]);
int alego_lm_set_params(AlegoHandle *h, int seq, const double params[6]);
/* The four VoxelGrid outputs of downsampleCurrentScan (laser_corner_ds_, laser_surf_ds_,
 * laser_outlier_ds_, laser_surf_total_ds_). Any pointer may be NULL. */
int alego_lm_get_downsampled(AlegoHandle *h, int seq, float *corner_ds, int32_t *n_corner_ds, float *surf_ds,
                             int32_t *n_surf_ds, float *outlier_ds, int32_t *n_outlier_ds, float *surf_total_ds,
                             int32_t *n_surf_total_ds);

/* ---- whole path: one sweep per sequence through IP -> LO -> LM ------------------------------------ */
/* poses_out: [n_seq][12] doubles: t_map2laser[3], then params_ of LM [6] (x,y,z,roll,pitch,yaw),
 * then t_w_cur_ of LO [3].  Host buffers in, host buffers out (H2D + D2H inside the call).
 * xyzi_host == NULL means "inputs already uploaded with alego_ip_upload". poses_out may be NULL. */
int alego_pipeline_step(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points, double *poses_out);
/* Asynchronous form of alego_pipeline_step for streaming ingestion (a rosbag / sensor thread feeding the nodelets,
 * imageProjection.cpp:45 subscriber queue): submit enqueues the H2D copy of the sweeps on a copy stream into one of
 * three device staging buffers and the whole IP -> LO -> LM pass behind it, and returns without waiting, so the copy
 * of sweep t+1 overlaps the pass over sweep t.  xyzi_host must be pinned (alego_host_alloc) and stay untouched until
 * the step is collected.  At most three steps may be in flight; collect waits for the OLDEST one and returns its
 * poses (layout as alego_pipeline_step; poses_out may be NULL). */
int alego_pipeline_submit(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points);
int alego_pipeline_collect(AlegoHandle *h, double *poses_out);
/* Diagnostics of the asynchronous form (no reference counterpart; the reference logs per-stage TicToc times instead,
 * laserMapping.cpp:344-477): with on != 0 every submitted step records CUDA timing events, and t_ms (may be NULL) receives the
 * timeline of the step collected LAST, in ms since the first timed submit: [0] H2D copy starts, [1] H2D copy done,
 * [2] front end (ImageProjection + LaserOdometry) starts on the main stream, [3] front end done, [4] poses in host memory. */
int alego_pipeline_timeline(AlegoHandle *h, int on, float *t_ms);
/* lm_every: run LM on every k-th sweep (reference: 2, laserMapping.cpp:112); 0 disables LM.
 * rebuild_map_index_every_step: rebuild the local-map search index on every mapped sweep, like the reference's kd-tree
 * builds (laserMapping.cpp:356-357), instead of only after alego_lm_set_map.  options: 0 default — the LaserMapping stage
 * of sweep t (index build + scan-to-map) runs on a second stream and overlaps ImageProjection + LaserOdometry of sweep
 * t+1, like the reference's separate nodes; a negative value keeps everything on one stream (debugging); bit 1 (value 2):
 * graph mode for the launch-latency regime (one or a few sequences) — alego_pipeline_step captures the pass once per
 * buffer parity / LM schedule / input buffer into a CUDA graph and replays it with a single launch. */
int alego_pipeline_config(AlegoHandle *h, int lm_every, int rebuild_map_index_every_step, int options);

/* ---- stand-alone operators (used by LO/LM internally, exposed for tests and callers) ------------- */
/* pcl::VoxelGrid<PointXYZI> equivalent on host data (setLeafSize(leaf,leaf,leaf); filter()).
 * out_xyzi needs room for n points. */
int alego_voxel_grid(AlegoHandle *h, const float *xyzi, int32_t n, float leaf, float *out_xyzi, int32_t *n_out);

/* ---- measurement hooks ----------------------------------------------------------------------------
 * CUDA-event timing on the handle's own stream (torch.cuda.Event only sees torch's current stream). */
int alego_timer_mark(AlegoHandle *h, int slot);                         /* cudaEventRecord(slot)  */
int alego_timer_elapsed_ms(AlegoHandle *h, int slot_a, int slot_b, float *ms);
/* Per-kernel accumulation: when enabled every kernel launch is bracketed by events. */
int alego_profile_enable(AlegoHandle *h, int on);
int alego_profile_reset(AlegoHandle *h);
int alego_profile_count(const AlegoHandle *h);
/* name_out: at least 64 bytes. total_ms summed over `launches` launches since the last reset. */
int alego_profile_get(AlegoHandle *h, int i, char *name_out, int64_t *launches, double *total_ms);
int64_t alego_launch_count(const AlegoHandle *h);                       /* kernels launched since create */

/* ---- test hook: copy a named device array of sequence `seq` to host ---------------------------------
 * Returns bytes written (>=0) or a negative status.  Names are listed in DESIGN.md ("debug arrays"). */
int64_t alego_debug_get(AlegoHandle *h, const char *name, int seq, void *dst, size_t capacity_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ALEGO_B200_H_ */
