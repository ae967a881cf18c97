// common.cuh — handle layout, launch/profiling helpers and small device utilities shared by every
// translation unit of libalego_b200.so.  sm_100a only; the whole library is compiled with --fmad=false so
// that float/double expressions round exactly like the reference's baseline-x86-64 build (no FMA
// contraction, CMakeLists.txt:4-5) — the path is HBM/latency bound, the lost FMA throughput is irrelevant.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/alego_b200.h"

#define ALEGO_MAX_RINGS 128
#define ALEGO_MAX_DEVICES 64
#define ALEGO_INFLIGHT 3  // steps alego_pipeline_submit keeps in flight (device staging buffers, pinned result slots)
#define ALEGO_EMPTY_RANGE 3.402823466e+38f  // FLT_MAX: "no return" marker of the range image (reference: DBL_MAX, imageProjection.cpp:33)
#define ALEGO_LABEL_INVALID 999999          // imageProjection.cpp:311

struct Pose {  // translation + row-major rotation
  double t[3];
  double R[9];
};

// Hashed uniform grid over one point cloud per sequence (replaces pcl::KdTreeFLANN, see grid.cuh)
struct GridIndex {
  int table_size = 0;  // power of two
  int cap = 0;         // points per sequence
  float cell = 1.0f;
  int *cell_start = nullptr;  // [B][table_size+4]  entry h = first slot of bucket h, entry table_size = #points (+3 pad)
  float4 *sorted = nullptr;   // [B][cap]  xyz + original index (int bits in .w)
};

// Voxel-row index over a local-map cloud that is pcl::VoxelGrid output (see map_rows.cuh)
struct MapFrame {
  float inv;      // 1 / leaf
  int min_b[3];   // floor(min * inv) per axis
  int dim[3];     // voxels per axis
  int W;          // 32-voxel words per row
  int valid;      // 1: the table describes the cloud (voxel order, one point per voxel, fits the table)
  int n;
};

struct MapRows {
  uint2 *tab = nullptr;      // [B][cap] x = occupancy mask of the word, y = index of its first point
  MapFrame *frame = nullptr; // [B]
  int *bbox = nullptr;       // [B][6] min xyz / max xyz as order-preserving ints
  int cap = 0;               // entries per sequence
  float leaf = 0.f;
  bool usable = false;       // host: every sequence's cloud is in voxel order and fits (decided by map_rows_validate)
};

struct KernelProfile {
  std::string name;
  int64_t launches = 0;
  double total_ms = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// One captured IP -> LO -> LM pass (alego_pipeline_step in graph mode): valid for exactly the buffers it was captured with
struct PipelineGraph {
  int par = 0, stride = 4;
  bool run_lm = false, rebuild = false;
  const void *raw = nullptr, *n_pts = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t n_launches = 0;
};

struct AlegoHandle {
  AlegoParams P;
  int dev = 0, B = 0, Nmax = 0, R = 0, C = 0, RC = 0;
  int in_stride = 4;  // floats per input point: 4 = x,y,z,intensity; 3 = packed x,y,z (alego_set_point_stride)
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;    // LaserMapping stage of the pipeline (map-index build + scan-to-map of sweep t) overlapped
                                         // with ImageProjection + LaserOdometry of sweep t+1, like the reference's three nodes
  cudaStream_t copy_stream = nullptr;    // H2D of the next sweep overlapped with the current pass (alego_pipeline_submit)
  cudaStream_t launch_stream = nullptr;  // when set, LAUNCH() / grid_build() target this stream instead of `stream`
  cudaEvent_t ev_lo_done = nullptr;              // main stream: LaserOdometry of the sweep finished (its clouds are final)
  cudaEvent_t ev_lm_done[2] = {nullptr, nullptr};  // side stream: LaserMapping has consumed the clouds of buffer parity k
  bool lm_done_valid[2] = {false, false};
  cudaEvent_t ev_side_tail = nullptr;            // side stream: everything enqueued there so far
  bool side_busy = false;                        // work was enqueued on the side stream since the last join
  bool overlap_lm = true;
  // CUDA graphs for the launch-latency regime (few sequences): alego_pipeline_config options bit 1
  bool use_graphs = false;
  bool capturing = false;          // pipeline_enqueue is being recorded into a graph
  bool pose_on_main = false;       // the last pass left d_pose on the main stream (graph mode / no overlap)
  unsigned graph_epoch = 0, graphs_epoch = 0;  // bumped whenever a captured pointer / setting may have changed
  std::vector<PipelineGraph> graphs;
  cudaEvent_t ev_g_fork = nullptr, ev_g_lo = nullptr, ev_g_tail = nullptr;  // fork / join inside a capture
  cudaEvent_t ev_copied[ALEGO_INFLIGHT] = {}, ev_consumed[ALEGO_INFLIGHT] = {}, ev_pose[ALEGO_INFLIGHT] = {};
  bool consumed_valid[ALEGO_INFLIGHT] = {};
  // grow-only device scratch of the per-keyframe / stand-alone entry points (local-map assembly, alego_voxel_grid)
  void *scratch[5] = {};
  size_t scratch_bytes[5] = {};
  // timeline of the asynchronous steps (alego_pipeline_timeline): timing events per slot, ms since the first submit
  cudaEvent_t ev_t_origin = nullptr, ev_t[ALEGO_INFLIGHT][5] = {};
  bool timeline_on = false;
  float last_timeline[5] = {0, 0, 0, 0, 0};
  long long n_submitted = 0, n_collected = 0;
  std::string err;
  int64_t launches = 0;
  bool profiling = false;
  std::vector<KernelProfile> prof;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t timer[16] = {};
  int scan_count = 0;
  int lm_every = 1;
  bool rebuild_map_every_step = true;
  bool stage_ip_done = false, stage_feat_done = false;
  bool want_labels = false;  // materialise label_mat_ (imageProjection.h:25) on every sweep instead of on demand
  bool label_valid = false;  // h->label matches the last ImageProjection pass
  int feat_buf = -1;        // buffer index holding the most recent feature clouds
  std::vector<uint8_t> lm_scan_is_external;  // per seq: inputs set through alego_lm_set_scan

  // precomputed on the host with the host libm so that they match the oracle bit for bit
  double seg_sin_x, seg_cos_x, seg_sin_y, seg_cos_y;

  // ---------------- ImageProjection ----------------
  float4 *raw = nullptr;       // [B][Nmax]  (points at raw_own or at a staged sweep)
  int *n_pts = nullptr;        // [B]
  float4 *raw_own = nullptr;
  int *n_pts_own = nullptr;
  float4 *raw_slot[ALEGO_INFLIGHT] = {};  // device staging of alego_pipeline_submit (slot 0 aliases raw_own)
  int *n_pts_slot[ALEGO_INFLIGHT] = {};
  int32_t *h_n_pts_slot[ALEGO_INFLIGHT] = {};  // pinned copies of the caller's n_points
  std::vector<float4 *> stage_raw;  // sweeps pre-staged in HBM (alego_stage_*)
  std::vector<int *> stage_n;
  int *winner = nullptr;       // [B][RC]  index of the last input point that fell in the cell (-1 none)
  float4 *cloud = nullptr;     // [B][RC]  full_cloud_
  float *range = nullptr;      // [B][RC]  range_mat_ (f32 is lossless: the reference stores a float sqrt)
  uint8_t *ground = nullptr;   // [B][RC]  ground_mat_
  uint8_t *cell_flags = nullptr;  // [B][RC]  validity + join flags of every cell (ip_image -> ccl_*)
  uint8_t *cell_class = nullptr;  // [B][RC] keep / outlier / feasible-root verdict of every cell (ip_rowcount -> ip_compact)
  int *parent = nullptr;       // [B][RC]  union-find forest; -1 = not a segmentation candidate
  int2 *comp_stat = nullptr;   // [B][RC]  at roots: (size, max row)
  int *comp_id = nullptr;      // [B][RC]  at roots: final label 1..K
  int *label = nullptr;        // [B][RC]  label_mat_
  int4 *rowcnt = nullptr;      // [B][R]   (kept, outliers, feasible roots, -)
  float4 *seg_cloud = nullptr; // [B][RC]
  uint8_t *seg_ground = nullptr;
  int *seg_col = nullptr;
  float *seg_range = nullptr;
  int *start_ring = nullptr;   // [B][R]
  int *end_ring = nullptr;     // [B][R]
  int *M = nullptr;            // [B]
  float4 *outlier = nullptr;   // [B][out_cap]  (points at outlier_buf[cur] of the sweep ImageProjection last processed)
  float4 *outlier_buf[2] = {nullptr, nullptr};
  int *n_outlier_buf[2] = {nullptr, nullptr};
  int out_cap = 0;
  int *n_outlier = nullptr;    // [B]
  float *orient = nullptr;     // [B][4] start, end, diff

  // ---------------- LaserOdometry features ----------------
  float *curv = nullptr;       // [B][RC]  |diff_range| (float); cloud_curvature_ == (double)c*(double)c exactly
  uint8_t *picked0 = nullptr;  // [B][RC]  cloud_neighbor_picked_ after markOccludedPoints
  uint8_t *picked = nullptr;   // [B][RC]  ... after extractFeatures
  int *flabel = nullptr;       // [B][RC]  cloud_label_
  int *sort_idx = nullptr;     // [B][RC]  cloud_sort_idx_
  unsigned long long *sort_scratch = nullptr;  // [B][RC]  sort words (segment sort slow path; less-flat voxel keys, buffer A)
  unsigned long long *lfv_keys = nullptr;      // [B][RC]  less-flat voxel keys, buffer B
  struct VoxState *lfv_state = nullptr;        // [B][R]   per ring: what the key stage of the VoxelGrid leaves for the next two
  struct VoxState *lmv_state = nullptr;        // [B][4]   the same for LaserMapping's four VoxelGrid filters
  int *ring_feat_cnt = nullptr;  // [B][R][4]  sharp, less_sharp, flat, less_flat_ds per ring
  int *ring_sharp = nullptr;     // [B][R][12]
  int *ring_less_sharp = nullptr;// [B][R][120]
  int *ring_flat = nullptr;      // [B][R][24]
  int *sharp_idx = nullptr, *less_sharp_idx = nullptr, *flat_idx = nullptr;  // [B][R*12], [B][R*120], [B][R*24]
  int *n_feat = nullptr;         // [B][4] sharp, less_sharp, flat, less_flat
  float4 *sharp = nullptr, *flat = nullptr;  // [B][R*12], [B][R*24]
  float4 *lf_stage = nullptr;    // [B][RC] per-ring voxel output staged at the ring's first kept index
  unsigned long long *vox_sort = nullptr;  // [B][vox_cap] composite (voxel key << 32 | point) sort buffer
  int vox_cap = 0;
  // double-buffered "last" clouds (index = scan parity)
  float4 *less_sharp[2] = {nullptr, nullptr};  // [B][R*120]  corner_last_ / less_sharp
  float4 *less_flat[2] = {nullptr, nullptr};   // [B][RC]     surf_last_ / less_flat
  int *ls_ring_off[2] = {nullptr, nullptr};    // [B][R+1]
  int *lf_ring_off[2] = {nullptr, nullptr};    // [B][R+1]
  float4 *az_stage = nullptr;                  // [B][RC]     per-ring azimuth-binned less-flat points (staging)
  float4 *az_pts[2] = {nullptr, nullptr};      // [B][RC]     ... aligned with less_flat[k]: (x, y, z, position in ring)
  int *az_off[2] = {nullptr, nullptr};         // [B][R][AZ_BINS+1] bin starts inside each ring
  int cur = 0;                                 // which buffer holds the CURRENT scan's clouds

  // ---------------- LaserOdometry distortion correction (optional stage) ----------------
  double *imu_q = nullptr;     // [B][10][imu_len]  time, roll, pitch, yaw, shift xyz, velocity xyz (laserOdometry.h:36-46)
  int imu_len = 0;
  int *imu_ptr = nullptr;      // [3][B]  imu_ptr_last_, imu_ptr_last_iter_, points visited
  double *imu_t0 = nullptr;    // [B]  scan_time
  float *imu_start = nullptr;  // [B][16]  r_s_i (9), shift_start (3), velo_start (3) of the sweep (laserOdometry.cpp:633-639)

  // ---------------- LaserOdometry scan-to-scan ----------------
  GridIndex g_surf_last, g_corner_last;
  double *lo_params = nullptr;  // [B][6]
  Pose *lo_pose = nullptr;      // [B] transformToStart of the current params_ (refreshed before each association)
  double *t_w = nullptr;        // [B][3]
  double *r_w = nullptr;        // [B][9]
  int *lo_init = nullptr;       // [B]
  float *lo_surf_res = nullptr;   // [B][R*24][12]  cp, lpj, lpl, lpm
  int *lo_surf_corr = nullptr;    // [B][R*24][4]   j, closest, idx2, idx3 (closest<0: no residual)
  float *lo_corner_res = nullptr; // [B][R*12][9]
  int *lo_corner_corr = nullptr;  // [B][R*12][3]
  AlegoSolveReport *lo_report = nullptr;  // [B]
  double *lo_trace = nullptr;     // [B][2*(iters+1)][7]
  int *lo_trace_n = nullptr;      // [B]

  // ---------------- LaserMapping ----------------
  int map_cap_c = 0, map_cap_s = 0;
  float4 *map_corner = nullptr, *map_surf = nullptr;  // [B][cap]
  int *n_map_corner = nullptr, *n_map_surf = nullptr; // [B]
  GridIndex g_map_corner, g_map_surf;
  MapRows rows_map_surf;          // voxel-row index of the surf map when it is pcl::VoxelGrid output (map_rows.cuh)
  bool map_rows_checked = false;  // rows_map_surf.usable is up to date with the current map clouds
  bool map_index_valid = false;
  int lm_cap_c = 0, lm_cap_s = 0, lm_cap_o = 0;       // capacities of stand-alone inputs
  float4 *lm_in_corner = nullptr, *lm_in_surf = nullptr, *lm_in_outlier = nullptr;  // [B][cap]
  int *lm_in_n = nullptr;      // [B][4] corner, surf, outlier counts of stand-alone inputs
  int *lm_use_ext = nullptr;   // [B] 1 = read lm_in_*, 0 = read LO outputs
  float4 *lm_corner_ds = nullptr, *lm_surf_ds = nullptr, *lm_outlier_ds = nullptr, *lm_surf_total = nullptr,
         *lm_surf_total_ds = nullptr;
  int ds_cap_c = 0, ds_cap_s = 0, ds_cap_o = 0;
  int *lm_n = nullptr;         // [B][8] corner_ds, surf_ds, outlier_ds, surf_total, surf_total_ds
  double *lm_params = nullptr; // [B][6]
  Pose *m2o = nullptr, *o2l = nullptr, *m2l = nullptr;  // [B]
  Pose *o2l_lo[2] = {nullptr, nullptr};  // [B] (t_w_cur_, r_w_cur_) of LaserOdometry after the sweep held in buffer parity k
  double *lm_edge = nullptr;   // [B][ds_cap_c][10]: valid, cp3, lpj3, lpl3
  double *lm_plane = nullptr;  // [B][ds_cap_s+o][8]: valid, cp3, n3, d
  int *lm_nn_c = nullptr, *lm_nn_s = nullptr;  // [B][cap][5] map indices of the gated 5-NN (first = -1: none)
  AlegoSolveReport *lm_report = nullptr;
  int *lm_guard = nullptr;     // [B] scan2MapOptimization guard (laserMapping.cpp:350)
  double *lm_trace = nullptr;  // [B][outer*(iters+1)][7]
  int *lm_trace_n = nullptr;
  int lm_trace_cap = 0, lo_trace_cap = 0;

  // ---------------- loop-closure ICP (one cloud pair at a time, independent of B) ----------------
  float4 *icp_src = nullptr, *icp_src0 = nullptr, *icp_tgt = nullptr;  // current / original source, target
  int icp_cap_src = 0, icp_cap_tgt = 0, icp_trace_cap = 0;
  GridIndex g_icp;
  double *icp_partials = nullptr, *icp_trace = nullptr;
  void *icp_state = nullptr;  // IcpState (lc_icp_kernels.cu)
  int *icp_n = nullptr;       // [2] source / target point counts

  // host staging for small D2H results
  double *h_pose = nullptr;  // pinned [B][12]
  double *h_pose_slot[ALEGO_INFLIGHT] = {};  // pinned, per in-flight step
  double *d_pose = nullptr;  // [B][12]
};

// ------------------------------------------------------------------------------------------------
#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                            \
      return ALEGO_CUDA_ERROR;                                                                  \
    }                                                                                           \
  } while (0)

// Bracket a launch with events when profiling is on; always count it.
struct LaunchScope {
  AlegoHandle *h;
  int id = -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  LaunchScope(AlegoHandle *hh, const char *name) : h(hh) {
    ++h->launches;
    if (!h->profiling) return;
    for (size_t i = 0; i < h->prof.size(); ++i)
      if (h->prof[i].name == name) { id = (int)i; break; }
    if (id < 0) {
      h->prof.push_back(KernelProfile());
      h->prof.back().name = name;
      id = (int)h->prof.size() - 1;
    }
    auto get = [&]() {
      cudaEvent_t e;
      if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    e0 = get(); e1 = get();
    cudaEventRecord(e0, h->launch_stream ? h->launch_stream : h->stream);
  }
  ~LaunchScope() {
    if (id < 0) return;
    cudaEventRecord(e1, h->launch_stream ? h->launch_stream : h->stream);
    h->prof[id].pending.emplace_back(e0, e1);
  }
};
#define ALEGO_CAT2(a, b) a##b
#define ALEGO_CAT(a, b) ALEGO_CAT2(a, b)
#define LAUNCH(h, name) LaunchScope ALEGO_CAT(_ls_, __LINE__)((h), (name))

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exclusive block scan of one int per thread; returns the exclusive prefix, *total gets the block sum.
// smem: at least 33 ints. All threads of the block must call it.
__device__ __forceinline__ int block_excl_scan(int v, int *smem, int *total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  if (total) *total = smem[32];
  return smem[wid] + inc - v;
}

// Approximate atan2f for decisions that are re-checked exactly (ImageProjection's row / column binning) or only need a
// conservative bound (azimuth bins of LaserOdometry's ring search): |error| < 1e-6 rad (degree-15 odd minimax polynomial on [0,1], 1.5e-7, plus the
// approximate division and the quadrant folds).  Returns false for operands it does not cover (zero / denormal / huge).
// single MUFU instructions (2-ulp approximations) for operands known to be normal floats
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define ALEGO_FAST_ATAN_ERR 1e-6
__device__ __forceinline__ bool fast_atan2(float y, float x, float &r) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  if (!(mx > 1e-30f && ax + ay < 1e30f)) return false;  // also rejects NaN / inf operands (fmaxf would drop a NaN)
  const float t = mn * rcp_approx(mx);  // mx is in [1e-30, 1e30]: no denormal / overflow handling needed
  const float u = t * t;
  float p = -0.00405456405133009f;
  p = __fmaf_rn(p, u, 0.021862948313355446f);
  p = __fmaf_rn(p, u, -0.055912312120199203f);
  p = __fmaf_rn(p, u, 0.09642196446657181f);
  p = __fmaf_rn(p, u, -0.1390862911939621f);
  p = __fmaf_rn(p, u, 0.19946566224098206f);
  p = __fmaf_rn(p, u, -0.33329859375953674f);
  p = __fmaf_rn(p, u, 0.9999993443489075f);
  float a = p * t;
  if (ay > ax) a = 1.57079637f - a;
  if (x < 0.f) a = 3.14159274f - a;
  r = copysignf(a, y);
  return true;
}

// Azimuth bins of a ring-ordered feature cloud (LaserOdometry's adjacent-ring search): 64 bins of 5.625 degrees, bin index
// wraps (azimuth pi and -pi are the same direction).  The 1e-6 rad error of fast_atan2 is covered by the search margins.
#define AZ_BINS 64
__device__ __forceinline__ float az_angle(float x, float y) {
  float a;
  if (!fast_atan2(y, x, a)) a = 0.f;  // on the axis of rotation (or absurd magnitudes): any bin, see lo_assoc
  return a;
}
__device__ __forceinline__ int az_bin_unwrapped(float a) { return (int)floorf((a + 3.14159274f) * (AZ_BINS / 6.28318531f)); }

// Rz(yaw)*Ry(pitch)*Rx(roll) and the trig terms, as in every cost function of utility.h:128-158
struct PoseTrig {
  double sr, cr, sp, cp, sy, cy;
  double R[9];
  __device__ __forceinline__ explicit PoseTrig(const double *x) {
    sr = sin(x[3]); cr = cos(x[3]);
    sp = sin(x[4]); cp = cos(x[4]);
    sy = sin(x[5]); cy = cos(x[5]);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
  }
};
