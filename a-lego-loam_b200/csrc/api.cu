// api.cu — the extern "C" boundary declared in include/alego_b200.h: handle lifecycle, host<->device
// transfers, stage sequencing, measurement hooks.  No numerics live here.
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "common.cuh"
#include "sort_voxel.cuh"
#include "vox_order.cuh"
#include "grid.cuh"
#include "ip_kernels.cuh"
#include "lm_kernels.cuh"
#include "map_rows.cuh"
#include "lo_kernels.cuh"

namespace {

template <typename T>
int dmalloc(AlegoHandle *h, T **p, size_t count) {
  CUDA_TRY(h, cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
  return ALEGO_OK;
}
#define DMALLOC(h, p, n)                              \
  do {                                                \
    int _rc = dmalloc((h), &(p), (size_t)(n));        \
    if (_rc != ALEGO_OK) return _rc;                  \
  } while (0)

__global__ void init_state_kernel(double *lo_params, double *t_w, double *r_w, int *lo_init, double *lm_params, Pose *m2o, Pose *o2l,
                                  Pose *m2l, int *use_ext, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int q = 0; q < 6; ++q) { lo_params[b * 6 + q] = 0; lm_params[b * 6 + q] = 0; }
  for (int q = 0; q < 3; ++q) { t_w[b * 3 + q] = 0; m2o[b].t[q] = 0; o2l[b].t[q] = 0; m2l[b].t[q] = 0; }
  for (int q = 0; q < 9; ++q) {
    const double e = (q % 4 == 0) ? 1.0 : 0.0;
    r_w[b * 9 + q] = e; m2o[b].R[q] = e; o2l[b].R[q] = e; m2l[b].R[q] = e;
  }
  lo_init[b] = 0;
  use_ext[b] = 0;
}

__global__ void pose_pack_kernel(const Pose *m2l, const double *lm_params, const Pose *lo_pose, double *pose_out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double *po = pose_out + b * 12;
  for (int q = 0; q < 3; ++q) po[q] = m2l[b].t[q];
  for (int q = 0; q < 6; ++q) po[3 + q] = lm_params[b * 6 + q];
  for (int q = 0; q < 3; ++q) po[9 + q] = lo_pose[b].t[q];
}

int check(AlegoHandle *h, int seq) {
  if (!h) return ALEGO_BAD_ARG;
  if (seq < 0 || seq >= h->B) { h->err = "sequence index out of range"; return ALEGO_BAD_ARG; }
  return ALEGO_OK;
}

// The pipeline runs its LaserMapping stage on the side stream: anything that reads results or continues stage by stage on
// the main stream first makes the main stream wait for what the side stream has been given.
int join_side(AlegoHandle *h) {
  if (!h->side_busy) return ALEGO_OK;
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_side_tail, 0));
  h->side_busy = false;
  return ALEGO_OK;
}

// LaserMapping's lazily sized buffers, and — once per map change, with the side stream joined — the decision whether the surf map
// can use the voxel-row index (map_rows.cuh); both outside any stream capture
static int lm_ready(AlegoHandle *h) {
  int rc = lm_ensure_buffers(h, 0, 0, 0);
  if (rc != ALEGO_OK) return rc;
  if (!h->map_rows_checked && h->map_surf) {
    if ((rc = join_side(h)) != ALEGO_OK) return rc;
    rc = lm_validate_map_rows(h);
  }
  return rc;
}

int d2h(AlegoHandle *h, void *dst, const void *src, size_t bytes) {
  if (!bytes) return ALEGO_OK;
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  CUDA_TRY(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}
int fetch_int(AlegoHandle *h, const int *src, int *v) { return d2h(h, v, src, sizeof(int)); }

void resolve_profile(AlegoHandle *h) {
  cudaStreamSynchronize(h->stream);
  for (auto &k : h->prof) {
    for (auto &pr : k.pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        k.total_ms += ms;
        ++k.launches;
      }
      h->event_pool.push_back(pr.first);
      h->event_pool.push_back(pr.second);
    }
    k.pending.clear();
  }
}

}  // namespace

extern "C" {

int alego_default_params(AlegoParams *p, int preset) {
  if (!p) return ALEGO_BAD_ARG;
  std::memset(p, 0, sizeof *p);
  p->seg_valid_point_num = 5;   // utility.h:64
  p->seg_valid_line_num = 3;    // utility.h:65
  p->seg_min_cluster = 30;      // imageProjection.cpp:283
  p->lo_surf_iters = 5;         // laserOdometry.cpp:415
  p->lo_corner_iters = 5;       // laserOdometry.cpp:489
  p->lm_outer_iters = 2;        // laserMapping.cpp:360
  p->lm_max_iters = 20;         // laserMapping.cpp:470
  p->sensor_mount_ang = 0.;     // utility.h:58
  p->seg_theta = 1.047;         // utility.h:63
  p->nearest_feature_dist = 25.;// utility.h:73
  p->huber_delta = 0.1;
  p->less_flat_leaf = 0.4;
  p->lm_corner_leaf = 0.4;
  p->lm_surf_leaf = 0.8;
  p->lm_outlier_leaf = 1.0;
  switch (preset) {
    case ALEGO_PRESET_VLP16_1800: p->n_scan = 16; p->ang_res_x = 0.2; p->ang_res_y = 2.0; p->ang_bottom = 15.0; p->ground_scan_id = 7; break;
    case ALEGO_PRESET_HDL64_1800: p->n_scan = 64; p->ang_res_x = 0.2; p->ang_res_y = 0.427; p->ang_bottom = 24.9; p->ground_scan_id = 50; break;
    case ALEGO_PRESET_HDL64_2048: p->n_scan = 64; p->ang_res_x = 360.0 / 2048.0; p->ang_res_y = 0.427; p->ang_bottom = 24.9; p->ground_scan_id = 50; break;
    case ALEGO_PRESET_REFERENCE: p->n_scan = 16; p->ang_res_x = 0.09; p->ang_res_y = 2.0; p->ang_bottom = 15.0; p->ground_scan_id = 10; break;
    default: return ALEGO_BAD_ARG;
  }
  p->horizon_scan = (int)(360.0 / p->ang_res_x + 0.5);  // utility.h:55
  return ALEGO_OK;
}

int alego_create(const AlegoParams *p, int device, int n_seq, int max_points_per_scan, AlegoHandle **out) {
  if (!p || !out || n_seq < 1 || max_points_per_scan < 1) return ALEGO_BAD_ARG;
  if (p->n_scan < 1 || p->n_scan > ALEGO_MAX_RINGS || p->horizon_scan < 12 || p->horizon_scan > 8192) return ALEGO_BAD_ARG;
  if (!(p->ang_res_x > 0) || !(p->ang_res_y > 0) || p->ground_scan_id < 0 || !(p->less_flat_leaf > 0) || !(p->lm_corner_leaf > 0) ||
      !(p->lm_surf_leaf > 0) || !(p->lm_outlier_leaf > 0) || p->lo_surf_iters < 0 || p->lo_corner_iters < 0 || p->lm_outer_iters < 0 ||
      p->lm_max_iters < 0)
    return ALEGO_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || device >= ALEGO_MAX_DEVICES) return ALEGO_CUDA_ERROR;
  AlegoHandle *h = new AlegoHandle();
  *out = h;
  h->P = *p;
  h->dev = device;
  h->B = n_seq;
  h->Nmax = max_points_per_scan;
  h->R = p->n_scan;
  h->C = p->horizon_scan;
  h->RC = h->R * h->C;
  // sin/cos of the segmentation angles with the HOST libm, exactly what the reference evaluates (:269)
  const double ax = p->ang_res_x / 180.0 * M_PI, ay = p->ang_res_y / 180.0 * M_PI;  // ANGLE2RAD, utility.h:48,60-61
  h->seg_sin_x = std::sin(ax); h->seg_cos_x = std::cos(ax);
  h->seg_sin_y = std::sin(ay); h->seg_cos_y = std::cos(ay);
  CUDA_TRY(h, cudaSetDevice(device));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto &e : h->timer) CUDA_TRY(h, cudaEventCreate(&e));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_lo_done, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_side_tail, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_g_fork, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_g_lo, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_g_tail, cudaEventDisableTiming));
  for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_lm_done[k], cudaEventDisableTiming));
  for (int k = 0; k < ALEGO_INFLIGHT; ++k) {
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_copied[k], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_consumed[k], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_pose[k], cudaEventDisableTiming));
  }
  const size_t B = n_seq, RC = h->RC, R = h->R;
  DMALLOC(h, h->raw_own, B * h->Nmax);
  DMALLOC(h, h->n_pts_own, B);
  h->raw = h->raw_own;
  h->n_pts = h->n_pts_own;
  h->raw_slot[0] = h->raw_own;
  h->n_pts_slot[0] = h->n_pts_own;
  DMALLOC(h, h->winner, B * RC);
  DMALLOC(h, h->cloud, B * RC);
  DMALLOC(h, h->range, B * RC);
  DMALLOC(h, h->ground, B * RC);
  DMALLOC(h, h->cell_flags, B * RC);
  DMALLOC(h, h->cell_class, B * RC);
  DMALLOC(h, h->parent, B * RC);
  DMALLOC(h, h->comp_stat, B * RC);
  DMALLOC(h, h->comp_id, B * RC);
  DMALLOC(h, h->label, B * RC);
  DMALLOC(h, h->rowcnt, B * R);
  DMALLOC(h, h->seg_cloud, B * RC);
  DMALLOC(h, h->seg_ground, B * RC);
  DMALLOC(h, h->seg_col, B * RC);
  DMALLOC(h, h->seg_range, B * RC);
  DMALLOC(h, h->start_ring, B * R);
  DMALLOC(h, h->end_ring, B * R);
  DMALLOC(h, h->M, B);
  h->out_cap = std::max(1, (h->R - std::min(h->P.ground_scan_id + 1, h->R))) * ((h->C + 4) / 5);
  for (int k = 0; k < 2; ++k) {
    DMALLOC(h, h->outlier_buf[k], B * h->out_cap);
    DMALLOC(h, h->n_outlier_buf[k], B);
    CUDA_TRY(h, cudaMemsetAsync(h->n_outlier_buf[k], 0, B * sizeof(int), h->stream));
    DMALLOC(h, h->o2l_lo[k], B);
  }
  h->outlier = h->outlier_buf[0];
  h->n_outlier = h->n_outlier_buf[0];
  DMALLOC(h, h->orient, B * 4);
  CUDA_TRY(h, cudaMemsetAsync(h->winner, 0xFF, B * RC * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->M, 0, B * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->n_pts, 0, B * sizeof(int), h->stream));
  // features
  DMALLOC(h, h->curv, B * RC);
  DMALLOC(h, h->picked0, B * RC);
  DMALLOC(h, h->picked, B * RC);
  DMALLOC(h, h->flabel, B * RC);
  DMALLOC(h, h->sort_idx, B * RC);
  DMALLOC(h, h->sort_scratch, B * RC);
  DMALLOC(h, h->lfv_keys, B * RC);
  CUDA_TRY(h, cudaMalloc(&h->lfv_state, (size_t)B * R * sizeof(VoxState)));
  CUDA_TRY(h, cudaMalloc(&h->lmv_state, (size_t)B * 4 * sizeof(VoxState)));
  DMALLOC(h, h->ring_feat_cnt, B * R * 4);
  DMALLOC(h, h->ring_sharp, B * R * 12);
  DMALLOC(h, h->ring_less_sharp, B * R * 120);
  DMALLOC(h, h->ring_flat, B * R * 24);
  DMALLOC(h, h->sharp_idx, B * R * 12);
  DMALLOC(h, h->less_sharp_idx, B * R * 120);
  DMALLOC(h, h->flat_idx, B * R * 24);
  DMALLOC(h, h->n_feat, B * 4);
  DMALLOC(h, h->sharp, B * R * 12);
  DMALLOC(h, h->flat, B * R * 24);
  DMALLOC(h, h->lf_stage, B * RC);
  DMALLOC(h, h->az_stage, B * RC);
  CUDA_TRY(h, cudaMemsetAsync(h->n_feat, 0, B * 4 * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->picked, 0, B * RC, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->flabel, 0, B * RC * sizeof(int), h->stream));
  for (int k = 0; k < 2; ++k) {
    DMALLOC(h, h->less_sharp[k], B * R * 120);
    DMALLOC(h, h->less_flat[k], B * RC);
    DMALLOC(h, h->ls_ring_off[k], B * (R + 1));
    DMALLOC(h, h->lf_ring_off[k], B * (R + 1));
    DMALLOC(h, h->az_pts[k], B * RC);
    DMALLOC(h, h->az_off[k], B * R * (AZ_BINS + 1));
    CUDA_TRY(h, cudaMemsetAsync(h->az_off[k], 0, B * R * (AZ_BINS + 1) * sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->ls_ring_off[k], 0, B * (R + 1) * sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->lf_ring_off[k], 0, B * (R + 1) * sizeof(int), h->stream));
  }
  // scan-to-scan
  int rc = grid_alloc(h, &h->g_surf_last, (int)RC, 0.5f);  // dense cloud (every ring voxelised on its own): small cells keep the buckets short
  if (rc != ALEGO_OK) return rc;
  rc = grid_alloc(h, &h->g_corner_last, (int)(R * 120), 2.0f);  // sparse cloud: bigger cells settle the 1-NN in the first 3x3x3 pass
  if (rc != ALEGO_OK) return rc;
  DMALLOC(h, h->lo_params, B * 6);
  DMALLOC(h, h->lo_pose, B);
  DMALLOC(h, h->t_w, B * 3);
  DMALLOC(h, h->r_w, B * 9);
  DMALLOC(h, h->lo_init, B);
  DMALLOC(h, h->lo_surf_res, B * R * 24 * 12);
  DMALLOC(h, h->lo_surf_corr, B * R * 24 * 4);
  DMALLOC(h, h->lo_corner_res, B * R * 12 * 9);
  DMALLOC(h, h->lo_corner_corr, B * R * 12 * 3);
  DMALLOC(h, h->lo_report, B);
  h->lo_trace_cap = p->lo_surf_iters + p->lo_corner_iters + 4;
  DMALLOC(h, h->lo_trace, B * h->lo_trace_cap * 7);
  DMALLOC(h, h->lo_trace_n, B);
  CUDA_TRY(h, cudaMemsetAsync(h->lo_report, 0, B * sizeof(AlegoSolveReport), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->lo_trace_n, 0, B * sizeof(int), h->stream));
  // mapping
  DMALLOC(h, h->n_map_corner, B);
  DMALLOC(h, h->n_map_surf, B);
  CUDA_TRY(h, cudaMemsetAsync(h->n_map_corner, 0, B * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->n_map_surf, 0, B * sizeof(int), h->stream));
  DMALLOC(h, h->lm_in_n, B * 4);
  DMALLOC(h, h->lm_use_ext, B);
  DMALLOC(h, h->lm_n, B * 8);
  CUDA_TRY(h, cudaMemsetAsync(h->lm_in_n, 0, B * 4 * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->lm_n, 0, B * 8 * sizeof(int), h->stream));
  DMALLOC(h, h->lm_params, B * 6);
  DMALLOC(h, h->m2o, B);
  DMALLOC(h, h->o2l, B);
  DMALLOC(h, h->m2l, B);
  DMALLOC(h, h->lm_report, B);
  CUDA_TRY(h, cudaMemsetAsync(h->lm_report, 0, B * sizeof(AlegoSolveReport), h->stream));
  h->lm_trace_cap = std::max(1, p->lm_outer_iters) * (p->lm_max_iters + 2) + 2;
  DMALLOC(h, h->lm_trace, B * h->lm_trace_cap * 7);
  DMALLOC(h, h->lm_trace_n, B);
  CUDA_TRY(h, cudaMemsetAsync(h->lm_trace_n, 0, B * sizeof(int), h->stream));
  DMALLOC(h, h->lm_guard, B);
  DMALLOC(h, h->d_pose, B * 12);
  CUDA_TRY(h, cudaMallocHost((void **)&h->h_pose, B * 12 * sizeof(double)));
  h->lm_scan_is_external.assign(B, 0);
  init_state_kernel<<<div_up(n_seq, 128), 128, 0, h->stream>>>(h->lo_params, h->t_w, h->r_w, h->lo_init, h->lm_params, h->m2o, h->o2l,
                                                               h->m2l, h->lm_use_ext, n_seq);
  for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaMemcpyAsync(h->o2l_lo[k], h->o2l, B * sizeof(Pose), cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}

void alego_destroy(AlegoHandle *h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->side_stream) cudaStreamSynchronize(h->side_stream);
  for (int k = 1; k < ALEGO_INFLIGHT; ++k) {  // slot 0 aliases raw_own / n_pts_own
    if (h->raw_slot[k]) cudaFree(h->raw_slot[k]);
    if (h->n_pts_slot[k]) cudaFree(h->n_pts_slot[k]);
  }
  for (int k = 0; k < ALEGO_INFLIGHT; ++k) {
    if (h->h_n_pts_slot[k]) cudaFreeHost(h->h_n_pts_slot[k]);
    if (h->h_pose_slot[k]) cudaFreeHost(h->h_pose_slot[k]);
    if (h->ev_copied[k]) cudaEventDestroy(h->ev_copied[k]);
    if (h->ev_consumed[k]) cudaEventDestroy(h->ev_consumed[k]);
    if (h->ev_pose[k]) cudaEventDestroy(h->ev_pose[k]);
    for (int e = 0; e < 5; ++e)
      if (h->ev_t[k][e]) cudaEventDestroy(h->ev_t[k][e]);
  }
  if (h->ev_t_origin) cudaEventDestroy(h->ev_t_origin);
  for (void *p : h->scratch)
    if (p) cudaFree(p);
  if (h->ev_lo_done) cudaEventDestroy(h->ev_lo_done);
  if (h->ev_side_tail) cudaEventDestroy(h->ev_side_tail);
  for (auto &g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaEvent_t e : {h->ev_g_fork, h->ev_g_lo, h->ev_g_tail})
    if (e) cudaEventDestroy(e);
  for (int k = 0; k < 2; ++k)
    if (h->ev_lm_done[k]) cudaEventDestroy(h->ev_lm_done[k]);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto p : h->stage_raw) cudaFree(p);
  for (auto p : h->stage_n) cudaFree(p);
  void *ptrs[] = {h->raw_own, h->n_pts_own, h->winner, h->cloud, h->range, h->ground, h->cell_flags, h->cell_class, h->parent, h->comp_stat,
                  h->comp_id, h->label, h->rowcnt, h->seg_cloud, h->seg_ground, h->seg_col, h->seg_range, h->start_ring, h->end_ring,
                  h->M, h->outlier_buf[0], h->outlier_buf[1], h->n_outlier_buf[0], h->n_outlier_buf[1], h->o2l_lo[0], h->o2l_lo[1], h->orient, h->curv, h->picked0, h->picked, h->flabel, h->sort_idx, h->sort_scratch, h->lfv_keys, h->lfv_state, h->lmv_state,
                  h->ring_feat_cnt, h->ring_sharp, h->ring_less_sharp, h->ring_flat, h->sharp_idx, h->less_sharp_idx, h->flat_idx,
                  h->n_feat, h->sharp, h->flat, h->lf_stage, h->vox_sort, h->less_sharp[0], h->less_sharp[1], h->less_flat[0],
                  h->less_flat[1], h->ls_ring_off[0], h->ls_ring_off[1], h->lf_ring_off[0], h->lf_ring_off[1], h->lo_pose, h->az_stage, h->az_pts[0], h->az_pts[1], h->az_off[0], h->az_off[1], h->lo_params, h->t_w,
                  h->r_w, h->lo_init, h->lo_surf_res, h->lo_surf_corr, h->lo_corner_res, h->lo_corner_corr, h->lo_report, h->lo_trace,
                  h->lo_trace_n, h->map_corner, h->map_surf, h->n_map_corner, h->n_map_surf, h->lm_in_corner, h->lm_in_surf,
                  h->lm_in_outlier, h->lm_in_n, h->lm_use_ext, h->lm_corner_ds, h->lm_surf_ds, h->lm_outlier_ds, h->lm_surf_total,
                  h->lm_surf_total_ds, h->lm_n, h->lm_params, h->m2o, h->o2l, h->m2l, h->lm_edge, h->lm_plane, h->lm_report,
                  h->lm_trace, h->lm_trace_n, h->lm_guard, h->d_pose, h->lm_nn_c, h->lm_nn_s, h->imu_q, h->imu_ptr, h->imu_t0, h->imu_start, h->icp_src, h->icp_src0, h->icp_tgt,
                  h->icp_partials, h->icp_trace, h->icp_state, h->icp_n};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  grid_free(&h->g_surf_last);
  grid_free(&h->g_corner_last);
  grid_free(&h->g_map_corner);
  grid_free(&h->g_map_surf);
  map_rows_free(&h->rows_map_surf);
  grid_free(&h->g_icp);
  if (h->h_pose) cudaFreeHost(h->h_pose);
  for (auto &k : h->prof)
    for (auto &pr : k.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (auto e : h->timer)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *alego_last_error(const AlegoHandle *h) { return h ? h->err.c_str() : "null handle"; }
int alego_synchronize(AlegoHandle *h) {
  if (!h) return ALEGO_BAD_ARG;
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}
int alego_get_params(const AlegoHandle *h, AlegoParams *out) {
  if (!h || !out) return ALEGO_BAD_ARG;
  *out = h->P;
  return ALEGO_OK;
}
int alego_n_seq(const AlegoHandle *h) { return h ? h->B : ALEGO_BAD_ARG; }

int alego_set_point_stride(AlegoHandle *h, int floats_per_point) {
  if (!h || (floats_per_point != 3 && floats_per_point != 4)) return ALEGO_BAD_ARG;
  if (h->n_submitted != h->n_collected) { h->err = "alego_set_point_stride while submitted steps are in flight"; return ALEGO_NOT_READY; }
  h->in_stride = floats_per_point;
  return ALEGO_OK;
}

void *alego_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void alego_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------
static int upload_into(AlegoHandle *h, float4 *raw, int *n_dev, const float *xyzi_host, const int32_t *n_points, cudaStream_t st = nullptr);

int alego_ip_upload(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points) {
  if (!h || !xyzi_host || !n_points) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  h->raw = h->raw_own;
  h->n_pts = h->n_pts_own;
  return upload_into(h, h->raw, h->n_pts, xyzi_host, n_points);
}

// Pre-stage sweeps in HBM: slot k holds one sweep per sequence; alego_stage_select makes it the input of the
// next alego_ip_run / alego_pipeline_step(NULL, ...) without any copy.
int alego_stage_upload(AlegoHandle *h, int slot, const float *xyzi_host, const int32_t *n_points) {
  if (!h || slot < 0 || slot > 4096 || !xyzi_host || !n_points) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  while ((int)h->stage_raw.size() <= slot) {
    float4 *r = nullptr;
    int *n = nullptr;
    CUDA_TRY(h, cudaMalloc(&r, (size_t)h->B * h->Nmax * sizeof(float4)));
    CUDA_TRY(h, cudaMalloc(&n, (size_t)h->B * sizeof(int)));
    h->stage_raw.push_back(r);
    h->stage_n.push_back(n);
  }
  int rc = upload_into(h, h->stage_raw[slot], h->stage_n[slot], xyzi_host, n_points);
  if (rc != ALEGO_OK) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}
int alego_stage_select(AlegoHandle *h, int slot) {
  if (!h || slot < 0 || slot >= (int)h->stage_raw.size()) return ALEGO_BAD_ARG;
  h->raw = h->stage_raw[slot];
  h->n_pts = h->stage_n[slot];
  return ALEGO_OK;
}

static int upload_into(AlegoHandle *h, float4 *raw, int *n_dev, const float *xyzi_host, const int32_t *n_points, cudaStream_t st) {
  if (!st) st = h->stream;
  size_t total = 0;
  for (int b = 0; b < h->B; ++b) {
    if (n_points[b] < 0 || n_points[b] > h->Nmax) { h->err = "n_points out of range"; return ALEGO_BAD_ARG; }
    total += n_points[b];
  }
  CUDA_TRY(h, cudaMemcpyAsync(n_dev, n_points, h->B * sizeof(int), cudaMemcpyHostToDevice, st));
  const size_t pt_bytes = (size_t)h->in_stride * sizeof(float), row = (size_t)h->Nmax * h->in_stride;  // floats per sequence
  float *dst = reinterpret_cast<float *>(raw);
  if (total * 10 >= (size_t)h->B * h->Nmax * 9) {  // nearly full rows: one strided DMA of the longest row's width
    int widest = 0;
    for (int b = 0; b < h->B; ++b) widest = std::max(widest, n_points[b]);
    if (widest > 0)
      CUDA_TRY(h, cudaMemcpy2DAsync(dst, (size_t)h->Nmax * pt_bytes, xyzi_host, (size_t)h->Nmax * pt_bytes, (size_t)widest * pt_bytes, (size_t)h->B,
                                    cudaMemcpyHostToDevice, st));
  } else {
    for (int b = 0; b < h->B; ++b)
      if (n_points[b] > 0)
        CUDA_TRY(h, cudaMemcpyAsync(dst + (size_t)b * row, xyzi_host + (size_t)b * row, (size_t)n_points[b] * pt_bytes, cudaMemcpyHostToDevice, st));
  }
  return ALEGO_OK;
}

static int ip_run_internal(AlegoHandle *h) {
  // the outlier cloud travels to LaserMapping with the feature clouds of the same sweep: same buffer parity
  h->outlier = h->outlier_buf[h->cur];
  h->n_outlier = h->n_outlier_buf[h->cur];
  const int rc = ip_run_device(h, h->want_labels);
  if (rc == ALEGO_OK) { h->stage_ip_done = true; h->stage_feat_done = false; }
  return rc;
}

int alego_ip_run(AlegoHandle *h) {
  if (!h) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  return ip_run_internal(h);
}

int alego_ip_process(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points) {
  const int rc = alego_ip_upload(h, xyzi_host, n_points);
  if (rc != ALEGO_OK) return rc;
  return alego_ip_run(h);
}

int alego_ip_get(AlegoHandle *h, int seq, AlegoCloudInfo *info, float *segmented_xyzi, float *outlier_xyzi, int32_t *n_outlier,
                 int32_t *label_image) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (!h->stage_ip_done) { h->err = "alego_ip_get before alego_ip_process"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  const size_t base = (size_t)seq * h->RC;
  int M = 0, no = 0;
  if ((rc = fetch_int(h, h->M + seq, &M)) != ALEGO_OK) return rc;
  if ((rc = fetch_int(h, h->n_outlier + seq, &no)) != ALEGO_OK) return rc;
  if (info) {
    info->size = M;
    float o[4];
    if ((rc = d2h(h, o, h->orient + seq * 4, sizeof o)) != ALEGO_OK) return rc;
    info->startOrientation = o[0]; info->endOrientation = o[1]; info->orientationDiff = o[2];
    if (info->startRingIndex && (rc = d2h(h, info->startRingIndex, h->start_ring + seq * h->R, h->R * sizeof(int))) != ALEGO_OK) return rc;
    if (info->endRingIndex && (rc = d2h(h, info->endRingIndex, h->end_ring + seq * h->R, h->R * sizeof(int))) != ALEGO_OK) return rc;
    if (info->segmentedCloudGroundFlag && (rc = d2h(h, info->segmentedCloudGroundFlag, h->seg_ground + base, M)) != ALEGO_OK) return rc;
    if (info->segmentedCloudColInd && (rc = d2h(h, info->segmentedCloudColInd, h->seg_col + base, M * sizeof(int))) != ALEGO_OK) return rc;
    if (info->segmentedCloudRange && (rc = d2h(h, info->segmentedCloudRange, h->seg_range + base, M * sizeof(float))) != ALEGO_OK) return rc;
  }
  if (segmented_xyzi && (rc = d2h(h, segmented_xyzi, h->seg_cloud + base, (size_t)M * sizeof(float4))) != ALEGO_OK) return rc;
  if (outlier_xyzi && (rc = d2h(h, outlier_xyzi, h->outlier + (size_t)seq * h->out_cap, (size_t)no * sizeof(float4))) != ALEGO_OK) return rc;
  if (n_outlier) *n_outlier = no;
  if (label_image) {
    if ((rc = ip_label_device(h)) != ALEGO_OK) return rc;
    if ((rc = d2h(h, label_image, h->label + base, (size_t)h->RC * sizeof(int))) != ALEGO_OK) return rc;
  }
  return ALEGO_OK;
}

// ---------------------------------------------------------------------------------------------------
int alego_lo_extract(AlegoHandle *h) {
  if (!h) return ALEGO_BAD_ARG;
  if (!h->stage_ip_done) { h->err = "alego_lo_extract before alego_ip_process"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  const int rc = lo_extract_device(h);
  if (rc == ALEGO_OK) { h->stage_feat_done = true; h->feat_buf = h->cur; }
  return rc;
}

// adjustDistortion (laserOdometry.cpp:557-726, IMU branch): between ImageProjection and the feature stage, in place on the
// segmented cloud of every sequence.
int alego_lo_adjust_distortion(AlegoHandle *h, const double *scan_time, AlegoImuQueue *queues, double scan_period, int32_t *n_adjusted) {
  if (!h || !scan_time || !queues || !(scan_period > 0.)) return ALEGO_BAD_ARG;
  if (!h->stage_ip_done) { h->err = "alego_lo_adjust_distortion before alego_ip_process"; return ALEGO_NOT_READY; }
  const int B = h->B, len = queues[0].length;
  if (len < 1) { h->err = "alego_lo_adjust_distortion: empty IMU queue"; return ALEGO_BAD_ARG; }
  std::vector<double> q((size_t)B * 10 * len);
  std::vector<int> ptr(3 * (size_t)B, 0);
  for (int b = 0; b < B; ++b) {
    const AlegoImuQueue &Q = queues[b];
    const double *src[10] = {Q.time, Q.roll, Q.pitch, Q.yaw, Q.shift_x, Q.shift_y, Q.shift_z, Q.velo_x, Q.velo_y, Q.velo_z};
    if (Q.length != len || Q.ptr_last >= len || Q.ptr_last_iter < 0 || Q.ptr_last_iter >= len) {
      h->err = "alego_lo_adjust_distortion: inconsistent IMU queue"; return ALEGO_BAD_ARG;
    }
    for (int k = 0; k < 10; ++k) {
      if (!src[k]) { h->err = "alego_lo_adjust_distortion: null IMU array"; return ALEGO_BAD_ARG; }
      std::memcpy(q.data() + ((size_t)b * 10 + k) * len, src[k], (size_t)len * sizeof(double));
    }
    // the forward-only pointer walk of :587-595 equals a search only for stamps that do not run backwards over the live
    // entries imu_ptr_last_iter_ .. imu_ptr_last_ (any IMU driver; the reference assumes it as well)
    if (Q.ptr_last > 0)
      for (int k = Q.ptr_last_iter; k != Q.ptr_last; k = (k + 1) % len)
        if (Q.time[(k + 1) % len] < Q.time[k]) { h->err = "alego_lo_adjust_distortion: IMU stamps run backwards"; return ALEGO_BAD_ARG; }
    ptr[b] = Q.ptr_last;
    ptr[B + b] = Q.ptr_last_iter;
  }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  if (len != h->imu_len) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->imu_q) cudaFree(h->imu_q);
    h->imu_q = nullptr;
    h->imu_len = 0;
    DMALLOC(h, h->imu_q, (size_t)B * 10 * len);
    if (!h->imu_ptr) DMALLOC(h, h->imu_ptr, 3 * (size_t)B);
    if (!h->imu_t0) DMALLOC(h, h->imu_t0, B);
    if (!h->imu_start) DMALLOC(h, h->imu_start, 16 * (size_t)B);
    h->imu_len = len;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->imu_q, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->imu_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->imu_t0, scan_time, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = lo_adjust_distortion_device(h, h->imu_q, len, h->imu_ptr, h->imu_ptr + B, h->imu_t0, scan_period, h->imu_ptr + 2 * B, h->imu_start);
  if (rc != ALEGO_OK) return rc;
  if ((rc = d2h(h, ptr.data(), h->imu_ptr, ptr.size() * sizeof(int))) != ALEGO_OK) return rc;
  for (int b = 0; b < B; ++b) {
    queues[b].ptr_last_iter = ptr[B + b];
    if (n_adjusted) n_adjusted[b] = ptr[2 * B + b];
  }
  return ALEGO_OK;
}

int alego_lo_get_features(AlegoHandle *h, int seq, int32_t *sharp_idx, int32_t *n_sharp, int32_t *less_sharp_idx, int32_t *n_less_sharp,
                          int32_t *flat_idx, int32_t *n_flat, float *less_flat_xyzi, int32_t *n_less_flat, int32_t *cloud_label) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (h->feat_buf < 0) { h->err = "alego_lo_get_features before alego_lo_extract"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  int nf[4], M = 0;
  if ((rc = d2h(h, nf, h->n_feat + seq * 4, sizeof nf)) != ALEGO_OK) return rc;
  if ((rc = fetch_int(h, h->M + seq, &M)) != ALEGO_OK) return rc;
  const int R = h->R;
  if (n_sharp) *n_sharp = nf[0];
  if (n_less_sharp) *n_less_sharp = nf[1];
  if (n_flat) *n_flat = nf[2];
  if (n_less_flat) *n_less_flat = nf[3];
  if (sharp_idx && (rc = d2h(h, sharp_idx, h->sharp_idx + (size_t)seq * R * 12, nf[0] * sizeof(int))) != ALEGO_OK) return rc;
  if (less_sharp_idx && (rc = d2h(h, less_sharp_idx, h->less_sharp_idx + (size_t)seq * R * 120, nf[1] * sizeof(int))) != ALEGO_OK) return rc;
  if (flat_idx && (rc = d2h(h, flat_idx, h->flat_idx + (size_t)seq * R * 24, nf[2] * sizeof(int))) != ALEGO_OK) return rc;
  if (less_flat_xyzi && (rc = d2h(h, less_flat_xyzi, h->less_flat[h->feat_buf] + (size_t)seq * h->RC, (size_t)nf[3] * sizeof(float4))) != ALEGO_OK)
    return rc;
  if (cloud_label && (rc = d2h(h, cloud_label, h->flabel + (size_t)seq * h->RC, (size_t)M * sizeof(int))) != ALEGO_OK) return rc;
  return ALEGO_OK;
}

int alego_lo_scan2scan(AlegoHandle *h, AlegoSolveReport *reports) {
  if (!h) return ALEGO_BAD_ARG;
  if (!h->stage_feat_done) { h->err = "alego_lo_scan2scan before alego_lo_extract"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  int rc = lo_scan2scan_device(h);
  if (rc != ALEGO_OK) return rc;
  h->stage_feat_done = false;  // the features were consumed (they are now the "last" clouds)
  if (reports) {
    if ((rc = d2h(h, reports, h->lo_report, h->B * sizeof(AlegoSolveReport))) != ALEGO_OK) return rc;
    for (int b = 0; b < h->B; ++b)
      if (reports[b].status != ALEGO_OK) return ALEGO_FEW_FEATURES;
  }
  return ALEGO_OK;
}

int alego_lo_get_state(AlegoHandle *h, int seq, double params[6], double t_w[3], double r_w[9]) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (params && (rc = d2h(h, params, h->lo_params + seq * 6, 6 * sizeof(double))) != ALEGO_OK) return rc;
  if (t_w && (rc = d2h(h, t_w, h->t_w + seq * 3, 3 * sizeof(double))) != ALEGO_OK) return rc;
  if (r_w && (rc = d2h(h, r_w, h->r_w + seq * 9, 9 * sizeof(double))) != ALEGO_OK) return rc;
  return ALEGO_OK;
}

int alego_lo_set_params(AlegoHandle *h, int seq, const double params[6]) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK || !params) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  CUDA_TRY(h, cudaMemcpyAsync(h->lo_params + seq * 6, params, 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}

// ---------------------------------------------------------------------------------------------------
static int realloc_keep(AlegoHandle *h, float4 **buf, int *cap, int need) {
  // an EMPTY cloud still gets a buffer: the first mapped frames of a run have no keyframe yet, the map clouds are empty and
  // scan2MapOptimization's guard skips the solve (laserMapping.cpp:196-199, :350-354) — that is a valid state, not "no map"
  if (*buf && need <= *cap) return ALEGO_OK;
  ++h->graph_epoch;  // captured graphs hold the old pointer
  const int ncap = std::max(need, 1024);
  float4 *nb = nullptr;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMalloc(&nb, (size_t)h->B * ncap * sizeof(float4)));
  if (*buf) {
    for (int b = 0; b < h->B; ++b)
      CUDA_TRY(h, cudaMemcpyAsync(nb + (size_t)b * ncap, *buf + (size_t)b * *cap, (size_t)*cap * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(*buf);
  }
  *buf = nb;
  *cap = ncap;
  return ALEGO_OK;
}

int alego_lm_set_map(AlegoHandle *h, int seq, const float *corner_xyzi, int32_t n_corner, const float *surf_xyzi, int32_t n_surf) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (n_corner < 0 || n_surf < 0 || (n_corner && !corner_xyzi) || (n_surf && !surf_xyzi)) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  const int old_c = h->map_cap_c, old_s = h->map_cap_s;
  if ((rc = realloc_keep(h, &h->map_corner, &h->map_cap_c, n_corner)) != ALEGO_OK) return rc;
  if ((rc = realloc_keep(h, &h->map_surf, &h->map_cap_s, n_surf)) != ALEGO_OK) return rc;
  if (h->map_cap_c != old_c && (rc = grid_alloc(h, &h->g_map_corner, h->map_cap_c, 1.01f, 1)) != ALEGO_OK) return rc;
  if (h->map_cap_s != old_s && (rc = grid_alloc(h, &h->g_map_surf, h->map_cap_s, 1.01f, 1)) != ALEGO_OK) return rc;
  if (n_corner) CUDA_TRY(h, cudaMemcpyAsync(h->map_corner + (size_t)seq * h->map_cap_c, corner_xyzi, (size_t)n_corner * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  if (n_surf) CUDA_TRY(h, cudaMemcpyAsync(h->map_surf + (size_t)seq * h->map_cap_s, surf_xyzi, (size_t)n_surf * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->n_map_corner + seq, &n_corner, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->n_map_surf + seq, &n_surf, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->map_index_valid = false;
  h->map_rows_checked = false;
  return ALEGO_OK;
}

// Rotation of a keyframe pose the way transformPointCloud builds it (laserMapping.h:163-177):
// (AngleAxisf(yaw, Z) * AngleAxisf(pitch, Y) * AngleAxisf(roll, X)).toRotationMatrix() — Eigen turns every AngleAxis into a
// quaternion (w = cos(a/2), axis * sin(a/2)), multiplies the quaternions and expands the product; all in float, evaluated
// on the host like the reference does.  M: row-major 3x4 (rotation | translation).
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

int alego_lm_assemble_map(AlegoHandle *h, int seq, int n_keyframes, const float *const *corner_xyzi, const int32_t *n_corner,
                          const float *const *surf_xyzi, const int32_t *n_surf, const float *const *outlier_xyzi,
                          const int32_t *n_outlier, const float *poses6) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (n_keyframes < 0 || (n_keyframes && (!corner_xyzi || !n_corner || !surf_xyzi || !n_surf || !outlier_xyzi || !n_outlier || !poses6)))
    return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  const int K = n_keyframes;
  std::vector<float> M((size_t)std::max(K, 1) * 12);
  std::vector<const float *> cp(K), sp(2 * (size_t)K);
  std::vector<int> cn(K), sn(2 * (size_t)K);
  long long tot_c = 0, tot_s = 0;
  for (int k = 0; k < K; ++k) {
    if (n_corner[k] < 0 || n_surf[k] < 0 || n_outlier[k] < 0 || (n_corner[k] && !corner_xyzi[k]) || (n_surf[k] && !surf_xyzi[k]) ||
        (n_outlier[k] && !outlier_xyzi[k]))
      return ALEGO_BAD_ARG;
    keyframe_matrix(poses6 + (size_t)k * 6, M.data() + (size_t)k * 12);
    cp[k] = corner_xyzi[k]; cn[k] = n_corner[k];
    // surf_from_map_ += surf keyframe, then the outlier keyframe (:239-243)
    sp[2 * k] = surf_xyzi[k]; sn[2 * k] = n_surf[k];
    sp[2 * k + 1] = outlier_xyzi[k]; sn[2 * k + 1] = n_outlier[k];
    tot_c += n_corner[k];
    tot_s += (long long)n_surf[k] + n_outlier[k];
  }
  if (tot_c > 0x3fffffff || tot_s > 0x3fffffff) { h->err = "alego_lm_assemble_map: too many points"; return ALEGO_BAD_ARG; }
  const int old_c = h->map_cap_c, old_s = h->map_cap_s;
  if ((rc = realloc_keep(h, &h->map_corner, &h->map_cap_c, (int)tot_c)) != ALEGO_OK) return rc;
  if ((rc = realloc_keep(h, &h->map_surf, &h->map_cap_s, (int)tot_s)) != ALEGO_OK) return rc;
  if (h->map_cap_c != old_c && (rc = grid_alloc(h, &h->g_map_corner, h->map_cap_c, 1.01f, 1)) != ALEGO_OK) return rc;
  if (h->map_cap_s != old_s && (rc = grid_alloc(h, &h->g_map_surf, h->map_cap_s, 1.01f, 1)) != ALEGO_OK) return rc;
  // ds_corner_ / ds_surf_ (:316-319)
  if ((rc = lm_assemble_cloud(h, cp.data(), cn.data(), K, M.data(), std::max(K, 1), 0, (float)h->P.lm_corner_leaf,
                              h->map_corner + (size_t)seq * h->map_cap_c, h->n_map_corner + seq)) != ALEGO_OK) return rc;
  if ((rc = lm_assemble_cloud(h, sp.data(), sn.data(), 2 * K, M.data(), std::max(K, 1), 1, (float)h->P.lm_surf_leaf,
                              h->map_surf + (size_t)seq * h->map_cap_s, h->n_map_surf + seq)) != ALEGO_OK) return rc;
  h->map_index_valid = false;
  h->map_rows_checked = false;
  return ALEGO_OK;
}

int alego_lm_get_map(AlegoHandle *h, int seq, float *corner_xyzi, int32_t *n_corner, float *surf_xyzi, int32_t *n_surf) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (!h->map_corner || !h->map_surf) { h->err = "alego_lm_get_map: no local map"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  int nc = 0, ns = 0;
  if ((rc = fetch_int(h, h->n_map_corner + seq, &nc)) != ALEGO_OK) return rc;
  if ((rc = fetch_int(h, h->n_map_surf + seq, &ns)) != ALEGO_OK) return rc;
  if (n_corner) *n_corner = nc;
  if (n_surf) *n_surf = ns;
  if (corner_xyzi && (rc = d2h(h, corner_xyzi, h->map_corner + (size_t)seq * h->map_cap_c, (size_t)nc * sizeof(float4))) != ALEGO_OK) return rc;
  if (surf_xyzi && (rc = d2h(h, surf_xyzi, h->map_surf + (size_t)seq * h->map_cap_s, (size_t)ns * sizeof(float4))) != ALEGO_OK) return rc;
  return ALEGO_OK;
}

// The ICP of performLoopClosure (laserMapping.cpp:667-688)
int alego_lc_icp(AlegoHandle *h, const float *source_xyzi, int32_t n_source, const float *target_xyzi, int32_t n_target,
                 double max_correspondence_distance, int32_t max_iterations, double transformation_epsilon,
                 double euclidean_fitness_epsilon, AlegoIcpResult *out, double *trace) {
  if (!h || !out || n_source < 0 || n_target < 0 || (n_source && !source_xyzi) || (n_target && !target_xyzi) || max_iterations < 1 ||
      !(max_correspondence_distance > 0.))
    return ALEGO_BAD_ARG;
  if (n_source == 0 || n_target == 0) {  // PCL: "Invalid or empty point cloud dataset given" — align() returns without converging
    std::memset(out, 0, sizeof *out);
    for (int k = 0; k < 4; ++k) out->final_transformation[k * 5] = 1.f;
    out->convergence_state = 5;
    out->fitness_score = 1.7976931348623157e308;
    return ALEGO_FEW_FEATURES;
  }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  return lc_icp_device(h, source_xyzi, n_source, target_xyzi, n_target, max_correspondence_distance, max_iterations,
                       transformation_epsilon, euclidean_fitness_epsilon, out, trace);
}

int alego_lm_set_scan(AlegoHandle *h, int seq, const float *corner_xyzi, int32_t n_corner, const float *surf_xyzi, int32_t n_surf,
                      const float *outlier_xyzi, int32_t n_outlier) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (n_corner < 0 || n_surf < 0 || n_outlier < 0) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  if ((rc = realloc_keep(h, &h->lm_in_corner, &h->lm_cap_c, n_corner)) != ALEGO_OK) return rc;
  if ((rc = realloc_keep(h, &h->lm_in_surf, &h->lm_cap_s, n_surf)) != ALEGO_OK) return rc;
  if ((rc = realloc_keep(h, &h->lm_in_outlier, &h->lm_cap_o, n_outlier)) != ALEGO_OK) return rc;
  if (n_corner) CUDA_TRY(h, cudaMemcpyAsync(h->lm_in_corner + (size_t)seq * h->lm_cap_c, corner_xyzi, (size_t)n_corner * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  if (n_surf) CUDA_TRY(h, cudaMemcpyAsync(h->lm_in_surf + (size_t)seq * h->lm_cap_s, surf_xyzi, (size_t)n_surf * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  if (n_outlier) CUDA_TRY(h, cudaMemcpyAsync(h->lm_in_outlier + (size_t)seq * h->lm_cap_o, outlier_xyzi, (size_t)n_outlier * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  const int n4[4] = {n_corner, n_surf, n_outlier, 0};
  const int one = 1;
  CUDA_TRY(h, cudaMemcpyAsync(h->lm_in_n + seq * 4, n4, sizeof n4, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->lm_use_ext + seq, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->lm_scan_is_external[seq] = 1;
  return ALEGO_OK;
}

int alego_lm_set_odom(AlegoHandle *h, int seq, const double t[3], const double r[9]) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK || !t || !r) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  Pose p;
  std::memcpy(p.t, t, sizeof p.t);
  std::memcpy(p.R, r, sizeof p.R);
  CUDA_TRY(h, cudaMemcpyAsync(h->o2l + seq, &p, sizeof p, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}

int alego_lm_scan2map(AlegoHandle *h, AlegoSolveReport *reports) {
  if (!h) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  int rc = lm_ready(h);
  if (rc != ALEGO_OK) return rc;
  rc = lm_scan2map_device(h, h->lm_guard, true, false);
  if (rc != ALEGO_OK) return rc;
  if (reports) {
    if ((rc = d2h(h, reports, h->lm_report, h->B * sizeof(AlegoSolveReport))) != ALEGO_OK) return rc;
    for (int b = 0; b < h->B; ++b)
      if (reports[b].status != ALEGO_OK) return ALEGO_FEW_FEATURES;
  }
  return ALEGO_OK;
}

int alego_lm_get_state(AlegoHandle *h, int seq, double params[6], double t_m2l[3], double r_m2l[9], double t_m2o[3], double r_m2o[9]) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  Pose ml, mo;
  if ((rc = d2h(h, &ml, h->m2l + seq, sizeof ml)) != ALEGO_OK) return rc;
  if ((rc = d2h(h, &mo, h->m2o + seq, sizeof mo)) != ALEGO_OK) return rc;
  if (params && (rc = d2h(h, params, h->lm_params + seq * 6, 6 * sizeof(double))) != ALEGO_OK) return rc;
  if (t_m2l) std::memcpy(t_m2l, ml.t, sizeof ml.t);
  if (r_m2l) std::memcpy(r_m2l, ml.R, sizeof ml.R);
  if (t_m2o) std::memcpy(t_m2o, mo.t, sizeof mo.t);
  if (r_m2o) std::memcpy(r_m2o, mo.R, sizeof mo.R);
  return ALEGO_OK;
}

int alego_lm_set_params(AlegoHandle *h, int seq, const double params[6]) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK || !params) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  CUDA_TRY(h, cudaMemcpyAsync(h->lm_params + seq * 6, params, 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return ALEGO_OK;
}

int alego_lm_get_downsampled(AlegoHandle *h, int seq, float *corner_ds, int32_t *n_corner_ds, float *surf_ds, int32_t *n_surf_ds,
                             float *outlier_ds, int32_t *n_outlier_ds, float *surf_total_ds, int32_t *n_surf_total_ds) {
  int rc = check(h, seq);
  if (rc != ALEGO_OK) return rc;
  if (!h->lm_corner_ds) { h->err = "alego_lm_get_downsampled before alego_lm_scan2map"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  int n[8];
  if ((rc = d2h(h, n, h->lm_n + seq * 8, sizeof n)) != ALEGO_OK) return rc;
  if (n_corner_ds) *n_corner_ds = n[0];
  if (n_surf_ds) *n_surf_ds = n[1];
  if (n_outlier_ds) *n_outlier_ds = n[2];
  if (n_surf_total_ds) *n_surf_total_ds = n[4];
  if (corner_ds && (rc = d2h(h, corner_ds, h->lm_corner_ds + (size_t)seq * h->ds_cap_c, (size_t)n[0] * sizeof(float4))) != ALEGO_OK) return rc;
  if (surf_ds && (rc = d2h(h, surf_ds, h->lm_surf_ds + (size_t)seq * h->ds_cap_s, (size_t)n[1] * sizeof(float4))) != ALEGO_OK) return rc;
  if (outlier_ds && (rc = d2h(h, outlier_ds, h->lm_outlier_ds + (size_t)seq * h->ds_cap_o, (size_t)n[2] * sizeof(float4))) != ALEGO_OK) return rc;
  if (surf_total_ds &&
      (rc = d2h(h, surf_total_ds, h->lm_surf_total_ds + (size_t)seq * (h->ds_cap_s + h->ds_cap_o), (size_t)n[4] * sizeof(float4))) != ALEGO_OK)
    return rc;
  return ALEGO_OK;
}

// ---------------------------------------------------------------------------------------------------
int alego_pipeline_config(AlegoHandle *h, int lm_every, int rebuild_map_index_every_step, int use_cuda_graph) {
  if (!h || lm_every < 0) return ALEGO_BAD_ARG;
  h->lm_every = lm_every;
  h->rebuild_map_every_step = rebuild_map_index_every_step != 0;
  h->overlap_lm = use_cuda_graph >= 0;  // third argument: < 0 keeps LaserMapping on the main stream (no overlap with the next sweep's front end)
  h->use_graphs = use_cuda_graph > 0 && (use_cuda_graph & 2) != 0;  // bit 1: alego_pipeline_step replays a captured CUDA graph
  ++h->graph_epoch;
  return ALEGO_OK;
}

// IP -> LO -> LM on whatever h->raw / h->n_pts point at; everything is enqueued, nothing waits.
//
// Two streams, like the reference's separate nodes: ImageProjection + LaserOdometry of sweep t on the main stream,
// LaserMapping (local-map index build + scan-to-map) of sweep t on the side stream, so that mapping of sweep t overlaps
// the front end of sweep t+1.  What crosses from the front end to mapping is double-buffered by sweep parity (the
// less-sharp / less-flat clouds, the outlier cloud, LaserOdometry's pose); the front end of sweep t+2 waits for the
// mapping of sweep t before it overwrites that parity.
static int pipeline_enqueue(AlegoHandle *h, cudaEvent_t consumed_ev = nullptr) {
  int rc;
  const bool cap = h->capturing;  // recorded into a CUDA graph: the side stream forks from and joins the main stream inside
  if (!cap) {
    bool any_ext = false;
    for (auto v : h->lm_scan_is_external) any_ext |= v != 0;
    if (any_ext) {  // the pipeline feeds LaserMapping from LaserOdometry's device clouds
      if ((rc = join_side(h)) != ALEGO_OK) return rc;
      CUDA_TRY(h, cudaMemsetAsync(h->lm_use_ext, 0, h->B * sizeof(int), h->stream));
      std::fill(h->lm_scan_is_external.begin(), h->lm_scan_is_external.end(), 0);
    }
  }
  const bool run_lm = h->lm_every > 0 && h->map_corner && h->map_surf && (h->scan_count % h->lm_every == 0);
  const bool overlap = h->overlap_lm && !h->profiling;
  const int par = h->cur;  // buffer parity of this sweep's clouds
  if (!cap && run_lm && (rc = lm_ready(h)) != ALEGO_OK) return rc;
  // ---- front end (main stream)
  if (cap) {
    if (overlap) {
      CUDA_TRY(h, cudaEventRecord(h->ev_g_fork, h->stream));
      CUDA_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_g_fork, 0));
    }
  } else {
    if (overlap && h->lm_done_valid[par]) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_lm_done[par], 0));
    if (!overlap && (rc = join_side(h)) != ALEGO_OK) return rc;
  }
  if ((rc = ip_run_internal(h)) != ALEGO_OK) return rc;
  // ImageProjection is the only reader of the raw sweep: its staging buffer may be refilled from here on
  if (consumed_ev) CUDA_TRY(h, cudaEventRecord(consumed_ev, h->stream));
  if ((rc = lo_extract_device(h)) != ALEGO_OK) return rc;
  h->feat_buf = h->cur;
  if ((rc = lo_scan2scan_device(h)) != ALEGO_OK) return rc;
  h->stage_feat_done = false;
  // ---- mapping (side stream when overlapped)
  cudaStream_t ms = h->stream;
  cudaEvent_t ev_lo = cap ? h->ev_g_lo : h->ev_lo_done;
  if (overlap) {
    CUDA_TRY(h, cudaEventRecord(ev_lo, h->stream));
    ms = h->side_stream;
    h->launch_stream = ms;
  }
  // The local-map index does not depend on the sweep: the reference rebuilds its kd-trees every mapped frame
  // (laserMapping.cpp:356-357); on the side stream that happens before the wait for the front end.
  if (run_lm && (h->rebuild_map_every_step || !h->map_index_valid)) {
    rc = lm_build_map_index(h);
    if (rc != ALEGO_OK) { h->launch_stream = nullptr; return rc; }
  }
  if (overlap) CUDA_TRY(h, cudaStreamWaitEvent(ms, ev_lo, 0));
  if (run_lm) {
    rc = lm_scan2map_device(h, h->lm_guard, true, true);
  } else {
    LAUNCH(h, "pose_pack");
    pose_pack_kernel<<<div_up(h->B, 128), 128, 0, ms>>>(h->m2l, h->lm_params, h->o2l_lo[par], h->d_pose, h->B);
    rc = ALEGO_OK;
  }
  h->launch_stream = nullptr;
  if (rc != ALEGO_OK) return rc;
  if (overlap && cap) {  // join: the graph is complete when the main stream is
    CUDA_TRY(h, cudaEventRecord(h->ev_g_tail, ms));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_g_tail, 0));
  } else if (overlap) {
    CUDA_TRY(h, cudaEventRecord(h->ev_lm_done[par], ms));
    h->lm_done_valid[par] = true;
    CUDA_TRY(h, cudaEventRecord(h->ev_side_tail, ms));
    h->side_busy = true;
  }
  h->pose_on_main = cap || !overlap;
  ++h->scan_count;
  if (!cap) CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

static void drop_graphs(AlegoHandle *h) {
  for (auto &g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
  h->graphs_epoch = h->graph_epoch;
}

// Graph mode of the synchronous step (launch-latency regime: one or a few sequences, ~40 launches of a few microseconds of
// work each): the pass is captured once per (buffer parity, LM schedule, input buffer) and replayed with a single launch.
static int pipeline_enqueue_graph(AlegoHandle *h) {
  int rc;
  // what the eager path does outside the streams
  bool any_ext = false;
  for (auto v : h->lm_scan_is_external) any_ext |= v != 0;
  const bool run_lm = h->lm_every > 0 && h->map_corner && h->map_surf && (h->scan_count % h->lm_every == 0);
  if (any_ext || h->profiling || h->scan_count < 2) return pipeline_enqueue(h);  // first sweeps (lazy set-up) run eagerly
  if (run_lm && (rc = lm_ready(h)) != ALEGO_OK) return rc;
  if ((rc = join_side(h)) != ALEGO_OK) return rc;
  if (h->graphs_epoch != h->graph_epoch || h->graphs.size() > 64) drop_graphs(h);
  const int par = h->cur;
  const bool rebuild = run_lm && (h->rebuild_map_every_step || !h->map_index_valid);
  PipelineGraph *g = nullptr;
  for (auto &e : h->graphs)
    if (e.par == par && e.run_lm == run_lm && e.rebuild == rebuild && e.raw == h->raw && e.n_pts == h->n_pts && e.stride == h->in_stride) g = &e;
  if (g) {
    CUDA_TRY(h, cudaGraphLaunch(g->exec, h->stream));
    // host-side state the recorded calls would have advanced
    h->outlier = h->outlier_buf[par];
    h->n_outlier = h->n_outlier_buf[par];
    h->stage_ip_done = true;
    h->label_valid = false;
    h->feat_buf = par;
    h->cur = 1 - par;
    h->stage_feat_done = false;
    if (rebuild) h->map_index_valid = true;
    h->launches += g->n_launches;
    ++h->scan_count;
  } else {
    PipelineGraph e;
    e.par = par; e.run_lm = run_lm; e.rebuild = rebuild; e.raw = h->raw; e.n_pts = h->n_pts; e.stride = h->in_stride;
    const int64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    h->capturing = true;
    rc = pipeline_enqueue(h);
    h->capturing = false;
    h->launch_stream = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc != ALEGO_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      if (rc == ALEGO_OK) { h->err = std::string("graph capture: ") + cudaGetErrorString(ce); rc = ALEGO_CUDA_ERROR; }
      return rc;
    }
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie); return ALEGO_CUDA_ERROR; }
    e.n_launches = h->launches - l0;
    CUDA_TRY(h, cudaGraphLaunch(e.exec, h->stream));
    h->graphs.push_back(e);
  }
  h->lm_done_valid[0] = h->lm_done_valid[1] = false;  // every pass is joined into the main stream
  h->side_busy = false;
  h->pose_on_main = true;
  return ALEGO_OK;
}

// stream that produced d_pose of the pass enqueued last
static cudaStream_t pose_stream(AlegoHandle *h) { return h->pose_on_main ? h->stream : h->side_stream; }

int alego_pipeline_step(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points, double *poses_out) {
  if (!h) return ALEGO_BAD_ARG;
  if (h->n_submitted != h->n_collected) { h->err = "alego_pipeline_step while submitted steps are in flight"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  int rc;
  if (xyzi_host) {
    if ((rc = alego_ip_upload(h, xyzi_host, n_points)) != ALEGO_OK) return rc;
  }
  if ((rc = h->use_graphs ? pipeline_enqueue_graph(h) : pipeline_enqueue(h)) != ALEGO_OK) return rc;
  if (poses_out) {
    cudaStream_t ps = pose_stream(h);
    CUDA_TRY(h, cudaMemcpyAsync(h->h_pose, h->d_pose, (size_t)h->B * 12 * sizeof(double), cudaMemcpyDeviceToHost, ps));
    CUDA_TRY(h, cudaStreamSynchronize(ps));
    std::memcpy(poses_out, h->h_pose, (size_t)h->B * 12 * sizeof(double));
  }
  return ALEGO_OK;
}

// Asynchronous form: the H2D copy of sweep t+1 (copy stream, second device staging buffer) overlaps the pass over
// sweep t.  At most ALEGO_INFLIGHT (3) steps in flight; alego_pipeline_collect returns them in submission order.
int alego_pipeline_submit(AlegoHandle *h, const float *xyzi_host, const int32_t *n_points) {
  if (!h || !xyzi_host || !n_points) return ALEGO_BAD_ARG;
  if (h->n_submitted - h->n_collected >= ALEGO_INFLIGHT) { h->err = "alego_pipeline_submit: three steps already in flight (collect first)"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  const int slot = (int)(h->n_submitted % ALEGO_INFLIGHT);
  if (!h->raw_slot[slot]) {
    CUDA_TRY(h, cudaMalloc(&h->raw_slot[slot], (size_t)h->B * h->Nmax * sizeof(float4)));
    CUDA_TRY(h, cudaMalloc(&h->n_pts_slot[slot], (size_t)h->B * sizeof(int)));
  }
  for (int k = 0; k < ALEGO_INFLIGHT; ++k) {
    if (!h->h_n_pts_slot[k]) CUDA_TRY(h, cudaMallocHost((void **)&h->h_n_pts_slot[k], (size_t)h->B * sizeof(int32_t)));
    if (!h->h_pose_slot[k]) CUDA_TRY(h, cudaMallocHost((void **)&h->h_pose_slot[k], (size_t)h->B * 12 * sizeof(double)));
  }
  std::memcpy(h->h_n_pts_slot[slot], n_points, (size_t)h->B * sizeof(int32_t));
  // the staging buffer may still be read by the pass submitted two steps ago
  if (h->consumed_valid[slot]) CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0));
  const bool tl = h->timeline_on;
  if (tl && !h->ev_t_origin) {
    CUDA_TRY(h, cudaEventCreate(&h->ev_t_origin));
    for (int k = 0; k < ALEGO_INFLIGHT; ++k)
      for (int e = 0; e < 5; ++e) CUDA_TRY(h, cudaEventCreate(&h->ev_t[k][e]));
    CUDA_TRY(h, cudaEventRecord(h->ev_t_origin, h->copy_stream));
  }
  if (tl) CUDA_TRY(h, cudaEventRecord(h->ev_t[slot][0], h->copy_stream));
  int rc = upload_into(h, h->raw_slot[slot], h->n_pts_slot[slot], xyzi_host, h->h_n_pts_slot[slot], h->copy_stream);
  if (rc != ALEGO_OK) return rc;
  if (tl) CUDA_TRY(h, cudaEventRecord(h->ev_t[slot][1], h->copy_stream));
  CUDA_TRY(h, cudaEventRecord(h->ev_copied[slot], h->copy_stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
  if (tl) CUDA_TRY(h, cudaEventRecord(h->ev_t[slot][2], h->stream));
  h->raw = h->raw_slot[slot];
  h->n_pts = h->n_pts_slot[slot];
  if ((rc = pipeline_enqueue(h, h->ev_consumed[slot])) != ALEGO_OK) return rc;
  h->consumed_valid[slot] = true;
  if (tl) CUDA_TRY(h, cudaEventRecord(h->ev_t[slot][3], h->stream));  // front end (IP + LO) of this step done
  cudaStream_t ps = pose_stream(h);
  CUDA_TRY(h, cudaMemcpyAsync(h->h_pose_slot[slot], h->d_pose, (size_t)h->B * 12 * sizeof(double), cudaMemcpyDeviceToHost, ps));
  if (tl) CUDA_TRY(h, cudaEventRecord(h->ev_t[slot][4], ps));
  CUDA_TRY(h, cudaEventRecord(h->ev_pose[slot], ps));
  if (ps == h->side_stream) CUDA_TRY(h, cudaEventRecord(h->ev_side_tail, ps));
  ++h->n_submitted;
  return ALEGO_OK;
}

int alego_pipeline_collect(AlegoHandle *h, double *poses_out) {
  if (!h) return ALEGO_BAD_ARG;
  if (h->n_submitted == h->n_collected) { h->err = "alego_pipeline_collect: nothing in flight"; return ALEGO_NOT_READY; }
  CUDA_TRY(h, cudaSetDevice(h->dev));
  const int slot = (int)(h->n_collected % ALEGO_INFLIGHT);
  CUDA_TRY(h, cudaEventSynchronize(h->ev_pose[slot]));
  if (poses_out) std::memcpy(poses_out, h->h_pose_slot[slot], (size_t)h->B * 12 * sizeof(double));
  if (h->timeline_on && h->ev_t_origin)
    for (int e = 0; e < 5; ++e) CUDA_TRY(h, cudaEventElapsedTime(&h->last_timeline[e], h->ev_t_origin, h->ev_t[slot][e]));
  ++h->n_collected;
  return ALEGO_OK;
}

int alego_pipeline_timeline(AlegoHandle *h, int on, float *t_ms) {
  if (!h) return ALEGO_BAD_ARG;
  if (h->n_submitted != h->n_collected && (on != 0) != h->timeline_on) {
    h->err = "alego_pipeline_timeline: switch while steps are in flight";
    return ALEGO_NOT_READY;
  }
  h->timeline_on = on != 0;
  if (t_ms) std::memcpy(t_ms, h->last_timeline, sizeof h->last_timeline);
  return ALEGO_OK;
}

// ---------------------------------------------------------------------------------------------------
// Stand-alone VoxelGrid on host data: the same block-wide routine the LM stage uses, on one cloud.
int alego_voxel_grid(AlegoHandle *h, const float *xyzi, int32_t n, float leaf, float *out_xyzi, int32_t *n_out) {
  if (!h || n < 0 || (n && !xyzi) || !n_out || !(leaf > 0)) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaSetDevice(h->dev));
  return voxel_grid_host(h, xyzi, n, leaf, out_xyzi, n_out);
}

int alego_timer_mark(AlegoHandle *h, int slot) {
  if (!h || slot < 0 || slot >= 16) return ALEGO_BAD_ARG;
  if (join_side(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;  // the mark covers both streams
  CUDA_TRY(h, cudaEventRecord(h->timer[slot], h->stream));
  return ALEGO_OK;
}
int alego_timer_elapsed_ms(AlegoHandle *h, int a, int b, float *ms) {
  if (!h || !ms || a < 0 || a >= 16 || b < 0 || b >= 16) return ALEGO_BAD_ARG;
  CUDA_TRY(h, cudaEventSynchronize(h->timer[b]));
  CUDA_TRY(h, cudaEventElapsedTime(ms, h->timer[a], h->timer[b]));
  return ALEGO_OK;
}
int alego_profile_enable(AlegoHandle *h, int on) {
  if (!h) return ALEGO_BAD_ARG;
  if (!on) resolve_profile(h);
  h->profiling = on != 0;
  return ALEGO_OK;
}
int alego_profile_reset(AlegoHandle *h) {
  if (!h) return ALEGO_BAD_ARG;
  resolve_profile(h);
  for (auto &k : h->prof) { k.launches = 0; k.total_ms = 0; }
  return ALEGO_OK;
}
int alego_profile_count(const AlegoHandle *h) { return h ? (int)h->prof.size() : ALEGO_BAD_ARG; }
int alego_profile_get(AlegoHandle *h, int i, char *name_out, int64_t *launches, double *total_ms) {
  if (!h || i < 0 || i >= (int)h->prof.size()) return ALEGO_BAD_ARG;
  resolve_profile(h);
  if (name_out) { std::strncpy(name_out, h->prof[i].name.c_str(), 63); name_out[63] = 0; }
  if (launches) *launches = h->prof[i].launches;
  if (total_ms) *total_ms = h->prof[i].total_ms;
  return ALEGO_OK;
}
int64_t alego_launch_count(const AlegoHandle *h) { return h ? h->launches : ALEGO_BAD_ARG; }

// ---------------------------------------------------------------------------------------------------
int64_t alego_debug_get(AlegoHandle *h, const char *name, int seq, void *dst, size_t cap) {
  if (check(h, seq) != ALEGO_OK || !name) return ALEGO_BAD_ARG;
  if (cudaSetDevice(h->dev) != cudaSuccess) return ALEGO_CUDA_ERROR;
  const std::string s(name);
  const size_t RC = h->RC, R = h->R;
  const size_t base = (size_t)seq * RC;
  int M = 0, nf[4] = {0, 0, 0, 0}, lmn[8] = {0}, no = 0;
  if (fetch_int(h, h->M + seq, &M) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  if (d2h(h, nf, h->n_feat + seq * 4, sizeof nf) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  if (d2h(h, lmn, h->lm_n + seq * 8, sizeof lmn) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  if (fetch_int(h, h->n_outlier + seq, &no) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  const void *src = nullptr;
  size_t bytes = 0;
  const int fb = h->feat_buf < 0 ? 0 : h->feat_buf;
  Pose pose;
  bool from_pose = false;
  double pose_buf[9];
  if (s == "range_mat") { src = h->range + base; bytes = RC * 4; }
  else if (s == "full_cloud") { src = h->cloud + base; bytes = RC * 16; }
  else if (s == "ground_mat") { src = h->ground + base; bytes = RC; }
  else if (s == "label_mat") {
    if (!h->stage_ip_done) return ALEGO_NOT_READY;
    if (ip_label_device(h) != ALEGO_OK) return ALEGO_CUDA_ERROR;
    src = h->label + base; bytes = RC * 4;
  }
  else if (s == "startRingIndex") { src = h->start_ring + seq * R; bytes = R * 4; }
  else if (s == "endRingIndex") { src = h->end_ring + seq * R; bytes = R * 4; }
  else if (s == "segmentedCloudGroundFlag") { src = h->seg_ground + base; bytes = M; }
  else if (s == "segmentedCloudColInd") { src = h->seg_col + base; bytes = (size_t)M * 4; }
  else if (s == "segmentedCloudRange") { src = h->seg_range + base; bytes = (size_t)M * 4; }
  else if (s == "segmented_cloud") { src = h->seg_cloud + base; bytes = (size_t)M * 16; }
  else if (s == "outlier_cloud") { src = h->outlier + (size_t)seq * h->out_cap; bytes = (size_t)no * 16; }
  else if (s == "orientation") { src = h->orient + seq * 4; bytes = 12; }
  else if (s == "cloud_curvature_abs") { src = h->curv + base; bytes = (size_t)M * 4; }
  else if (s == "cloud_neighbor_picked") { src = h->picked + base; bytes = M; }
  else if (s == "cloud_neighbor_picked_occl") { src = h->picked0 + base; bytes = M; }
  else if (s == "cloud_label") { src = h->flabel + base; bytes = (size_t)M * 4; }
  else if (s == "cloud_sort_idx") { src = h->sort_idx + base; bytes = (size_t)M * 4; }
  else if (s == "sharp_idx") { src = h->sharp_idx + (size_t)seq * R * 12; bytes = (size_t)nf[0] * 4; }
  else if (s == "less_sharp_idx") { src = h->less_sharp_idx + (size_t)seq * R * 120; bytes = (size_t)nf[1] * 4; }
  else if (s == "flat_idx") { src = h->flat_idx + (size_t)seq * R * 24; bytes = (size_t)nf[2] * 4; }
  else if (s == "sharp") { src = h->sharp + (size_t)seq * R * 12; bytes = (size_t)nf[0] * 16; }
  else if (s == "flat") { src = h->flat + (size_t)seq * R * 24; bytes = (size_t)nf[2] * 16; }
  else if (s == "less_sharp" || s == "corner_last") { src = h->less_sharp[fb] + (size_t)seq * R * 120; bytes = (size_t)nf[1] * 16; }
  else if (s == "less_flat" || s == "surf_last") { src = h->less_flat[fb] + base; bytes = (size_t)nf[3] * 16; }
  else if (s == "lo_params") { src = h->lo_params + seq * 6; bytes = 48; }
  else if (s == "t_w_cur") { src = h->t_w + seq * 3; bytes = 24; }
  else if (s == "r_w_cur") { src = h->r_w + seq * 9; bytes = 72; }
  else if (s == "lo_surf_corr") { src = h->lo_surf_corr + (size_t)seq * R * 24 * 4; bytes = (size_t)nf[2] * 16; }
  else if (s == "lo_corner_corr") { src = h->lo_corner_corr + (size_t)seq * R * 12 * 3; bytes = (size_t)nf[0] * 12; }
  else if (s == "lo_trace") {
    int n = 0;
    if (fetch_int(h, h->lo_trace_n + seq, &n) != ALEGO_OK) return ALEGO_CUDA_ERROR;
    src = h->lo_trace + (size_t)seq * h->lo_trace_cap * 7; bytes = (size_t)n * 56;
  } else if (s == "lm_trace") {
    int n = 0;
    if (fetch_int(h, h->lm_trace_n + seq, &n) != ALEGO_OK) return ALEGO_CUDA_ERROR;
    src = h->lm_trace + (size_t)seq * h->lm_trace_cap * 7; bytes = (size_t)n * 56;
  }
  else if (s == "lm_params") { src = h->lm_params + seq * 6; bytes = 48; }
  else if (s == "lm_report") { src = h->lm_report + seq; bytes = sizeof(AlegoSolveReport); }  // of the last mapped sweep
  else if (s == "lo_report") { src = h->lo_report + seq; bytes = sizeof(AlegoSolveReport); }
  else if (s == "lm_corner_ds") { src = h->lm_corner_ds + (size_t)seq * h->ds_cap_c; bytes = (size_t)lmn[0] * 16; }
  else if (s == "lm_surf_ds") { src = h->lm_surf_ds + (size_t)seq * h->ds_cap_s; bytes = (size_t)lmn[1] * 16; }
  else if (s == "lm_outlier_ds") { src = h->lm_outlier_ds + (size_t)seq * h->ds_cap_o; bytes = (size_t)lmn[2] * 16; }
  else if (s == "lm_surf_total_ds") { src = h->lm_surf_total_ds + (size_t)seq * (h->ds_cap_s + h->ds_cap_o); bytes = (size_t)lmn[4] * 16; }
  else if (s == "lm_edge") { src = h->lm_edge + (size_t)seq * h->ds_cap_c * 10; bytes = (size_t)lmn[0] * 80; }
  else if (s == "lm_plane") { src = h->lm_plane + (size_t)seq * (h->ds_cap_s + h->ds_cap_o) * 8; bytes = (size_t)lmn[4] * 64; }
  else if (s == "t_map2laser" || s == "r_map2laser" || s == "t_map2odom" || s == "r_map2odom") {
    const Pose *p = (s.find("map2laser") != std::string::npos) ? h->m2l + seq : h->m2o + seq;
    if (d2h(h, &pose, p, sizeof pose) != ALEGO_OK) return ALEGO_CUDA_ERROR;
    from_pose = true;
    if (s[0] == 't') { std::memcpy(pose_buf, pose.t, 24); bytes = 24; }
    else { std::memcpy(pose_buf, pose.R, 72); bytes = 72; }
  } else if (s == "map_index_kind") {  // 1: the surf map is searched through the voxel-row index, 0: through the hashed grid
    const int kind = h->map_rows_checked && h->rows_map_surf.usable ? 1 : 0;
    std::memcpy(pose_buf, &kind, sizeof kind);
    from_pose = true;
    bytes = sizeof kind;
  } else {
    h->err = "unknown debug array: " + s;
    return ALEGO_BAD_ARG;
  }
  if (!dst) return (int64_t)bytes;
  if (bytes > cap) { h->err = "debug_get: capacity too small"; return ALEGO_BAD_ARG; }
  if (from_pose) { std::memcpy(dst, pose_buf, bytes); return (int64_t)bytes; }
  if (bytes && !src) return ALEGO_NOT_READY;
  if (d2h(h, dst, src, bytes) != ALEGO_OK) return ALEGO_CUDA_ERROR;
  return (int64_t)bytes;
}

}  // extern "C"
