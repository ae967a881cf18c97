// lo_distortion_kernels.cu — SURVEY §8f row N2: LaserOdometry::adjustDistortion (src/laserOdometry.cpp:557-726, IMU branch
// :581-657; the call site :115 is commented out in the reference, so this stage is optional and off by default).
//
// The reference walks the segmented cloud once, point by point, with a forward-only pointer into the IMU ring buffer
// (imu_ptr_front_ starts at imu_ptr_last_iter_ and only advances while cur_time >= imu_time_[front], :587-595).  With
// non-decreasing IMU stamps over the live part of the queue (checked on the host before the launch) the pointer after
// point i is F(max_{j<=i} cur_time_j), F(t) = first live entry with t < imu_time_ — and cur_time is a monotone function of
// the point's column index, so the pointer is a PREFIX MAXIMUM of the column indices followed by a binary search.  The
// `return` of the unsynchronised case (:596-600) leaves the points before it adjusted and the rest untouched: the first
// such point is a min-reduction.  That walk is one CTA per sequence (lo_adjust_walk, integer work); everything else is per
// point and runs over the whole GPU (lo_adjust_apply).
//
// Arithmetic follows the reference: times, ratios and the interpolated roll / pitch / yaw / shift / velocity in double,
// stored into Eigen float vectors; rotation = (AngleAxisf(yaw,Z) * AngleAxisf(pitch,Y) * AngleAxisf(roll,X)) through float
// quaternions; r_s_i by Eigen's 3x3 cofactor inverse; products in float, left to right.  sinf / cosf are evaluated in
// double and rounded (equal to glibc's float routines except for rare 1-ulp cases; the parity test states the tolerance).
#include "common.cuh"
#include "lo_kernels.cuh"

namespace {

#define DIST_THREADS 1024
#define DIST_MAX_IMU 2048  // ring-buffer entries held in shared memory (reference: imu_queue_length = 200, utility.h:70)

struct ImuSample {
  float rpy[3], shift[3], velo[3];
};

// entry layout of one queue: [10][len] doubles — time, roll, pitch, yaw, shift x y z, velocity x y z
__device__ __forceinline__ void imu_sample(const double *__restrict__ q, const double *s_time, int len, int front, double cur_time,
                                           ImuSample &o) {
  if (cur_time > s_time[front]) {  // :602-613
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o.rpy[k] = (float)q[(1 + k) * len + front];
      o.shift[k] = (float)q[(4 + k) * len + front];
      o.velo[k] = (float)q[(7 + k) * len + front];
    }
  } else {  // :614-629
    const int back = (front - 1 + len) % len;
    const double ratio_front = (cur_time - s_time[back]) / (s_time[front] - s_time[back]);
    const double ratio_back = 1. - ratio_front;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o.rpy[k] = (float)(q[(1 + k) * len + front] * ratio_front + q[(1 + k) * len + back] * ratio_back);
      o.shift[k] = (float)(q[(4 + k) * len + front] * ratio_front + q[(4 + k) * len + back] * ratio_back);
      o.velo[k] = (float)(q[(7 + k) * len + front] * ratio_front + q[(7 + k) * len + back] * ratio_back);
    }
  }
}

__device__ __forceinline__ void quat_mul(const float a[4], const float b[4], float o[4]) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

// (AngleAxisf(yaw, Z) * AngleAxisf(pitch, Y) * AngleAxisf(roll, X)).toRotationMatrix() (:631), row-major
__device__ __forceinline__ void rpy_matrix(const float rpy[3], float M[9]) {
  float qx[4] = {0.f, 0.f, 0.f, 0.f}, qy[4] = {0.f, 0.f, 0.f, 0.f}, qz[4] = {0.f, 0.f, 0.f, 0.f}, qzy[4], q[4];
  const float hr = 0.5f * rpy[0], hp = 0.5f * rpy[1], hy = 0.5f * rpy[2];
  qx[0] = (float)cos((double)hr); qx[1] = (float)sin((double)hr);
  qy[0] = (float)cos((double)hp); qy[2] = (float)sin((double)hp);
  qz[0] = (float)cos((double)hy); qz[3] = (float)sin((double)hy);
  quat_mul(qz, qy, qzy);
  quat_mul(qzy, qx, q);
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M[0] = 1.f - (tyy + tzz); M[1] = txy - twz;         M[2] = txz + twy;
  M[3] = txy + twz;         M[4] = 1.f - (txx + tzz); M[5] = tyz - twx;
  M[6] = txz - twy;         M[7] = tyz + twx;         M[8] = 1.f - (txx + tyy);
}

// Eigen's fixed-size 3x3 inverse (cofactors of column 0 -> determinant -> scaled cofactor matrix)
__device__ __forceinline__ float cof(const float *m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
__device__ __forceinline__ void inverse3(const float *m, float *r) {
  const float c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
  const float det = c0 * m[0] + (c1 * m[3] + c2 * m[6]);
  const float invdet = 1.f / det;
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  r[3] = cof(m, 0, 1) * invdet; r[4] = cof(m, 1, 1) * invdet; r[5] = cof(m, 2, 1) * invdet;
  r[6] = cof(m, 0, 2) * invdet; r[7] = cof(m, 1, 2) * invdet; r[8] = cof(m, 2, 2) * invdet;
}

__device__ __forceinline__ double point_rel_time(int col, int start_ori, int ori_diff, double scan_period) {
  return (col - start_ori) * scan_period / ori_diff;  // :579
}

__global__ void __launch_bounds__(DIST_THREADS)
lo_adjust_walk_kernel(const int *__restrict__ seg_col, const int *__restrict__ Mv, const float *__restrict__ orient, int RC, int C,
                      double scan_period, const double *__restrict__ queues, int len, const int *__restrict__ ptr_last,
                      int *__restrict__ ptr_last_iter, const double *__restrict__ scan_time, int *__restrict__ n_done,
                      int *__restrict__ front_of, float *__restrict__ start_pose) {
  const int b = blockIdx.x;
  const int M = Mv[b];
  const double *q = queues + (size_t)b * 10 * len;
  const int *col = seg_col + (size_t)b * RC;
  int *front = front_of + (size_t)b * RC;
  __shared__ double s_time[DIST_MAX_IMU];
  __shared__ int s_warp[32];
  __shared__ int s_carry, s_stop;

  const int last = ptr_last[b], iter0 = ptr_last_iter[b];
  if (last <= 0 || M <= 0) {  // :583: nothing is touched before the second IMU message
    if (threadIdx.x == 0) n_done[b] = 0;
    return;
  }
  for (int k = threadIdx.x; k < len; k += blockDim.x) s_time[k] = q[k];
  if (threadIdx.x == 0) { s_carry = -2147483647 - 1; s_stop = M; }
  __syncthreads();

  // seg_info->startOrientation + 2*M_PI is a double sum, divided by the int Horizon_SCAN, truncated (:562-563)
  int start_ori = (int)(((double)orient[b * 4 + 0] + 2 * 3.14159265358979323846) / C);
  int end_ori = (int)(((double)orient[b * 4 + 1] + 2 * 3.14159265358979323846) / C);
  if (start_ori >= C) start_ori -= C;
  if (end_ori >= C) end_ori -= C;
  int ori_diff = end_ori - start_ori;
  if (ori_diff <= 0) ori_diff = C;  // :573-577
  const double t0 = scan_time[b];
  const int live = (last - iter0 + len) % len;  // ring positions iter0 .. last

  // ---- pass 1: pointer position after every point + first unsynchronised point --------------------------------------
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < M; base += DIST_THREADS) {
    const int i = base + threadIdx.x;
    int v = i < M ? col[i] : (-2147483647 - 1);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v = max(v, t);
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    int pre = s_carry;
    for (int w = 0; w < wid; ++w) pre = max(pre, s_warp[w]);
    v = max(v, pre);
    if (i < M) {
      const double t_max = t0 + point_rel_time(v, start_ori, ori_diff, scan_period);
      // first live ring position k with t_max < time (stamps non-decreasing over the live range), else `last`
      int lo = 0, hi = live;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t_max < s_time[(iter0 + mid) % len]) hi = mid; else lo = mid + 1;
      }
      const int f = (iter0 + lo) % len;
      front[i] = f;
      const double cur_time = t0 + point_rel_time(col[i], start_ori, ori_diff, scan_period);
      if (fabs(cur_time - s_time[f]) > scan_period) atomicMin(&s_stop, i);  // :596-600
    }
    __syncthreads();
    if (threadIdx.x == DIST_THREADS - 1) s_carry = v;
    __syncthreads();
    if (s_stop < base + DIST_THREADS) break;
  }
  const int S = s_stop;
  if (S <= 0) {
    if (threadIdx.x == 0) n_done[b] = 0;
    return;
  }

  // ---- point 0 fixes the start pose (:633-639) ----------------------------------------------------------------------
  if (threadIdx.x == 0) {
    ImuSample s0;
    imu_sample(q, s_time, len, front[0], t0 + point_rel_time(col[0], start_ori, ori_diff, scan_period), s0);
    float rc[9], inv[9];
    rpy_matrix(s0.rpy, rc);
    inverse3(rc, inv);
    float *o = start_pose + (size_t)b * 16;
    for (int k = 0; k < 9; ++k) o[k] = inv[k];
    for (int k = 0; k < 3; ++k) { o[9 + k] = s0.shift[k]; o[12 + k] = s0.velo[k]; }
    n_done[b] = S;
    ptr_last_iter[b] = front[S - 1];  // :656
  }
}

// ---- every other point before the stop (:640-655), the whole GPU over all sequences ----------------------------------------
__global__ void __launch_bounds__(256)
lo_adjust_apply_kernel(float4 *__restrict__ seg_cloud, const int *__restrict__ seg_col, const float *__restrict__ orient, int RC, int C,
                       double scan_period, const double *__restrict__ queues, int len, const double *__restrict__ scan_time,
                       const int *__restrict__ n_done, const int *__restrict__ front_of, const float *__restrict__ start_pose) {
  const int b = blockIdx.y;
  const int S = n_done[b];
  if (S <= 1) return;
  const double *q = queues + (size_t)b * 10 * len;
  float4 *cloud = seg_cloud + (size_t)b * RC;
  const int *col = seg_col + (size_t)b * RC;
  const int *front = front_of + (size_t)b * RC;
  __shared__ float s_rsi[9], s_shift0[3], s_velo0[3];
  if (threadIdx.x < 9) s_rsi[threadIdx.x] = start_pose[(size_t)b * 16 + threadIdx.x];
  else if (threadIdx.x < 12) s_shift0[threadIdx.x - 9] = start_pose[(size_t)b * 16 + threadIdx.x];
  else if (threadIdx.x < 15) s_velo0[threadIdx.x - 12] = start_pose[(size_t)b * 16 + threadIdx.x];
  __syncthreads();
  int start_ori = (int)(((double)orient[b * 4 + 0] + 2 * 3.14159265358979323846) / C);
  int end_ori = (int)(((double)orient[b * 4 + 1] + 2 * 3.14159265358979323846) / C);
  if (start_ori >= C) start_ori -= C;
  if (end_ori >= C) end_ori -= C;
  int ori_diff = end_ori - start_ori;
  if (ori_diff <= 0) ori_diff = C;
  const double t0 = scan_time[b];
  for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    const double rel_time = point_rel_time(col[i], start_ori, ori_diff, scan_period);
    ImuSample s;
    imu_sample(q, q, len, front[i], t0 + rel_time, s);
    float rc[9];
    rpy_matrix(s.rpy, rc);
    const float relf = (float)rel_time;  // Eigen promotes the double scalar to the vector's float
    float sh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sh[k] = (s.shift[k] - s_shift0[k]) - s_velo0[k] * relf;
    float4 p = cloud[i];
    const float ax = ((rc[0] * p.x + rc[1] * p.y) + rc[2] * p.z) + sh[0];
    const float ay = ((rc[3] * p.x + rc[4] * p.y) + rc[5] * p.z) + sh[1];
    const float az = ((rc[6] * p.x + rc[7] * p.y) + rc[8] * p.z) + sh[2];
    p.x = (s_rsi[0] * ax + s_rsi[1] * ay) + s_rsi[2] * az;
    p.y = (s_rsi[3] * ax + s_rsi[4] * ay) + s_rsi[5] * az;
    p.z = (s_rsi[6] * ax + s_rsi[7] * ay) + s_rsi[8] * az;
    cloud[i] = p;
  }
}

}  // namespace

int lo_adjust_distortion_device(AlegoHandle *h, const double *queues_dev, int len, const int *ptr_last_dev, int *ptr_last_iter_dev,
                                const double *scan_time_dev, double scan_period, int *n_done_dev, float *start_pose_dev) {
  if (len < 1 || len > DIST_MAX_IMU) { h->err = "alego_lo_adjust_distortion: queue length must be 1..2048"; return ALEGO_BAD_ARG; }
  // sort_idx is free between ImageProjection and lo_curv_occl (which rewrites it): pointer position of every point
  { LAUNCH(h, "lo_adjust_walk");
    lo_adjust_walk_kernel<<<h->B, DIST_THREADS, 0, h->stream>>>(h->seg_col, h->M, h->orient, h->RC, h->C, scan_period, queues_dev, len,
                                                                ptr_last_dev, ptr_last_iter_dev, scan_time_dev, n_done_dev, h->sort_idx,
                                                                start_pose_dev); }
  { LAUNCH(h, "lo_adjust_apply");
    lo_adjust_apply_kernel<<<dim3(min(div_up(h->RC, 256), 96), h->B), 256, 0, h->stream>>>(h->seg_cloud, h->seg_col, h->orient, h->RC, h->C,
                                                                                        scan_period, queues_dev, len, scan_time_dev,
                                                                                        n_done_dev, h->sort_idx, start_pose_dev); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
