// lc_icp_kernels.cu — SURVEY §8f row N4: the ICP of LaserMapping::performLoopClosure (src/laserMapping.cpp:652-711):
// pcl::IterativeClosestPoint<PointXYZI, PointXYZI> (point-to-point, SVD / Umeyama estimation, DefaultConvergenceCriteria),
// latest keyframe -> 1 m-voxelised history cloud, <= 100 iterations.  iSAM2 and the pose graph stay on the host (out of scope).
//
// One launch per ICP iteration, no host round trip inside the loop:
//   every thread owns one source point: applies the previous iteration's incremental transform in place (PCL's transformCloud),
//   finds its exact nearest target point in the hashed cell grid of grid.cu (27 cells, then the 5x5x5 shell; what is still
//   unresolved is ranked against the whole target cloud by the CTA's warps), and accumulates the 17 sums Umeyama needs
//   (n, sum d^2, sum s, sum t, sum t s^T) in double -> warp shuffles -> per-CTA partials;
//   the last CTA to finish (ticket) adds the partials in CTA order (deterministic), builds the float quantities Eigen::umeyama
//   holds (means, sigma), takes the 3x3 SVD, composes final_transformation_ and evaluates PCL's convergence criteria; a device
//   flag turns the remaining launches into no-ops.
// A last pass over the ORIGINAL source cloud gives getFitnessScore().
#include "common.cuh"
#include "grid.cuh"
#include "lm_kernels.cuh"

namespace {

#define ICP_THREADS 256
#define ICP_WARPS (ICP_THREADS / 32)
#define ICP_NSUM 17
#define ICP_RINGS 2  // cell rings searched before a query goes to the exhaustive pass

struct IcpState {
  float T[16];       // transformation_ of the last finished iteration (row-major)
  float Tfinal[16];  // final_transformation_
  double prev_mse;   // correspondences_prev_mse_
  double fitness;    // getFitnessScore()
  int iterations;    // nr_iterations_
  int state;         // 0 not converged, 1 iterations, 2 transform, 3 abs mse, 4 rel mse, 5 no correspondences
  int n_corr;        // correspondences of the last iteration
  unsigned ticket;
};

struct IcpParams {
  double max_dist_sqr, transformation_epsilon, fitness_epsilon;
  int max_iterations;
};

__device__ __forceinline__ void nn_offer(float d, int idx, float &bd, int &bi) {
  if (d < bd || (d == bd && idx < bi)) { bd = d; bi = idx; }
}

// exact nearest neighbour if it lies within ICP_RINGS cell rings; false = not settled (bd / bi hold the best seen so far)
__device__ __forceinline__ bool nn_grid(const GridIndex &g, float qx, float qy, float qz, float &bd, int &bi) {
  const int T = g.table_size;
  const int *cs = g.cell_start;
  const float inv = 1.0f / g.cell;
  const int cx = grid_coord(qx, inv), cy = grid_coord(qy, inv), cz = grid_coord(qz, inv);
#pragma unroll 1
  for (int k = 1; k <= ICP_RINGS; ++k) {
#pragma unroll 1
    for (int dz = -k; dz <= k; ++dz)
#pragma unroll 1
      for (int dy = -k; dy <= k; ++dy) {
        const bool inner_yz = k > 1 && abs(dz) < k && abs(dy) < k;  // ring 1 includes the query's own cell
#pragma unroll 1
        for (int dx = -k; dx <= k; ++dx) {
          if (inner_yz && abs(dx) < k) continue;  // visited with the previous ring
          const int hb = grid_hash(cx + dx, cy + dy, cz + dz, T);
          const int e = cs[hb + 1];
          for (int t = cs[hb]; t < e; ++t) {
            const float4 p = g.sorted[t];
            nn_offer(l2_simple(qx, qy, qz, p), __float_as_int(p.w), bd, bi);
          }
        }
      }
    // every point outside the (2k+1)^3 block is farther than k cells from the query (0.1 % margin for the float cell index)
    const float lim = (float)k * g.cell * 0.999f;
    if (bd <= lim * lim) return true;
  }
  return false;
}

__device__ void jacobi_eig3(const double A[9], double w[3], double V[9]) {  // symmetric; ascending eigenvalues, columns of V
  double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double dsum = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-40 * dsum || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
      }
  }
  int o0 = 0, o1 = 1, o2 = 2;
  if (a[o1][o1] < a[o0][o0]) { const int t = o0; o0 = o1; o1 = t; }
  if (a[o2][o2] < a[o1][o1]) { const int t = o1; o1 = o2; o2 = t; }
  if (a[o1][o1] < a[o0][o0]) { const int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
  for (int k = 0; k < 3; ++k) {
    w[k] = a[ord[k]][ord[k]];
    for (int r = 0; r < 3; ++r) V[r * 3 + k] = v[r][ord[k]];
  }
}

// R = U diag(1, 1, det(U) det(V)) V^T of sigma = U S V^T (Eigen::umeyama without scaling), SVD through the eigenvectors of
// sigma^T sigma in double; u2 = u0 x u1 makes det(U) = +1, which also covers a rank-2 sigma (planar correspondences)
__device__ void svd_rotation(const float sigma[9], float R[9]) {
  double A[9], AtA[9], w[3], V[9];
  for (int k = 0; k < 9; ++k) A[k] = sigma[k];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) AtA[i * 3 + j] = A[0 * 3 + i] * A[0 * 3 + j] + A[1 * 3 + i] * A[1 * 3 + j] + A[2 * 3 + i] * A[2 * 3 + j];
  jacobi_eig3(AtA, w, V);
  double v[3][3], u[3][3];  // [k] = k-th singular vector, singular values descending
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 3; ++r) v[k][r] = V[r * 3 + (2 - k)];
  for (int k = 0; k < 2; ++k) {
    double n2 = 0;
    for (int r = 0; r < 3; ++r) { u[k][r] = A[r * 3] * v[k][0] + A[r * 3 + 1] * v[k][1] + A[r * 3 + 2] * v[k][2]; n2 += u[k][r] * u[k][r]; }
    const double inv = n2 > 0 ? 1.0 / sqrt(n2) : 0.0;
    for (int r = 0; r < 3; ++r) u[k][r] *= inv;
  }
  {
    const double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) { u[1][r] -= d * u[0][r]; n2 += u[1][r] * u[1][r]; }
    const double inv = n2 > 0 ? 1.0 / sqrt(n2) : 0.0;
    for (int r = 0; r < 3; ++r) u[1][r] *= inv;
  }
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  const double detV = v[0][0] * (v[1][1] * v[2][2] - v[1][2] * v[2][1]) - v[0][1] * (v[1][0] * v[2][2] - v[1][2] * v[2][0]) +
                      v[0][2] * (v[1][0] * v[2][1] - v[1][1] * v[2][0]);
  const double s2 = detV < 0 ? -1.0 : 1.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = (float)(u[0][r] * v[0][c] + u[1][r] * v[1][c] + s2 * u[2][r] * v[2][c]);
}

// transformation estimation + bookkeeping of one finished iteration (icp.hpp computeTransformation loop body after the
// correspondence step; default_convergence_criteria.hpp hasConverged).  tot: n, sum d^2, sum s[3], sum t[3], sum t s^T [9]
__device__ void icp_finish_iteration(const double *tot, IcpState *st, const IcpParams &prm, double *trace) {
  const int n = (int)tot[0];
  st->n_corr = n;
  if (n < 3) { st->state = 5; return; }  // "Not enough correspondences found"
  const float one_over_n = 1.f / (float)n;
  float sm[3], dm[3], sigma[9], R[9];
  for (int c = 0; c < 3; ++c) { sm[c] = (float)tot[2 + c] * one_over_n; dm[c] = (float)tot[5 + c] * one_over_n; }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      sigma[r * 3 + c] = one_over_n * (float)(tot[8 + r * 3 + c] - (double)dm[r] * tot[2 + c] - tot[5 + r] * (double)sm[c] +
                                              (double)n * (double)dm[r] * (double)sm[c]);
  svd_rotation(sigma, R);
  float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = dm[r] - ((R[r * 3] * sm[0] + R[r * 3 + 1] * sm[1]) + R[r * 3 + 2] * sm[2]);
  }
  float Tn[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      Tn[r * 4 + c] = ((T[r * 4] * st->Tfinal[c] + T[r * 4 + 1] * st->Tfinal[4 + c]) + T[r * 4 + 2] * st->Tfinal[8 + c]) + T[r * 4 + 3] * st->Tfinal[12 + c];
  for (int k = 0; k < 16; ++k) { st->T[k] = T[k]; st->Tfinal[k] = Tn[k]; }
  const double mse = tot[1] / (double)n;
  if (trace) {
    double *tr = trace + (size_t)st->iterations * 14;
    tr[0] = n; tr[1] = mse;
    for (int k = 0; k < 9; ++k) tr[2 + k] = R[k];
    for (int k = 0; k < 3; ++k) tr[11 + k] = T[k * 4 + 3];
  }
  const int it = ++st->iterations;
  if (it >= prm.max_iterations) { st->state = 1; return; }
  const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1);
  const double translation_sqr = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
  if (cos_angle >= 1.0 - prm.transformation_epsilon && translation_sqr <= prm.transformation_epsilon) { st->state = 2; return; }
  if (fabs(mse - st->prev_mse) < 1e-12) { st->state = 3; return; }
  if (fabs(mse - st->prev_mse) / st->prev_mse < prm.fitness_epsilon) { st->state = 4; return; }
  st->prev_mse = mse;
}

// fitness_pass = 0: one ICP iteration on src_cur (in place).  fitness_pass = 1: getFitnessScore over src0 * Tfinal.
__global__ void __launch_bounds__(ICP_THREADS)
icp_iterate_kernel(float4 *__restrict__ src_cur, const float4 *__restrict__ src0, int n_src, const float4 *__restrict__ tgt, int n_tgt,
                   GridIndex g, IcpState *st, double *__restrict__ partials, IcpParams prm, double *trace, int fitness_pass) {
  __shared__ float s_T[12];
  __shared__ int s_flag[2];  // state, iterations
  __shared__ int s_unres[ICP_THREADS], s_n_unres;
  __shared__ float s_q[ICP_THREADS][3], s_bd[ICP_THREADS];
  __shared__ int s_bi[ICP_THREADS];
  __shared__ double s_red[ICP_WARPS][ICP_NSUM];
  __shared__ double s_tot[ICP_NSUM];
  __shared__ bool s_last;
  if (threadIdx.x == 0) { s_flag[0] = st->state; s_flag[1] = st->iterations; s_n_unres = 0; }
  if (threadIdx.x < 12) s_T[threadIdx.x] = fitness_pass ? st->Tfinal[threadIdx.x] : st->T[threadIdx.x];
  __syncthreads();
  if (!fitness_pass && s_flag[0] != 0) return;  // converged (or failed) in an earlier launch
  const bool apply = fitness_pass || s_flag[1] > 0;

  const int i = blockIdx.x * ICP_THREADS + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f, bd = 3.402823466e+38f;
  int bi = 0x7fffffff;
  bool valid = false, unresolved = false;
  if (i < n_src) {
    float4 p = fitness_pass ? src0[i] : src_cur[i];
    if (apply) {  // pcl::transformPointCloud: float, left to right
      const float x = ((s_T[0] * p.x + s_T[1] * p.y) + s_T[2] * p.z) + s_T[3];
      const float y = ((s_T[4] * p.x + s_T[5] * p.y) + s_T[6] * p.z) + s_T[7];
      const float z = ((s_T[8] * p.x + s_T[9] * p.y) + s_T[10] * p.z) + s_T[11];
      p.x = x; p.y = y; p.z = z;
      if (!fitness_pass) src_cur[i] = p;
    }
    qx = p.x; qy = p.y; qz = p.z;
    valid = isfinite(qx) && isfinite(qy) && isfinite(qz) && fabsf(qx) < 1e6f && fabsf(qy) < 1e6f && fabsf(qz) < 1e6f;
    if (valid && !nn_grid(g, qx, qy, qz, bd, bi)) {
      unresolved = true;
      const int slot = atomicAdd(&s_n_unres, 1);
      s_unres[slot] = threadIdx.x;
      s_q[threadIdx.x][0] = qx; s_q[threadIdx.x][1] = qy; s_q[threadIdx.x][2] = qz;
    }
  }
  __syncthreads();
  // exhaustive pass for the queries whose neighbour is farther than ICP_RINGS cells: one warp per query, lanes stride the target
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_unres = s_n_unres;
    for (int u = wid; u < n_unres; u += ICP_WARPS) {
      const int owner = s_unres[u];
      const float ux = s_q[owner][0], uy = s_q[owner][1], uz = s_q[owner][2];
      float d = 3.402823466e+38f;
      int id = 0x7fffffff;
      for (int t = lane; t < n_tgt; t += 32) nn_offer(l2_simple(ux, uy, uz, tgt[t]), t, d, id);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, id, o);
        nn_offer(od, oi, d, id);
      }
      if (lane == 0) { s_bd[owner] = d; s_bi[owner] = id; }
    }
  }
  __syncthreads();
  if (unresolved) { bd = s_bd[threadIdx.x]; bi = s_bi[threadIdx.x]; }

  double acc[ICP_NSUM];
#pragma unroll
  for (int k = 0; k < ICP_NSUM; ++k) acc[k] = 0.0;
  if (valid && bi != 0x7fffffff && (fitness_pass || !((double)bd > prm.max_dist_sqr))) {
    acc[0] = 1.0;
    acc[1] = (double)bd;
    if (!fitness_pass) {
      const float4 m = tgt[bi];
      const double s3[3] = {qx, qy, qz}, t3[3] = {m.x, m.y, m.z};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        acc[2 + r] = s3[r];
        acc[5 + r] = t3[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[8 + r * 3 + c] = t3[r] * s3[c];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < ICP_NSUM; ++k) acc[k] = warp_sum(acc[k]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < ICP_NSUM; ++k) s_red[threadIdx.x >> 5][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < ICP_NSUM) {
    double v = 0.0;
    for (int w = 0; w < ICP_WARPS; ++w) v += s_red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * ICP_NSUM + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&st->ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < ICP_NSUM) {
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(partials + (size_t)b * ICP_NSUM + threadIdx.x);
    s_tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ticket = 0;
    if (fitness_pass) {
      st->fitness = s_tot[0] > 0 ? s_tot[1] / s_tot[0] : 1.7976931348623157e308;
    } else {
      icp_finish_iteration(s_tot, st, prm, trace);
    }
  }
}

__global__ void icp_init_kernel(IcpState *st) {
  for (int k = 0; k < 16; ++k) { st->T[k] = (k % 5 == 0) ? 1.f : 0.f; st->Tfinal[k] = (k % 5 == 0) ? 1.f : 0.f; }
  st->prev_mse = 1.7976931348623157e308;  // std::numeric_limits<double>::max()
  st->fitness = 0.0;
  st->iterations = 0;
  st->state = 0;
  st->n_corr = 0;
  st->ticket = 0;
}

}  // namespace

int lc_icp_device(AlegoHandle *h, const float *src_host, int n_src, const float *tgt_host, int n_tgt, double max_corr_dist,
                  int max_iterations, double transformation_epsilon, double fitness_epsilon, AlegoIcpResult *out, double *trace_host) {
  cudaStream_t s = h->stream;
  int rc;
  if (n_src > h->icp_cap_src) {
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (h->icp_src) cudaFree(h->icp_src);
    if (h->icp_src0) cudaFree(h->icp_src0);
    if (h->icp_partials) cudaFree(h->icp_partials);
    h->icp_src = h->icp_src0 = nullptr; h->icp_partials = nullptr; h->icp_cap_src = 0;
    const int cap = std::max(n_src, 4096);
    CUDA_TRY(h, cudaMalloc(&h->icp_src, (size_t)cap * sizeof(float4)));
    CUDA_TRY(h, cudaMalloc(&h->icp_src0, (size_t)cap * sizeof(float4)));
    CUDA_TRY(h, cudaMalloc(&h->icp_partials, (size_t)div_up(cap, ICP_THREADS) * ICP_NSUM * sizeof(double)));
    h->icp_cap_src = cap;
  }
  if (n_tgt > h->icp_cap_tgt) {
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (h->icp_tgt) cudaFree(h->icp_tgt);
    h->icp_tgt = nullptr; h->icp_cap_tgt = 0;
    const int cap = std::max(n_tgt, 4096);
    CUDA_TRY(h, cudaMalloc(&h->icp_tgt, (size_t)cap * sizeof(float4)));
    // history cloud = VoxelGrid(1.0) output (laserMapping.cpp:41,811): 2 m cells hold a handful of points each
    if ((rc = grid_alloc(h, &h->g_icp, cap, 2.0f, 0, 1)) != ALEGO_OK) return rc;
    h->icp_cap_tgt = cap;
  }
  if (max_iterations > h->icp_trace_cap || !h->icp_state) {
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (h->icp_trace) cudaFree(h->icp_trace);
    h->icp_trace = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->icp_trace, (size_t)std::max(max_iterations, 1) * 14 * sizeof(double)));
    h->icp_trace_cap = std::max(max_iterations, 1);
    if (!h->icp_state) CUDA_TRY(h, cudaMalloc(&h->icp_state, sizeof(IcpState)));
    if (!h->icp_n) CUDA_TRY(h, cudaMalloc(&h->icp_n, 2 * sizeof(int)));
  }
  IcpState *st = static_cast<IcpState *>(h->icp_state);
  CUDA_TRY(h, cudaMemcpyAsync(h->icp_src, src_host, (size_t)n_src * sizeof(float4), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->icp_src0, h->icp_src, (size_t)n_src * sizeof(float4), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->icp_tgt, tgt_host, (size_t)n_tgt * sizeof(float4), cudaMemcpyHostToDevice, s));
  const int n2[2] = {n_src, n_tgt};
  CUDA_TRY(h, cudaMemcpyAsync(h->icp_n, n2, sizeof n2, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));  // n2 is a stack array
  if ((rc = grid_build(h, &h->g_icp, h->icp_tgt, 0, h->icp_n + 1, 0, "icp_target", false, 1)) != ALEGO_OK) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->icp_trace, 0, (size_t)h->icp_trace_cap * 14 * sizeof(double), s));
  { LAUNCH(h, "icp_init"); icp_init_kernel<<<1, 1, 0, s>>>(st); }
  IcpParams prm;
  prm.max_dist_sqr = max_corr_dist * max_corr_dist;
  prm.transformation_epsilon = transformation_epsilon;
  prm.fitness_epsilon = fitness_epsilon;
  prm.max_iterations = max_iterations;
  const int blocks = div_up(n_src, ICP_THREADS);
  IcpState hs;
  for (int it = 0; it < max_iterations;) {
    // launches are enqueued eight at a time; the device flag makes the ones after convergence return immediately
    const int burst = std::min(8, max_iterations - it);
    for (int k = 0; k < burst; ++k) {
      LAUNCH(h, "icp_iterate");
      icp_iterate_kernel<<<blocks, ICP_THREADS, 0, s>>>(h->icp_src, h->icp_src0, n_src, h->icp_tgt, n_tgt, h->g_icp, st, h->icp_partials, prm,
                                                       h->icp_trace, 0);
    }
    it += burst;
    CUDA_TRY(h, cudaMemcpyAsync(&hs, st, sizeof hs, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (hs.state != 0) break;
  }
  { LAUNCH(h, "icp_fitness");
    icp_iterate_kernel<<<blocks, ICP_THREADS, 0, s>>>(h->icp_src, h->icp_src0, n_src, h->icp_tgt, n_tgt, h->g_icp, st, h->icp_partials, prm,
                                                     nullptr, 1); }
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaMemcpyAsync(&hs, st, sizeof hs, cudaMemcpyDeviceToHost, s));
  if (trace_host)
    CUDA_TRY(h, cudaMemcpyAsync(trace_host, h->icp_trace, (size_t)max_iterations * 14 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  for (int k = 0; k < 16; ++k) out->final_transformation[k] = hs.Tfinal[k];
  out->fitness_score = hs.fitness;
  out->has_converged = (hs.state >= 1 && hs.state <= 4) ? 1 : 0;
  out->iterations = hs.iterations;
  out->convergence_state = hs.state;
  out->n_correspondences = hs.n_corr;
  return ALEGO_OK;
}
