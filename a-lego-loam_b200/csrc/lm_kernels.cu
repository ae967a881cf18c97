// lm_kernels.cu — LaserMapping scan-to-map on the device: replaces src/laserMapping.cpp:188-192, 325-489 and
// pointAssociateToMap (include/alego/laserMapping.h:187-194).
//
//       lm_prepare   : transformAssociateToMap (:188-192)
// K9    lm_voxel     : downsampleCurrentScan — four VoxelGrid filters (:325-346), one CTA per cloud
// K13   grid_build   : (grid.cu) index of the local map, replaces the kd-tree builds (:356-357)
// K14   lm_assoc<EDGE>  : per corner query — pointAssociateToMap, exact 5-NN, mean + 3x3 scatter, symmetric
//                         eigen-solve, lambda2 > 3*lambda1 test, line end points (:371-417)
// K15   lm_assoc<PLANE> : per surf query — 5-NN, 5x3 least squares A n = -1, normalise, 0.2 m test (:419-462)
// K16/K17 lm_solve   : one CTA per sequence — LidarEdge/LidarPlane residuals, Huber, 6x6 reduction, LM
//                      (solver.cuh), lm_outer_iters fresh solves (:360-478), transformUpdate (:481-489)
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "grid.cuh"
#include "lm_kernels.cuh"
#include "map_rows.cuh"
#include "solver.cuh"
#include "sort_voxel.cuh"
#include "vox_order.cuh"

namespace {

struct LmInputs {  // where the three input clouds of every sequence live
  const float4 *ext_corner, *ext_surf, *ext_outlier;
  int ext_cap_c, ext_cap_s, ext_cap_o;
  const int *ext_n;  // [B][4]
  const int *use_ext;  // [B]
  const float4 *lo_corner, *lo_surf, *lo_outlier;  // less_sharp, less_flat, outlier of the current sweep
  int lo_cap_c, lo_cap_s, lo_cap_o;
  const int *lo_n_corner, *lo_n_surf;  // ring_off arrays, entry [b*(R+1)+R]
  const int *lo_n_outlier;             // [B]
  int R;
};
__device__ __forceinline__ const float4 *lm_input(const LmInputs &in, int b, int kind, int *n) {
  if (in.use_ext[b]) {
    *n = in.ext_n[b * 4 + kind];
    if (kind == 0) { *n = min(*n, in.ext_cap_c); return in.ext_corner + (size_t)b * in.ext_cap_c; }
    if (kind == 1) { *n = min(*n, in.ext_cap_s); return in.ext_surf + (size_t)b * in.ext_cap_s; }
    *n = min(*n, in.ext_cap_o);
    return in.ext_outlier + (size_t)b * in.ext_cap_o;
  }
  if (kind == 0) { *n = in.lo_n_corner[b * (in.R + 1) + in.R]; return in.lo_corner + (size_t)b * in.lo_cap_c; }
  if (kind == 1) { *n = in.lo_n_surf[b * (in.R + 1) + in.R]; return in.lo_surf + (size_t)b * in.lo_cap_s; }
  *n = in.lo_n_outlier[b];
  return in.lo_outlier + (size_t)b * in.lo_cap_o;
}

__device__ __forceinline__ void mat3_mul_d(const double *A, const double *Bm, double *C) {
  double t[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[r * 3 + c] = A[r * 3] * Bm[c] + A[r * 3 + 1] * Bm[3 + c] + A[r * 3 + 2] * Bm[6 + c];
  for (int q = 0; q < 9; ++q) C[q] = t[q];
}

// transformAssociateToMap (:188-192); in pipeline mode odom2laser is LaserOdometry's (t_w_cur_, r_w_cur_)
__global__ void lm_prepare_kernel(const int *use_ext, const Pose *lo_pose, Pose *o2l, const Pose *m2o, Pose *m2l, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (!use_ext[b]) o2l[b] = lo_pose[b];  // LaserOdometry's (t_w_cur_, r_w_cur_) after the sweep being mapped
  const Pose &mo = m2o[b];
  const Pose &ol = o2l[b];
  for (int r = 0; r < 3; ++r) m2l[b].t[r] = mo.R[r * 3] * ol.t[0] + mo.R[r * 3 + 1] * ol.t[1] + mo.R[r * 3 + 2] * ol.t[2] + mo.t[r];
  mat3_mul_d(mo.R, ol.R, m2l[b].R);
}

#define LMV_THREADS 1024
#define LMV_WARPS (LMV_THREADS / 32)
// dynamic shared memory: the digit counters / centroid windows, then LMV_EXACT_RECORDS records of the list whose exact
// std::sort order is being established (clouds up to that size are partitioned in shared memory, longer ones in global
// memory); sized so that two CTAs still share an SM
#define LMV_EXACT_RECORDS 4608
#define LMV_VOX_BYTES ((sizeof(VoxShared<LMV_WARPS>) + 15) & ~(size_t)15)
// voxel_single (the one-CTA, all-in-one variant: alego_voxel_grid, local-map assembly): behind the counters either the records of a
// short list, or — for a longer one — the work-sharing scratch of its 32 warps (task queue, cooperative-phase ranges, per-warp range
// copies and swap positions)
#define LMV_WS_BYTES (((sizeof(IswShared) + 15) & ~(size_t)15) + ((sizeof(IswBig) + 15) & ~(size_t)15) + \
                      (size_t)LMV_WARPS * ISB_REG * (sizeof(u64) + sizeof(unsigned short)))
#define LMV_EXACT_BYTES ((size_t)LMV_EXACT_RECORDS * sizeof(u64))
#define LMV_SMEM_EXACT (LMV_VOX_BYTES + (LMV_WS_BYTES > LMV_EXACT_BYTES ? LMV_WS_BYTES : LMV_EXACT_BYTES))
#define LMV_SMEM LMV_VOX_BYTES
__device__ __forceinline__ VoxExact lmv_exact(uint8_t *smem, IsbShared *isb) {
  VoxExact ex;
  ex.e_smem = reinterpret_cast<u64 *>(smem + LMV_VOX_BYTES);
  ex.e_cap = LMV_EXACT_RECORDS;
  ex.pos = nullptr;  // positions and range lists come out of the idle second key buffer
  ex.lists = nullptr;
  ex.list_cap = 0;
  ex.isb = isb;
  ex.wpos = nullptr;
  // the same bytes, laid out for a list that does not fit them
  uint8_t *w = smem + LMV_VOX_BYTES;
  ex.wq = reinterpret_cast<IswShared *>(w);
  ex.wbig = reinterpret_cast<IswBig *>(w + ((sizeof(IswShared) + 15) & ~(size_t)15));
  ex.wbuf_all = reinterpret_cast<u64 *>(reinterpret_cast<uint8_t *>(ex.wbig) + ((sizeof(IswBig) + 15) & ~(size_t)15));
  ex.wpos_all = reinterpret_cast<unsigned short *>(ex.wbuf_all + (size_t)LMV_WARPS * ISB_REG);
  return ex;
}
// kind 0 corner (leaf lm_corner_leaf), 1 surf, 2 outlier : blockIdx.y selects; kind 3 = surf_total (own launches).
// stage 0: record lists (sort_voxel.cuh block_voxel_keys).  A cloud none of whose voxels holds three or more points (typically
//          surf_total: its inputs are already one point per voxel) is finished on the spot — the order of the records inside a
//          voxel cannot change a sum of at most two terms; every other cloud is left to the ordering kernel (vox_order.cu:
//          pcl::VoxelGrid's std::sort order) and finished by stage 1.
// stage 1: stable radix sort + centroids of the clouds stage 0 left pending.
__global__ void __launch_bounds__(LMV_THREADS)
lm_voxel_kernel(LmInputs in, int first_kind, int stage, float leaf_c, float leaf_s, float leaf_o, float4 *ds_c, float4 *ds_s,
                float4 *ds_o, float4 *total, float4 *ds_total, int cap_c, int cap_s, int cap_o, int *lm_n, u64 *sort_base, u64 *sort_c,
                u64 *sort_s, u64 *sort_o, int sort_cap_c, int sort_cap_s, int sort_cap_o, VoxState *states) {
  extern __shared__ __align__(16) uint8_t lmv_smem[];
  VoxShared<LMV_WARPS> *sh = reinterpret_cast<VoxShared<LMV_WARPS> *>(lmv_smem);
  __shared__ int s_flag;
  const int b = blockIdx.x, kind = first_kind + blockIdx.y;
  VoxState *state = states + b * 4 + kind;
  if (stage == 1 && state->done) return;
  const float4 *src;
  int n;
  float leaf;
  float4 *dst;
  u64 *keys;
  int sort_cap;
  if (kind == 3) {
    // laser_surf_total_ = laser_surf_ds_ + laser_outlier_ds_ (:337-340)
    const int ns = lm_n[b * 8 + 1], no = lm_n[b * 8 + 2];
    float4 *tot = total + (size_t)b * (cap_s + cap_o);
    if (stage == 0) {
      for (int t = threadIdx.x; t < ns; t += blockDim.x) tot[t] = ds_s[(size_t)b * cap_s + t];
      for (int t = threadIdx.x; t < no; t += blockDim.x) tot[ns + t] = ds_o[(size_t)b * cap_o + t];
      __syncthreads();
      if (threadIdx.x == 0) lm_n[b * 8 + 3] = ns + no;
    }
    src = tot; n = ns + no; leaf = leaf_s;
    dst = ds_total + (size_t)b * (cap_s + cap_o);
    keys = sort_s + (size_t)b * 2 * sort_cap_s; sort_cap = sort_cap_s;
  } else {
    src = lm_input(in, b, kind, &n);
    if (kind == 0) { leaf = leaf_c; dst = ds_c + (size_t)b * cap_c; keys = sort_c + (size_t)b * 2 * sort_cap_c; sort_cap = sort_cap_c; n = min(n, cap_c); }
    else if (kind == 1) { leaf = leaf_s; dst = ds_s + (size_t)b * cap_s; keys = sort_s + (size_t)b * 2 * sort_cap_s; sort_cap = sort_cap_s; n = min(n, cap_s); }
    else { leaf = leaf_o; dst = ds_o + (size_t)b * cap_o; keys = sort_o + (size_t)b * 2 * sort_cap_o; sort_cap = sort_cap_o; n = min(n, cap_o); }
  }
  n = min(n, sort_cap);
  int *n_out_ptr = lm_n + b * 8 + (kind == 3 ? 4 : kind);
  // ping-pong key buffers [2][sort_cap] in global memory (L2 resident), digit counters in shared memory
  if (stage == 1) {
    const VoxState st = *state;
    const int n_out = block_voxel_finish<LMV_WARPS>(src, keys, keys + sort_cap, dst, sh, st.n, st.nv, st.frame.key_bits);
    if (threadIdx.x == 0) { *n_out_ptr = n_out; state->done = 1; }  // nothing pending any more (the ordering kernel skips it)
    return;
  }
  const bool done = block_voxel_keys<LMV_WARPS, false>(src, n, [](int) { return true; }, leaf, keys, dst, sh, nullptr, nullptr);
  const int nl = max(sh->n, 0), nv = sh->nv;
  // lists with non-finite points (never on the hot path) keep the input order inside a voxel, as do lists that cannot care
  const bool need_order = !done && nv == nl && nl > 16 && block_any_voxel_ge3<LMV_WARPS>(keys, nl, reinterpret_cast<u64 *>(dst), &s_flag);
  if (!need_order) {
    const int n_out = done ? nl : block_voxel_finish<LMV_WARPS>(src, keys, keys + sort_cap, dst, sh, nl, nv, sh->frame.key_bits);
    if (threadIdx.x == 0) { *n_out_ptr = n_out; state->done = 1; }
    return;
  }
  if (threadIdx.x == 0) {
    VoxState st;
    st.frame = sh->frame;
    st.n = nl;
    st.nv = nv;
    st.done = 0;
    st.off = (long long)(keys - sort_base);
    st.off_b = st.off + sort_cap;
    *state = st;
  }
}

// stand-alone VoxelGrid of one device cloud (alego_voxel_grid)
__global__ void __launch_bounds__(LMV_THREADS)
voxel_single_kernel(const float4 *src, int n, float leaf, float4 *dst, u64 *keys, int *n_out) {
  extern __shared__ __align__(16) uint8_t lmv_smem[];
  VoxShared<LMV_WARPS> *sh = reinterpret_cast<VoxShared<LMV_WARPS> *>(lmv_smem);
  __shared__ IsbShared s_isb;
  const VoxExact ex = lmv_exact(lmv_smem, &s_isb);
  const int m = block_voxel_grid8<LMV_WARPS, false>(src, n, [](int) { return true; }, leaf, keys, keys + n, dst, sh, nullptr, nullptr, &ex);
  if (threadIdx.x == 0) *n_out = m;
}

// ---- small dense helpers (same algorithms as the oracle, so results agree to rounding) -------------
// symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues ascending (Eigen::SelfAdjointEigenSolver
// convention, laserMapping.cpp:397-403); V columns = eigenvectors
__device__ void eig3_sym_dev(const double A[9], double w[3], double V[9]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double dsum = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-40 * dsum || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int o0 = 0, o1 = 1, o2 = 2;  // stable ascending order of the diagonal
  if (a[o1][o1] < a[o0][o0]) { int t = o0; o0 = o1; o1 = t; }
  if (a[o2][o2] < a[o1][o1]) { int t = o1; o1 = o2; o2 = t; }
  if (a[o1][o1] < a[o0][o0]) { int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    w[k] = a[ord[k]][ord[k]];
#pragma unroll
    for (int r = 0; r < 3; ++r) V[r * 3 + k] = v[r][ord[k]];
  }
}

// least squares of the 5x3 system A n = b by column-pivoted Householder QR (colPivHouseholderQr().solve, :435)
__device__ void lstsq_5x3_dev(double A[5][3], double b[5], double n[3]) {
  int perm[3] = {0, 1, 2};
  int rank = 3;
  for (int c = 0; c < 3; ++c) {
    int best = c;
    double bn = -1;
    for (int c2 = c; c2 < 3; ++c2) {
      double q = 0;
      for (int i = c; i < 5; ++i) q += A[i][c2] * A[i][c2];
      if (q > bn) { bn = q; best = c2; }
    }
    if (best != c) {
      for (int i = 0; i < 5; ++i) { const double t = A[i][c]; A[i][c] = A[i][best]; A[i][best] = t; }
      const int t = perm[c]; perm[c] = perm[best]; perm[best] = t;
    }
    const double nrm = sqrt(bn);
    if (nrm < 1e-300) { rank = c; break; }
    const double alpha = A[c][c] > 0 ? -nrm : nrm;
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = c; i < 5; ++i) v[i] = A[i][c];
    v[c] -= alpha;
    double vn = 0;
    for (int i = c; i < 5; ++i) vn += v[i] * v[i];
    if (vn > 0) {
      for (int c2 = c; c2 < 3; ++c2) {
        double dot = 0;
        for (int i = c; i < 5; ++i) dot += v[i] * A[i][c2];
        const double f = 2.0 * dot / vn;
        for (int i = c; i < 5; ++i) A[i][c2] -= f * v[i];
      }
      double dot = 0;
      for (int i = c; i < 5; ++i) dot += v[i] * b[i];
      const double f = 2.0 * dot / vn;
      for (int i = c; i < 5; ++i) b[i] -= f * v[i];
    }
  }
  double y[3] = {0, 0, 0};
  for (int c = rank - 1; c >= 0; --c) {
    double acc = b[c];
    for (int c2 = c + 1; c2 < rank; ++c2) acc -= A[c][c2] * y[c2];
    y[c] = acc / A[c][c];
  }
  for (int c = 0; c < 3; ++c) n[perm[c]] = y[c];
}

// exact 5 nearest neighbours within squared distance < 1.0 (float), ascending (distance, index).
// A cell edge >= 1 m makes the 27 surrounding cells sufficient for every neighbour that can pass the gate.  The
// x-runs of 3 cells are contiguous in the blocked hash layout (grid.cuh), so a query reads 9 columns x (1 or 2)
// ranges; the range bounds of all columns are independent loads issued before any point is touched.
__device__ __forceinline__ void knn5_offer(float d, int idx, float *bd, int *bi, int &found) {
  if (d < bd[4] || (d == bd[4] && idx < bi[4])) {
    // two hash-colliding blocks can present the same bucket twice: a point already held is not offered again (a point
    // that was evicted can never re-enter: everything held is strictly better)
    if (idx == bi[0] || idx == bi[1] || idx == bi[2] || idx == bi[3] || idx == bi[4]) return;
    int pos = 4;
#pragma unroll
    for (int t = 4; t > 0; --t) {
      if (pos == t && (d < bd[t - 1] || (d == bd[t - 1] && idx < bi[t - 1]))) {
        bd[t] = bd[t - 1];
        bi[t] = bi[t - 1];
        pos = t - 1;
      }
    }
#pragma unroll
    for (int t = 0; t < 5; ++t)
      if (pos == t) { bd[t] = d; bi[t] = idx; }
    ++found;
  }
}

__device__ __forceinline__ int knn5_gate(const GridIndex &g, int b, float qx, float qy, float qz, float *bd, int *bi) {
  const int T = g.table_size;
  const int *cs = g.cell_start + (size_t)b * GRID_TABLE_STRIDE(T);
  const float4 *sp = g.sorted + (size_t)b * g.cap;
  const float inv = 1.0f / g.cell;
  const int cx = grid_coord(qx, inv), cy = grid_coord(qy, inv), cz = grid_coord(qz, inv);
#pragma unroll
  for (int t = 0; t < 5; ++t) { bd[t] = 3.402823466e+38f; bi[t] = 0x7fffffff; }
  int found = 0;
#pragma unroll 1
  for (int lz = 0; lz < 3; ++lz) {  // one z layer at a time: 3 columns = up to 12 independent bound loads in flight
    const int iz = cz + lz - 1;
    int r1s[3], r1e[3], r2s[3], r2e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int iy = cy + c - 1;
      const int h0 = grid_hash(cx - 1, iy, iz, T), h1 = grid_hash(cx, iy, iz, T), h2 = grid_hash(cx + 1, iy, iz, T);
      r1s[c] = cs[h0];
      r2e[c] = cs[h2 + 1];
      if (h2 == h0 + 2) {          // one block: [h0, h2] contiguous
        r1e[c] = r2e[c];
        r2s[c] = r2e[c];
      } else if (h1 == h0 + 1) {   // {cx-1, cx} | {cx+1}
        r1e[c] = cs[h1 + 1];
        r2s[c] = cs[h2];
      } else {                     // {cx-1} | {cx, cx+1}
        r1e[c] = cs[h0 + 1];
        r2s[c] = cs[h1];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      for (int t = r1s[c]; t < r1e[c]; ++t) {
        const float4 p = sp[t];
        const float d = l2_simple(qx, qy, qz, p);
        if (d < 1.0f) knn5_offer(d, __float_as_int(p.w), bd, bi, found);
      }
      for (int t = r2s[c]; t < r2e[c]; ++t) {
        const float4 p = sp[t];
        const float d = l2_simple(qx, qy, qz, p);
        if (d < 1.0f) knn5_offer(d, __float_as_int(p.w), bd, bi, found);
      }
    }
  }
  return found < 5 ? found : 5;
}

// K14a/K15a: per query — pointAssociateToMap, exact gated 5-NN; writes the 5 map indices (nn[0] = -1: no residual)
__global__ void __launch_bounds__(256, 4)
lm_knn_kernel(const float4 *__restrict__ query, int qcap, const int *__restrict__ lm_n, int n_slot, GridIndex g,
              const Pose *__restrict__ m2l, const int *__restrict__ guard, int *__restrict__ nn) {
  const int b = blockIdx.y;
  const int nq = min(lm_n[b * 8 + n_slot], qcap);
  const bool ok = guard[b] != 0;
  const Pose &P = m2l[b];
  // the buffers are sized for the largest possible cloud, the queries that exist are a few thousand: bounded grid + stride
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
    int *o = nn + ((size_t)b * qcap + i) * 5;
    if (!ok) { o[0] = -1; continue; }
    const float4 cp = query[(size_t)b * qcap + i];
    // pointAssociateToMap (laserMapping.h:187-194): double transform, float result
    const float sx = (float)(P.R[0] * cp.x + P.R[1] * cp.y + P.R[2] * cp.z + P.t[0]);
    const float sy = (float)(P.R[3] * cp.x + P.R[4] * cp.y + P.R[5] * cp.z + P.t[1]);
    const float sz = (float)(P.R[6] * cp.x + P.R[7] * cp.y + P.R[8] * cp.z + P.t[2]);
    float bd[5];
    int bi[5];
    if (knn5_gate(g, b, sx, sy, sz, bd, bi) < 5) { o[0] = -1; continue; }  // point_dist_[4] < 1.0 (:376, :426)
#pragma unroll
    for (int t = 0; t < 5; ++t) o[t] = bi[t];
  }
}

// K14a/K15a over the voxel-row index (map_rows.cuh) of a map that is pcl::VoxelGrid output: same exact gated 5-NN, same
// (distance, index) ranking — the index of a point is its position in the map cloud either way — but the candidates are read
// in place: the rows (y, z voxel pairs) and x voxels that a point within the gate can occupy, found by voxel arithmetic with a
// margin that covers float rounding of the coordinates (1.001 m for a 1 m gate).
__global__ void __launch_bounds__(256, 4)
lm_knn_rows_kernel(const float4 *__restrict__ query, int qcap, const int *__restrict__ lm_n, int n_slot, const float4 *__restrict__ map,
                   int map_cap, const MapFrame *__restrict__ frame, const uint2 *__restrict__ tab, int tab_cap,
                   const Pose *__restrict__ m2l, const int *__restrict__ guard, int *__restrict__ nn) {
  const int b = blockIdx.y;
  const int nq = min(lm_n[b * 8 + n_slot], qcap);
  const MapFrame F = frame[b];
  const bool ok = guard[b] != 0 && F.valid;
  const Pose &P = m2l[b];
  const float4 *mp = map + (size_t)b * map_cap;
  const uint2 *t = tab + (size_t)b * tab_cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
    int *o = nn + ((size_t)b * qcap + i) * 5;
    if (!ok) { o[0] = -1; continue; }
    const float4 cp = query[(size_t)b * qcap + i];
    // pointAssociateToMap (laserMapping.h:187-194): double transform, float result
    const float sx = (float)(P.R[0] * cp.x + P.R[1] * cp.y + P.R[2] * cp.z + P.t[0]);
    const float sy = (float)(P.R[3] * cp.x + P.R[4] * cp.y + P.R[5] * cp.z + P.t[1]);
    const float sz = (float)(P.R[6] * cp.x + P.R[7] * cp.y + P.R[8] * cp.z + P.t[2]);
    float bd[5];
    int bi[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { bd[k] = 3.402823466e+38f; bi[k] = 0x7fffffff; }
    int found = 0;
    const float G = 1.001f;
    const int x0 = max(mr_voxel(sx - G, F.inv, F.min_b[0]), 0), x1 = min(mr_voxel(sx + G, F.inv, F.min_b[0]), F.dim[0] - 1);
    const int y0 = max(mr_voxel(sy - G, F.inv, F.min_b[1]), 0), y1 = min(mr_voxel(sy + G, F.inv, F.min_b[1]), F.dim[1] - 1);
    const int z0 = max(mr_voxel(sz - G, F.inv, F.min_b[2]), 0), z1 = min(mr_voxel(sz + G, F.inv, F.min_b[2]), F.dim[2] - 1);
    if (x0 <= x1) {
      const int w0 = x0 >> 5, w1 = x1 >> 5;
      for (int iz = z0; iz <= z1; ++iz)
        for (int iy = y0; iy <= y1; ++iy) {
          const uint2 *row = t + (size_t)(iy + F.dim[1] * iz) * F.W;
          for (int w = w0; w <= w1; ++w) {
            const uint2 ent = row[w];
            const int lo = w == w0 ? (x0 & 31) : 0, hi = w == w1 ? (x1 & 31) : 31;
            unsigned m = ent.x & (0xffffffffu << lo) & (0xffffffffu >> (31 - hi));
            while (m) {
              const int bit = __ffs(m) - 1;
              m &= m - 1;
              const int idx = (int)ent.y + __popc(ent.x & ((1u << bit) - 1u));
              const float4 p = mp[idx];
              const float d = l2_simple(sx, sy, sz, p);
              if (d < 1.0f) knn5_offer(d, idx, bd, bi, found);
            }
          }
        }
    }
    if (found < 5) { o[0] = -1; continue; }  // point_dist_[4] < 1.0 (:376, :426)
#pragma unroll
    for (int k = 0; k < 5; ++k) o[k] = bi[k];
  }
}

// K14b/K15b: per query with 5 neighbours — line test (PCA) or plane fit in double; writes the residual block
template <bool EDGE>
__device__ __forceinline__ void lm_fit_one(const float4 *__restrict__ query, int qcap, const float4 *__restrict__ map, int map_cap,
                                           const int *__restrict__ nn, double *__restrict__ out, int out_w, int b, int i);
template <bool EDGE>
__global__ void __launch_bounds__(128)
lm_fit_kernel(const float4 *__restrict__ query, int qcap, const int *__restrict__ lm_n, int n_slot, const float4 *__restrict__ map,
              int map_cap, const int *__restrict__ nn, double *__restrict__ out, int out_w) {
  const int b = blockIdx.y;
  const int nq = min(lm_n[b * 8 + n_slot], qcap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x)
    lm_fit_one<EDGE>(query, qcap, map, map_cap, nn, out, out_w, b, i);
}

template <bool EDGE>
__device__ __forceinline__ void lm_fit_one(const float4 *__restrict__ query, int qcap, const float4 *__restrict__ map, int map_cap,
                                           const int *__restrict__ nn, double *__restrict__ out, int out_w, int b, int i) {
  double *o = out + ((size_t)b * qcap + i) * out_w;
  o[0] = 0.0;
  const int *bi = nn + ((size_t)b * qcap + i) * 5;
  if (bi[0] < 0) return;
  const float4 cp = query[(size_t)b * qcap + i];
  const float4 *M = map + (size_t)b * map_cap;
  double nb[5][3];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float4 p = M[bi[j]];
    nb[j][0] = p.x; nb[j][1] = p.y; nb[j][2] = p.z;
  }
  if (EDGE) {
    double center[3] = {0, 0, 0};
    for (int j = 0; j < 5; ++j)
      for (int c = 0; c < 3; ++c) center[c] = center[c] + nb[j][c];
    for (int c = 0; c < 3; ++c) center[c] = center[c] / 5.0;
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 5; ++j) {
      const double zm[3] = {nb[j][0] - center[0], nb[j][1] - center[1], nb[j][2] - center[2]};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cov[r * 3 + c] = cov[r * 3 + c] + zm[r] * zm[c];
    }
    double w[3], V[9];
    eig3_sym_dev(cov, w, V);
    if (!(w[2] > 3 * w[1])) return;  // (:403)
    o[1] = cp.x; o[2] = cp.y; o[3] = cp.z;
    for (int c = 0; c < 3; ++c) {
      const double u = V[c * 3 + 2];
      o[4 + c] = 0.1 * u + center[c];   // lpj (:406)
      o[7 + c] = -0.1 * u + center[c];  // lpl (:407)
    }
    o[0] = 1.0;
  } else {
    double A[5][3], rhs[5] = {-1, -1, -1, -1, -1}, nrm[3];
    for (int j = 0; j < 5; ++j)
      for (int c = 0; c < 3; ++c) A[j][c] = nb[j][c];
    lstsq_5x3_dev(A, rhs, nrm);
    const double nn2 = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    const double d = 1 / nn2;
    for (int c = 0; c < 3; ++c) nrm[c] /= nn2;
    for (int j = 0; j < 5; ++j)
      if (fabs(nrm[0] * nb[j][0] + nrm[1] * nb[j][1] + nrm[2] * nb[j][2] + d) > 0.2) return;  // (:441-452)
    o[1] = cp.x; o[2] = cp.y; o[3] = cp.z;
    o[4] = nrm[0]; o[5] = nrm[1]; o[6] = nrm[2];
    o[7] = d;
    o[0] = 1.0;
  }
}

// guard of scan2MapOptimization (:350-354)
__global__ void lm_guard_kernel(const int *lm_n, const int *n_map_corner, int *guard, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  guard[b] = !(lm_n[b * 8 + 0] < 10 || lm_n[b * 8 + 3] < 100 || n_map_corner[b] < 10);
}

// Shared-memory staging of the residual blocks.  The solver wants the full register budget per thread, so an SM holds 256 solver
// threads either way: one 256-thread CTA with 200 KB when the batch fits one wave of that shape (B <= SMs), otherwise two
// 128-thread CTAs with 100 KB each — a batch of 256 sequences is then a single wave instead of 148 + 108.
#define LM_STAGE_BYTES_WIDE (200 * 1024)
#define LM_STAGE_BYTES_PAIR (100 * 1024)
struct LmResidSet {
  const double *edge;
  int n_edge_slots;
  const double *plane;
  int n_plane_slots;
  __device__ int slots() const { return n_edge_slots + n_plane_slots; }
  __device__ bool load(int i, int &kind, double cp[3], double a[3], double b[3], double c[3], double &d) const {
    if (i < n_edge_slots) {
      const double *e = edge + (size_t)i * 10;
      if (e[0] == 0.0) return false;
      kind = 2;
      for (int q = 0; q < 3; ++q) { cp[q] = e[1 + q]; a[q] = e[4 + q]; b[q] = e[7 + q]; c[q] = 0; }
      d = 0;
      return true;
    }
    const double *p = plane + (size_t)(i - n_edge_slots) * 8;
    if (p[0] == 0.0) return false;
    kind = 3;
    for (int q = 0; q < 3; ++q) { cp[q] = p[1 + q]; a[q] = p[4 + q]; b[q] = 0; c[q] = 0; }
    d = p[7];
    return true;
  }
};

// the same residuals, compacted into shared memory by stage_blocks (edge blocks of 9 doubles, then plane blocks of 7)
struct LmStagedSet {
  const double *edge;
  int n_edge;
  const double *plane;
  int n_plane;
  __device__ int slots() const { return n_edge + n_plane; }
  __device__ bool load(int i, int &kind, double cp[3], double a[3], double b[3], double c[3], double &d) const {
    if (i < n_edge) {
      const double *e = edge + (size_t)i * 9;
      kind = 2;
      for (int q = 0; q < 3; ++q) { cp[q] = e[q]; a[q] = e[3 + q]; b[q] = e[6 + q]; c[q] = 0; }
      d = 0;
      return true;
    }
    const double *p = plane + (size_t)(i - n_edge) * 7;
    kind = 3;
    for (int q = 0; q < 3; ++q) { cp[q] = p[q]; a[q] = p[3 + q]; b[q] = 0; c[q] = 0; }
    d = p[6];
    return true;
  }
};

template <int NT, int NB>
__global__ void __launch_bounds__(NT, NB)
lm_solve_kernel(const double *__restrict__ edge, int ecap, const double *__restrict__ plane, int pcap, const int *__restrict__ lm_n,
                const int *__restrict__ guard, double *lm_params, Pose *m2o, const Pose *o2l, Pose *m2l, AlegoSolveReport *report,
                double *trace, int *trace_n, int trace_cap, int outer_iters, int max_iters, double huber_a, double *pose_out,
                const Pose *lo_pose, int stage_doubles) {
  extern __shared__ __align__(16) double lm_stage[];
  const int b = blockIdx.x;
  __shared__ LmShared sh;
  __shared__ int s_red[34];
  __shared__ int s_cnt[2];
  AlegoSolveReport *rep = report + b;
  double *x = lm_params + b * 6;
  LmResidSet rs;
  rs.edge = edge + (size_t)b * ecap * 10;
  rs.n_edge_slots = min(lm_n[b * 8 + 0], ecap);
  rs.plane = plane + (size_t)b * pcap * 8;
  rs.n_plane_slots = min(lm_n[b * 8 + 4], pcap);
  if (threadIdx.x == 0) trace_n[b] = 0;
  const bool ok = guard[b] != 0;
  int ne = 0, np = 0;
  if (ok) {
    for (int i = threadIdx.x; i < rs.n_edge_slots; i += blockDim.x) ne += rs.edge[(size_t)i * 10] != 0.0;
    for (int i = threadIdx.x; i < rs.n_plane_slots; i += blockDim.x) np += rs.plane[(size_t)i * 8] != 0.0;
  }
  int tot;
  block_excl_scan(ne, s_red, &tot);
  if (threadIdx.x == 0) s_cnt[0] = tot;
  block_excl_scan(np, s_red, &tot);
  if (threadIdx.x == 0) s_cnt[1] = tot;
  __syncthreads();
  ne = s_cnt[0];
  np = s_cnt[1];
  if (threadIdx.x == 0) {
    rep->status = ok ? ALEGO_OK : ALEGO_FEW_FEATURES;
    rep->n_corner = ne; rep->n_surf = np; rep->iterations = 0; rep->initial_cost = 0; rep->final_cost = 0;
  }
  __syncthreads();
  if (ok && ne + np > 0) {
    double *tr = trace ? trace + (size_t)b * trace_cap * 7 : nullptr;
    // staged copy of the valid blocks when they fit (a few thousand correspondences do); otherwise the global arrays are swept
    LmStagedSet ss{lm_stage, 0, lm_stage, 0};
    const bool staged = ne * 9 + np * 7 <= stage_doubles;
    if (staged) {
      const double *ge = rs.edge, *gp = rs.plane;
      ss.n_edge = stage_blocks<double, 9>(rs.n_edge_slots, lm_stage, [&](int i) { return ge[(size_t)i * 10] != 0.0; },
                                          [&](int i, double *d) { for (int q = 0; q < 9; ++q) d[q] = ge[(size_t)i * 10 + 1 + q]; }, s_red);
      ss.plane = lm_stage + (size_t)ss.n_edge * 9;
      ss.n_plane = stage_blocks<double, 7>(rs.n_plane_slots, lm_stage + (size_t)ss.n_edge * 9, [&](int i) { return gp[(size_t)i * 8] != 0.0; },
                                           [&](int i, double *d) { for (int q = 0; q < 7; ++q) d[q] = gp[(size_t)i * 8 + 1 + q]; }, s_red);
    }
    for (int outer = 0; outer < outer_iters; ++outer) {  // (:360) the association pose is frozen, so both outer
      // iterations see the same correspondences; the second Solve continues from the first
      const LmResult r = staged ? block_lm_solve(ss, x, max_iters, huber_a, &sh, tr, trace_n + b, trace_cap)
                                : block_lm_solve(rs, x, max_iters, huber_a, &sh, tr, trace_n + b, trace_cap);
      if (threadIdx.x == 0) {
        if (outer == 0) rep->initial_cost = r.initial_cost;
        rep->iterations += r.iterations;
        rep->final_cost = r.final_cost;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {  // transformUpdate (:481-489)
    const PoseTrig T(x);
    Pose &ml = m2l[b];
    for (int q = 0; q < 9; ++q) ml.R[q] = T.R[q];
    for (int q = 0; q < 3; ++q) ml.t[q] = x[q];
    const Pose &ol = o2l[b];
    const double inv[9] = {ol.R[0], ol.R[3], ol.R[6], ol.R[1], ol.R[4], ol.R[7], ol.R[2], ol.R[5], ol.R[8]};
    Pose &mo = m2o[b];
    mat3_mul_d(ml.R, inv, mo.R);
    for (int r = 0; r < 3; ++r) mo.t[r] = ml.t[r] - (mo.R[r * 3] * ol.t[0] + mo.R[r * 3 + 1] * ol.t[1] + mo.R[r * 3 + 2] * ol.t[2]);
    if (pose_out) {
      double *po = pose_out + b * 12;
      for (int q = 0; q < 3; ++q) po[q] = ml.t[q];
      for (int q = 0; q < 6; ++q) po[3 + q] = x[q];
      for (int q = 0; q < 3; ++q) po[9 + q] = lo_pose ? lo_pose[b].t[q] : 0.0;
    }
  }
}

}  // namespace

static int lm_ensure_ds_buffers(AlegoHandle *h, int need_c, int need_s, int need_o);

static LmInputs make_inputs(AlegoHandle *h) {
  LmInputs in;
  in.ext_corner = h->lm_in_corner; in.ext_surf = h->lm_in_surf; in.ext_outlier = h->lm_in_outlier;
  in.ext_cap_c = h->lm_cap_c; in.ext_cap_s = h->lm_cap_s; in.ext_cap_o = h->lm_cap_o;
  in.ext_n = h->lm_in_n;
  in.use_ext = h->lm_use_ext;
  const int buf = 1 - h->cur;  // lo_scan2scan_device flipped the buffers: the sweep just processed is in 1-cur
  in.lo_corner = h->less_sharp[buf]; in.lo_surf = h->less_flat[buf]; in.lo_outlier = h->outlier_buf[buf];
  in.lo_cap_c = h->R * 120; in.lo_cap_s = h->RC; in.lo_cap_o = h->out_cap;
  in.lo_n_corner = h->ls_ring_off[buf]; in.lo_n_surf = h->lf_ring_off[buf]; in.lo_n_outlier = h->n_outlier_buf[buf];
  in.R = h->R;
  return in;
}

// Decide (once per map change, synchronising) whether the surf map can use the voxel-row index: it can when every sequence's
// cloud is pcl::VoxelGrid output (one point per voxel, ascending voxel index) — what surf_from_map_ds_ is (laserMapping.cpp:316-319).
int lm_validate_map_rows(AlegoHandle *h) {
  if (h->map_rows_checked || !h->map_surf) return ALEGO_OK;
  const int rc = map_rows_validate(h, &h->rows_map_surf, h->map_surf, (size_t)h->map_cap_s, h->n_map_surf, (float)h->P.lm_surf_leaf, "map_surf");
  if (rc != ALEGO_OK) return rc;
  h->map_rows_checked = true;
  return ALEGO_OK;
}

int lm_build_map_index(AlegoHandle *h) {
  int rc = grid_build(h, &h->g_map_corner, h->map_corner, (size_t)h->map_cap_c, h->n_map_corner, 1, "map_corner");
  if (rc != ALEGO_OK) return rc;
  if (h->map_rows_checked && h->rows_map_surf.usable)
    rc = map_rows_build(h, &h->rows_map_surf, h->map_surf, (size_t)h->map_cap_s, h->n_map_surf, "map_surf");
  else
    rc = grid_build(h, &h->g_map_surf, h->map_surf, (size_t)h->map_cap_s, h->n_map_surf, 1, "map_surf");
  if (rc != ALEGO_OK) return rc;
  h->map_index_valid = true;
  return ALEGO_OK;
}

int lm_scan2map_device(AlegoHandle *h, int *guard_dev, bool write_pose, bool index_ready) {
  const int B = h->B;
  cudaStream_t s = h->launch_stream ? h->launch_stream : h->stream;
  if (!h->map_corner || !h->map_surf) { h->err = "alego_lm_scan2map: no local map (call alego_lm_set_map)"; return ALEGO_NOT_READY; }
  int rc = lm_ensure_ds_buffers(h, 0, 0, 0);
  if (rc != ALEGO_OK) return rc;
  const LmInputs in = make_inputs(h);
  { LAUNCH(h, "lm_prepare"); lm_prepare_kernel<<<div_up(B, 128), 128, 0, s>>>(h->lm_use_ext, h->o2l_lo[1 - h->cur], h->o2l, h->m2o, h->m2l, B); }
  static bool attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(lm_voxel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LMV_SMEM));
    attr_set[h->dev] = true;
  }
  const int cs = h->ds_cap_s, cc = h->ds_cap_c, co = h->ds_cap_o;
  u64 *sort_c = h->vox_sort, *sort_s = sort_c + (size_t)B * 2 * cc, *sort_o = sort_s + (size_t)B * 2 * (cs + co);
  // downsampleCurrentScan (:325-346): three filters, then the filter of their union; each = record lists -> std::sort's partition
  // phase (vox_order.cu) -> stable radix + centroids
  auto voxel_stage = [&](const char *tag, int first_kind, int n_kinds, int stage) {
    LAUNCH(h, tag);
    lm_voxel_kernel<<<dim3(B, n_kinds), LMV_THREADS, LMV_SMEM, s>>>(in, first_kind, stage, (float)h->P.lm_corner_leaf,
        (float)h->P.lm_surf_leaf, (float)h->P.lm_outlier_leaf, h->lm_corner_ds, h->lm_surf_ds, h->lm_outlier_ds, h->lm_surf_total,
        h->lm_surf_total_ds, cc, cs, co, h->lm_n, h->vox_sort, sort_c, sort_s, sort_o, cc, cs + co, co, h->lmv_state);
  };
  voxel_stage("lm_voxel_keys_3", 0, 3, 0);
  rc = vox_order_lists_by_cta(h, h->lmv_state, 4 * B, h->vox_sort, h->vox_sort, s, "lm_voxel_order_3", 4, 1);  // surf lists first
  if (rc != ALEGO_OK) return rc;
  voxel_stage("lm_voxel_finish_3", 0, 3, 1);
  voxel_stage("lm_voxel_keys_total", 3, 1, 0);
  rc = vox_order_lists_by_cta(h, h->lmv_state, 4 * B, h->vox_sort, h->vox_sort, s, "lm_voxel_order_total", 4, 3);
  if (rc != ALEGO_OK) return rc;
  voxel_stage("lm_voxel_finish_total", 3, 1, 1);
  if (!index_ready && (h->rebuild_map_every_step || !h->map_index_valid)) {  // the reference rebuilds both kd-trees every mapped frame (:356-357)
    rc = lm_build_map_index(h);
    if (rc != ALEGO_OK) return rc;
  }
  { LAUNCH(h, "lm_guard"); lm_guard_kernel<<<div_up(B, 128), 128, 0, s>>>(h->lm_n, h->n_map_corner, guard_dev, B); }
  // query slots actually populated are far fewer than the capacities: size the grids from the LO feature capacities
  { LAUNCH(h, "lm_knn_corner");
    lm_knn_kernel<<<dim3(min(div_up(cc, 256), 32), B), 256, 0, s>>>(h->lm_corner_ds, cc, h->lm_n, 0, h->g_map_corner, h->m2l, guard_dev, h->lm_nn_c); }
  { LAUNCH(h, "lm_fit_corner");
    lm_fit_kernel<true><<<dim3(min(div_up(cc, 128), 64), B), 128, 0, s>>>(h->lm_corner_ds, cc, h->lm_n, 0, h->map_corner, h->map_cap_c,
                                                                  h->lm_nn_c, h->lm_edge, 10); }
  if (h->map_rows_checked && h->rows_map_surf.usable) {
    LAUNCH(h, "lm_knn_surf");
    lm_knn_rows_kernel<<<dim3(min(div_up(cs + co, 256), 32), B), 256, 0, s>>>(h->lm_surf_total_ds, cs + co, h->lm_n, 4, h->map_surf,
        h->map_cap_s, h->rows_map_surf.frame, h->rows_map_surf.tab, h->rows_map_surf.cap, h->m2l, guard_dev, h->lm_nn_s);
  } else {
    LAUNCH(h, "lm_knn_surf");
    lm_knn_kernel<<<dim3(min(div_up(cs + co, 256), 32), B), 256, 0, s>>>(h->lm_surf_total_ds, cs + co, h->lm_n, 4, h->g_map_surf, h->m2l, guard_dev,
                                                                h->lm_nn_s);
  }
  { LAUNCH(h, "lm_fit_surf");
    lm_fit_kernel<false><<<dim3(min(div_up(cs + co, 128), 64), B), 128, 0, s>>>(h->lm_surf_total_ds, cs + co, h->lm_n, 4, h->map_surf, h->map_cap_s,
                                                                       h->lm_nn_s, h->lm_plane, 8); }
  static bool solve_attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  static int sm_count[ALEGO_MAX_DEVICES] = {};
  if (!solve_attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(lm_solve_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_STAGE_BYTES_WIDE));
    CUDA_TRY(h, cudaFuncSetAttribute(lm_solve_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_STAGE_BYTES_PAIR));
    CUDA_TRY(h, cudaDeviceGetAttribute(&sm_count[h->dev], cudaDevAttrMultiProcessorCount, h->dev));
    solve_attr_set[h->dev] = true;
  }
  { LAUNCH(h, "lm_solve");
    if (B <= sm_count[h->dev])
      lm_solve_kernel<256, 1><<<B, 256, LM_STAGE_BYTES_WIDE, s>>>(h->lm_edge, cc, h->lm_plane, cs + co, h->lm_n, guard_dev, h->lm_params, h->m2o,
          h->o2l, h->m2l, h->lm_report, h->lm_trace, h->lm_trace_n, h->lm_trace_cap, h->P.lm_outer_iters, h->P.lm_max_iters, h->P.huber_delta,
          write_pose ? h->d_pose : nullptr, h->o2l_lo[1 - h->cur], (int)(LM_STAGE_BYTES_WIDE / sizeof(double)));
    else
      lm_solve_kernel<128, 2><<<B, 128, LM_STAGE_BYTES_PAIR, s>>>(h->lm_edge, cc, h->lm_plane, cs + co, h->lm_n, guard_dev, h->lm_params, h->m2o,
          h->o2l, h->m2l, h->lm_report, h->lm_trace, h->lm_trace_n, h->lm_trace_cap, h->P.lm_outer_iters, h->P.lm_max_iters, h->P.huber_delta,
          write_pose ? h->d_pose : nullptr, h->o2l_lo[1 - h->cur], (int)(LM_STAGE_BYTES_PAIR / sizeof(double))); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

// downsampled-cloud / residual buffers, sized from the largest possible inputs
static int lm_ensure_ds_buffers(AlegoHandle *h, int need_c, int need_s, int need_o) {
  const int B = h->B;
  int want_c = std::max(std::max(h->R * 120, h->lm_cap_c), need_c);
  int want_s = std::max(std::max(h->RC, h->lm_cap_s), need_s);
  int want_o = std::max(std::max(h->out_cap, h->lm_cap_o), need_o);
  if (h->lm_corner_ds && want_c <= h->ds_cap_c && want_s <= h->ds_cap_s && want_o <= h->ds_cap_o) return ALEGO_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  ++h->graph_epoch;  // captured graphs hold the old pointers
  cudaFree(h->lm_corner_ds); cudaFree(h->lm_surf_ds); cudaFree(h->lm_outlier_ds); cudaFree(h->lm_surf_total);
  cudaFree(h->lm_surf_total_ds); cudaFree(h->lm_edge); cudaFree(h->lm_plane); cudaFree(h->vox_sort);
  cudaFree(h->lm_nn_c); cudaFree(h->lm_nn_s);
  h->ds_cap_c = want_c; h->ds_cap_s = want_s; h->ds_cap_o = want_o;
  CUDA_TRY(h, cudaMalloc(&h->lm_corner_ds, (size_t)B * want_c * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->lm_surf_ds, (size_t)B * want_s * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->lm_outlier_ds, (size_t)B * want_o * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->lm_surf_total, (size_t)B * (want_s + want_o) * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->lm_surf_total_ds, (size_t)B * (want_s + want_o) * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->lm_edge, (size_t)B * want_c * 10 * sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->lm_plane, (size_t)B * (want_s + want_o) * 8 * sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->lm_nn_c, (size_t)B * want_c * 5 * sizeof(int)));
  CUDA_TRY(h, cudaMalloc(&h->lm_nn_s, (size_t)B * (want_s + want_o) * 5 * sizeof(int)));
  const size_t sort_elems = (size_t)B * 2 * ((size_t)want_c + (size_t)(want_s + want_o) + (size_t)want_o);  // ping-pong per cloud
  CUDA_TRY(h, cudaMalloc(&h->vox_sort, sort_elems * sizeof(u64)));
  return ALEGO_OK;
}

// grow-only scratch slot k of the handle (cudaFree of the outgrown buffer waits for the work that still uses it)
static int scratch_get(AlegoHandle *h, int k, size_t bytes, void **out) {
  if (bytes > h->scratch_bytes[k]) {
    if (h->scratch[k]) cudaFree(h->scratch[k]);
    h->scratch[k] = nullptr;
    h->scratch_bytes[k] = 0;
    const size_t want = bytes + bytes / 4;
    CUDA_TRY(h, cudaMalloc(&h->scratch[k], want));
    h->scratch_bytes[k] = want;
  }
  *out = h->scratch[k];
  return ALEGO_OK;
}

int voxel_grid_host(AlegoHandle *h, const float *xyzi, int n, float leaf, float *out_xyzi, int *n_out) {
  cudaStream_t s = h->stream;
  if (n == 0) { *n_out = 0; return ALEGO_OK; }
  void *v_in, *v_out, *v_keys, *v_n;
  int rc;
  if ((rc = scratch_get(h, 0, (size_t)n * sizeof(float4), &v_in)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 1, (size_t)2 * n * sizeof(u64), &v_keys)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 2, (size_t)n * sizeof(float4), &v_out)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 3, 64, &v_n)) != ALEGO_OK) return rc;
  float4 *d_in = static_cast<float4 *>(v_in), *d_out = static_cast<float4 *>(v_out);
  u64 *d_keys = static_cast<u64 *>(v_keys);
  int *d_n = static_cast<int *>(v_n);
  CUDA_TRY(h, cudaMemcpyAsync(d_in, xyzi, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaFuncSetAttribute(voxel_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LMV_SMEM_EXACT));
  { LAUNCH(h, "voxel_single"); voxel_single_kernel<<<1, LMV_THREADS, LMV_SMEM_EXACT, s>>>(d_in, n, leaf, d_out, d_keys, d_n); }
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaMemcpyAsync(n_out, d_n, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  if (out_xyzi && *n_out > 0) {
    CUDA_TRY(h, cudaMemcpyAsync(out_xyzi, d_out, (size_t)*n_out * sizeof(float4), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
  }
  return ALEGO_OK;
}

int lm_ensure_buffers(AlegoHandle *h, int need_c, int need_s, int need_o) { return lm_ensure_ds_buffers(h, need_c, need_s, need_o); }

// ---------------------------------------------------------------------------------------------------
// N1: local-map assembly — the cloud side of extractSurroundingKeyFrames (laserMapping.cpp:194-323).
namespace {
// pcl::transformPointCloud (PCL 1.8 transforms.hpp) on a concatenation of segments: point i of segment s is mapped by
// matrix s >> mat_shift (row-major 3x4 floats), x' = m00*x + m01*y + m02*z + m03 evaluated left to right in float,
// intensity copied (laserMapping.h:163-177).
__global__ void __launch_bounds__(256) lm_kf_transform_kernel(float4 *__restrict__ pts, int n, const int *__restrict__ seg_off, int n_seg,
                                                              const float *__restrict__ M, int mat_shift) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = n_seg;  // segment s with seg_off[s] <= i < seg_off[s+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    const float *m = M + (size_t)(lo >> mat_shift) * 12;
    const float4 p = pts[i];
    float4 o;
    o.x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    o.y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    o.z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    o.w = p.w;
    pts[i] = o;
  }
}
}  // namespace

// segs: n_seg host clouds (pointer, count) concatenated in order, transformed by M[s >> mat_shift], VoxelGrid(leaf) into
// dst[0..*] on the device, count into n_dst (device).  Scratch is pooled in the handle (grow-only); no synchronisation.
int lm_assemble_cloud(AlegoHandle *h, const float *const *seg_ptr, const int *seg_n, int n_seg, const float *M_host, int n_mat,
                      int mat_shift, float leaf, float4 *dst, int *n_dst) {
  cudaStream_t s = h->stream;
  std::vector<int> off(n_seg + 1, 0);
  for (int k = 0; k < n_seg; ++k) off[k + 1] = off[k] + seg_n[k];
  const int n = off[n_seg];
  if (n == 0) {
    CUDA_TRY(h, cudaMemsetAsync(n_dst, 0, sizeof(int), s));
    return ALEGO_OK;
  }
  void *v_in, *v_keys, *v_off, *v_M;
  int rc;
  if ((rc = scratch_get(h, 0, (size_t)n * sizeof(float4), &v_in)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 1, (size_t)2 * n * sizeof(u64), &v_keys)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 3, (size_t)(n_seg + 1) * sizeof(int), &v_off)) != ALEGO_OK) return rc;
  if ((rc = scratch_get(h, 4, (size_t)n_mat * 12 * sizeof(float), &v_M)) != ALEGO_OK) return rc;
  float4 *d_in = static_cast<float4 *>(v_in);
  u64 *d_keys = static_cast<u64 *>(v_keys);
  int *d_off = static_cast<int *>(v_off);
  float *d_M = static_cast<float *>(v_M);
  for (int k = 0; k < n_seg; ++k)
    if (seg_n[k] > 0)
      CUDA_TRY(h, cudaMemcpyAsync(d_in + off[k], seg_ptr[k], (size_t)seg_n[k] * sizeof(float4), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(d_off, off.data(), (size_t)(n_seg + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(d_M, M_host, (size_t)n_mat * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
  { LAUNCH(h, "lm_kf_transform");
    lm_kf_transform_kernel<<<std::min(div_up(n, 256), 2048), 256, 0, s>>>(d_in, n, d_off, n_seg, d_M, mat_shift); }
  CUDA_TRY(h, cudaFuncSetAttribute(voxel_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LMV_SMEM_EXACT));
  { LAUNCH(h, "voxel_single"); voxel_single_kernel<<<1, LMV_THREADS, LMV_SMEM_EXACT, s>>>(d_in, n, leaf, dst, d_keys, n_dst); }
  CUDA_TRY(h, cudaGetLastError());
  // No synchronisation: the host clouds are pageable memory (copied to the driver's staging buffers before cudaMemcpyAsync
  // returns), the offsets / matrices likewise, and the scratch stays with the handle; everything later is ordered by the stream.
  return ALEGO_OK;
}
