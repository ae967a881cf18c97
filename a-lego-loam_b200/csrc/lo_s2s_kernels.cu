// lo_s2s_kernels.cu — LaserOdometry scan-to-scan on the device: replaces src/laserOdometry.cpp:316-535 and
// transformToStart (:728-740).
//
// K10 lo_assoc<SURF>  : one 8-lane group per feature point — transformToStart, exact 1-NN in the last cloud through
//                       the hashed grid (kd-tree replacement), then the adjacent-ring walk (:342-396, :432-470)
//                       as a lane-parallel argmin over (distance, walk position); surf: only inside the azimuth window
//                       that can hold a closer point (rings are binned by azimuth in lo_less_flat_voxel)
// K11/K12 lo_solve    : one CTA per sequence — Corner/Surf residuals, Huber, 6x6 reduction, LM (solver.cuh);
//                       phase 1 = surf blocks only (:410-421), phase 2 = surf + corner blocks (:484-495) and
//                       the pose integration (:504-508)
#include <algorithm>

#include "common.cuh"
#include "grid.cuh"
#include "lo_kernels.cuh"
#include "solver.cuh"

namespace {

// One GROUP of G lanes (a power of two <= 32) serves one query: when the candidate sets are small a full warp per query
// mostly idles and every reduction / control instruction is paid per warp.  All shuffles use the group's own lane mask,
// groups of one warp diverge freely.
template <int G> __device__ __forceinline__ unsigned group_mask() {
  return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (threadIdx.x & 31 & ~(G - 1)));
}
template <int G> __device__ __forceinline__ int group_lane() { return threadIdx.x & (G - 1); }

struct Best {
  float d;
  int i;
};
__device__ __forceinline__ bool better(float d, int i, const Best &b) { return d < b.d || (d == b.d && i < b.i); }
template <int G> __device__ __forceinline__ Best group_min_best(Best v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    Best w;
    w.d = __shfl_xor_sync(gm, v.d, o);
    w.i = __shfl_xor_sync(gm, v.i, o);
    if (better(w.d, w.i, v)) v = w;
  }
  return v;
}

__device__ __forceinline__ void nearest_visit_cell(const int *__restrict__ cs, const float4 *__restrict__ sp, int hsh, float qx, float qy,
                                                   float qz, Best &best) {
  const int e = cs[hsh + 1];
  for (int t = cs[hsh]; t < e; ++t) {
    const float4 p = sp[t];
    const float d = l2_simple(qx, qy, qz, p);
    const int idx = __float_as_int(p.w) & GRID_INDEX_MASK;
    if (better(d, idx, best)) { best.d = d; best.i = idx; }
  }
}

// exact nearest neighbour of (qx,qy,qz) by one lane group (kd_*_last_->nearestKSearch(point_sel, 1, ...), :341,:431):
// the 3x3x3 cube of cells first, then shell by shell until the best distance is covered (or the gate radius is
// exhausted).  Candidates are ranked by (float distance, index).  Returns index -1 when nothing lies within max_shell.
template <int G> __device__ Best group_nearest(const GridIndex &g, int b, float qx, float qy, float qz, int max_shell, unsigned gm) {
  const int gl = group_lane<G>();
  const int T = g.table_size;
  const int *cs = g.cell_start + (size_t)b * GRID_TABLE_STRIDE(T);
  const float4 *sp = g.sorted + (size_t)b * g.cap;
  const float inv = 1.0f / g.cell;
  const int cx = grid_coord(qx, inv), cy = grid_coord(qy, inv), cz = grid_coord(qz, inv);
  Best best{3.402823466e+38f, 0x7fffffff};
#pragma unroll
  for (int c = gl; c < 27; c += G) {
    const int dz = c / 9 - 1, dy = (c % 9) / 3 - 1, dx = c % 3 - 1;
    nearest_visit_cell(cs, sp, grid_hash(cx + dx, cy + dy, cz + dz, T), qx, qy, qz, best);
  }
  best = group_min_best<G>(best, gm);
  for (int r = 2; r <= max_shell; ++r) {
    const float covered = (float)(r - 1) * g.cell;
    if (best.i != 0x7fffffff && best.d <= covered * covered) break;
    const int side = 2 * r + 1, total = side * side * side;
    for (int c = gl; c < total; c += G) {
      const int dz = c / (side * side) - r, rem = c % (side * side), dy = rem / side - r, dx = rem % side - r;
      if (max(abs(dx), max(abs(dy), abs(dz))) != r) continue;  // only the new shell
      nearest_visit_cell(cs, sp, grid_hash(cx + dx, cy + dy, cz + dz, T), qx, qy, qz, best);
    }
    best = group_min_best<G>(best, gm);
  }
  if (best.i == 0x7fffffff) best.i = -1;
  return best;
}

struct WalkMin {
  double d;
  int pos;  // position in the sequential walk order (forward sweep first, then backward)
  int k;
};
template <int G> __device__ __forceinline__ WalkMin group_min_walk(WalkMin v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    WalkMin w;
    w.d = __shfl_xor_sync(gm, v.d, o);
    w.pos = __shfl_xor_sync(gm, v.pos, o);
    w.k = __shfl_xor_sync(gm, v.k, o);
    if (w.d < v.d || (w.d == v.d && w.pos < v.pos)) v = w;
  }
  return v;
}
__device__ __forceinline__ double sqdist_walk(const float4 &p, float sx, float sy, float sz) {
  // pow(float - float, 2) summed in double (:354, :445)
  const double dx = p.x - sx, dy = p.y - sy, dz = p.z - sz;
  return dx * dx + dy * dy + dz * dz;
}

struct WalkCtx {  // one query of the adjacent-ring walk
  float sx, sy, sz;
  int closest, nf;  // nf = number of forward-sweep positions (walk position of a backward candidate = nf + distance)
  __device__ __forceinline__ int pos(int k) const { return k > closest ? k - closest - 1 : nf + closest - 1 - k; }
};
__device__ __forceinline__ void walk_offer(WalkMin &m, double pd, int pos, int k) {
  if (pd < m.d || (pd == m.d && pos < m.pos)) { m.d = pd; m.pos = pos; m.k = k; }
}

// Azimuth window that holds every point closer than sqrt(d2) to a query at horizontal range rho and azimuth a_q:
// dist(p, q) >= rho * sin|daz| for |daz| <= 90 degrees, so |daz| < asin(D / rho) <= s + 0.571 s^3 (s = D / rho <= 1).
// Margins cover fast_atan2 (1e-6 rad) on both sides and the float evaluation.  full: the bound says nothing (D ~ rho).
__device__ __forceinline__ void az_window(double d2, float rho, float &w, bool &full) {
  const float D = sqrtf((float)d2) * 1.0001f + 1e-4f;
  const float s = D / rho;
  full = !(s < 0.99f);
  w = s + 0.571f * s * s * s + 2e-4f;
}
// bins [b0, b0 + nb) (wrapping) of the window around azimuth a_q
__device__ __forceinline__ void az_window_bins(float a_q, float w, bool full, int &b0, int &nb) {
  b0 = 0;
  nb = AZ_BINS;
  if (!full) {
    const int u0 = az_bin_unwrapped(a_q - w), u1 = az_bin_unwrapped(a_q + w);
    nb = min(u1 - u0 + 1, AZ_BINS);
    b0 = u0 & (AZ_BINS - 1);
  }
}

// candidates of the ring starting at ring_begin inside the bins, offered to the lane's private minimum
template <int G>
__device__ __forceinline__ void az_scan_ring(const float4 *__restrict__ az, const int *__restrict__ off, int ring_begin, int b0, int nb,
                                             const WalkCtx &c, WalkMin &m, int gl) {
  const int e = b0 + nb;  // exclusive end in unwrapped bins (< 2*AZ_BINS): at most two contiguous index ranges
  const int s0 = off[b0], e0 = off[min(e, AZ_BINS)];
  const int e1 = e > AZ_BINS ? off[e - AZ_BINS] : 0;
  const float4 *ring = az + ring_begin;
  for (int t = s0 + gl; t < e0; t += G) {
    const float4 p = ring[t];
    const int k = ring_begin + __float_as_int(p.w);
    if (k != c.closest) walk_offer(m, sqdist_walk(p, c.sx, c.sy, c.sz), c.pos(k), k);
  }
  for (int t = gl; t < e1; t += G) {
    const float4 p = ring[t];
    const int k = ring_begin + __float_as_int(p.w);
    if (k != c.closest) walk_offer(m, sqdist_walk(p, c.sx, c.sy, c.sz), c.pos(k), k);
  }
}

// (t_w_cur_, r_w_cur_) after this sweep, kept per buffer parity: LaserMapping of sweep t reads it while LaserOdometry of
// sweep t+1 already integrates the next step (laserMapping.cpp:154-164 receives it as the /odom/lidar message)
__global__ void lo_snapshot_kernel(const double *__restrict__ t_w, const double *__restrict__ r_w, Pose *__restrict__ snap, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int q = 0; q < 3; ++q) snap[b].t[q] = t_w[b * 3 + q];
  for (int q = 0; q < 9; ++q) snap[b].R[q] = r_w[b * 9 + q];
}

// params_ -> (R, t) of transformToStart, once per sequence instead of six double sin/cos per query
__global__ void lo_pose_kernel(const double *__restrict__ lo_params, Pose *__restrict__ lo_pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double *x = lo_params + b * 6;
  const PoseTrig T(x);
  for (int q = 0; q < 9; ++q) lo_pose[b].R[q] = T.R[q];
  for (int q = 0; q < 3; ++q) lo_pose[b].t[q] = x[q];
}

// One GROUP of G lanes serves one query (surf: 8 — a few cells and a few dozen ring points; corner: 32 — the walk covers
// up to 480 points)
#define ASSOC_THREADS 256
#define LO_G_SURF 8
#define LO_G_CORNER 32
template <bool SURF, int G>
__global__ void __launch_bounds__(ASSOC_THREADS)
lo_assoc_kernel(const float4 *__restrict__ feat, int feat_stride, const int *__restrict__ n_feat, const float4 *__restrict__ last,
                size_t last_stride, const int *__restrict__ ring_off, GridIndex g, const Pose *__restrict__ lo_pose,
                const int *__restrict__ lo_init, float *__restrict__ res, int *__restrict__ corr, int R, double gate,
                const float4 *__restrict__ az_pts, const int *__restrict__ az_off) {
  const int b = blockIdx.y, gl = group_lane<G>();
  const unsigned gm = group_mask<G>();
  const int q = blockIdx.x * (ASSOC_THREADS / G) + threadIdx.x / G;
  const int nq = n_feat[b * 4 + (SURF ? 2 : 0)];
  if (q >= nq || !lo_init[b]) return;
  const float4 cp = feat[(size_t)b * feat_stride + q];
  const Pose &T = lo_pose[b];  // Rz*Ry*Rx and translation of params_, evaluated once per sequence (lo_pose_kernel)
  // transformToStart (:728-740): double rotation + translation, stored as float
  const float sx = (float)(T.R[0] * cp.x + T.R[1] * cp.y + T.R[2] * cp.z + T.t[0]);
  const float sy = (float)(T.R[3] * cp.x + T.R[4] * cp.y + T.R[5] * cp.z + T.t[1]);
  const float sz = (float)(T.R[6] * cp.x + T.R[7] * cp.y + T.R[8] * cp.z + T.t[2]);
  const int CW = SURF ? 4 : 3;
  int *co = corr + ((size_t)b * feat_stride + q) * CW;
  const float4 *L = last + (size_t)b * last_stride;
  const int *ro = ring_off + b * (R + 1);
  const int n_last = ro[R];
  const int max_shell = (int)ceilf(sqrtf((float)gate) / g.cell) + 1;
  Best nn{0.f, -1};
  if (n_last > 0) nn = group_nearest<G>(g, b, sx, sy, sz, max_shell, gm);
  if (nn.i < 0 || !((double)nn.d < gate)) {  // search_dist[0] < nearest_feature_dist (:343, :433)
    if (gl == 0) co[1] = -1;
    return;
  }
  const int closest = nn.i;
  const int cs = (int)L[closest].w;  // ring id = int(intensity) (:347, :436)
  // rings cs-2 .. cs+2 take part (break at int(intensity) > cs+2.5 / < cs-2.5, resp. > cs+2 / < cs-2).  The cloud is ring
  // ordered, so the sequential walks decompose into contiguous index ranges: same ring above / below `closest` (surf:
  // min_idx2) and the neighbouring rings above / below (surf: min_idx3, corner: min_idx2).  The reference keeps the FIRST
  // candidate of minimal distance in walk order (strict `point_dist < min_dist`): candidates are ranked by
  // (distance, walk position).
  const int fwd_end = ro[min(cs + 3, R)];
  const int bwd_begin = ro[max(cs - 2, 0)];
  const int same_lo = ro[min(max(cs, 0), R)], same_hi = ro[min(max(cs + 1, 0), R)];
  const int nf = max(fwd_end - (closest + 1), 0);
  const WalkCtx c{sx, sy, sz, closest, nf};
  WalkMin m2{gate, 0x7fffffff, -1}, m3{gate, 0x7fffffff, -1};
  if (SURF) {
    // The walk visits every point of up to five rings (~1500 candidates).  Only points inside the azimuth window of the
    // current best distance can win, and the rings are indexed by azimuth bin (lo_less_flat_voxel): round 1 looks inside
    // the 1 m window; a class whose best stays above that is searched again inside the window of what round 1 established
    // (the gate when it found nothing).  Same (distance, walk position) minimum as the full walk.
    const float a_q = az_angle(sx, sy), rho = sqrtf(sx * sx + sy * sy);
    const float4 *AZ = az_pts + (size_t)b * last_stride;
    const int *AO = az_off + (size_t)b * R * (AZ_BINS + 1);
    const double d_round1 = fmin(1.0, gate);
    double bound2 = d_round1, bound3 = d_round1;
    for (int round = 0; round < 2; ++round) {
      WalkMin l2{gate, 0x7fffffff, -1}, l3{gate, 0x7fffffff, -1};
      float w;
      bool full;
      int b0 = 0, nb = AZ_BINS;
      if (bound2 > 0) {
        az_window(bound2, rho, w, full);
        az_window_bins(a_q, w, full, b0, nb);
        if (cs >= 0 && cs < R) az_scan_ring<G>(AZ, AO + cs * (AZ_BINS + 1), same_lo, b0, nb, c, l2, gl);
        m2 = group_min_walk<G>(l2, gm);
      }
      if (bound3 > 0) {
        if (bound3 != bound2) {
          az_window(bound3, rho, w, full);
          az_window_bins(a_q, w, full, b0, nb);
        }
        for (int r = max(cs - 2, 0); r <= min(cs + 2, R - 1); ++r)
          if (r != cs) az_scan_ring<G>(AZ, AO + r * (AZ_BINS + 1), ro[r], b0, nb, c, l3, gl);
        m3 = group_min_walk<G>(l3, gm);
      }
      // settled: the best lies strictly inside the searched radius (everything outside the window is farther)
      bound2 = (bound2 > 0 && !(m2.d < bound2 * 0.9999)) ? (round == 0 ? m2.d : 0.0) : 0.0;
      bound3 = (bound3 > 0 && !(m3.d < bound3 * 0.9999)) ? (round == 0 ? m3.d : 0.0) : 0.0;
      if (!(bound2 > 0) && !(bound3 > 0)) break;
    }
  } else {
    WalkMin l2{gate, 0x7fffffff, -1};
    for (int k = max(closest + 1, same_hi) + gl; k < fwd_end; k += G)  // forward, rings above (:439-454)
      walk_offer(l2, sqdist_walk(L[k], sx, sy, sz), c.pos(k), k);
    for (int k = min(closest - 1, same_lo - 1) - gl; k >= bwd_begin; k -= G)  // backward, rings below (:455-470)
      walk_offer(l2, sqdist_walk(L[k], sx, sy, sz), c.pos(k), k);
    m2 = group_min_walk<G>(l2, gm);
  }
  if (gl == 0) {
    const bool ok = m2.k >= 0 && (!SURF || m3.k >= 0);
    co[0] = q;
    co[1] = ok ? closest : -1;
    co[2] = m2.k;
    if (SURF) co[3] = m3.k;
    if (ok) {
      float *r = res + ((size_t)b * feat_stride + q) * (SURF ? 12 : 9);
      const float4 pj = L[closest], pl = L[m2.k];
      r[0] = cp.x; r[1] = cp.y; r[2] = cp.z;
      r[3] = pj.x; r[4] = pj.y; r[5] = pj.z;
      r[6] = pl.x; r[7] = pl.y; r[8] = pl.z;
      if (SURF) {
        const float4 pm = L[m3.k];
        r[9] = pm.x; r[10] = pm.y; r[11] = pm.z;
      }
    }
  }
}

// residual set of one sequence: surf slots first, then corner slots (the order blocks were added to the
// ceres::Problem, :403,:477)
struct LoResidSet {
  const float *surf;
  const int *surf_corr;
  int n_surf_slots;
  const float *corner;
  const int *corner_corr;
  int n_corner_slots;
  __device__ int slots() const { return n_surf_slots + n_corner_slots; }
  __device__ bool load(int i, int &kind, double cp[3], double a[3], double b[3], double c[3], double &d) const {
    d = 0;
    if (i < n_surf_slots) {
      if (surf_corr[i * 4 + 1] < 0) return false;
      const float *r = surf + (size_t)i * 12;
      kind = 1;
      for (int q = 0; q < 3; ++q) { cp[q] = r[q]; a[q] = r[3 + q]; b[q] = r[6 + q]; c[q] = r[9 + q]; }
      return true;
    }
    i -= n_surf_slots;
    if (corner_corr[i * 3 + 1] < 0) return false;
    const float *r = corner + (size_t)i * 9;
    kind = 0;
    for (int q = 0; q < 3; ++q) { cp[q] = r[q]; a[q] = r[3 + q]; b[q] = r[6 + q]; c[q] = 0; }
    return true;
  }
};

// the same residuals, compacted into shared memory by stage_blocks (surf blocks of 12 floats, then corner blocks of 9)
struct LoStagedSet {
  const float *surf;
  int n_surf;
  const float *corner;
  int n_corner;
  __device__ int slots() const { return n_surf + n_corner; }
  __device__ bool load(int i, int &kind, double cp[3], double a[3], double b[3], double c[3], double &d) const {
    d = 0;
    if (i < n_surf) {
      const float *r = surf + (size_t)i * 12;
      kind = 1;
      for (int q = 0; q < 3; ++q) { cp[q] = r[q]; a[q] = r[3 + q]; b[q] = r[6 + q]; c[q] = r[9 + q]; }
      return true;
    }
    const float *r = corner + (size_t)(i - n_surf) * 9;
    kind = 0;
    for (int q = 0; q < 3; ++q) { cp[q] = r[q]; a[q] = r[3 + q]; b[q] = r[6 + q]; c[q] = 0; }
    return true;
  }
};

// 128 threads with the full register budget: two sequences per SM solve side by side, so a batch of 256 is one wave (a
// scan-to-scan problem has a few hundred residual blocks; what costs is the latency of the serial trust-region logic, not the
// sweep over the blocks — capping the registers at 128 for four CTAs per SM spilled and was slower)
#define LO_SOLVE_THREADS 128
__global__ void __launch_bounds__(LO_SOLVE_THREADS, 2)
lo_solve_kernel(int phase, const float *__restrict__ surf_res, const int *__restrict__ surf_corr, const float *__restrict__ corner_res,
                const int *__restrict__ corner_corr, const int *__restrict__ n_feat, double *lo_params, double *t_w, double *r_w,
                int *lo_init, AlegoSolveReport *report, double *trace, int *trace_n, int trace_cap, int R, int surf_iters,
                int corner_iters, double huber_a, int stage_floats) {
  extern __shared__ __align__(16) float lo_stage[];
  const int b = blockIdx.x;
  __shared__ LmShared sh;
  __shared__ int s_cnt[2];
  __shared__ int s_red[34];
  AlegoSolveReport *rep = report + b;
  // block-uniform decision: thread 0 stores lo_init[b] = 1 below, so every thread must branch on the value read BEFORE that store
  __shared__ int s_init;
  if (threadIdx.x == 0) s_init = lo_init[b];
  __syncthreads();
  if (!s_init) {  // first frame of the sequence: only the targets are initialised (:316-324)
    if (phase == 2 && threadIdx.x == 0) {
      lo_init[b] = 1;
      rep->status = ALEGO_OK; rep->n_corner = 0; rep->n_surf = 0; rep->iterations = 0; rep->initial_cost = 0; rep->final_cost = 0;
      trace_n[b] = 0;
    }
    return;
  }
  const int sslots = n_feat[b * 4 + 2], cslots = n_feat[b * 4 + 0];
  const int sstride = R * 24, cstride = R * 12;
  const int *sc = surf_corr + (size_t)b * sstride * 4;
  const int *cc = corner_corr + (size_t)b * cstride * 3;
  // count correspondences
  int ns = 0, nc = 0;
  for (int i = threadIdx.x; i < sslots; i += blockDim.x) ns += sc[i * 4 + 1] >= 0;
  if (phase == 2)
    for (int i = threadIdx.x; i < cslots; i += blockDim.x) nc += cc[i * 3 + 1] >= 0;
  int tot;
  block_excl_scan(ns, s_red, &tot);
  if (threadIdx.x == 0) s_cnt[0] = tot;
  block_excl_scan(nc, s_red, &tot);
  if (threadIdx.x == 0) s_cnt[1] = tot;
  __syncthreads();
  ns = s_cnt[0];
  nc = s_cnt[1];
  LoResidSet rs;
  rs.surf = surf_res + (size_t)b * sstride * 12;
  rs.surf_corr = sc;
  rs.n_surf_slots = sslots;
  rs.corner = corner_res + (size_t)b * cstride * 9;
  rs.corner_corr = cc;
  rs.n_corner_slots = phase == 2 ? cslots : 0;
  double *x = lo_params + b * 6;
  double *tr = trace ? trace + (size_t)b * trace_cap * 7 : nullptr;
  // staged copy of the valid blocks (always fits for the supported ring counts; otherwise the global arrays are swept)
  LoStagedSet ss{lo_stage, 0, lo_stage, 0};
  const bool staged = ns * 12 + nc * 9 <= stage_floats && (phase == 1 ? ns >= 10 : nc >= 10);
  if (staged) {
    const float *gs = rs.surf, *gc = rs.corner;
    ss.n_surf = stage_blocks<float, 12>(sslots, lo_stage, [&](int i) { return sc[i * 4 + 1] >= 0; },
                                        [&](int i, float *d) { for (int q = 0; q < 12; ++q) d[q] = gs[(size_t)i * 12 + q]; }, s_red);
    ss.corner = lo_stage + (size_t)ss.n_surf * 12;
    if (phase == 2)
      ss.n_corner = stage_blocks<float, 9>(cslots, lo_stage + (size_t)ss.n_surf * 12, [&](int i) { return cc[i * 3 + 1] >= 0; },
                                           [&](int i, float *d) { for (int q = 0; q < 9; ++q) d[q] = gc[(size_t)i * 9 + q]; }, s_red);
  }
  if (phase == 1) {
    if (threadIdx.x == 0) {
      trace_n[b] = 0;
      rep->n_surf = ns; rep->n_corner = 0; rep->iterations = 0; rep->initial_cost = 0; rep->final_cost = 0;
      rep->status = ns >= 10 ? ALEGO_OK : ALEGO_FEW_FEATURES;
    }
    __syncthreads();
    if (ns >= 10) {  // (:410)
      const LmResult r = staged ? block_lm_solve(ss, x, surf_iters, huber_a, &sh, tr, trace_n + b, trace_cap)
                                : block_lm_solve(rs, x, surf_iters, huber_a, &sh, tr, trace_n + b, trace_cap);
      if (threadIdx.x == 0) { rep->iterations = r.iterations; rep->initial_cost = r.initial_cost; rep->final_cost = r.final_cost; }
    }
    return;
  }
  // phase 2
  if (nc >= 10) {  // (:484)
    const LmResult r = staged ? block_lm_solve(ss, x, corner_iters, huber_a, &sh, tr, trace_n + b, trace_cap)
                              : block_lm_solve(rs, x, corner_iters, huber_a, &sh, tr, trace_n + b, trace_cap);
    if (threadIdx.x == 0) {
      if (rep->iterations == 0 && ns < 10) rep->initial_cost = r.initial_cost;
      rep->iterations += r.iterations;
      rep->final_cost = r.final_cost;
    }
  }
  if (threadIdx.x == 0) {
    rep->n_corner = nc;
    if (nc < 10) rep->status = ALEGO_FEW_FEATURES;
    // pose integration (:504-508): t_w += R_w * t ; R_w *= Rz(yaw)
    const double cy = cos(x[5]), sy = sin(x[5]);
    double *Rw = r_w + b * 9, *tw = t_w + b * 3;
    const double t0 = x[0], t1 = x[1], t2 = x[2];
    tw[0] += Rw[0] * t0 + Rw[1] * t1 + Rw[2] * t2;
    tw[1] += Rw[3] * t0 + Rw[4] * t1 + Rw[5] * t2;
    tw[2] += Rw[6] * t0 + Rw[7] * t1 + Rw[8] * t2;
    const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double N[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) N[r * 3 + c] = Rw[r * 3] * Rz[c] + Rw[r * 3 + 1] * Rz[3 + c] + Rw[r * 3 + 2] * Rz[6 + c];
    for (int q = 0; q < 9; ++q) Rw[q] = N[q];
  }
}

}  // namespace

int lo_scan2scan_device(AlegoHandle *h) {
  const int B = h->B, R = h->R, RC = h->RC;
  cudaStream_t s = h->stream;
  const int cur = h->cur, prev = 1 - cur;
  const double gate = h->P.nearest_feature_dist, hub = h->P.huber_delta;
  // shared-memory staging of the valid residual blocks (surf 12 floats, corner 9 floats), capped at 48 KB so that four CTAs share
  // an SM (a sweep with more correspondences than that is solved from global memory)
  const size_t lo_stage_bytes = std::min<size_t>((size_t)R * (24 * 12 + 12 * 9) * sizeof(float), 48 * 1024);
  static bool lo_attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!lo_attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(lo_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    lo_attr_set[h->dev] = true;
  }
  { LAUNCH(h, "lo_pose"); lo_pose_kernel<<<div_up(B, 128), 128, 0, s>>>(h->lo_params, h->lo_pose, B); }
  { LAUNCH(h, "lo_assoc_surf");
    lo_assoc_kernel<true, LO_G_SURF><<<dim3(div_up(R * 24, ASSOC_THREADS / LO_G_SURF), B), ASSOC_THREADS, 0, s>>>(
        h->flat, R * 24, h->n_feat, h->less_flat[prev], (size_t)RC, h->lf_ring_off[prev], h->g_surf_last, h->lo_pose, h->lo_init,
        h->lo_surf_res, h->lo_surf_corr, R, gate, h->az_pts[prev], h->az_off[prev]); }
  { LAUNCH(h, "lo_solve_surf");
    lo_solve_kernel<<<B, LO_SOLVE_THREADS, lo_stage_bytes, s>>>(1, h->lo_surf_res, h->lo_surf_corr, h->lo_corner_res, h->lo_corner_corr, h->n_feat,
                                      h->lo_params, h->t_w, h->r_w, h->lo_init, h->lo_report, h->lo_trace, h->lo_trace_n,
                                      h->lo_trace_cap, R, h->P.lo_surf_iters, h->P.lo_corner_iters, hub, (int)(lo_stage_bytes / sizeof(float))); }
  { LAUNCH(h, "lo_pose"); lo_pose_kernel<<<div_up(B, 128), 128, 0, s>>>(h->lo_params, h->lo_pose, B); }
  { LAUNCH(h, "lo_assoc_corner");
    lo_assoc_kernel<false, LO_G_CORNER><<<dim3(div_up(R * 12, ASSOC_THREADS / LO_G_CORNER), B), ASSOC_THREADS, 0, s>>>(
        h->sharp, R * 12, h->n_feat, h->less_sharp[prev], (size_t)R * 120, h->ls_ring_off[prev], h->g_corner_last, h->lo_pose,
        h->lo_init, h->lo_corner_res, h->lo_corner_corr, R, gate, nullptr, nullptr); }
  { LAUNCH(h, "lo_solve_corner");
    lo_solve_kernel<<<B, LO_SOLVE_THREADS, lo_stage_bytes, s>>>(2, h->lo_surf_res, h->lo_surf_corr, h->lo_corner_res, h->lo_corner_corr, h->n_feat,
                                      h->lo_params, h->t_w, h->r_w, h->lo_init, h->lo_report, h->lo_trace, h->lo_trace_n,
                                      h->lo_trace_cap, R, h->P.lo_surf_iters, h->P.lo_corner_iters, hub, (int)(lo_stage_bytes / sizeof(float))); }
  { LAUNCH(h, "lo_snapshot"); lo_snapshot_kernel<<<div_up(B, 128), 128, 0, s>>>(h->t_w, h->r_w, h->o2l_lo[cur], B); }
  // the current clouds become the targets of the next sweep (:531-534): index them, then flip the buffers
  int rc = grid_build(h, &h->g_surf_last, h->less_flat[cur], (size_t)RC, h->lf_ring_off[cur] + R, R + 1, "surf_last");
  if (rc != ALEGO_OK) return rc;
  rc = grid_build(h, &h->g_corner_last, h->less_sharp[cur], (size_t)R * 120, h->ls_ring_off[cur] + R, R + 1, "corner_last");
  if (rc != ALEGO_OK) return rc;
  h->cur = prev;
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
