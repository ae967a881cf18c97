// lo_s2s_kernels.cu — LaserOdometry scan-to-scan on the device: replaces src/laserOdometry.cpp:316-535 and
// transformToStart (:728-740).
//
// K10 lo_assoc<SURF>  : one warp per feature point — transformToStart, exact 1-NN in the last cloud through
//                       the hashed grid (kd-tree replacement), then the adjacent-ring walk (:342-396, :432-470)
//                       as a lane-parallel argmin that keeps the sequential walk's tie rule
// K11/K12 lo_solve    : one CTA per sequence — Corner/Surf residuals, Huber, 6x6 reduction, LM (solver.cuh);
//                       phase 1 = surf blocks only (:410-421), phase 2 = surf + corner blocks (:484-495) and
//                       the pose integration (:504-508)
#include "common.cuh"
#include "grid.cuh"
#include "lo_kernels.cuh"
#include "solver.cuh"

namespace {

struct Best {
  float d;
  int i;
};
__device__ __forceinline__ bool better(float d, int i, const Best &b) { return d < b.d || (d == b.d && i < b.i); }
__device__ __forceinline__ Best warp_min_best(Best v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best w;
    w.d = __shfl_xor_sync(0xffffffffu, v.d, o);
    w.i = __shfl_xor_sync(0xffffffffu, v.i, o);
    if (better(w.d, w.i, v)) v = w;
  }
  return v;
}

// exact nearest neighbour of (qx,qy,qz) by one warp: grows the searched cube shell by shell until the best
// distance is covered (or the gate radius is exhausted).  Returns index -1 when nothing lies within max_shell.
__device__ Best warp_nearest(const GridIndex &g, int b, float qx, float qy, float qz, int max_shell) {
  const int lane = threadIdx.x & 31;
  const int T = g.table_size;
  const int *cs = g.cell_start + (size_t)b * (T + 4);
  const float4 *sp = g.sorted + (size_t)b * g.cap;
  const float inv = 1.0f / g.cell;
  const int cx = grid_coord(qx, inv), cy = grid_coord(qy, inv), cz = grid_coord(qz, inv);
  Best best{3.402823466e+38f, 0x7fffffff};
  for (int r = 1; r <= max(max_shell, 1); ++r) {
    const int side = 2 * r + 1, total = side * side * side;
    for (int c = lane; c < total; c += 32) {
      const int dz = c / (side * side) - r, rem = c % (side * side), dy = rem / side - r, dx = rem % side - r;
      if (r > 1 && max(abs(dx), max(abs(dy), abs(dz))) != r) continue;  // only the new shell (the first pass takes the whole 3x3x3 cube)
      const int hsh = grid_hash(cx + dx, cy + dy, cz + dz, T);
      const int e = cs[hsh + 1];
      for (int t = cs[hsh]; t < e; ++t) {
        const float4 p = sp[t];
        const float d = l2_simple(qx, qy, qz, p);
        const int idx = __float_as_int(p.w) & GRID_INDEX_MASK;
        if (better(d, idx, best)) { best.d = d; best.i = idx; }
      }
    }
    best = warp_min_best(best);
    const float covered = (float)r * g.cell;
    if (best.i != 0x7fffffff && best.d <= covered * covered) break;
  }
  if (best.i == 0x7fffffff) best.i = -1;
  return best;
}

struct WalkMin {
  double d;
  int pos;  // position in the sequential walk order (forward sweep first, then backward)
  int k;
};
__device__ __forceinline__ WalkMin warp_min_walk(WalkMin v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    WalkMin w;
    w.d = __shfl_xor_sync(0xffffffffu, v.d, o);
    w.pos = __shfl_xor_sync(0xffffffffu, v.pos, o);
    w.k = __shfl_xor_sync(0xffffffffu, v.k, o);
    if (w.d < v.d || (w.d == v.d && w.pos < v.pos)) v = w;
  }
  return v;
}
__device__ __forceinline__ double sqdist_walk(const float4 &p, float sx, float sy, float sz) {
  // pow(float - float, 2) summed in double (:354, :445)
  const double dx = p.x - sx, dy = p.y - sy, dz = p.z - sz;
  return dx * dx + dy * dy + dz * dz;
}

#define ASSOC_WARPS 8
template <bool SURF>
__global__ void __launch_bounds__(ASSOC_WARPS * 32)
lo_assoc_kernel(const float4 *__restrict__ feat, int feat_stride, const int *__restrict__ n_feat, const float4 *__restrict__ last,
                size_t last_stride, const int *__restrict__ ring_off, GridIndex g, const double *__restrict__ lo_params,
                const int *__restrict__ lo_init, float *__restrict__ res, int *__restrict__ corr, int R, double gate) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * ASSOC_WARPS + warp;
  const int nq = n_feat[b * 4 + (SURF ? 2 : 0)];
  if (q >= nq || !lo_init[b]) return;
  const float4 cp = feat[(size_t)b * feat_stride + q];
  const double *x = lo_params + b * 6;
  const PoseTrig T(x);
  // transformToStart (:728-740): double rotation + translation, stored as float
  const float sx = (float)(T.R[0] * cp.x + T.R[1] * cp.y + T.R[2] * cp.z + x[0]);
  const float sy = (float)(T.R[3] * cp.x + T.R[4] * cp.y + T.R[5] * cp.z + x[1]);
  const float sz = (float)(T.R[6] * cp.x + T.R[7] * cp.y + T.R[8] * cp.z + x[2]);
  const int CW = SURF ? 4 : 3;
  int *co = corr + ((size_t)b * feat_stride + q) * CW;
  const float4 *L = last + (size_t)b * last_stride;
  const int *ro = ring_off + b * (R + 1);
  const int n_last = ro[R];
  int max_shell = (int)ceilf(sqrtf((float)gate) / g.cell) + 1;
  Best nn{0.f, -1};
  if (n_last > 0) nn = warp_nearest(g, b, sx, sy, sz, max_shell);
  if (nn.i < 0 || !((double)nn.d < gate)) {  // search_dist[0] < nearest_feature_dist (:343, :433)
    if (lane == 0) co[1] = -1;
    return;
  }
  const int closest = nn.i;
  const int cs = (int)L[closest].w;  // ring id = int(intensity) (:347, :436)
  // rings cs-2 .. cs+2 take part (break at int(intensity) > cs+2.5 / < cs-2.5, resp. > cs+2 / < cs-2).  The cloud is ring
  // ordered, so the sequential walks decompose into contiguous index ranges: same ring above / below `closest` (surf:
  // min_idx2) and the neighbouring rings above / below (surf: min_idx3, corner: min_idx2).  A lane meets its candidates
  // in walk order, so the reference's strict `point_dist < min_dist` keeps the earliest of equal distances inside the
  // lane; the walk position only enters the cross-lane reduction.
  const int fwd_end = ro[min(cs + 3, R)];
  const int bwd_begin = ro[max(cs - 2, 0)];
  const int same_lo = ro[min(max(cs, 0), R)], same_hi = ro[min(max(cs + 1, 0), R)];
  const int nf = max(fwd_end - (closest + 1), 0);
  double d2 = gate, d3 = gate;
  int k2 = -1, k3 = -1;
  if (SURF) {
#pragma unroll 4
    for (int k = closest + 1 + lane; k < same_hi; k += 32) {  // forward, same ring (:348-371)
      const double pd = sqdist_walk(L[k], sx, sy, sz);
      if (pd < d2) { d2 = pd; k2 = k; }
    }
#pragma unroll 4
    for (int k = closest - 1 - lane; k >= same_lo; k -= 32) {  // backward, same ring (:372-395)
      const double pd = sqdist_walk(L[k], sx, sy, sz);
      if (pd < d2) { d2 = pd; k2 = k; }
    }
  }
  {
    double &dn = SURF ? d3 : d2;
    int &kn = SURF ? k3 : k2;
#pragma unroll 4
    for (int k = max(closest + 1, same_hi) + lane; k < fwd_end; k += 32) {  // forward, rings above (:439-454)
      const double pd = sqdist_walk(L[k], sx, sy, sz);
      if (pd < dn) { dn = pd; kn = k; }
    }
#pragma unroll 4
    for (int k = min(closest - 1, same_lo - 1) - lane; k >= bwd_begin; k -= 32) {  // backward, rings below (:455-470)
      const double pd = sqdist_walk(L[k], sx, sy, sz);
      if (pd < dn) { dn = pd; kn = k; }
    }
  }
  auto walk_pos = [&](int k) { return k < 0 ? 0x7fffffff : (k > closest ? k - closest - 1 : nf + closest - 1 - k); };
  WalkMin m2{d2, walk_pos(k2), k2}, m3{d3, walk_pos(k3), k3};
  m2 = warp_min_walk(m2);
  if (SURF) m3 = warp_min_walk(m3);
  if (lane == 0) {
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

__global__ void __launch_bounds__(256)
lo_solve_kernel(int phase, const float *__restrict__ surf_res, const int *__restrict__ surf_corr, const float *__restrict__ corner_res,
                const int *__restrict__ corner_corr, const int *__restrict__ n_feat, double *lo_params, double *t_w, double *r_w,
                int *lo_init, AlegoSolveReport *report, double *trace, int *trace_n, int trace_cap, int R, int surf_iters,
                int corner_iters, double huber_a) {
  const int b = blockIdx.x;
  __shared__ LmShared sh;
  __shared__ int s_cnt[2];
  __shared__ int s_red[34];
  AlegoSolveReport *rep = report + b;
  if (!lo_init[b]) {  // first frame of the sequence: only the targets are initialised (:316-324)
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
  if (phase == 1) {
    if (threadIdx.x == 0) {
      trace_n[b] = 0;
      rep->n_surf = ns; rep->n_corner = 0; rep->iterations = 0; rep->initial_cost = 0; rep->final_cost = 0;
      rep->status = ns >= 10 ? ALEGO_OK : ALEGO_FEW_FEATURES;
    }
    __syncthreads();
    if (ns >= 10) {  // (:410)
      const LmResult r = block_lm_solve(rs, x, surf_iters, huber_a, &sh, tr, trace_n + b, trace_cap);
      if (threadIdx.x == 0) { rep->iterations = r.iterations; rep->initial_cost = r.initial_cost; rep->final_cost = r.final_cost; }
    }
    return;
  }
  // phase 2
  if (nc >= 10) {  // (:484)
    const LmResult r = block_lm_solve(rs, x, corner_iters, huber_a, &sh, tr, trace_n + b, trace_cap);
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
  { LAUNCH(h, "lo_assoc_surf");
    lo_assoc_kernel<true><<<dim3(div_up(R * 24, ASSOC_WARPS), B), ASSOC_WARPS * 32, 0, s>>>(
        h->flat, R * 24, h->n_feat, h->less_flat[prev], (size_t)RC, h->lf_ring_off[prev], h->g_surf_last, h->lo_params, h->lo_init,
        h->lo_surf_res, h->lo_surf_corr, R, gate); }
  { LAUNCH(h, "lo_solve_surf");
    lo_solve_kernel<<<B, 256, 0, s>>>(1, h->lo_surf_res, h->lo_surf_corr, h->lo_corner_res, h->lo_corner_corr, h->n_feat,
                                      h->lo_params, h->t_w, h->r_w, h->lo_init, h->lo_report, h->lo_trace, h->lo_trace_n,
                                      h->lo_trace_cap, R, h->P.lo_surf_iters, h->P.lo_corner_iters, hub); }
  { LAUNCH(h, "lo_assoc_corner");
    lo_assoc_kernel<false><<<dim3(div_up(R * 12, ASSOC_WARPS), B), ASSOC_WARPS * 32, 0, s>>>(
        h->sharp, R * 12, h->n_feat, h->less_sharp[prev], (size_t)R * 120, h->ls_ring_off[prev], h->g_corner_last, h->lo_params,
        h->lo_init, h->lo_corner_res, h->lo_corner_corr, R, gate); }
  { LAUNCH(h, "lo_solve_corner");
    lo_solve_kernel<<<B, 256, 0, s>>>(2, h->lo_surf_res, h->lo_surf_corr, h->lo_corner_res, h->lo_corner_corr, h->n_feat,
                                      h->lo_params, h->t_w, h->r_w, h->lo_init, h->lo_report, h->lo_trace, h->lo_trace_n,
                                      h->lo_trace_cap, R, h->P.lo_surf_iters, h->P.lo_corner_iters, hub); }
  // the current clouds become the targets of the next sweep (:531-534): index them, then flip the buffers
  int rc = grid_build(h, &h->g_surf_last, h->less_flat[cur], (size_t)RC, h->lf_ring_off[cur] + R, R + 1, "surf_last");
  if (rc != ALEGO_OK) return rc;
  rc = grid_build(h, &h->g_corner_last, h->less_sharp[cur], (size_t)R * 120, h->ls_ring_off[cur] + R, R + 1, "corner_last");
  if (rc != ALEGO_OK) return rc;
  h->cur = prev;
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
