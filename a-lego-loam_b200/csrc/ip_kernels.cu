// ip_kernels.cu — ImageProjection on the device: replaces ImageProjection::pcCB + labelComponents
// (src/imageProjection.cpp:49-208, 210-316).  All kernels carry the sequence index in blockIdx.y so one
// launch serves the whole batch of independent sequences.
//
// K1a ip_project   : per input point row/col binning (:76-104), deterministic "last writer wins" by atomicMax
// K1b ip_gather    : per cell — organised cloud, range image, resets of the per-scan images (:24-33,:197-205)
// K2  ip_ground    : vertical-neighbour slope test (:106-132)
// K3  ccl_*        : connected components: per-ring runs by scan, vertical joins by lock-free union-find (:134-156, 210-280)
// K4/K5 ip_rowcount + ip_compact : feasibility (:282-315), raster-order label numbering, ring-major stream
//                    compaction into the cloud_info arrays (:158-191)
#include <math_constants.h>

#include "common.cuh"
#include "ip_kernels.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// row / column of one point.  The reference evaluates atan2f / hypotf (float overloads, SURVEY Appendix
// A.1) then scales in double and truncates.  Fast path: float intrinsics, accepted when the fractional
// cell coordinate is > 2e-3 away from an integer (their error is < 3e-4 cell); otherwise the slow path
// recomputes with correctly rounded float results obtained through double precision.
// ---------------------------------------------------------------------------------------------------

// Exact forms: what the reference evaluates (float atan2f / hypotf results, double scaling, :79-80, :87-88)
__device__ __noinline__ double exact_row_f(float x, float y, float z, const IpDev &P) {
  const float hyp = (float)sqrt((double)x * (double)x + (double)y * (double)y);
  const float va = (float)atan2((double)z, (double)hyp);
  return ((double)va * 180.0 / CUDART_PI + P.ang_bottom) / P.ang_res_y + 0.5;
}
__device__ __noinline__ double exact_col_f(float x, float y, const IpDev &P) {
  const float ha = (float)atan2((double)y, (double)x);
  return (((double)(-ha) + 2 * CUDART_PI) * 180.0 / CUDART_PI) / P.ang_res_x;
}

__device__ __forceinline__ bool project_rowcol(float x, float y, float z, const IpDev &P, int &row, int &col) {
  // ---- row (imageProjection.cpp:79-85).  Fast path: approximate float angle scaled by one precomputed double factor;
  // accepted only when the cell coordinate is farther from an integer than the fast path's error bound (P.eps_row /
  // P.eps_col, >= 2e-3 cell), otherwise the exact form decides.
  const float s2 = x * x + y * y;
  const float hyp = s2 * rsqrt_approx(fmaxf(s2, 1e-30f));
  float va = 0.f, ha = 0.f;
  const bool fast_row = s2 < 1e30f && fast_atan2(z, hyp, va), fast_col = fast_atan2(y, x, ha);
  // float cell coordinates: the float scaling adds < 1e-6 * |coordinate| to the angle error (covered by eps_row / eps_col)
  const float row_ff = __fmaf_rn(va, P.row_scale_f, P.row_off_f);
  const float col_ff = (6.283185307f - ha) * P.col_scale_f;
  const bool row_sure = fast_row && fabsf(row_ff - rintf(row_ff)) >= P.eps_row && fabsf(row_ff) < 1.0e6f;
  const bool col_sure = fast_col && fabsf(col_ff - rintf(col_ff)) >= P.eps_col;
  if (row_sure) {
    row = (int)row_ff;
  } else {
    const double row_f = exact_row_f(x, y, z, P);
    if (!(row_f > -1.0e6 && row_f < 1.0e6)) return false;
    row = (int)row_f;
  }
  if (row < 0 || row >= P.R) return false;
  // ---- column (:87-97)
  col = col_sure ? (int)col_ff : (int)exact_col_f(x, y, P);
  if (col >= P.C) col -= P.C;
  if (col < 0 || col >= P.C) return false;
  return true;
}

__device__ __forceinline__ bool finite3(const float4 &p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// Input point i of a sweep buffer holding `stride` floats per point: 4 = (x, y, z, intensity), one 16-byte load;
// 3 = (x, y, z) packed, 25 % fewer bytes over PCIe.  The intensity of the input is never read by the reference's path —
// pcCB overwrites it with row + col/10000 (imageProjection.cpp:101).
__device__ __forceinline__ float4 load_point(const float *__restrict__ sweep, int i, int stride) {
  if (stride == 4) return ldg_f4(reinterpret_cast<const float4 *>(sweep) + i);
  const float *p = sweep + (size_t)i * 3;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
}

// K1a.  A spinning sensor emits its points firing by firing: ring index fastest, azimuth slowest.  A warp that processed
// 32 CONSECUTIVE points would therefore scatter its 32 winner updates over 32 image rows (32 L2 sectors per instruction,
// measured as the limiter of this kernel).  Instead a CTA stages a tile of 32*R consecutive points in shared memory with
// coalesced 16-byte loads and lane l of a warp then takes point  l*R + j  of the tile: for ring-fastest input the lanes
// of one instruction hit consecutive columns of ONE row (4 sectors); for any other order the mapping is merely as
// scattered as the naive one.  The staging rows are padded to R+1 points so the transposed reads are conflict free.
__global__ void __launch_bounds__(256) ip_project_kernel(const float *__restrict__ raw, const int *__restrict__ n_pts,
                                                         int *__restrict__ winner, int Nmax, int stride, IpDev P) {
  extern __shared__ float4 s_tile[];  // stride 4: [32][R+1] points; stride 3: [32][3R+1] floats
  const int b = blockIdx.y;
  const int n = min(n_pts[b], Nmax);
  const float *src = raw + (size_t)b * Nmax * stride;
  int *win = winner + (size_t)b * P.RC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int tile_pts = 32 * P.R;
  float *sf = reinterpret_cast<float *>(s_tile);
  const int frow = 3 * P.R + 1;
  for (int tile = blockIdx.x * tile_pts; tile < n; tile += gridDim.x * tile_pts) {
    __syncthreads();
    // staging row l = points [tile + l*R, tile + (l+1)*R): contiguous in the sweep, so every warp load is coalesced
    for (int l = warp; l < 32; l += nwarp) {
      const int first = tile + l * P.R, cnt = min(P.R, n - first);
      if (stride == 4) {
        for (int j = lane; j < cnt; j += 32) s_tile[l * (P.R + 1) + j] = ldg_f4(reinterpret_cast<const float4 *>(src) + first + j);
      } else {
        const float *rf = src + (size_t)first * 3;
        for (int f = lane; f < cnt * 3; f += 32) sf[l * frow + f] = __ldg(rf + f);
      }
    }
    __syncthreads();
    for (int j = warp; j < P.R; j += nwarp) {
      const int i = tile + lane * P.R + j;
      if (i >= n) continue;
      float4 p;
      if (stride == 4) {
        p = s_tile[lane * (P.R + 1) + j];
      } else {
        const float *q = sf + lane * frow + 3 * j;
        p = make_float4(q[0], q[1], q[2], 0.f);
      }
      if (!finite3(p)) continue;  // pcl::removeNaNFromPointCloud (:59)
      int row, col;
      if (!project_rowcol(p.x, p.y, p.z, P, row, col)) continue;
      atomicMax(win + row * P.C + col, i);  // the later point of a cell wins (:103)
    }
  }
}

// range-angle criterion (:255-270): atan2(d2*sin(alpha), d1 - d2*cos(alpha)) > seg_theta
__device__ __forceinline__ bool seg_join(float ra, float rb, bool horizontal, const IpDev &P) {
  const float d1f = fmaxf(ra, rb), d2f = fminf(ra, rb);
  const float sf = horizontal ? (float)P.sin_x : (float)P.sin_y, cf = horizontal ? (float)P.cos_x : (float)P.cos_y;
  float af = 0.f;  // approximate angle (1e-6 rad) decides unless it is within 1e-4 rad of the threshold
  if (fast_atan2(d2f * sf, d1f - d2f * cf, af) && fabsf(af - (float)P.seg_theta) > 1e-4f) return af > (float)P.seg_theta;
  const double d1 = d1f, d2 = d2f;
  const double angle = atan2(d2 * (horizontal ? P.sin_x : P.sin_y), d1 - d2 * (horizontal ? P.cos_x : P.cos_y));
  return angle > P.seg_theta;
}

// cell flags written by ip_image, read by the connected-component kernels
#define CELL_VALID 1      // has a return and is not ground: a segmentation candidate (label_mat_ == 0, :134-143)
#define CELL_GROUND 2
#define CELL_JOIN_LEFT 4  // joined with the cell one column to the left (column 0: with column C-1, :241-248)
#define CELL_JOIN_DOWN 8  // joined with the cell one row up in the image (row + 1)
#define CELL_STRIP_ROOT 16  // set by ccl_strip: the cell is the root of its component inside its column strip

// K1b + K2 + the join tests of K3 in one pass over the image.  A CTA owns a strip of IMG_W columns x all rows, staged in
// shared memory together with the column to its left:
//   * per cell: the winner of ip_project -> organised cloud + range image (:99-103, fill values :24-33);
//   * per column: ground test of every vertical pair below ground_scan_id (:106-132) — a cell's ground flag depends only on
//     its own column, which the CTA holds completely;
//   * per cell: validity and the two range-angle join tests (:255-270) towards the left and the upper neighbour, kept as
//     one flag byte, so that the union-find kernels touch 1 byte per cell and evaluate no atan2.
#define IMG_W 32
__global__ void __launch_bounds__(256) ip_image_kernel(const float *__restrict__ raw, int stride, const int *__restrict__ winner,
                                                       float4 *__restrict__ cloud, float *__restrict__ range,
                                                       uint8_t *__restrict__ ground, uint8_t *__restrict__ flags, int Nmax, IpDev P) {
  extern __shared__ float4 s_pt[];  // [R][IMG_W + 1]: x, y, z, range (ALEGO_EMPTY_RANGE: no return); column 0 = left neighbour strip
  const int SW = IMG_W + 1;
  uint8_t *s_ground = reinterpret_cast<uint8_t *>(s_pt + (size_t)P.R * SW);  // [R][SW]
  const int b = blockIdx.y, col0 = blockIdx.x * IMG_W;
  const size_t base = (size_t)b * P.RC;
  const float *sweep = raw + (size_t)b * Nmax * stride;
  const int ncell = P.R * SW;
  // winners of the strip, read along the image rows (coalesced) ...
  int *s_win = reinterpret_cast<int *>(s_ground + (((size_t)ncell + 3) & ~(size_t)3));  // [R][SW]
  for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
    const int row = t / SW, c = t - row * SW;
    int col = col0 + c - 1;
    if (col < 0) col += P.C;
    s_win[t] = col < P.C ? winner[base + (size_t)row * P.C + col] : -1;
    s_ground[t] = 0;
  }
  __syncthreads();
  // ... and gathered along the image COLUMNS: for a ring-fastest sweep the points of one column are consecutive in the input,
  // so the lanes of a warp (consecutive rows) read a contiguous span instead of 32 points a firing apart
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = warp; c < SW; c += nwarp)
    for (int row = lane; row < P.R; row += 32) {
      const int w = s_win[row * SW + c];
      float4 v = make_float4(0.f, 0.f, 0.f, ALEGO_EMPTY_RANGE);
      if (w >= 0) {
        const float4 p = load_point(sweep, w, stride);
        v = make_float4(p.x, p.y, p.z, sqrtf(p.x * p.x + p.y * p.y + p.z * p.z));  // (:99) float sum, float sqrt
      }
      s_pt[row * SW + c] = v;
    }
  __syncthreads();
  // ---- ground (:106-132): one thread per vertical pair (i, i+1), i < ground_scan_id
  const int grows = min(P.ground_scan_id, P.R - 1);
  for (int t = threadIdx.x; t < grows * SW; t += blockDim.x) {
    const float4 lo = s_pt[t], up = s_pt[t + SW];
    if (lo.w == ALEGO_EMPTY_RANGE || up.w == ALEGO_EMPTY_RANGE) continue;
    const float fx = up.x - lo.x, fy = up.y - lo.y, fz = up.z - lo.z;  // float differences, then widened
    // fast float estimate; the exact double evaluation only near the 10 degree threshold
    float af = 0.f;
    const float h2 = fx * fx + fy * fy;
    const bool fast = h2 < 1e30f && fast_atan2(fz, h2 * rsqrt_approx(fmaxf(h2, 1e-30f)), af);
    bool is_ground;
    const float da = fabsf(af * 57.29577951f - (float)P.sensor_mount_ang);
    if (fast && fabsf(da - 10.f) > 1e-2f) {
      is_ground = da < 10.f;
    } else {
      const double dx = fx, dy = fy, dz = fz;
      const double angle = atan2(dz, hypot(dx, dy)) * 180.0 / CUDART_PI;
      is_ground = fabs(angle - P.sensor_mount_ang) < 10.;
    }
    if (is_ground) {
      s_ground[t] = 1;
      s_ground[t + SW] = 1;
    }
  }
  __syncthreads();
  // ---- outputs + flags of the strip's own columns
  for (int t = threadIdx.x; t < P.R * IMG_W; t += blockDim.x) {
    const int row = t / IMG_W, c = (t & (IMG_W - 1)) + 1, col = col0 + c - 1;
    if (col >= P.C) continue;
    const int si = row * SW + c;
    const float4 v = s_pt[si];
    const bool has = v.w != ALEGO_EMPTY_RANGE, g = s_ground[si] != 0;
    const bool valid = has && !g;
    int f = (valid ? CELL_VALID : 0) | (g ? CELL_GROUND : 0);
    if (valid) {
      const float4 l = s_pt[si - 1];
      if (P.C > 1 && l.w != ALEGO_EMPTY_RANGE && s_ground[si - 1] == 0 && seg_join(l.w, v.w, true, P)) f |= CELL_JOIN_LEFT;
      if (row + 1 < P.R) {
        const float4 u = s_pt[si + SW];
        if (u.w != ALEGO_EMPTY_RANGE && s_ground[si + SW] == 0 && seg_join(v.w, u.w, false, P)) f |= CELL_JOIN_DOWN;
      }
    }
    const size_t cell = base + (size_t)row * P.C + col;
    // nan_p (:24-27) for cells without a return; intensity = row + col/10000 (:101)
    cloud[cell] = has ? make_float4(v.x, v.y, v.z, (float)((double)row + (double)col / 10000.0)) : make_float4(0.f, 0.f, 0.f, -1.f);
    range[cell] = v.w;
    ground[cell] = g ? 1 : 0;
    flags[cell] = (uint8_t)f;
  }
}

// ---------------------------------------------------------------------------------------------------
// connected components.  parent[] holds cell indices local to the sequence; the root of a component is
// its smallest raster index (links always point to smaller indices → no cycles).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int x) {
  const volatile int *V = L;  // other threads relink roots concurrently (atomicMin in uf_unite)
  int p = V[x];
  while (p != x) {
    x = p;
    p = V[x];
  }
  return x;
}
__device__ __forceinline__ void uf_unite(int *L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      const int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// ---------------------------------------------------------------------------------------------------
// labelComponents (:210-316) as connected components in SHARED MEMORY, strip by strip: one CTA owns all R rows of CCL_W
// columns (R * CCL_W <= 4096 cells).  Inside the strip: per-row runs by a warp max-scan (no atomics), vertical joins by
// union-find on 32-bit labels in shared memory, flatten, per-component size / highest row by shared-memory atomics.  What
// leaves the CTA: parent[cell] = the strip-local root (global raster index: the smallest one of the component's cells in the
// strip) and (size, highest row) at the roots.  ccl_seam then unites strip roots across the strip seams and the column wrap
// (:241-248) — a few hundred joins per image — and hands every absorbed root's statistics to the root that survives, so a
// final root is at most ONE hop from any cell's parent (ccl_root).  Nothing is resolved through global-memory pointer chasing.
// ---------------------------------------------------------------------------------------------------
#define CCL_CELLS 4096
#define CCL_SMEM (CCL_CELLS * 13)
// strip width: the largest power of two with R * W <= CCL_CELLS (>= 32), so that cell <-> (row, column) is shift and mask
__host__ __device__ __forceinline__ int ccl_strip_shift(int R) {
  int sh = 5;
  while (sh < 12 && (R << (sh + 1)) <= CCL_CELLS) ++sh;
  return sh;
}
__device__ __forceinline__ int ccl_strip_width(int R) { return 1 << ccl_strip_shift(R); }

__global__ void __launch_bounds__(256) ccl_strip_kernel(uint8_t *flags, int *__restrict__ parent, int2 *__restrict__ comp_stat, IpDev P) {
  extern __shared__ __align__(16) unsigned char ccl_smem[];  // CCL_SMEM bytes
  int *s_lab = reinterpret_cast<int *>(ccl_smem);
  int *s_cnt = s_lab + CCL_CELLS;
  int *s_mxr = s_cnt + CCL_CELLS;
  uint8_t *s_fl = reinterpret_cast<uint8_t *>(s_mxr + CCL_CELLS);
  const int b = blockIdx.y, WS = ccl_strip_shift(P.R), W = 1 << WS, c0 = blockIdx.x * W, w = min(W, P.C - c0), n = P.R * W;
  const size_t base = (size_t)b * P.RC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((P.C & 3) == 0 && (w & 3) == 0) {  // four flag bytes per load (rows start 4-byte aligned when C is a multiple of 4)
    const uint32_t *f4 = reinterpret_cast<const uint32_t *>(flags + base + c0);
    uint32_t *s4 = reinterpret_cast<uint32_t *>(s_fl);
    for (int q = threadIdx.x; q < (n >> 2); q += blockDim.x) {
      const int idx = q << 2, r = idx >> WS, c = idx & (W - 1);
      s4[q] = c < w ? f4[((size_t)r * P.C + c) >> 2] : 0u;
    }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) { s_cnt[idx] = 0; s_mxr[idx] = 0; }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const int r = idx >> WS, c = idx & (W - 1);
      s_fl[idx] = c < w ? flags[base + (size_t)r * P.C + c0 + c] : 0;
      s_cnt[idx] = 0;
      s_mxr[idx] = 0;
    }
  }
  __syncthreads();
  // runs of every row: a cell joined to its left neighbour continues the run (the strip's first column starts one: the join
  // across the seam is ccl_seam's)
  for (int r = warp; r < P.R; r += (int)(blockDim.x >> 5)) {
    int carry = -1;
    for (int cc = 0; cc < W; cc += 32) {
      const int c = cc + lane, f = s_fl[r * W + c];
      const bool valid = (f & CELL_VALID) != 0, join_left = c > 0 && (f & CELL_JOIN_LEFT) != 0;
      int run = (valid && !join_left) ? c : -1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run = max(run, t);
      }
      run = max(run, carry);
      s_lab[r * W + c] = valid ? r * W + run : -1;
      carry = __shfl_sync(0xffffffffu, run, 31);
    }
  }
  __syncthreads();
  // joins between a cell and the cell one row up in the image (rows do not wrap, :237)
  for (int idx = threadIdx.x; idx < n - W; idx += blockDim.x) {
    const int f = s_fl[idx];
    if (!(f & CELL_JOIN_DOWN)) continue;
    // two runs in contact along several columns: only the first column of the contact touches the forest
    const int c = idx & (W - 1);
    if (c > 0 && (f & CELL_JOIN_LEFT) && (s_fl[idx + W] & CELL_JOIN_LEFT) && (s_fl[idx - 1] & CELL_JOIN_DOWN)) continue;
    uf_unite(s_lab, idx, idx + W);
  }
  __syncthreads();
  // flatten by pointer jumping: the vertical joins chain the runs of a column row by row (up to R hops), so every cell chasing
  // its own root would walk those chains serially; halving all paths together needs about log2(R) sweeps
  // (Jacobi sweeps: all reads of a sweep before its writes, so that no thread reads a label another one is replacing)
  {
    bool changed;
    do {
      changed = false;
      int nl[CCL_CELLS / 256];
#pragma unroll
      for (int k = 0; k < CCL_CELLS / 256; ++k) {
        const int idx = threadIdx.x + k * 256;
        nl[k] = -2;
        if (idx < n) {
          const int l = s_lab[idx];
          if (l >= 0) {
            const int ll = s_lab[s_lab[s_lab[l]]];  // three jumps per sweep: a third of the barriers
            if (ll != l) nl[k] = ll;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CCL_CELLS / 256; ++k)
        if (nl[k] != -2) { s_lab[threadIdx.x + k * 256] = nl[k]; changed = true; }
      changed = __syncthreads_or(changed) != 0;
    } while (changed);
  }
  for (int idx0 = 0; idx0 < n; idx0 += blockDim.x) {  // whole warps take part in the match
    const int idx = idx0 + threadIdx.x;
    const int root = idx < n ? s_lab[idx] : -1;
    // the cells a warp holds mostly share a component: one shared-memory atomic per (warp, component)
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (root >= 0 && (peers & ((1u << lane) - 1u)) == 0u) {
      atomicAdd(&s_cnt[root], __popc(peers));
      atomicMax(&s_mxr[root], (idx | 31) >> WS);  // a warp's 32 cells share a row when W >= 32: its highest row
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int r = idx >> WS, c = idx & (W - 1);
    if (c >= w) continue;
    const size_t cell = base + (size_t)r * P.C + c0 + c;
    const int lab = s_lab[idx];
    int g = -1;
    if (lab >= 0) {
      const int root = lab;  // flattened above
      g = (root >> WS) * P.C + c0 + (root & (W - 1));
      if (root == idx) {
        comp_stat[cell] = make_int2(s_cnt[idx], s_mxr[idx]);
        flags[cell] = (uint8_t)(s_fl[idx] | CELL_STRIP_ROOT);
      }
    }
    parent[cell] = g;
  }
}

// joins across the strip seams and the column wrap: one CTA per image.  Only strip roots are linked (their entries of parent[]).
__global__ void __launch_bounds__(256) ccl_seam_kernel(const uint8_t *__restrict__ flags, int *parent, int2 *comp_stat, IpDev P) {
  const int b = blockIdx.x, W = ccl_strip_width(P.R);
  const size_t base = (size_t)b * P.RC;
  int *L = parent + base;
  int2 *S = comp_stat + base;
  const uint8_t *fl = flags + base;
  const int n_seams = (P.C + W - 1) / W;  // seam 0 = the wrap between column C-1 and column 0
  const int n_edges = P.C > 1 ? n_seams * P.R : 0;
  for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
    const int s = e / P.R, r = e - s * P.R;
    const int col = s * W, left = col == 0 ? P.C - 1 : col - 1;
    if (fl[r * P.C + col] & CELL_JOIN_LEFT) uf_unite(L, L[r * P.C + col], L[r * P.C + left]);
  }
  __syncthreads();
  // every strip root that was absorbed: point it straight at the surviving root and hand over its statistics (once)
  for (int e = threadIdx.x; e < 2 * n_edges; e += blockDim.x) {
    const int ee = e >> 1, s = ee / P.R, r = ee - s * P.R;
    const int col = s * W, left = col == 0 ? P.C - 1 : col - 1;
    if (!(fl[r * P.C + col] & CELL_JOIN_LEFT)) continue;
    const int cell = r * P.C + ((e & 1) ? left : col);
    const int x = (fl[cell] & CELL_STRIP_ROOT) ? cell : L[cell];  // a non-root cell still points at its strip root
    const int f = uf_find(L, x);
    if (f == x) continue;
    L[x] = f;
    const int old = atomicExch(&S[x].x, 0);
    if (old > 0) {
      atomicAdd(&S[f].x, old);
      atomicMax(&S[f].y, S[x].y);
    }
  }
}

// final root of a cell's component: its strip root, or what ccl_seam pointed that root at
__device__ __forceinline__ int ccl_root(const int *L, int cell) {
  int root = L[cell];
  if (root >= 0) {
    const int up = L[root];
    if (up != root) {
      root = up;
      int up2 = L[root];
      while (up2 != root) { root = up2; up2 = L[root]; }  // only when a hand-over has not been flattened (never after ccl_seam)
    }
  }
  return root;
}

struct CellClass {
  bool keep, outl, rootflag, valid, feasible;
};
__device__ __forceinline__ CellClass classify(int cell, int row, int col, const int *L, const int2 *stat, const uint8_t *gr,
                                              const IpDev &P) {
  CellClass c;
  const int root = ccl_root(L, cell);
  c.valid = root >= 0;
  c.feasible = false;
  if (c.valid) {
    const int2 s = stat[root];
    // (:282-301) the distinct-row count (an integer division) only matters for the small clusters
    c.feasible = s.x >= P.seg_min_cluster || (s.x >= P.seg_valid_point_num && s.y - root / P.C + 1 >= P.seg_valid_line_num);
  }
  const bool g = gr[cell] == 1;
  // (:164-181) kept: feasible cluster cells, and ground cells on every 5th column or the 5-column borders
  c.keep = (c.valid && c.feasible) || (g && (col % 5 == 0 || col <= 4 || col >= P.C - 5));
  c.outl = c.valid && !c.feasible && row > P.ground_scan_id && col % 5 == 0;  // (:167-173)
  c.rootflag = c.valid && c.feasible && root == cell;
  return c;
}

__global__ void __launch_bounds__(256) ip_rowcount_kernel(const int *__restrict__ parent, const int2 *__restrict__ comp_stat,
                                                          const uint8_t *__restrict__ ground, uint8_t *__restrict__ cell_class,
                                                          int4 *__restrict__ rowcnt,
                                                          const float *__restrict__ raw, int stride, const int *__restrict__ n_pts,
                                                          float *orient, int Nmax, IpDev P) {
  const int b = blockIdx.y, row = blockIdx.x;
  const size_t base = (size_t)b * P.RC;
  int nk = 0, no = 0, nr = 0;
  for (int col = threadIdx.x; col < P.C; col += blockDim.x) {
    const CellClass c = classify(row * P.C + col, row, col, parent + base, comp_stat + base, ground + base, P);
    // the compaction kernel reads the verdict back instead of classifying every cell two more times
    cell_class[base + row * P.C + col] = (uint8_t)((c.keep ? 1 : 0) | (c.outl ? 2 : 0) | (c.rootflag ? 4 : 0));
    nk += c.keep;
    no += c.outl;
    nr += c.rootflag;
  }
  __shared__ int s[3][8];
  __shared__ int s_fv, s_lv;
  nk = warp_sum_i(nk); no = warp_sum_i(no); nr = warp_sum_i(nr);
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = nk; s[1][threadIdx.x >> 5] = no; s[2][threadIdx.x >> 5] = nr; }
  if (threadIdx.x == 0) { s_fv = 0x7fffffff; s_lv = -1; }
  __syncthreads();
  if (row == 0) {  // first / last point that survives removeNaNFromPointCloud (:59,:62-63): scan inwards from both ends
    const int n = min(n_pts[b], Nmax);
    const float *src = raw + (size_t)b * Nmax * stride;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
      const int i = c0 + threadIdx.x;
      const bool f = i < n && finite3(load_point(src, i, stride));
      if (f) atomicMin(&s_fv, i);
      if (__syncthreads_or(f)) break;
    }
    for (int c0 = n - 1; c0 >= 0; c0 -= blockDim.x) {
      const int i = c0 - threadIdx.x;
      const bool f = i >= 0 && finite3(load_point(src, i, stride));
      if (f) atomicMax(&s_lv, i);
      if (__syncthreads_or(f)) break;
    }
  }
  if (threadIdx.x == 0) {
    int a = 0, o = 0, r = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s[0][w]; o += s[1][w]; r += s[2][w]; }
    rowcnt[b * P.R + row] = make_int4(a, o, r, 0);
    if (row == 0) {  // start/end orientation (:62-72): float atan2, float32 message fields
      float so = 0.f, eo = 0.f;
      const int fv = s_fv, lv = s_lv;
      if (lv >= 0 && fv <= lv) {
        const float *sw = raw + (size_t)b * Nmax * stride;
        const float4 p0 = load_point(sw, fv, stride), p1 = load_point(sw, lv, stride);
        so = -(float)atan2((double)p0.y, (double)p0.x);
        eo = (float)((double)(-(float)atan2((double)p1.y, (double)p1.x)) + 2 * CUDART_PI);
        if ((double)(eo - so) > 3 * CUDART_PI) eo = (float)((double)eo - 2 * CUDART_PI);
        else if ((double)(eo - so) < CUDART_PI) eo = (float)((double)eo + 2 * CUDART_PI);
      }
      orient[b * 4 + 0] = so;
      orient[b * 4 + 1] = eo;
      orient[b * 4 + 2] = eo - so;
    }
  }
}

__global__ void __launch_bounds__(256)
ip_compact_kernel(const uint8_t *__restrict__ cell_class, const uint8_t *__restrict__ ground, const int4 *__restrict__ rowcnt, const float4 *__restrict__ cloud, const float *__restrict__ range,
                  int *__restrict__ comp_id, float4 *__restrict__ seg_cloud, uint8_t *__restrict__ seg_ground,
                  int *__restrict__ seg_col, float *__restrict__ seg_range, int *__restrict__ start_ring,
                  int *__restrict__ end_ring, int *__restrict__ Mout, float4 *__restrict__ outlier, int *__restrict__ n_outlier,
                  int out_cap, IpDev P) {
  const int b = blockIdx.y, row = blockIdx.x;
  const size_t base = (size_t)b * P.RC;
  __shared__ int s_base[3];
  __shared__ int s_scan[34];
  if (threadIdx.x < 32) {
    int a = 0, o = 0, r = 0;
    for (int t = threadIdx.x; t < row; t += 32) {
      const int4 c = rowcnt[b * P.R + t];
      a += c.x; o += c.y; r += c.z;
    }
    a = warp_sum_i(a); o = warp_sum_i(o); r = warp_sum_i(r);
    if (threadIdx.x == 0) {
      s_base[0] = a; s_base[1] = o; s_base[2] = r;
      const int4 mine = rowcnt[b * P.R + row];
      start_ring[b * P.R + row] = a + 5;             // (:161)
      end_ring[b * P.R + row] = a + mine.x - 1 - 5;  // (:190)
      if (row == P.R - 1) {
        Mout[b] = a + mine.x;
        n_outlier[b] = min(o + mine.y, out_cap);
      }
    }
  }
  __syncthreads();
  // every thread owns `per` consecutive columns: count, ONE block scan for the whole ring, then the ordered writes (the class
  // bytes ip_rowcount left are read twice — cached loads — instead of scanning the ring in eight 256-column chunks)
  const int per = (P.C + (int)blockDim.x - 1) / (int)blockDim.x;
  const int c_lo = min((int)threadIdx.x * per, P.C), c_hi = min(c_lo + per, P.C);
  int nk = 0, no = 0, nr = 0;
  const uint8_t *cls = cell_class + base + (size_t)row * P.C;
  for (int col = c_lo; col < c_hi; ++col) {
    const int c = cls[col];
    nk += c & 1;
    no += (c >> 1) & 1;
    nr += (c >> 2) & 1;
  }
  int total;
  const int ex_ko = block_excl_scan(nk | (no << 16), s_scan, &total);  // a ring holds < 65536 cells
  const int ex_r = block_excl_scan(nr, s_scan, &total);
  int dk = s_base[0] + (ex_ko & 0xffff), dout = s_base[1] + (ex_ko >> 16), dr = s_base[2] + ex_r;
  for (int col = c_lo; col < c_hi; ++col) {
    const int cell = row * P.C + col;
    const int c = cls[col];
    if (c & 1) {
      seg_cloud[base + dk] = cloud[base + cell];
      seg_ground[base + dk] = ground[base + cell] == 1;  // (:183)
      seg_col[base + dk] = col;                          // (:184)
      seg_range[base + dk] = range[base + cell];         // (:185)
      ++dk;
    }
    if (c & 2) {
      if (dout < out_cap) outlier[(size_t)b * out_cap + dout] = cloud[base + cell];
      ++dout;
    }
    if (c & 4) comp_id[base + cell] = ++dr;  // label_cnt_ in raster-seed order (:303-306)
  }
}

__global__ void __launch_bounds__(256) ip_label_kernel(const int *__restrict__ parent, const int2 *__restrict__ comp_stat,
                                                       const int *__restrict__ comp_id, int *__restrict__ label, IpDev P) {
  const int b = blockIdx.y;
  const size_t base = (size_t)b * P.RC;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < P.RC; cell += gridDim.x * blockDim.x) {
    const int root = ccl_root(parent + base, cell);
    int lab = -1;
    if (root >= 0) {
      const int2 s = comp_stat[base + root];
      const bool feas = s.x >= P.seg_min_cluster || (s.x >= P.seg_valid_point_num && s.y - root / P.C + 1 >= P.seg_valid_line_num);
      lab = feas ? comp_id[base + root] : ALEGO_LABEL_INVALID;
    }
    label[base + cell] = lab;
  }
}

}  // namespace

IpDev make_ip_dev(const AlegoHandle *h) {
  IpDev d;
  d.R = h->R; d.C = h->C; d.RC = h->RC;
  d.ground_scan_id = h->P.ground_scan_id;
  d.seg_valid_point_num = h->P.seg_valid_point_num;
  d.seg_valid_line_num = h->P.seg_valid_line_num;
  d.seg_min_cluster = h->P.seg_min_cluster;
  d.ang_res_x = h->P.ang_res_x; d.ang_res_y = h->P.ang_res_y; d.ang_bottom = h->P.ang_bottom;
  d.sensor_mount_ang = h->P.sensor_mount_ang;
  d.seg_theta = h->P.seg_theta;
  d.row_scale = 180.0 / M_PI / h->P.ang_res_y;
  d.row_off = h->P.ang_bottom / h->P.ang_res_y + 0.5;
  d.col_scale = 180.0 / M_PI / h->P.ang_res_x;
  d.row_scale_f = (float)d.row_scale; d.row_off_f = (float)d.row_off; d.col_scale_f = (float)d.col_scale;
  // uncertainty band of the fast path in cells: 4x (angle error bound x scale + float rounding of a coordinate < 1.5 C)
  d.eps_row = (float)std::max(2e-3, 4.0 * (ALEGO_FAST_ATAN_ERR * d.row_scale + 2e-7 * (h->R + std::fabs(d.row_off))));
  d.eps_col = (float)std::max(2e-3, 4.0 * (2.0 * ALEGO_FAST_ATAN_ERR * d.col_scale + 2e-7 * 1.5 * h->C));
  d.sin_x = h->seg_sin_x; d.cos_x = h->seg_cos_x; d.sin_y = h->seg_sin_y; d.cos_y = h->seg_cos_y;
  return d;
}

int ip_run_device(AlegoHandle *h, bool want_labels) {
  const IpDev P = make_ip_dev(h);
  const int B = h->B;
  cudaStream_t s = h->stream;
  const int pt_blocks = min(div_up(h->Nmax, 32 * P.R), 4096);
  const int cell_blocks = min(div_up(h->RC, 256), 4096);
  static bool proj_attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!proj_attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(ip_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(32 * (ALEGO_MAX_RINGS + 1) * sizeof(float4))));
    CUDA_TRY(h, cudaFuncSetAttribute(ip_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(ALEGO_MAX_RINGS * (IMG_W + 1) * (sizeof(float4) + 1 + sizeof(int)) + 16)));
    proj_attr_set[h->dev] = true;
  }
  { LAUNCH(h, "ip_project");
    ip_project_kernel<<<dim3(pt_blocks, B), 256, (size_t)32 * (P.R + 1) * sizeof(float4), s>>>(reinterpret_cast<const float *>(h->raw), h->n_pts, h->winner, h->Nmax, h->in_stride, P); }
  const size_t img_smem = (size_t)P.R * (IMG_W + 1) * (sizeof(float4) + 1 + sizeof(int)) + 16;
  { LAUNCH(h, "ip_image");
    ip_image_kernel<<<dim3(div_up(P.C, IMG_W), B), 256, img_smem, s>>>(reinterpret_cast<const float *>(h->raw), h->in_stride, h->winner,
                                                                   h->cloud, h->range, h->ground, h->cell_flags, h->Nmax, P); }
  // the winner image is consumed: empty it for the next sweep (a cell is reset by memset rather than by its reader because the
  // neighbouring strip reads it too)
  CUDA_TRY(h, cudaMemsetAsync(h->winner, 0xFF, (size_t)B * h->RC * sizeof(int), s));
  const int ccl_w = 1 << ccl_strip_shift(P.R);
  static bool ccl_attr_set[ALEGO_MAX_DEVICES] = {};
  if (!ccl_attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(ccl_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CCL_SMEM));
    ccl_attr_set[h->dev] = true;
  }
  { LAUNCH(h, "ccl_strip"); ccl_strip_kernel<<<dim3(div_up(P.C, ccl_w), B), 256, CCL_SMEM, s>>>(h->cell_flags, h->parent, h->comp_stat, P); }
  { LAUNCH(h, "ccl_seam"); ccl_seam_kernel<<<B, 256, 0, s>>>(h->cell_flags, h->parent, h->comp_stat, P); }
  { LAUNCH(h, "ip_rowcount");
    ip_rowcount_kernel<<<dim3(P.R, B), 256, 0, s>>>(h->parent, h->comp_stat, h->ground, h->cell_class, h->rowcnt, reinterpret_cast<const float *>(h->raw), h->in_stride, h->n_pts, h->orient,
                                                   h->Nmax, P); }
  { LAUNCH(h, "ip_compact");
    ip_compact_kernel<<<dim3(P.R, B), 256, 0, s>>>(h->cell_class, h->ground, h->rowcnt, h->cloud, h->range, h->comp_id,
                                                  h->seg_cloud, h->seg_ground, h->seg_col, h->seg_range, h->start_ring,
                                                  h->end_ring, h->M, h->outlier, h->n_outlier, h->out_cap, P); }
  h->label_valid = false;
  if (want_labels) return ip_label_device(h);
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

// label_mat_ (imageProjection.h:25) is internal to pcCB — nothing downstream reads it — so it is materialised from the
// component forest only when a caller asks for it (alego_ip_get(label_image), tests).
int ip_label_device(AlegoHandle *h) {
  if (h->label_valid) return ALEGO_OK;
  const IpDev P = make_ip_dev(h);
  const int cell_blocks = min(div_up(h->RC, 256), 4096);
  { LAUNCH(h, "ip_label");
    ip_label_kernel<<<dim3(cell_blocks, h->B), 256, 0, h->stream>>>(h->parent, h->comp_stat, h->comp_id, h->label, P); }
  CUDA_TRY(h, cudaGetLastError());
  h->label_valid = true;
  return ALEGO_OK;
}
