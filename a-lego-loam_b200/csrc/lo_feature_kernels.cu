// lo_feature_kernels.cu — LaserOdometry feature extraction on the device: replaces steps 2-4 of
// LaserOdometry::mainLoop (src/laserOdometry.cpp:118-297).
//
// K6+K7 lo_curv_occl      : 11-tap float range stencil (:122-129) fused with markOccludedPoints (:131-159);
//                           the ±6 window of ranges / columns is staged in shared memory
// K8a   lo_sort_segments  : per ring segment, the exact permutation std::sort would produce (:185)
// K8b   lo_select         : one warp per ring walks its 6 segments in order (:188-286)
// K9    lo_less_flat_voxel: per-ring VoxelGrid(0.4) of the less-flat points (:279-293) + azimuth bins of the result
//       lo_finalize       : ring-major concatenation of the per-ring lists / clouds
#include "common.cuh"
#include "lo_kernels.cuh"
#include "sort_voxel.cuh"
#include "stdsort_clone.cuh"
#include "vox_order.cuh"

namespace {

#define CURV_TILE 256

__global__ void __launch_bounds__(CURV_TILE)
lo_curv_occl_kernel(const float *__restrict__ seg_range, const int *__restrict__ seg_col, const int *__restrict__ Mdev,
                    float *__restrict__ curv, uint8_t *__restrict__ picked0, int *__restrict__ flabel,
                    int *__restrict__ sort_idx, int RC) {
  const int b = blockIdx.y;
  const int M = Mdev[b];
  const size_t base = (size_t)b * RC;
  __shared__ float sr[CURV_TILE + 12];
  __shared__ int sc[CURV_TILE + 12];
  __shared__ uint8_t sf[CURV_TILE + 12];  // per loop iteration j of markOccludedPoints: 1 = marks j-5..j, 2 = marks j+1..j+5
  // the segmented cloud holds M of the R*C cells (about a third): bounded grid, tiles taken with a stride
  for (int tile_lo = blockIdx.x * CURV_TILE; tile_lo < M; tile_lo += gridDim.x * CURV_TILE) {
  __syncthreads();
  for (int t = threadIdx.x; t < CURV_TILE + 12; t += CURV_TILE) {
    const int g = tile_lo - 6 + t;
    const bool ok = g >= 0 && g < M;
    sr[t] = ok ? seg_range[base + g] : 0.f;
    sc[t] = ok ? seg_col[base + g] : 0;
  }
  __syncthreads();
  const int lo = 5, hi = M - 5;  // loop bounds of both reference loops: i in [5, M-5)
  // the occlusion test of iteration j (:134-150) is evaluated once, by the thread that stages j, and shared
  for (int t = threadIdx.x; t < CURV_TILE + 11; t += CURV_TILE) {
    const int j = tile_lo - 6 + t;
    int f = 0;
    if (j >= lo && j < hi && abs(sc[t] - sc[t + 1]) < 10) {
      const double d1 = sr[t], d2 = sr[t + 1];
      if (d1 - d2 > 0.5) f = 1;        // (:140-145), `continue`s past the parallel-beam test
      else if (d2 - d1 > 0.5) f = 2;   // (:146-150)
    }
    sf[t] = (uint8_t)f;
  }
  __syncthreads();
  const int i = tile_lo + threadIdx.x;
  if (i >= M) continue;
  const int li = threadIdx.x + 6;
  float c = 0.f;
  // markOccludedPoints as a gather: iteration j = i..i+5 with f == 1 marks i, iteration j = i-5..i-1 with f == 2 marks i
  bool pk = ((sf[li] | sf[li + 1] | sf[li + 2] | sf[li + 3] | sf[li + 4] | sf[li + 5]) & 1) != 0 ||
            ((sf[li - 1] | sf[li - 2] | sf[li - 3] | sf[li - 4] | sf[li - 5]) & 2) != 0;
  if (i >= lo && i < hi) {
    // (:124) float sum, strictly left to right, r[i]*10 is a float product; compiled without FMA contraction
    const float d = sr[li - 5] + sr[li - 4] + sr[li - 3] + sr[li - 2] + sr[li - 1] - sr[li] * 10 + sr[li + 1] + sr[li + 2] +
                    sr[li + 3] + sr[li + 4] + sr[li + 5];
    c = fabsf(d);  // cloud_curvature_ = double(d)*double(d) is recovered exactly as (double)c*(double)c
    if (sf[li] != 1) {  // parallel-beam test, skipped by the `continue` of the first branch (:144,152-158)
      const double d1 = sr[li], d2 = sr[li + 1];
      const double diff1 = fabs((double)sr[li - 1] - d1), diff2 = fabs(d2 - d1), thr = 0.02 * d1;
      if (diff1 > thr && diff2 > thr) pk = true;
    }
  }
  curv[base + i] = c;
  picked0[base + i] = pk ? 1 : 0;
  flabel[base + i] = 0;
  sort_idx[base + i] = i;
  }
}

// segment bounds (:177-178)
__device__ __forceinline__ void segment_bounds(int start, int end, int j, int &sp, int &ep) {
  sp = (start * (6 - j) + end * j) / 6;
  ep = (start * (5 - j) + end * (j + 1)) / 6 - 1;
}

// ---------------------------------------------------------------------------------------------------
// K8a: one WARP per (sequence, ring, segment) reproduces the permutation libstdc++'s std::sort leaves with the
// reference's curvature-only comparator (:185), ties included — as a data-parallel algorithm:
//  * introsort's partition loop keeps its exact structure (explicit stack, median-of-3 by lane 0), but each
//    __unguarded_partition is evaluated with prefix counts instead of two walking pointers.  With pivot p,
//    "left stoppers" L = positions (ascending) whose key >= p, "right stoppers" R = positions (descending) whose
//    key <= p: the sequential loop swaps L[k] <-> R[k] while L[k] < R[k] and returns min(L[K], R[K-1]) (K = number
//    of swaps).  An L element at t swaps iff #R after t > #L before t; an R element swaps iff #L before t > #R after
//    t; its partner is the stopper of equal rank.  Three ballot sweeps over the range, no divergence.
//  * __final_insertion_sort is a stable sort of what the partitions leave; ranges <= 16 are mutually ordered, so
//    it reduces to a stable rank sort inside every leaf, done for all elements at once.
//  * depth-limit heapsort (never reached on non-adversarial data) runs on lane 0 with the sequential clone.
// tools/proto_parallel_introsort.py checks this formulation against the real std::sort; tests/test_gpu_parity.py
// checks the kernel's cloud_sort_idx_ against the oracle's on tie-heavy sweeps.
// Fast path of K8a: when all curvatures of a segment are distinct the permutation std::sort leaves is simply the ascending
// order, which a register bitonic network over (key, position) produces in a quarter of the instructions of the introsort
// clone (NR elements per lane, partners >= 32 apart live in the same lane).  Returns false — nothing written — when two
// equal keys meet, i.e. when the result depends on introsort's tie behaviour.
template <int NR>
__device__ __forceinline__ bool warp_sort_distinct(const float *__restrict__ keys_in, int n, int *__restrict__ out, int out_base) {
  const int lane = threadIdx.x & 31;
  unsigned key[NR];
  int idx[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int e = lane + 32 * r;
    key[r] = e < n ? __float_as_uint(keys_in[e]) : 0xffffffffu;
    idx[r] = e < n ? e : 0x7fffffff;
  }
#pragma unroll
  for (int k = 2; k <= 32 * NR; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int e = lane + 32 * r;
        unsigned ok;
        int oi;
        if (j >= 32) {
          if ((r & (j >> 5)) != 0) continue;  // handled together with the partner register below
          const int r2 = r | (j >> 5);
          const bool up = (e & k) == 0;
          const bool gt = key[r] > key[r2] || (key[r] == key[r2] && idx[r] > idx[r2]);
          if (gt == up) {
            const unsigned tk = key[r]; key[r] = key[r2]; key[r2] = tk;
            const int ti = idx[r]; idx[r] = idx[r2]; idx[r2] = ti;
          }
          continue;
        }
        ok = __shfl_xor_sync(0xffffffffu, key[r], j);
        oi = __shfl_xor_sync(0xffffffffu, idx[r], j);
        const bool up = (e & k) == 0, lower = (e & j) == 0;
        const bool other_less = ok < key[r] || (ok == key[r] && oi < idx[r]);
        // the lower element of an ascending pair keeps the minimum, the upper one the maximum (reversed when descending)
        if ((up == lower) == other_less) { key[r] = ok; idx[r] = oi; }
      }
    }
  }
  // ties between real elements?  (element e+1 lives in lane+1, or in lane 0 of the next register)
  bool tie = false;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    unsigned nk = __shfl_down_sync(0xffffffffu, key[r], 1);
    const unsigned wrap = r + 1 < NR ? __shfl_sync(0xffffffffu, key[r + 1 < NR ? r + 1 : r], 0) : 0xffffffffu;
    if (lane == 31) nk = wrap;
    const int e = lane + 32 * r;
    tie |= e + 1 < n && nk == key[r];
  }
  if (__any_sync(0xffffffffu, tie)) return false;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int e = lane + 32 * r;
    if (e < n) out[e] = out_base + idx[r];
  }
  return true;
}

#define SORT_WARPS 8
__device__ __forceinline__ unsigned ss_key(unsigned long long e) { return (unsigned)(e >> 32); }
__device__ __forceinline__ unsigned ss_pack(int f, int l, int depth) { return (unsigned)f | ((unsigned)l << 11) | ((unsigned)depth << 22); }

__global__ void __launch_bounds__(SORT_WARPS * 32)
lo_sort_segments_kernel(const float *__restrict__ curv, const int *__restrict__ start_ring, const int *__restrict__ end_ring,
                        unsigned long long *__restrict__ scratch, int *__restrict__ sort_idx, int B, int R, int RC, int cap) {
  extern __shared__ __align__(16) unsigned char ss_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * SORT_WARPS + warp;
  if (seg >= B * R * 6) return;
  const int b = seg / (R * 6), rs = seg - b * R * 6, ring = rs / 6, j = rs - ring * 6;
  int sp, ep;
  segment_bounds(start_ring[b * R + ring], end_ring[b * R + ring], j, sp, ep);
  if (sp >= ep) return;
  const size_t base = (size_t)b * RC;
  const int n = ep - sp + 1;
  if (n > cap) {  // cannot happen for cap = C/6 + 8 (a ring holds <= C points); kept as a correct slow path
    if (lane == 0) {
      unsigned long long *e = scratch + base + sp;
      for (int k = 0; k < n; ++k) e[k] = ((unsigned long long)__float_as_uint(curv[base + sp + k]) << 32) | (unsigned)(sp + k);
      ssc::sort(e, n);
      for (int k = 0; k < n; ++k) sort_idx[base + sp + k] = (int)(e[k] & 0xffffffffu);
    }
    return;
  }
  {  // distinct keys (the common case): plain ascending order
    const float *kin = curv + base + sp;
    int *o = sort_idx + base + sp;
    bool done = false;
    if (n <= 32) done = warp_sort_distinct<1>(kin, n, o, sp);
    else if (n <= 64) done = warp_sort_distinct<2>(kin, n, o, sp);
    else if (n <= 128) done = warp_sort_distinct<4>(kin, n, o, sp);
    if (done) return;
  }
  const size_t per_warp = (size_t)cap * 14 + 16;
  unsigned long long *e = reinterpret_cast<unsigned long long *>(ss_smem + warp * per_warp);
  unsigned *bounds = reinterpret_cast<unsigned *>(e + cap);
  unsigned short *posL = reinterpret_cast<unsigned short *>(bounds + cap);
  unsigned short *posR = posL + cap / 2 + 4;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
  for (int k = lane; k < n; k += 32) e[k] = ((unsigned long long)__float_as_uint(curv[base + sp + k]) << 32) | (unsigned)k;
  __syncwarp();
  unsigned stack[24];
  int spn = 0;
  stack[spn++] = ss_pack(0, n, 2 * (31 - __clz(n)));
  while (spn > 0) {
    const unsigned fr = stack[--spn];
    const int f = (int)(fr & 2047u);
    int l = (int)((fr >> 11) & 2047u), depth = (int)(fr >> 22);
    bool heap = false;
    while (l - f > 16) {
      if (depth == 0) {
        if (lane == 0) ssc::heap_sort(e + f, (long)(l - f));
        __syncwarp();
        heap = true;
        break;
      }
      --depth;
      if (lane == 0) ssc::median_to_first(e + f, e + f + 1, e + f + (l - f) / 2, e + l - 1);
      __syncwarp();
      const unsigned p = ss_key(e[f]);
      const int lo = f + 1;
      int totR = 0;
      for (int c = lo; c < l; c += 32) {
        const int t = c + lane;
        totR += __popc(__ballot_sync(full, t < l && ss_key(e[t]) <= p));
      }
      int runL = 0, runR = 0, K = 0, first_keep_L = 0x7fffffff, min_swap_R = l;
      for (int c = lo; c < l; c += 32) {
        const int t = c + lane;
        const bool v = t < l;
        const unsigned kt = v ? ss_key(e[t]) : 0u;
        const bool isL = v && kt >= p, isR = v && kt <= p;
        const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
        const int cL = runL + __popc(mL & lt_mask);           // left stoppers before t
        const int cR = totR - (runR + __popc(mR & le_mask));  // right stoppers after t
        const bool sL = isL && cR > cL, sR = isR && cL > cR;
        if (sL) posL[cL] = (unsigned short)t;
        if (sR) posR[cR] = (unsigned short)t;
        if (isL && !sL) first_keep_L = min(first_keep_L, t);
        if (sR) min_swap_R = min(min_swap_R, t);
        K += __popc(__ballot_sync(full, sL));
        runL += __popc(mL);
        runR += __popc(mR);
      }
      first_keep_L = __reduce_min_sync(full, first_keep_L);
      min_swap_R = __reduce_min_sync(full, min_swap_R);
      __syncwarp();
      for (int k = lane; k < K; k += 32) {
        const int a = posL[k], c2 = posR[k];
        const unsigned long long ea = e[a];
        e[a] = e[c2];
        e[c2] = ea;
      }
      __syncwarp();
      const int cut = min(first_keep_L, min_swap_R);
      if (spn < 24) stack[spn++] = ss_pack(cut, l, depth);
      l = cut;
    }
    if (heap) {
      for (int t = f + lane; t < l; t += 32) bounds[t] = (unsigned)t | ((unsigned)(t + 1) << 16);
    } else {
      for (int t = f + lane; t < l; t += 32) bounds[t] = (unsigned)f | ((unsigned)l << 16);
    }
  }
  __syncwarp();
  for (int t = lane; t < n; t += 32) {
    const unsigned bd = bounds[t];
    const int a = (int)(bd & 0xffffu), c2 = (int)(bd >> 16);
    const unsigned long long my = e[t];
    const unsigned mk = ss_key(my);
    int rank = a;
    for (int u = a; u < c2; ++u) {
      const unsigned ku = ss_key(e[u]);
      rank += (ku < mk) || (ku == mk && u < t);
    }
    sort_idx[base + sp + rank] = sp + (int)(my & 0xffffu);
  }
}

// neighbour suppression (:211-234, 252-275) executed by one lane; pk is the ring's picked window
__device__ __forceinline__ void suppress(const int *__restrict__ col, uint8_t *pk, int idx, int lo) {
  for (int l = 1; l <= 5; ++l) {
    if (abs(col[idx + l] - col[idx + l - 1]) > 10) break;
    pk[idx + l - lo] = 1;
  }
  for (int l = -1; l >= -5; --l) {
    if (abs(col[idx + l] - col[idx + l + 1]) > 10) break;
    pk[idx + l - lo] = 1;
  }
}

#define SEL_WARPS 4
__global__ void __launch_bounds__(SEL_WARPS * 32)
lo_select_kernel(const int *__restrict__ seg_col, const uint8_t *__restrict__ seg_ground, const float *__restrict__ curv,
                 const int *__restrict__ sort_idx, const int *__restrict__ start_ring, const int *__restrict__ end_ring,
                 const uint8_t *__restrict__ picked0, uint8_t *__restrict__ picked, int *__restrict__ flabel,
                 int *__restrict__ ring_feat_cnt, int *__restrict__ ring_sharp, int *__restrict__ ring_less_sharp,
                 int *__restrict__ ring_flat, int R, int RC, int pkcap) {
  extern __shared__ uint8_t sel_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ring = blockIdx.x * SEL_WARPS + warp, b = blockIdx.y;
  if (ring >= R) return;
  const size_t base = (size_t)b * RC;
  const int br = b * R + ring;
  const int start = start_ring[br], end = end_ring[br];
  const int lo = start - 5, len = end + 5 - lo + 1;  // the ring's kept points are [lo, lo+len)
  uint8_t *pk = sel_smem + (size_t)warp * pkcap;
  for (int t = lane; t < len; t += 32) pk[t] = picked0[base + lo + t];
  __syncwarp();
  const int *col = seg_col + base;
  int n_sharp = 0, n_less = 0, n_flat = 0;
  for (int j = 0; j < 6; ++j) {
    int sp, ep;
    segment_bounds(start, end, j, sp, ep);
    if (sp >= ep) continue;
    // ---- sharp / less sharp: descending curvature (:188-236)
    int picked_num = 0;
    bool stop = false;
    for (int kk = ep; kk >= sp && !stop; kk -= 32) {
      const int k = kk - lane;
      const bool inr = k >= sp;
      const int idx = inr ? sort_idx[base + k] : sp;
      const float c = inr ? curv[base + idx] : 0.f;
      const bool thr = inr && (double)c * (double)c > 0.1;
      const bool nong = inr && seg_ground[base + idx] == 0;
      if (__ballot_sync(0xffffffffu, thr) == 0) break;  // sorted: nothing further can exceed the threshold
      int cursor = 0;
      while (true) {
        const bool cand = thr && nong && lane >= cursor && pk[idx - lo] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        __syncwarp();  // every lane has read its flag before the chosen lane updates the window
        if (!m) break;
        const int first = __ffs(m) - 1;
        ++picked_num;
        if (lane == first) {
          pk[idx - lo] = 1;
          if (picked_num <= 2) {
            flabel[base + idx] = 2;
            ring_sharp[br * 12 + n_sharp] = idx;
            ring_less_sharp[br * 120 + n_less] = idx;
          } else if (picked_num <= 20) {
            flabel[base + idx] = 1;
            ring_less_sharp[br * 120 + n_less] = idx;
          }
        }
        if (picked_num <= 2) { ++n_sharp; ++n_less; }
        else if (picked_num <= 20) ++n_less;
        else { stop = true; __syncwarp(); break; }  // 21st candidate: marked picked, no label, walk ends (:207-210)
        if (lane == first) suppress(col, pk, idx, lo);
        __syncwarp();
        cursor = first + 1;
      }
    }
    __syncwarp();
    // ---- flat: ascending curvature (:238-277)
    picked_num = 0;
    stop = false;
    for (int kk = sp; kk <= ep && !stop; kk += 32) {
      const int k = kk + lane;
      const bool inr = k <= ep;
      const int idx = inr ? sort_idx[base + k] : sp;
      const float c = inr ? curv[base + idx] : 1e30f;
      const bool thr = inr && (double)c * (double)c < 0.1;
      const bool gr = inr && seg_ground[base + idx] == 1;
      if (__ballot_sync(0xffffffffu, thr) == 0) break;
      int cursor = 0;
      while (true) {
        const bool cand = thr && gr && lane >= cursor && pk[idx - lo] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        __syncwarp();
        if (!m) break;
        const int first = __ffs(m) - 1;
        ++picked_num;
        if (lane == first) {
          flabel[base + idx] = -1;
          ring_flat[br * 24 + n_flat] = idx;
          pk[idx - lo] = 1;
        }
        ++n_flat;
        if (picked_num >= 4) { stop = true; __syncwarp(); break; }  // breaks before the suppression (:248-251)
        if (lane == first) suppress(col, pk, idx, lo);
        __syncwarp();
        cursor = first + 1;
      }
    }
    __syncwarp();
  }
  for (int t = lane; t < len; t += 32) picked[base + lo + t] = pk[t];
  if (lane == 0) {
    ring_feat_cnt[br * 4 + 0] = n_sharp;
    ring_feat_cnt[br * 4 + 1] = n_less;
    ring_feat_cnt[br * 4 + 2] = n_flat;
  }
}

// per-ring less_flat_scan (:279-285) + VoxelGrid (:288-293): one CTA of LFV_WARPS warps per (ring, sequence)
// (sort_voxel.cuh).  Output staged at lf_stage[lo...] of the ring.
//
// Then the azimuth index of the ring's output (the next sweep's surf_last_): a stable counting sort of the ring's points
// into AZ_BINS azimuth bins.  lo_assoc<SURF> walks "every point of rings cs-2..cs+2" (laserOdometry.cpp:348-395) only
// inside the azimuth window that can hold a point closer than its current best — same result, ~10x fewer candidates.
// az_stage[j] = (x, y, z, position of the point inside the ring); az_off[0..AZ_BINS] = bin starts.
//
// Three launches (one stream): lo_lfv_keys (membership, bounding box, (voxel, point) records), lo_lfv_order (vox_order.cu: the
// partition phase of std::sort on every ring's list, one warp per ring — no block-wide step, where a CTA-per-ring,
// barrier-per-level version left three of four warps idle), lo_lfv_finish (stable radix sort, centroids, azimuth bins).
#define LFV_WARPS 4
#define LFV_MAX_CHUNKS 256  // 8192 columns / 32
__global__ void __launch_bounds__(LFV_WARPS * 32)
lo_lfv_keys_kernel(const float4 *__restrict__ seg_cloud, const int *__restrict__ flabel, const int *__restrict__ start_ring,
                   const int *__restrict__ end_ring, float4 *__restrict__ lf_stage, u64 *__restrict__ keys_a, VoxState *__restrict__ state,
                   int R, int RC, float leaf) {
  __shared__ VoxShared<LFV_WARPS> sh;
  __shared__ unsigned s_member[LFV_MAX_CHUNKS];
  __shared__ int s_chunk_base[LFV_MAX_CHUNKS];
  const int ring = blockIdx.x, b = blockIdx.y;
  const int br = b * R + ring;
  const size_t base = (size_t)b * RC;
  const int start = start_ring[br], end = end_ring[br];
  // {k in a processed segment : cloud_label_[k] <= 0}; the six segments tile [start, end-1], a segment with sp >= ep
  // (at most one point) is skipped by the reference (:179)
  int sp[6], ep[6];
  bool all_segments = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    segment_bounds(start, end, j, sp[j], ep[j]);
    all_segments &= sp[j] < ep[j];
  }
  const int lo = max(start - 5, 0);
  const float4 *pts = seg_cloud + base + start;  // item i = kept point start + i
  const int *lab = flabel + base + start;
  float4 *out = lf_stage + base + lo;
  auto member = [&](int i) {
    bool m = all_segments;
    if (!m) {
      const int k = start + i;
#pragma unroll
      for (int j = 0; j < 6; ++j) m |= (sp[j] < ep[j] && k >= sp[j] && k <= ep[j]);
    }
    return m && lab[i] <= 0;
  };
  const bool done = block_voxel_keys<LFV_WARPS, true>(pts, max(end - start, 0), member, leaf, keys_a + base + lo, out, &sh, s_member,
                                                      s_chunk_base);
  if (threadIdx.x == 0) {
    VoxState st;
    st.frame = sh.frame;
    st.n = max(sh.n, 0);
    st.nv = sh.nv;
    st.done = done ? 1 : 0;
    st.off = st.off_b = (long long)(base + lo);
    state[br] = st;
  }
}

__global__ void __launch_bounds__(LFV_WARPS * 32)
lo_lfv_finish_kernel(const float4 *__restrict__ seg_cloud, const int *__restrict__ start_ring, float4 *__restrict__ lf_stage,
                     int *__restrict__ ring_feat_cnt, u64 *__restrict__ keys_a, u64 *__restrict__ keys_b,
                     const VoxState *__restrict__ state, float4 *__restrict__ az_stage, int *__restrict__ az_off, int R, int RC) {
  __shared__ VoxShared<LFV_WARPS> sh;
  const int ring = blockIdx.x, b = blockIdx.y;
  const int br = b * R + ring;
  const size_t base = (size_t)b * RC;
  const int start = start_ring[br];
  const int lo = max(start - 5, 0);
  const float4 *pts = seg_cloud + base + start;
  float4 *out = lf_stage + base + lo;
  const VoxState st = state[br];
  const int n_out = st.done ? st.n : block_voxel_finish<LFV_WARPS>(pts, keys_a + base + lo, keys_b + base + lo, out, &sh, st.n, st.nv,
                                                                    st.frame.key_bits);
  if (threadIdx.x == 0) ring_feat_cnt[br * 4 + 3] = n_out;
  // ---- azimuth bins of the ring's output
  int *azo = az_off + (size_t)br * (AZ_BINS + 1);
  if (n_out == 0) {
    for (int t = threadIdx.x; t <= AZ_BINS; t += blockDim.x) azo[t] = 0;
    return;
  }
  float4 *azs = az_stage + base + lo;
  block_counting_pass<LFV_WARPS>(
      n_out, &sh,
      [&](int i) {
        const float4 p = out[i];
        return az_bin_unwrapped(az_angle(p.x, p.y)) & (AZ_BINS - 1);
      },
      [&](int i, int d) {
        const float4 p = out[i];
        azs[d] = make_float4(p.x, p.y, p.z, __int_as_float(i));
      },
      false);
  for (int t = threadIdx.x; t <= AZ_BINS; t += blockDim.x) azo[t] = sh.dbase[t];
}

// ring-major concatenation: index lists, feature clouds, ring offsets of the clouds that become the next
// scan's targets (corner_last_ = less_sharp, surf_last_ = less_flat, :531-534)
__global__ void __launch_bounds__(128)
lo_finalize_kernel(const float4 *__restrict__ seg_cloud, const int *__restrict__ ring_feat_cnt, const int *__restrict__ ring_sharp,
                   const int *__restrict__ ring_less_sharp, const int *__restrict__ ring_flat, const float4 *__restrict__ lf_stage,
                   const int *__restrict__ start_ring, int *__restrict__ sharp_idx, int *__restrict__ less_sharp_idx,
                   int *__restrict__ flat_idx, float4 *__restrict__ sharp, float4 *__restrict__ flat,
                   float4 *__restrict__ less_sharp, float4 *__restrict__ less_flat, int *__restrict__ ls_ring_off,
                   int *__restrict__ lf_ring_off, int *__restrict__ n_feat, const float4 *__restrict__ az_stage,
                   float4 *__restrict__ az_pts, int R, int RC) {
  const int ring = blockIdx.x, b = blockIdx.y, br = b * R + ring;
  const size_t base = (size_t)b * RC;
  __shared__ int off[4], cnt[4];
  if (threadIdx.x < 32) {
    int a[4] = {0, 0, 0, 0};
    for (int t = threadIdx.x; t < ring; t += 32)
      for (int q = 0; q < 4; ++q) a[q] += ring_feat_cnt[(b * R + t) * 4 + q];
    for (int q = 0; q < 4; ++q) a[q] = warp_sum_i(a[q]);
    if (threadIdx.x == 0)
      for (int q = 0; q < 4; ++q) { off[q] = a[q]; cnt[q] = ring_feat_cnt[br * 4 + q]; }
  }
  __syncthreads();
  const int cS = R * 12, cL = R * 120, cF = R * 24;
  for (int t = threadIdx.x; t < cnt[0]; t += blockDim.x) {
    const int idx = ring_sharp[br * 12 + t];
    sharp_idx[b * cS + off[0] + t] = idx;
    sharp[(size_t)b * cS + off[0] + t] = seg_cloud[base + idx];
  }
  for (int t = threadIdx.x; t < cnt[1]; t += blockDim.x) {
    const int idx = ring_less_sharp[br * 120 + t];
    less_sharp_idx[b * cL + off[1] + t] = idx;
    less_sharp[(size_t)b * cL + off[1] + t] = seg_cloud[base + idx];
  }
  for (int t = threadIdx.x; t < cnt[2]; t += blockDim.x) {
    const int idx = ring_flat[br * 24 + t];
    flat_idx[b * cF + off[2] + t] = idx;
    flat[(size_t)b * cF + off[2] + t] = seg_cloud[base + idx];
  }
  const int lo = max(start_ring[br] - 5, 0);
  for (int t = threadIdx.x; t < cnt[3]; t += blockDim.x) {
    less_flat[base + off[3] + t] = lf_stage[base + lo + t];
    az_pts[base + off[3] + t] = az_stage[base + lo + t];  // azimuth-binned copy of the ring, aligned with less_flat
  }
  if (threadIdx.x == 0) {
    ls_ring_off[b * (R + 1) + ring] = off[1];
    lf_ring_off[b * (R + 1) + ring] = off[3];
    if (ring == R - 1) {
      ls_ring_off[b * (R + 1) + R] = off[1] + cnt[1];
      lf_ring_off[b * (R + 1) + R] = off[3] + cnt[3];
      n_feat[b * 4 + 0] = off[0] + cnt[0];
      n_feat[b * 4 + 1] = off[1] + cnt[1];
      n_feat[b * 4 + 2] = off[2] + cnt[2];
      n_feat[b * 4 + 3] = off[3] + cnt[3];
    }
  }
}

}  // namespace

int lo_extract_device(AlegoHandle *h) {
  const int B = h->B, R = h->R, C = h->C, RC = h->RC;
  cudaStream_t s = h->stream;
  { LAUNCH(h, "lo_curv_occl");
    lo_curv_occl_kernel<<<dim3(min(div_up(RC, CURV_TILE), 192), B), CURV_TILE, 0, s>>>(h->seg_range, h->seg_col, h->M, h->curv, h->picked0,
                                                                             h->flabel, h->sort_idx, RC); }
  const int sort_cap = ((C / 6 + 8) + 3) & ~3;
  const size_t sort_smem = (size_t)SORT_WARPS * ((size_t)sort_cap * 14 + 16);
  static bool sort_attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!sort_attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(lo_sort_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    sort_attr_set[h->dev] = true;
  }
  { LAUNCH(h, "lo_sort_segments");
    lo_sort_segments_kernel<<<div_up(B * R * 6, SORT_WARPS), SORT_WARPS * 32, sort_smem, s>>>(h->curv, h->start_ring, h->end_ring,
                                                                                          h->sort_scratch, h->sort_idx, B, R, RC, sort_cap); }
  const int pkcap = (C + 32 + 15) & ~15;
  { LAUNCH(h, "lo_select");
    lo_select_kernel<<<dim3(div_up(R, SEL_WARPS), B), SEL_WARPS * 32, (size_t)SEL_WARPS * pkcap, s>>>(
        h->seg_col, h->seg_ground, h->curv, h->sort_idx, h->start_ring, h->end_ring, h->picked0, h->picked, h->flabel,
        h->ring_feat_cnt, h->ring_sharp, h->ring_less_sharp, h->ring_flat, R, RC, pkcap); }
  { LAUNCH(h, "lo_lfv_keys");
    lo_lfv_keys_kernel<<<dim3(R, B), LFV_WARPS * 32, 0, s>>>(h->seg_cloud, h->flabel, h->start_ring, h->end_ring, h->lf_stage,
                                                             h->sort_scratch, h->lfv_state, R, RC, (float)h->P.less_flat_leaf); }
  const int rc_q = vox_order_lists_by_warp(h, h->lfv_state, B * R, h->sort_scratch, h->lfv_keys, s, "lo_lfv_order", R, C);
  if (rc_q != ALEGO_OK) return rc_q;
  { LAUNCH(h, "lo_lfv_finish");
    lo_lfv_finish_kernel<<<dim3(R, B), LFV_WARPS * 32, 0, s>>>(h->seg_cloud, h->start_ring, h->lf_stage, h->ring_feat_cnt, h->sort_scratch,
                                                               h->lfv_keys, h->lfv_state, h->az_stage, h->az_off[h->cur], R, RC); }
  const int cur = h->cur;
  { LAUNCH(h, "lo_finalize");
    lo_finalize_kernel<<<dim3(R, B), 128, 0, s>>>(h->seg_cloud, h->ring_feat_cnt, h->ring_sharp, h->ring_less_sharp, h->ring_flat,
                                                 h->lf_stage, h->start_ring, h->sharp_idx, h->less_sharp_idx, h->flat_idx, h->sharp,
                                                 h->flat, h->less_sharp[cur], h->less_flat[cur], h->ls_ring_off[cur],
                                                 h->lf_ring_off[cur], h->n_feat, h->az_stage, h->az_pts[cur], R, RC); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
