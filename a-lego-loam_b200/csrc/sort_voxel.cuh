// sort_voxel.cuh — block-wide stable radix sort of (voxel key, input position) words and the
// pcl::VoxelGrid<PointXYZI> equivalent built on it (call sites in the reference: laserOdometry.cpp:288-293,
// laserMapping.cpp:325-342).
//
// VoxelGrid semantics restated (PCL 1.8-1.10 voxel_grid.hpp): bounding box in float, inverse leaf 1.0f/leaf,
// min_b = floor(min*inv), div_b = max_b-min_b+1, voxel key ijk0 + ijk1*div0 + ijk2*div0*div1 with
// ijk = int(floor(p*inv) - float(min_b)); points sorted by key; one output per occupied voxel in ascending key
// order = float sums of x,y,z,intensity divided by the float count; if the index space overflows int32 the
// input is returned unchanged.  PCL sorts with std::sort on the key alone, so the summation order inside a
// voxel is an accident of introsort; here points of a voxel are summed in ascending input order (a STABLE sort by
// key of the position-ordered list), which is deterministic and equals the oracle's "stable" variant bit for bit.
// One CTA handles one cloud.
//
// The sort is an LSD radix sort, 4 bits per pass over only the significant bits of the key (a cloud spans a few
// hundred cells per axis: 4-6 passes): each thread owns a contiguous slice of the list (stability), counts its digits
// in a private shared-memory column, a register-level vector scan turns the 16 x nthreads counts into destinations,
// and the slice is scattered in order.  3 barriers per pass instead of one per bitonic stage (55-120 of them).
#pragma once
#include "common.cuh"

typedef unsigned long long u64;
#define VOX_RADIX_BITS 4
#define VOX_RADIX 16

// shared-memory footprint of the sort scratch for a block of `nt` threads with counter type CT
template <typename CT>
__host__ __device__ constexpr size_t radix_scratch_bytes(int nt) {
  return (size_t)VOX_RADIX * nt * sizeof(CT) + (size_t)(VOX_RADIX * 33) * sizeof(int);
}

// Stable sort of a[0..n) by bits [32, 32+key_bits).  a, b: ping-pong buffers (shared or global memory).  scratch:
// radix_scratch_bytes<CT>(blockDim.x) bytes of shared memory, 4-byte aligned.  CT must hold n.  Returns the buffer
// that holds the sorted list.  All threads of the block must call; ends with a barrier.
template <typename CT>
__device__ u64 *block_radix_sort(u64 *a, u64 *b, int n, int key_bits, void *scratch) {
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
  CT *cnt = reinterpret_cast<CT *>(scratch);                           // [16][nt], column tid is private
  int *wtot = reinterpret_cast<int *>(cnt + (size_t)VOX_RADIX * nt);   // [16][32] per-warp totals -> exclusive prefixes
  int *dbase = wtot + VOX_RADIX * 32;                                  // [16] first destination of each digit
  const int E = ((n + nt - 1) / nt) | 1;  // odd slice length: conflict-free 8-byte shared-memory accesses
  const int lo = min(tid * E, n), hi = min(lo + E, n);
  for (int shift = 32; shift < 32 + key_bits; shift += VOX_RADIX_BITS) {
#pragma unroll
    for (int d = 0; d < VOX_RADIX; ++d) cnt[d * nt + tid] = 0;
    for (int i = lo; i < hi; ++i) ++cnt[(int)((a[i] >> shift) & (VOX_RADIX - 1)) * nt + tid];
    int c[VOX_RADIX], inc[VOX_RADIX];
#pragma unroll
    for (int d = 0; d < VOX_RADIX; ++d) {
      c[d] = (int)cnt[d * nt + tid];
      int v = c[d];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      inc[d] = v;
    }
    if (lane == 31) {
#pragma unroll
      for (int d = 0; d < VOX_RADIX; ++d) wtot[d * 32 + wid] = inc[d];
    }
    __syncthreads();
    if (wid == 0) {
      int run = 0;
      if (lane < VOX_RADIX)
        for (int w = 0; w < nw; ++w) {
          const int t = wtot[lane * 32 + w];
          wtot[lane * 32 + w] = run;
          run += t;
        }
      int ex = run;  // digit totals -> exclusive scan over the 16 digits
#pragma unroll
      for (int o = 1; o < VOX_RADIX; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ex, o);
        if (lane >= o) ex += t;
      }
      if (lane < VOX_RADIX) dbase[lane] = ex - run;
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < VOX_RADIX; ++d) cnt[d * nt + tid] = (CT)(dbase[d] + wtot[d * 32 + wid] + inc[d] - c[d]);
    for (int i = lo; i < hi; ++i) {
      const u64 e = a[i];
      const int d = (int)((e >> shift) & (VOX_RADIX - 1));
      const int dst = (int)cnt[d * nt + tid];
      cnt[d * nt + tid] = (CT)(dst + 1);
      b[dst] = e;
    }
    __syncthreads();
    u64 *t = a;
    a = b;
    b = t;
  }
  return a;
}

struct VoxFrame {
  float inv;
  int min_b[3];
  int mul[3];
  int overflow;
  int n_valid;
  int key_bits;  // bits that hold every voxel index AND the index one past the last (the key of non-finite points)
  int n_cells;
};

// Block-wide VoxelGrid.  pts: n input points (shared or global memory).  keys_a / keys_b: two buffers of n words
// (shared or global memory).  scratch: radix_scratch_bytes<CT>(blockDim.x) bytes of shared memory.  out: room for n
// points.  redf: >= 6*warps floats, redi: >= 40 ints of shared scratch.  Returns the number of output points (same
// value in every thread).
template <typename CT>
static __device__ int block_voxel_grid(const float4 *pts, int n, float leaf, u64 *keys_a, u64 *keys_b, void *scratch, float4 *out,
                                       float *redf, int *redi, VoxFrame *frame) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (n <= 0) return 0;
  // ---- bounding box over finite points (getMinMax3D)
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const float4 p = pts[t];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  __syncthreads();
  if (lane == 0)
    for (int a = 0; a < 3; ++a) { redf[wid * 6 + a] = mn[a]; redf[wid * 6 + 3 + a] = mx[a]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w)
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], redf[w * 6 + a]); mx[a] = fmaxf(mx[a], redf[w * 6 + 3 + a]); }
    const float inv = 1.0f / leaf;
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                    dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    frame->overflow = (dx * dy * dz > 2147483647ll) || !(mx[0] >= mn[0]);
    frame->inv = inv;
    int div_b[3];
    for (int a = 0; a < 3; ++a) {
      frame->min_b[a] = (int)floorf(mn[a] * inv);
      div_b[a] = (int)floorf(mx[a] * inv) - frame->min_b[a] + 1;
    }
    frame->mul[0] = 1; frame->mul[1] = div_b[0]; frame->mul[2] = div_b[0] * div_b[1];
    const long long cells = (long long)div_b[0] * div_b[1] * div_b[2];
    if (cells > 0 && cells < 2147483647ll) {
      frame->n_cells = (int)cells;
      int kb = 1;
      while ((1ll << kb) <= cells) ++kb;
      frame->key_bits = kb;
    } else {  // degenerate extents (the reference's index arithmetic wraps): sort all 32 key bits
      frame->n_cells = -1;
      frame->key_bits = 32;
    }
  }
  __syncthreads();
  if (frame->overflow) {  // "leaf size is too small": output = input
    for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = pts[t];
    __syncthreads();
    return n;
  }
  const float inv = frame->inv;
  // ---- (voxel key, input position) words; non-finite points get the key one past the last voxel and sort to the end
  int nvalid_local = 0;
  const unsigned pad_key = (unsigned)frame->n_cells;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    unsigned vk = pad_key;
    const float4 p = pts[t];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const int i0 = (int)(floorf(p.x * inv) - (float)frame->min_b[0]);
      const int i1 = (int)(floorf(p.y * inv) - (float)frame->min_b[1]);
      const int i2 = (int)(floorf(p.z * inv) - (float)frame->min_b[2]);
      vk = (unsigned)(i0 * frame->mul[0] + i1 * frame->mul[1] + i2 * frame->mul[2]);
      ++nvalid_local;
    }
    keys_a[t] = ((u64)vk << 32) | (unsigned)t;
  }
  nvalid_local = warp_sum_i(nvalid_local);
  if (lane == 0) redi[wid] = nvalid_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < nw; ++w) s += redi[w];
    frame->n_valid = s;
  }
  __syncthreads();
  const int nv = frame->n_valid;
  // ---- stable sort by voxel key
  const u64 *keys = block_radix_sort<CT>(keys_a, keys_b, n, frame->key_bits, scratch);
  // ---- one output per run of equal voxel keys, in key order
  int run_base = 0;
  for (int c0 = 0; c0 < nv; c0 += blockDim.x) {
    const int t = c0 + threadIdx.x;
    bool head = false;
    u64 k = 0;
    if (t < nv) {
      k = keys[t];
      head = (t == 0) || ((unsigned)(keys[t - 1] >> 32) != (unsigned)(k >> 32));
    }
    int total;
    const int ex = block_excl_scan(head ? 1 : 0, redi, &total);
    if (head) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      const unsigned vk = (unsigned)(k >> 32);
      for (int u = t; u < nv; ++u) {
        const u64 ku = keys[u];
        if ((unsigned)(ku >> 32) != vk) break;
        const float4 p = pts[(unsigned)(ku & 0xffffffffu)];
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
        ++cnt;
      }
      const float fn = (float)cnt;
      out[run_base + ex] = make_float4(sx / fn, sy / fn, sz / fn, si / fn);
    }
    run_base += total;
  }
  __syncthreads();
  return run_base;
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-synchronous variant for small clouds (one ring of the less-flat cloud, laserOdometry.cpp:288-293): ONE WARP does
// the whole VoxelGrid of a point subset, no block barriers, so a CTA packs several independent clouds.
//  * digit ranks come from __match_any_sync (peers sharing a digit) instead of per-thread counter columns: a pass costs
//    ~15 warp instructions per 32 words, independent of the block size;
//  * the digit histograms of ALL passes are taken in the sweep that builds the keys, so every pass is a single
//    read + scatter; passes whose digit is constant over the cloud (the high bits, usually) are skipped;
//  * the (key, position) words ping-pong between two global buffers (L2 resident: a ring is a few KB).
// Same ordering contract as block_voxel_grid: stable by voxel key, sums in ascending input position.
struct WarpVoxFrame {
  float inv;
  int min_b[3];
  int mul[3];
  int overflow;
  int key_bits;
  int n_cells;
};

__device__ __forceinline__ WarpVoxFrame warp_vox_frame(const float mn[3], const float mx[3], float leaf) {
  WarpVoxFrame f;
  const float inv = 1.0f / leaf;
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                  dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  f.overflow = (dx * dy * dz > 2147483647ll) || !(mx[0] >= mn[0]);
  f.inv = inv;
  int div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    f.min_b[a] = (int)floorf(mn[a] * inv);
    div_b[a] = (int)floorf(mx[a] * inv) - f.min_b[a] + 1;
  }
  f.mul[0] = 1; f.mul[1] = div_b[0]; f.mul[2] = div_b[0] * div_b[1];
  const long long cells = (long long)div_b[0] * div_b[1] * div_b[2];
  if (cells > 0 && cells < 2147483647ll) {
    f.n_cells = (int)cells;
    int kb = 1;
    while ((1ll << kb) <= cells) ++kb;
    f.key_bits = kb;
  } else {
    f.n_cells = -1;
    f.key_bits = 32;
  }
  return f;
}

__device__ __forceinline__ unsigned warp_vox_key(const float4 &p, const WarpVoxFrame &f) {
  const int i0 = (int)(floorf(p.x * f.inv) - (float)f.min_b[0]);
  const int i1 = (int)(floorf(p.y * f.inv) - (float)f.min_b[1]);
  const int i2 = (int)(floorf(p.z * f.inv) - (float)f.min_b[2]);
  return (unsigned)(i0 * f.mul[0] + i1 * f.mul[1] + i2 * f.mul[2]);
}

#define WVOX_BITS 8
#define WVOX_RADIX 256
#define WVOX_MAX_PASSES 4
// hist: [WVOX_MAX_PASSES][256] ints of shared memory private to the warp, zeroed by the caller before the key sweep and
// filled by warp_vox_hist_add.  8-bit digits: a 21-bit voxel index needs 3 passes; __match_any_sync ranks any digit width.
__device__ __forceinline__ void warp_vox_hist_add(int *hist, unsigned vk, bool valid, int npass) {
  const unsigned lt_mask = (1u << (threadIdx.x & 31)) - 1u;
  for (int p = 0; p < npass; ++p) {
    const int d = (int)((vk >> (p * WVOX_BITS)) & (WVOX_RADIX - 1));
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : WVOX_RADIX);
    if (valid && (peers & lt_mask) == 0) hist[p * WVOX_RADIX + d] += __popc(peers);
  }
  __syncwarp();
}

// Stable LSD sort of a[0..n) by bits [32, 32 + 8*npass) using the precomputed histograms (which it turns into digit
// offsets in place).  Returns the sorted buffer.
__device__ __forceinline__ u64 *warp_radix_sort(u64 *a, u64 *b, int n, int npass, int *hist) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int p = 0; p < npass; ++p) {
    int *dbase = hist + p * WVOX_RADIX;
    // exclusive scan of the 256 digit counts (8 per lane); a digit that holds every word makes the pass the identity
    int c[8], sum = 0;
    bool all_in_one = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      c[q] = dbase[lane * 8 + q];
      all_in_one |= c[q] == n;
      sum += c[q];
    }
    if (__any_sync(0xffffffffu, all_in_one)) continue;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int run = inc - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      dbase[lane * 8 + q] = run;
      run += c[q];
    }
    __syncwarp();
    const int shift = 32 + p * WVOX_BITS;
    for (int i0 = 0; i0 < n; i0 += 128) {
      u64 e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        e[u] = i < n ? a[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i0 + u * 32 >= n) break;
        const bool valid = i < n;
        const int d = (int)((e[u] >> shift) & (WVOX_RADIX - 1));
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : WVOX_RADIX);
        const int rank = __popc(peers & lt_mask);
        const int dst = valid ? dbase[d] + rank : 0;
        __syncwarp();
        if (valid && rank == 0) dbase[d] += __popc(peers);
        if (valid) b[dst] = e[u];
        __syncwarp();
      }
    }
    u64 *t = a;
    a = b;
    b = t;
  }
  __syncwarp();
  return a;
}

// One output per run of equal voxel keys among keys[0..nv) (sorted), in key order; the low word of a key is the index of
// its point in pts.  Returns the number of outputs.
__device__ __forceinline__ int warp_vox_centroids(const u64 *keys, int nv, const float4 *__restrict__ pts, float4 *out) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int run_base = 0;
  for (int t0 = 0; t0 < nv; t0 += 32) {
    const int t = t0 + lane;
    bool head = false;
    u64 k = 0;
    if (t < nv) {
      k = keys[t];
      head = (t == 0) || ((unsigned)(keys[t - 1] >> 32) != (unsigned)(k >> 32));
    }
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (head) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      const unsigned vk = (unsigned)(k >> 32);
      u64 ku = k;
      for (int u = t;;) {
        const float4 p = pts[(unsigned)(ku & 0xffffffffu)];
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
        ++cnt;
        if (++u >= nv) break;
        ku = keys[u];
        if ((unsigned)(ku >> 32) != vk) break;
      }
      const float fn = (float)cnt;
      out[run_base + __popc(hm & lt_mask)] = make_float4(sx / fn, sy / fn, sz / fn, si / fn);
    }
    run_base += __popc(hm);
  }
  return run_base;
}
