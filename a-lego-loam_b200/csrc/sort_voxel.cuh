// sort_voxel.cuh — block-wide bitonic sort of 64-bit composite keys and the pcl::VoxelGrid<PointXYZI>
// equivalent built on it (call sites in the reference: laserOdometry.cpp:288-293, laserMapping.cpp:325-342).
//
// VoxelGrid semantics restated (PCL 1.8-1.10 voxel_grid.hpp): bounding box in float, inverse leaf 1.0f/leaf,
// min_b = floor(min*inv), div_b = max_b-min_b+1, voxel key ijk0 + ijk1*div0 + ijk2*div0*div1 with
// ijk = int(floor(p*inv) - float(min_b)); points sorted by key; one output per occupied voxel in ascending key
// order = float sums of x,y,z,intensity divided by the float count; if the index space overflows int32 the
// input is returned unchanged.  PCL sorts with std::sort on the key alone, so the summation order inside a
// voxel is an accident of introsort; here points of a voxel are summed in ascending input order (composite
// key = voxel<<32 | input position), which is deterministic and equals the oracle's "stable" variant bit for
// bit.  One CTA handles one cloud.
#pragma once
#include "common.cuh"

typedef unsigned long long u64;
#define VOX_PAD 0xFFFFFFFFFFFFFFFFull

// one compare-exchange stage (k = merge size, j = stride) on s[0..n) where element t has global position
// gbase + t (direction depends on the global position)
__device__ __forceinline__ void bitonic_stage(u64 *s, int n, int k, int j, int gbase) {
  for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int p = i | j;
    const bool up = (((gbase + i) & k) == 0);
    const u64 a = s[i], b = s[p];
    if ((a > b) == up) {
      s[i] = b;
      s[p] = a;
    }
  }
}

// sort npad (power of two) keys resident in shared memory, ascending
__device__ __forceinline__ void block_bitonic_smem(u64 *s, int npad) {
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      bitonic_stage(s, npad, k, j, 0);
      __syncthreads();
    }
}

// sort npad (power of two) keys in global memory with one CTA, staging chunks of `ch` keys in shared memory
__device__ __forceinline__ void block_bitonic_global(u64 *g, int npad, u64 *s, int ch) {
  if (npad <= ch) {
    for (int t = threadIdx.x; t < npad; t += blockDim.x) s[t] = g[t];
    __syncthreads();
    block_bitonic_smem(s, npad);
    for (int t = threadIdx.x; t < npad; t += blockDim.x) g[t] = s[t];
    __syncthreads();
    return;
  }
  // phase 1: every chunk fully sorted in shared memory (direction alternates with the chunk's global position)
  for (int c0 = 0; c0 < npad; c0 += ch) {
    for (int t = threadIdx.x; t < ch; t += blockDim.x) s[t] = g[c0 + t];
    __syncthreads();
    for (int k = 2; k <= ch; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        bitonic_stage(s, ch, k, j, c0);
        __syncthreads();
      }
    for (int t = threadIdx.x; t < ch; t += blockDim.x) g[c0 + t] = s[t];
    __syncthreads();
  }
  // phase 2: merges larger than a chunk — wide strides in global memory, the rest per chunk in shared memory
  for (int k = ch << 1; k <= npad; k <<= 1) {
    for (int j = k >> 1; j >= ch; j >>= 1) {
      bitonic_stage(g, npad, k, j, 0);
      __syncthreads();
    }
    for (int c0 = 0; c0 < npad; c0 += ch) {
      for (int t = threadIdx.x; t < ch; t += blockDim.x) s[t] = g[c0 + t];
      __syncthreads();
      for (int j = ch >> 1; j > 0; j >>= 1) {
        bitonic_stage(s, ch, k, j, c0);
        __syncthreads();
      }
      for (int t = threadIdx.x; t < ch; t += blockDim.x) g[c0 + t] = s[t];
      __syncthreads();
    }
  }
}

struct VoxFrame {
  float inv;
  int min_b[3];
  int mul[3];
  int overflow;
  int n_valid;
};

// Block-wide VoxelGrid.  pts: n input points (shared or global memory).  keys: npad >= n composite keys, in
// shared memory when keys_in_smem (then npad <= stage capacity) else in global memory with `stage` (ch keys
// of shared memory) as the staging buffer.  out: room for n points.  red: >= 40 floats + 40 ints of shared
// scratch.  Returns the number of output points (same value in every thread).
static __device__ int block_voxel_grid(const float4 *pts, int n, float leaf, u64 *keys, int npad, bool keys_in_smem, u64 *stage,
                                int ch, float4 *out, float *redf, int *redi, VoxFrame *frame) {
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
  }
  __syncthreads();
  if (frame->overflow) {  // "leaf size is too small": output = input
    for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = pts[t];
    __syncthreads();
    return n;
  }
  const float inv = frame->inv;
  // ---- composite keys
  int nvalid_local = 0;
  for (int t = threadIdx.x; t < npad; t += blockDim.x) {
    u64 key = VOX_PAD;
    if (t < n) {
      const float4 p = pts[t];
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int i0 = (int)(floorf(p.x * inv) - (float)frame->min_b[0]);
        const int i1 = (int)(floorf(p.y * inv) - (float)frame->min_b[1]);
        const int i2 = (int)(floorf(p.z * inv) - (float)frame->min_b[2]);
        const int idx = i0 * frame->mul[0] + i1 * frame->mul[1] + i2 * frame->mul[2];
        key = ((u64)(unsigned)idx << 32) | (unsigned)t;
        ++nvalid_local;
      }
    }
    keys[t] = key;
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
  // ---- sort
  if (keys_in_smem) block_bitonic_smem(keys, npad);
  else block_bitonic_global(keys, npad, stage, ch);
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
