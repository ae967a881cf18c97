// sort_voxel.cuh — block-wide stable radix sort of (voxel key, input position) words and the
// pcl::VoxelGrid<PointXYZI> equivalent built on it (call sites in the reference: laserOdometry.cpp:288-293,
// laserMapping.cpp:325-342).
//
// VoxelGrid semantics restated (PCL 1.8-1.10 voxel_grid.hpp): bounding box in float, inverse leaf 1.0f/leaf,
// min_b = floor(min*inv), div_b = max_b-min_b+1, voxel key ijk0 + ijk1*div0 + ijk2*div0*div1 with
// ijk = int(floor(p*inv) - float(min_b)); points sorted by key; one output per occupied voxel in ascending key
// order = float sums of x,y,z,intensity divided by the float count; if the index space overflows int32 the
// input is returned unchanged.  PCL sorts with std::sort on the key alone, so the summation order inside a
// voxel is whatever libstdc++'s introsort leaves.  With a VoxExact scratch the record list first goes through the
// partition phase of that very algorithm (introsort_block.cuh) and then through the stable radix sort below, which
// together give std::sort's exact record order — centroids equal PCL's bit for bit.  Without it (and for inputs that
// contain non-finite points, which never occur on the hot path) the points of a voxel are summed in ascending input
// order: deterministic, equal to the oracle's "stable" variant, within 2e-5 m of PCL's order.
//
// One CTA of NW warps handles one cloud (NW = 4 for one ring of the less-flat cloud, 32 for LaserMapping's clouds).
// The sort is an LSD radix sort with 8-bit digits over only the significant bits of the key (a 21-bit voxel index:
// 3 passes; a pass whose digit is constant over the cloud is skipped).  Every warp owns a contiguous slice of the list
// (stability); the rank of a word inside its 32-word chunk comes from __match_any_sync (the lanes that share its digit),
// so a pass costs ~15 warp instructions per chunk, and the per-warp digit counts (NW x 256 in shared memory) are turned
// into destinations by one column sweep + one 256-entry scan.  The (key, position) words ping-pong between two global
// buffers (L2 resident).
#pragma once
#include "common.cuh"
#include "introsort_block.cuh"

typedef unsigned long long u64;
#define VOX_BITS 8
#define VOX_RADIX 256
#define VOX_WIN 64

struct VoxFrame {
  float inv;
  int min_b[3];
  int mul[3];
  int overflow;
  int key_bits;  // bits that hold every voxel index AND the index one past the last (the key of non-finite points)
  int n_cells;
};

// Scratch of the exact (PCL == std::sort) record order.  e_smem: optional shared-memory home of the record list while it is
// partitioned (e_cap records; longer lists stay in global memory).  pos / lists: n ints and 4 x list_cap range records; when
// pos is null both are carved out of the second ping-pong key buffer, which is idle during the partition phase.
struct VoxExact {
  u64 *e_smem;
  int e_cap;
  int *pos;
  uint2 *lists;
  int list_cap;
  IsbShared *isb;
  unsigned short *wpos;  // optional shared memory: (warps of the CTA) x ISB_REG swap positions of short ranges
  // optional: work-sharing scratch (introsort_block.cuh block_introsort_ws) for lists longer than e_cap — one long list ordered
  // by all the warps of the CTA through a task queue instead of level by level
  IswShared *wq;
  IswBig *wbig;
  unsigned short *wpos_all;
  u64 *wbuf_all;
};

template <int NW>
struct VoxShared {
  float4 win_pt[NW][VOX_WIN];    // per-warp window of the sorted list (centroid pass)
  unsigned win_key[NW][VOX_WIN];
  int hist[NW][VOX_RADIX];       // per-warp digit counts, then the warp's offset inside each digit
  int dbase[VOX_RADIX + 1];      // first destination of each digit
  int wsum[NW];
  float red[NW][6];
  VoxFrame frame;
  int flag, n, nv;
};

__device__ __forceinline__ VoxFrame vox_frame(const float mn[3], const float mx[3], float leaf) {
  VoxFrame f;
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
  } else {  // degenerate extents (the reference's index arithmetic wraps): sort all 32 key bits
    f.n_cells = -1;
    f.key_bits = 32;
  }
  return f;
}

__device__ __forceinline__ unsigned vox_key(const float4 &p, const VoxFrame &f) {
  const int i0 = (int)(floorf(p.x * f.inv) - (float)f.min_b[0]);
  const int i1 = (int)(floorf(p.y * f.inv) - (float)f.min_b[1]);
  const int i2 = (int)(floorf(p.z * f.inv) - (float)f.min_b[2]);
  return (unsigned)(i0 * f.mul[0] + i1 * f.mul[1] + i2 * f.mul[2]);
}

// slice of warp w when n items are split into NW contiguous, chunk-aligned slices
template <int NW>
__device__ __forceinline__ void vox_slice(int n, int w, int &lo, int &hi) {
  const int S = (((n + NW - 1) / NW) + 31) & ~31;
  lo = min(w * S, n);
  hi = min(lo + S, n);
}

// One stable counting-sort pass over items [0, n).  digit(i) in [0, VOX_RADIX) is evaluated twice per item (count, then
// scatter); emit(i, dst) stores item i at its destination.  After the call sh->dbase[d] is the first destination of
// digit d (dbase[VOX_RADIX] = n).  With allow_skip a pass whose items all share one digit is the identity: nothing is
// emitted and false is returned.  All threads of the block must call; ends with a barrier.
template <int NW, class Digit, class Emit>
__device__ __forceinline__ bool block_counting_pass(int n, VoxShared<NW> *sh, Digit digit, Emit emit, bool allow_skip) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int lo, hi;
  vox_slice<NW>(n, w, lo, hi);
  int *hist = sh->hist[w];
  for (int t = lane; t < VOX_RADIX; t += 32) hist[t] = 0;
  __syncwarp();
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < hi;
    const int d = valid ? digit(i) : VOX_RADIX;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt_mask) == 0) hist[d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps (a warp's words of a digit follow those of the warps before it) + total
  for (int d = threadIdx.x; d < VOX_RADIX; d += NW * 32) {
    int run = 0;
#pragma unroll 4
    for (int ww = 0; ww < NW; ++ww) {
      const int t = sh->hist[ww][d];
      sh->hist[ww][d] = run;
      run += t;
    }
    sh->dbase[d] = run;
  }
  __syncthreads();
  if (w == 0) {  // exclusive scan of the 256 digit totals, 8 per lane
    int c[8], sum = 0;
    bool one = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      c[q] = sh->dbase[lane * 8 + q];
      one |= c[q] == n;
      sum += c[q];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int run = inc - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sh->dbase[lane * 8 + q] = run;
      run += c[q];
    }
    const bool any_one = __any_sync(0xffffffffu, one);
    if (lane == 31) { sh->dbase[VOX_RADIX] = inc; sh->flag = any_one ? 1 : 0; }
  }
  __syncthreads();
  if (allow_skip && sh->flag) return false;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < hi;
    const int d = valid ? digit(i) : VOX_RADIX;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & lt_mask);
    const int dst = valid ? sh->dbase[d] + hist[d] + rank : 0;
    __syncwarp();
    if (valid && rank == 0) hist[d] += __popc(peers);
    if (valid) emit(i, dst);
    __syncwarp();
  }
  __syncthreads();
  return true;
}

// Stable sort of a[0..n) by bits [32, 32+key_bits).  a, b: ping-pong buffers in global memory.  Returns the buffer that
// holds the sorted list.  All threads of the block must call; ends with a barrier.
template <int NW>
__device__ u64 *block_radix_sort8(u64 *a, u64 *b, int n, int key_bits, VoxShared<NW> *sh) {
  for (int shift = 32; shift < 32 + key_bits; shift += VOX_BITS) {
    const u64 *src = a;
    u64 *dst = b;
    const bool moved = block_counting_pass<NW>(
        n, sh, [&](int i) { return (int)((src[i] >> shift) & (VOX_RADIX - 1)); }, [&](int i, int d) { dst[d] = src[i]; }, true);
    if (moved) {
      u64 *t = a;
      a = b;
      b = t;
    }
  }
  return a;
}

// One output per run of equal voxel keys among the sorted keys[0..nv); this warp handles the runs that START in
// [lo, hi) (a chunk-aligned slice) and writes them from out[run_base] on, in key order.  The low word of a key is the
// index of its point in pts.  A run is summed by the lane that owns its first element, strictly in list order (float
// sums).  The 64 list entries starting at the chunk are staged in shared memory first (independent, coalesced loads), so
// the sequential run walk does not chain two global-memory round trips per element; only a run reaching beyond the
// window continues from global memory.  count_only: just count the runs.  Returns the number of runs of the slice.
__device__ __forceinline__ int warp_vox_centroids(const u64 *keys, int lo, int hi, int nv, const float4 *__restrict__ pts, float4 *out,
                                                  int run_base, unsigned *win_key, float4 *win_pt, bool count_only) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int runs = 0;
  // voxel key of the element before the slice (no voxel has key 0xffffffff: pad keys sort behind nv)
  unsigned prev_last = lo > 0 && lo < nv ? (unsigned)(keys[lo - 1] >> 32) : 0xffffffffu;
  for (int t0 = lo; t0 < hi; t0 += 32) {
    const int t = t0 + lane;
    if (count_only) {
      const unsigned vk = t < hi ? (unsigned)(keys[t] >> 32) : 0u;
      unsigned pv = __shfl_up_sync(0xffffffffu, vk, 1);
      if (lane == 0) pv = prev_last;
      runs += __popc(__ballot_sync(0xffffffffu, t < hi && vk != pv));
      prev_last = __shfl_sync(0xffffffffu, vk, 31);
      continue;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u = t0 + h * 32 + lane;
      if (u < nv) {
        const u64 k = keys[u];
        win_key[h * 32 + lane] = (unsigned)(k >> 32);
        win_pt[h * 32 + lane] = pts[(unsigned)(k & 0xffffffffu)];
      }
    }
    __syncwarp();
    const int wend = min(nv - t0, VOX_WIN);  // valid window entries
    bool head = false;
    unsigned vk = 0;
    if (t < hi) {
      vk = win_key[lane];
      head = vk != (lane == 0 ? prev_last : win_key[lane - 1]);
    }
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (head) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      int w = lane;
      for (; w < wend && win_key[w] == vk; ++w) {
        const float4 p = win_pt[w];
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
        ++cnt;
      }
      if (w == VOX_WIN) {  // the run may continue beyond the window
        for (int u = t0 + VOX_WIN; u < nv; ++u) {
          const u64 ku = keys[u];
          if ((unsigned)(ku >> 32) != vk) break;
          const float4 p = pts[(unsigned)(ku & 0xffffffffu)];
          sx += p.x; sy += p.y; sz += p.z; si += p.w;
          ++cnt;
        }
      }
      const float fn = (float)cnt;
      out[run_base + runs + __popc(hm & lt_mask)] = make_float4(sx / fn, sy / fn, sz / fn, si / fn);
    }
    runs += __popc(hm);
    prev_last = win_key[min(31, hi - t0 - 1)];
  }
  return runs;
}

// What the key stage of a VoxelGrid leaves for the later stages when they run in separate kernels (per cloud)
struct VoxState {
  VoxFrame frame;
  int n;      // records in the list
  int nv;     // records with a finite point (nv == n: the list holds no pad keys)
  int done;   // 1: the output is already final (empty input, or PCL's "leaf size too small" pass-through), n_out = n
  long long off;    // where the cloud's record list starts in the batch-wide key buffer A ...
  long long off_b;  // ... and its scratch (the second ping-pong buffer) in buffer B
};

// Stage 1 of the block-wide VoxelGrid of the points {pts[i] : i < n_src, member(i)} in ascending i: membership, bounding box,
// (voxel key, input position) records into ka.  MASKED = false: every i < n_src takes part (member is not called).  MASKED = true:
// member_bits / chunk_base are ceil(n_src/32) words of shared memory.  Returns through sh: frame, n, nv; the function result is
// true when the output is already final (see VoxState::done; out then holds sh->n points).  All threads of the block must call.
template <int NW, bool MASKED, class Member>
__device__ bool block_voxel_keys(const float4 *__restrict__ pts, int n_src, Member member, float leaf, u64 *ka, float4 *out,
                                 VoxShared<NW> *sh, unsigned *member_bits, int *chunk_base) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int nch = (n_src + 31) >> 5;
  // ---- membership + bounding box over the finite members (getMinMax3D)
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (int ch = w; ch < nch; ch += NW) {
    const int i = ch * 32 + lane;
    bool m = i < n_src && (!MASKED || member(i));
    if (m) {
      const float4 p = pts[i];
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      } else if (MASKED) {
        m = false;  // PCL skips non-finite points of a non-dense cloud before it sorts: they are not members
      }
    }
    if (MASKED) {
      const unsigned mm = __ballot_sync(0xffffffffu, m);
      if (lane == 0) { member_bits[ch] = mm; chunk_base[ch] = __popc(mm); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if (lane == 0)
    for (int a = 0; a < 3; ++a) { sh->red[w][a] = mn[a]; sh->red[w][3 + a] = mx[a]; }
  __syncthreads();
  if (w == 0) {
    int n = n_src;
    if (MASKED) {  // exclusive scan of the chunk populations
      int carry = 0;
      for (int c0 = 0; c0 < nch; c0 += 32) {
        const int v = c0 + lane < nch ? chunk_base[c0 + lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (c0 + lane < nch) chunk_base[c0 + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      n = carry;
    }
    if (lane == 0) {
      for (int ww = 1; ww < NW; ++ww)
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], sh->red[ww][a]); mx[a] = fmaxf(mx[a], sh->red[ww][3 + a]); }
      sh->frame = vox_frame(mn, mx, leaf);
      sh->n = n;
      sh->nv = n;
    }
  }
  __syncthreads();
  const int n = sh->n;
  if (n <= 0) return true;
  const VoxFrame frame = sh->frame;
  if (frame.overflow) {  // "leaf size is too small": output = input
    for (int ch = w; ch < nch; ch += NW) {
      const int i = ch * 32 + lane;
      if (MASKED) {
        const unsigned mm = member_bits[ch];
        if ((mm >> lane) & 1u) out[chunk_base[ch] + __popc(mm & lt_mask)] = pts[i];
      } else if (i < n_src) {
        out[i] = pts[i];
      }
    }
    __syncthreads();
    return true;
  }
  // ---- (voxel key, input position) words; non-finite points get the key one past the last voxel and sort to the end
  const unsigned pad_key = (unsigned)frame.n_cells;
  int nfin = 0;
  for (int ch = w; ch < nch; ch += NW) {
    const int i = ch * 32 + lane;
    unsigned mm = 0xffffffffu;
    bool m = i < n_src;
    if (MASKED) {
      mm = member_bits[ch];
      m = (mm >> lane) & 1u;
    }
    if (m) {
      const float4 p = pts[i];
      unsigned vk = pad_key;
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        vk = vox_key(p, frame);
        ++nfin;
      }
      const int pos = MASKED ? chunk_base[ch] + __popc(mm & lt_mask) : i;
      ka[pos] = ((u64)vk << 32) | (unsigned)i;
    }
  }
  nfin = warp_sum_i(nfin);
  if (lane == 0) sh->wsum[w] = nfin;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int ww = 0; ww < NW; ++ww) s += sh->wsum[ww];
    sh->nv = s;
  }
  __syncthreads();
  return false;
}

// Stage 3: stable sort of the records by voxel key, then one centroid per run, in key order.  n / nv / frame as left by the key
// stage.  Returns the number of output points (same value in every thread).  All threads of the block must call.
template <int NW>
__device__ int block_voxel_finish(const float4 *__restrict__ pts, u64 *ka, u64 *kb, float4 *out, VoxShared<NW> *sh, int n, int nv,
                                  int key_bits) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u64 *keys = block_radix_sort8<NW>(ka, kb, n, key_bits, sh);
  // ---- centroids: every warp counts the runs that start in its slice, then writes them behind those of the warps before
  int lo, hi;
  vox_slice<NW>(nv, w, lo, hi);
  const int my_runs = warp_vox_centroids(keys, lo, hi, nv, pts, out, 0, sh->win_key[w], sh->win_pt[w], true);
  if (lane == 0) sh->wsum[w] = my_runs;
  __syncthreads();
  int before = 0, total = 0;
  for (int ww = 0; ww < NW; ++ww) {
    const int t = sh->wsum[ww];
    if (ww < w) before += t;
    total += t;
  }
  warp_vox_centroids(keys, lo, hi, nv, pts, out, before, sh->win_key[w], sh->win_pt[w], false);
  __syncthreads();
  return total;
}

// Does any voxel of the record list hold three or more points?  Only then does the order of the records inside a voxel matter:
// a float sum of one or two terms (starting from 0.f) is the same in either order.  Open-addressing table of 2n (key + 1, count)
// slots in `table` (16n bytes of scratch: the output buffer, idle until the centroid stage).  All threads of the block must call.
template <int NW>
__device__ bool block_any_voxel_ge3(const u64 *keys, int n, u64 *table, int *flag_smem) {
  const unsigned T = 2u * (unsigned)n;
  if (threadIdx.x == 0) *flag_smem = 0;
  for (unsigned t = threadIdx.x; t < T; t += NW * 32) table[t] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NW * 32) {
    const unsigned key1 = (unsigned)(keys[i] >> 32) + 1u;
    unsigned s = (unsigned)(((u64)(key1 * 2654435761u) * T) >> 32);
    for (unsigned probe = 0; probe < T; ++probe) {
      u64 cur = *reinterpret_cast<volatile u64 *>(table + s);
      if ((unsigned)(cur >> 32) == 0u) {
        const u64 old = atomicCAS(table + s, 0ull, ((u64)key1 << 32) | 1ull);
        if (old == 0ull) break;
        cur = old;
      }
      if ((unsigned)(cur >> 32) == key1) {
        const u64 before = atomicAdd(table + s, 1ull);
        if ((unsigned)(before & 0xffffffffu) + 1u >= 3u) *flag_smem = 1;
        break;
      }
      s = s + 1u == T ? 0u : s + 1u;
    }
  }
  __syncthreads();
  return *flag_smem != 0;
}

// The whole VoxelGrid in one call (one CTA per cloud).  ka / kb: two buffers of (number of members) words in global memory.
// out: room for as many points.  ex: scratch of the exact record order (stage 2 = introsort partition phase), or nullptr.
template <int NW, bool MASKED, class Member>
__device__ int block_voxel_grid8(const float4 *__restrict__ pts, int n_src, Member member, float leaf, u64 *ka, u64 *kb, float4 *out,
                                 VoxShared<NW> *sh, unsigned *member_bits, int *chunk_base, const VoxExact *ex = nullptr) {
  if (block_voxel_keys<NW, MASKED>(pts, n_src, member, leaf, ka, out, sh, member_bits, chunk_base)) return max(sh->n, 0);
  const int n = sh->n, nv = sh->nv;
  u64 *list = ka;
  // ---- std::sort's record order = its partition phase, then a stable sort by key (see introsort_block.cuh)
  if (ex && nv == n && n > 16) {
    if (ex->e_smem && n <= ex->e_cap) {  // partition in shared memory (the first radix pass then reads it from there)
      for (int t = threadIdx.x; t < n; t += NW * 32) ex->e_smem[t] = ka[t];
      __syncthreads();
      list = ex->e_smem;
    }
    int *pos = ex->pos;
    uint2 *lists = ex->lists;
    int list_cap = ex->list_cap;
    if (!pos) {  // carve the scratch out of the idle second key buffer: n ints, then 4 lists of n / 17 + 1 ranges
      pos = reinterpret_cast<int *>(kb);
      lists = reinterpret_cast<uint2 *>(kb + (n + 1) / 2);
      list_cap = n / 17 + 1;
    }
    if (ex->wq && list == ka) block_introsort_ws<NW>(list, pos, n, ex->wq, ex->wbig, ex->wpos_all, ex->wbuf_all);
    else block_introsort_partitions<NW>(list, pos, n, lists, list_cap, ex->isb, ex->wpos);
    __syncthreads();
  }
  return block_voxel_finish<NW>(pts, list, kb, out, sh, n, nv, sh->frame.key_bits);
}
