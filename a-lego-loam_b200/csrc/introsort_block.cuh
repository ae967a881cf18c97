// introsort_block.cuh — the PARTITION PHASE of libstdc++'s std::sort (std::__introsort_loop: median-of-3 to the front,
// __unguarded_partition, recurse right / iterate left, depth limit 2*floor(log2 n) with a heapsort fallback, ranges of <= 16
// left untouched) evaluated by a whole CTA on a list of 64-bit words whose HIGH 32 bits are the key.
//
// Why: pcl::VoxelGrid sorts its (voxel index, point index) records with std::sort and a key-only operator< (PCL
// voxel_grid.hpp; call sites laserOdometry.cpp:288-293, laserMapping.cpp:325-342), so the order of the points INSIDE a voxel —
// and with it the last bits of the float centroid sums — is whatever introsort leaves.  std::__final_insertion_sort is a stable
// sort of what the partition phase leaves, so
//     std::sort(list)  ==  stable_sort_by_key( partition_phase(list) )
// and the device gets PCL's exact record order from this routine followed by the stable radix sort of sort_voxel.cuh.
//
// Data-parallel form (same formulation as lo_sort_segments, checked against the real std::sort in
// tools/proto_parallel_introsort.py): all ranges of one recursion level are disjoint, so a level is processed as a work list.
// With pivot p = e[f] and the scan range (f, l): "left stoppers" are the positions whose key >= p in ascending order, "right
// stoppers" those whose key <= p in descending order; the sequential loop swaps the k-th left stopper with the k-th right
// stopper while the former lies left of the latter.  A left stopper at t is swapped iff (#right stoppers after t) > (#left
// stoppers before t); likewise for right stoppers; the cut is min(first left stopper that stays, leftmost right stopper that
// moves).  Counts come from ballots (ranges handled by one warp) or from warp-slice counts + a block scan (ranges >=
// ISB_BIG, handled by the whole CTA).  Every thread of the CTA must call block_introsort_partitions.
#pragma once
#include "common.cuh"
#include "stdsort_clone.cuh"

#define ISB_BIG 1536  // ranges at least this long are partitioned by the whole CTA, shorter ones by one warp each

struct IsbShared {
  int cnt_big[2], cnt_small[2];  // list lengths of the current / next level
  int next;                      // work counter of the small list
  int wl[32], wr[32];            // per-warp stopper counts of a cooperative partition
  int red_keep, red_swap, K;
};

__device__ __forceinline__ unsigned isb_key(unsigned long long e) { return (unsigned)(e >> 32); }

// children of a partitioned range: (f, cut) and (cut, l); only ranges longer than 16 are partitioned again
__device__ __forceinline__ void isb_push(IsbShared *sh, uint2 *big, uint2 *small_, int nxt, int f, int l) {
  const int len = l - f;
  if (len <= 16) return;
  if (len >= ISB_BIG) big[atomicAdd(&sh->cnt_big[nxt], 1)] = make_uint2((unsigned)f, (unsigned)l);
  else small_[atomicAdd(&sh->cnt_small[nxt], 1)] = make_uint2((unsigned)f, (unsigned)l);
}

// one warp partitions [f, l): returns the cut (same value in every lane).  pos: scratch of l - f ints owned by the range.
__device__ __forceinline__ int isb_warp_partition(unsigned long long *e, int *pos, int f, int l) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
  if (lane == 0) ssc::median_to_first(e + f, e + f + 1, e + f + (l - f) / 2, e + l - 1);
  __syncwarp();
  const unsigned p = isb_key(e[f]);
  const int lo = f + 1, m = l - lo;
  int *posL = pos + f + 1, *posR = pos + f + 1 + (m - m / 2);  // at most m / 2 swaps
  int totR = 0;
  for (int c = lo; c < l; c += 32) {
    const int t = c + lane;
    totR += __popc(__ballot_sync(full, t < l && isb_key(e[t]) <= p));
  }
  int runL = 0, runR = 0, K = 0, first_keep_L = 0x7fffffff, min_swap_R = l;
  for (int c = lo; c < l; c += 32) {
    const int t = c + lane;
    const bool v = t < l;
    const unsigned kt = v ? isb_key(e[t]) : 0u;
    const bool isL = v && kt >= p, isR = v && kt <= p;
    const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
    const int cL = runL + __popc(mL & lt_mask);           // left stoppers before t
    const int cR = totR - (runR + __popc(mR & le_mask));  // right stoppers after t
    const bool sL = isL && cR > cL, sR = isR && cL > cR;
    if (sL) posL[cL] = t;
    if (sR) posR[cR] = t;
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
  return min(first_keep_L, min_swap_R);
}

// the whole CTA (NW warps) partitions [f, l): every warp owns a contiguous, 32-aligned slice of (f, l)
template <int NW>
__device__ __forceinline__ int isb_block_partition(unsigned long long *e, int *pos, int f, int l, IsbShared *sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
  if (threadIdx.x == 0) {
    ssc::median_to_first(e + f, e + f + 1, e + f + (l - f) / 2, e + l - 1);
    sh->red_keep = 0x7fffffff;
    sh->red_swap = l;
    sh->K = 0;
  }
  __syncthreads();
  const unsigned p = isb_key(e[f]);
  const int lo = f + 1, m = l - lo;
  int *posL = pos + f + 1, *posR = pos + f + 1 + (m - m / 2);
  const int S = (((m + NW - 1) / NW) + 31) & ~31;
  const int slo = min(lo + w * S, l), shi = min(slo + S, l);
  int nl = 0, nr = 0;
  for (int c = slo; c < shi; c += 32) {
    const int t = c + lane;
    const bool v = t < shi;
    const unsigned kt = v ? isb_key(e[t]) : 0u;
    nl += __popc(__ballot_sync(full, v && kt >= p));
    nr += __popc(__ballot_sync(full, v && kt <= p));
  }
  if (lane == 0) { sh->wl[w] = nl; sh->wr[w] = nr; }
  __syncthreads();
  int runL = 0, afterR = 0;  // left stoppers in the slices before mine, right stoppers in the slices after mine
  for (int ww = 0; ww < NW; ++ww) {
    if (ww < w) runL += sh->wl[ww];
    if (ww > w) afterR += sh->wr[ww];
  }
  int runR = 0, K = 0, first_keep_L = 0x7fffffff, min_swap_R = l;
  for (int c = slo; c < shi; c += 32) {
    const int t = c + lane;
    const bool v = t < shi;
    const unsigned kt = v ? isb_key(e[t]) : 0u;
    const bool isL = v && kt >= p, isR = v && kt <= p;
    const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
    const int cL = runL + __popc(mL & lt_mask);
    const int cR = afterR + nr - (runR + __popc(mR & le_mask));
    const bool sL = isL && cR > cL, sR = isR && cL > cR;
    if (sL) posL[cL] = t;
    if (sR) posR[cR] = t;
    if (isL && !sL) first_keep_L = min(first_keep_L, t);
    if (sR) min_swap_R = min(min_swap_R, t);
    K += __popc(__ballot_sync(full, sL));
    runL += __popc(mL);
    runR += __popc(mR);
  }
  first_keep_L = __reduce_min_sync(full, first_keep_L);
  min_swap_R = __reduce_min_sync(full, min_swap_R);
  if (lane == 0) {
    atomicMin(&sh->red_keep, first_keep_L);
    atomicMin(&sh->red_swap, min_swap_R);
    atomicAdd(&sh->K, K);
  }
  __syncthreads();
  const int Ktot = sh->K;
  for (int k = threadIdx.x; k < Ktot; k += NW * 32) {
    const int a = posL[k], c2 = posR[k];
    const unsigned long long ea = e[a];
    e[a] = e[c2];
    e[c2] = ea;
  }
  const int cut = min(sh->red_keep, sh->red_swap);
  __syncthreads();
  return cut;
}

// e[0..n): the list (shared or global memory).  pos: n ints of scratch.  lists: 4 x list_cap range records (two levels x big /
// small), list_cap >= n / 17 + 2.  Afterwards e holds exactly what std::__introsort_loop(e, e + n, 2 * lg(n)) leaves.
template <int NW>
__device__ void block_introsort_partitions(unsigned long long *e, int *pos, int n, uint2 *lists, int list_cap, IsbShared *sh) {
  const int lane = threadIdx.x & 31;
  if (n <= 16) return;
  int depth = 2 * (31 - __clz(n));
  uint2 *big[2] = {lists, lists + list_cap}, *small_[2] = {lists + 2 * list_cap, lists + 3 * list_cap};
  if (threadIdx.x == 0) {
    sh->cnt_big[0] = sh->cnt_big[1] = sh->cnt_small[0] = sh->cnt_small[1] = 0;
    sh->next = 0;
    if (n >= ISB_BIG) { big[0][0] = make_uint2(0u, (unsigned)n); sh->cnt_big[0] = 1; }
    else { small_[0][0] = make_uint2(0u, (unsigned)n); sh->cnt_small[0] = 1; }
  }
  __syncthreads();
  int cur = 0;
  while (true) {
    const int nb = sh->cnt_big[cur], ns = sh->cnt_small[cur];
    if (nb + ns == 0) break;
    const int nxt = cur ^ 1;
    if (depth == 0) {  // depth limit: std::__partial_sort(first, last, last) == heapsort, one range per thread (never reached on
                       // non-adversarial data)
      for (int r = threadIdx.x; r < nb + ns; r += NW * 32) {
        const uint2 fl = r < nb ? big[cur][r] : small_[cur][r - nb];
        ssc::heap_sort(e + fl.x, (long)(fl.y - fl.x));
      }
      __syncthreads();
      break;
    }
    for (int r = 0; r < nb; ++r) {  // long ranges: the whole CTA, one after the other
      const uint2 fl = big[cur][r];
      const int cut = isb_block_partition<NW>(e, pos, (int)fl.x, (int)fl.y, sh);
      if (threadIdx.x == 0) {
        isb_push(sh, big[nxt], small_[nxt], nxt, (int)fl.x, cut);
        isb_push(sh, big[nxt], small_[nxt], nxt, cut, (int)fl.y);
      }
    }
    for (;;) {  // short ranges: one warp each, taken from a shared counter
      int r = 0;
      if (lane == 0) r = atomicAdd(&sh->next, 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= ns) break;
      const uint2 fl = small_[cur][r];
      const int cut = isb_warp_partition(e, pos, (int)fl.x, (int)fl.y);
      if (lane == 0) {
        isb_push(sh, big[nxt], small_[nxt], nxt, (int)fl.x, cut);
        isb_push(sh, big[nxt], small_[nxt], nxt, cut, (int)fl.y);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { sh->cnt_big[cur] = 0; sh->cnt_small[cur] = 0; sh->next = 0; }
    --depth;
    cur = nxt;
    __syncthreads();
  }
}
