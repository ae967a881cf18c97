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

#define ISB_BIG 1024  // ranges at least this long are partitioned by the whole CTA, shorter ones by one warp each

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

__device__ __forceinline__ unsigned isb_key_at(const unsigned long long *e, int t) {
  return reinterpret_cast<const unsigned *>(e)[2 * t + 1];  // high word of a little-endian 64-bit record
}

// swaps recorded in posL / posR (k-th left stopper <-> k-th right stopper); wpos: optional per-warp shared-memory home of the
// positions (16-bit offsets from f, for ranges up to ISB_REG elements), else the global scratch owned by the range
__device__ __forceinline__ void isb_warp_swaps(unsigned long long *e, const int *posL, const int *posR, const unsigned short *wpos, int f,
                                               int half, int K) {
  const int lane = threadIdx.x & 31;
  for (int k = lane; k < K; k += 32) {
    const int a = wpos ? f + wpos[k] : posL[k], c2 = wpos ? f + wpos[half + k] : posR[k];
    const unsigned long long ea = e[a], ec = e[c2];
    e[a] = ec;
    e[c2] = ea;
  }
}

#define ISB_REG 256  // ranges up to this many elements keep their keys in registers (16 per lane): one load round trip

// register path of the warp partition: the scan range (f, l) has at most 32 * NR elements
template <int NR>
__device__ __forceinline__ int isb_warp_partition_regs(unsigned long long *e, int *posL, int *posR, unsigned short *wpos, int f, int l,
                                                       unsigned p, int half) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
  const int lo = f + 1;
  unsigned kq[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    const int t = lo + q * 32 + lane;
    kq[q] = t < l ? isb_key_at(e, t) : 0u;
  }
  int totR = 0;
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    const int t = lo + q * 32 + lane;
    totR += __popc(__ballot_sync(full, t < l && kq[q] <= p));
  }
  int runL = 0, runR = 0, K = 0, first_keep_L = 0x7fffffff, min_swap_R = l;
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    if (NR > 2 && lo + q * 32 >= l) break;  // warp-uniform
    const int t = lo + q * 32 + lane;
    const bool v = t < l;
    const unsigned kt = kq[q];
    const bool isL = v && kt >= p, isR = v && kt <= p;
    const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
    const int cL = runL + __popc(mL & lt_mask);           // left stoppers before t
    const int cR = totR - (runR + __popc(mR & le_mask));  // right stoppers after t
    const bool sL = isL && cR > cL, sR = isR && cL > cR;
    if (wpos) {
      if (sL) wpos[cL] = (unsigned short)(t - f);
      if (sR) wpos[half + cR] = (unsigned short)(t - f);
    } else {
      if (sL) posL[cL] = t;
      if (sR) posR[cR] = t;
    }
    if (isL && !sL) first_keep_L = min(first_keep_L, t);
    if (sR) min_swap_R = min(min_swap_R, t);
    if (sL) K = cL + 1;  // the swapping left stoppers are the first K of them: the last one seen carries the count
    runL += __popc(mL);
    runR += __popc(mR);
  }
  first_keep_L = __reduce_min_sync(full, first_keep_L);
  min_swap_R = __reduce_min_sync(full, min_swap_R);
  K = __reduce_max_sync(full, K);
  __syncwarp();
  isb_warp_swaps(e, posL, posR, wpos, f, half, K);
  __syncwarp();
  return min(first_keep_L, min_swap_R);
}

// one warp partitions [f, l): returns the cut (same value in every lane).  pos: scratch of l - f ints owned by the range.
// wpos: per-warp shared scratch of ISB_REG 16-bit positions, or nullptr.
__device__ __forceinline__ int isb_warp_partition(unsigned long long *e, int *pos, int f, int l, unsigned short *wpos) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
  if (lane == 0) ssc::median_to_first(e + f, e + f + 1, e + f + (l - f) / 2, e + l - 1);
  __syncwarp();
  const unsigned p = isb_key_at(e, f);
  const int lo = f + 1, m = l - lo;
  const int half = m - m / 2;
  int *posL = pos + f + 1, *posR = pos + f + 1 + half;  // at most m / 2 swaps
  if (m <= 32) return isb_warp_partition_regs<1>(e, posL, posR, wpos, f, l, p, half);
  if (m <= 64) return isb_warp_partition_regs<2>(e, posL, posR, wpos, f, l, p, half);
  if (m <= 128) return isb_warp_partition_regs<4>(e, posL, posR, wpos, f, l, p, half);
  if (m <= ISB_REG) return isb_warp_partition_regs<ISB_REG / 32>(e, posL, posR, wpos, f, l, p, half);
  if (m <= 0xffff) {
    // One sweep: the offsets of ALL left stoppers (keys >= pivot) and of all right stoppers (keys <= pivot), both in ascending
    // order, as 16-bit offsets from f — 2m of them fit the m ints of scratch the range owns.  The k-th exchange pairs the k-th left
    // stopper with the k-th right stopper from the top; exchanges happen while the former lies before the latter, which is
    // monotone in k, so their number is found by one ballot per 32 pairs (no second pass over the keys to count the right
    // stoppers first).
    unsigned short *PL = reinterpret_cast<unsigned short *>(pos + f + 1), *PR = PL + m;
    int nL = 0, nR = 0;
    for (int c = lo; c < l; c += 128) {
      unsigned k4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = c + q * 32 + lane;
        k4[q] = t < l ? isb_key_at(e, t) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = c + q * 32 + lane;
        const bool v = t < l;
        const bool isL = v && k4[q] >= p, isR = v && k4[q] <= p;
        const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
        if (isL) PL[nL + __popc(mL & lt_mask)] = (unsigned short)(t - f);
        if (isR) PR[nR + __popc(mR & lt_mask)] = (unsigned short)(t - f);
        nL += __popc(mL);
        nR += __popc(mR);
      }
    }
    __syncwarp();
    const int np = min(nL, nR);
    int K = 0;
    for (int k0 = 0; k0 < np; k0 += 32) {
      const int k = k0 + lane;
      const unsigned ok = __ballot_sync(full, k < np && PL[k] < PR[nR - 1 - k]);
      K += __popc(ok);
      if (ok != full) break;
    }
    const int keep = K < nL ? f + PL[K] : 0x7fffffff, swp = K > 0 ? f + PR[nR - K] : l;
    for (int k = lane; k < K; k += 32) {
      const int a = f + PL[k], c2 = f + PR[nR - 1 - k];
      const unsigned long long ea = e[a], ec = e[c2];
      e[a] = ec;
      e[c2] = ea;
    }
    __syncwarp();
    return min(keep, swp);
  }
  int runL = 0, runR = 0, K = 0, first_keep_L = 0x7fffffff, min_swap_R = l, totR = 0;
  // longer ranges: two sweeps, several chunks of loads in flight at a time
  for (int c = lo; c < l; c += 128) {
    unsigned k4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = c + q * 32 + lane;
      k4[q] = t < l ? isb_key_at(e, t) : 0xffffffffu;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = c + q * 32 + lane;
      totR += __popc(__ballot_sync(full, t < l && k4[q] <= p));
    }
  }
  for (int c = lo; c < l; c += 128) {
    unsigned k4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = c + q * 32 + lane;
      k4[q] = t < l ? isb_key_at(e, t) : 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = c + q * 32 + lane;
      const bool v = t < l;
      const unsigned kt = k4[q];
      const bool isL = v && kt >= p, isR = v && kt <= p;
      const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
      const int cL = runL + __popc(mL & lt_mask);
      const int cR = totR - (runR + __popc(mR & le_mask));
      const bool sL = isL && cR > cL, sR = isR && cL > cR;
      if (sL) posL[cL] = t;
      if (sR) posR[cR] = t;
      if (isL && !sL) first_keep_L = min(first_keep_L, t);
      if (sR) min_swap_R = min(min_swap_R, t);
      K += __popc(__ballot_sync(full, sL));
      runL += __popc(mL);
      runR += __popc(mR);
    }
  }
  first_keep_L = __reduce_min_sync(full, first_keep_L);
  min_swap_R = __reduce_min_sync(full, min_swap_R);
  __syncwarp();
  isb_warp_swaps(e, posL, posR, nullptr, f, half, K);
  __syncwarp();
  return min(first_keep_L, min_swap_R);
}

__device__ __forceinline__ void isw_heap(unsigned long long *e, int f, int l, unsigned long long *wheap);

// A whole range of at most 32 records, finished in registers: one record per lane, median-of-3 and the partition exchanges by
// shuffles (the partner of the k-th exchanging left stopper is the k-th right stopper from the top: __fns on the stopper mask),
// no memory traffic and no per-range bookkeeping until the final store.  Of the two parts of such a range at most one is longer
// than 16, so the introsort loop needs no stack here.  About half of all partitions of a list happen on ranges this short.
__device__ __forceinline__ void isb_warp_small(unsigned long long *e, int f, int l, int depth, unsigned long long *wheap,
                                               unsigned short *wpos) {
  const int lane = threadIdx.x & 31, len = l - f;
  const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u, gt_mask = ~(lt_mask | (1u << lane));
  unsigned long long rec = lane < len ? e[f + lane] : ~0ull;
  int a = 0, b = len;
  while (b - a > 16) {
    if (depth == 0) {  // depth limit: heapsort of what is left of the range
      if (lane < len) e[f + lane] = rec;
      __syncwarp();
      isw_heap(e, f + a, f + b, wheap);
      return;
    }
    --depth;
    // std::__move_median_to_first(a, a + 1, mid, b - 1)
    const int ia = a + 1, ib = a + (b - a) / 2, ic = b - 1;
    const unsigned key = (unsigned)(rec >> 32);
    const unsigned ka = __shfl_sync(full, key, ia), kb = __shfl_sync(full, key, ib), kc = __shfl_sync(full, key, ic);
    int sidx;
    if (ka < kb) sidx = kb < kc ? ib : (ka < kc ? ic : ia);
    else sidx = ka < kc ? ia : (kb < kc ? ic : ib);
    // exchange records a <-> sidx with one shuffle (every other lane reads itself); the pivot key is the median just chosen
    rec = __shfl_sync(full, rec, lane == a ? sidx : (lane == sidx ? a : lane));
    const unsigned p = sidx == ia ? ka : (sidx == ib ? kb : kc), k = (unsigned)(rec >> 32);
    // std::__unguarded_partition(a + 1, b, pivot at a)
    const bool in = lane > a && lane < b;
    const bool isL = in && k >= p, isR = in && k <= p;
    const unsigned mL = __ballot_sync(full, isL), mR = __ballot_sync(full, isR);
    const int cL = __popc(mL & lt_mask), cR = __popc(mR & gt_mask);
    const bool sL = isL && cR > cL, sR = isR && cL > cR;
    int src = lane;
    if (wpos) {
      // stoppers listed by rank in the warp's shared scratch (right stoppers from the top, left stoppers from the bottom): a
      // swapping lane looks its partner up — two 16-bit stores and loads instead of two software __fns (37 instructions each)
      if (isR) wpos[cR] = (unsigned short)lane;
      if (isL) wpos[32 + cL] = (unsigned short)lane;
      __syncwarp();
      if (sL) src = wpos[cL];
      if (sR) src = wpos[32 + cR];
      __syncwarp();
    } else {
      if (sL) src = (int)__fns(mR, 31, -(cL + 1));
      if (sR) src = (int)__fns(mL, 0, cR + 1);
    }
    rec = __shfl_sync(full, rec, src);
    const unsigned keepL = __ballot_sync(full, isL && !sL), swapR = __ballot_sync(full, sR);
    const int cut = min(keepL ? __ffs(keepL) - 1 : 0x7fffffff, swapR ? __ffs(swapR) - 1 : b);
    if (cut - a > 16) b = cut;
    else if (b - cut > 16) a = cut;
    else break;
  }
  if (lane < len) e[f + lane] = rec;
  __syncwarp();
}

// one warp finishes the introsort loop of [f, l) on its own (explicit stack: the right part is deferred, the left part continued —
// the parts are disjoint, so the order in which they are processed does not matter).  LOCAL = true: e is the warp's shared-memory
// copy of a range of at most ISB_REG records (pos unused: the exchange positions live in wpos).
template <bool LOCAL>
__device__ __forceinline__ void isb_warp_finish_t(unsigned long long *e, int *pos, int f, int l, int depth, unsigned short *wpos,
                                                  unsigned long long *wheap, unsigned long long *wbuf) {
  const int lane = threadIdx.x & 31;
  uint2 st[40];  // y = l | depth << 24
  int sp = 0;
  st[sp++] = make_uint2((unsigned)f, (unsigned)l | ((unsigned)depth << 24));
  while (sp > 0) {
    const uint2 fr = st[--sp];
    f = (int)fr.x;
    l = (int)(fr.y & 0xffffffu);
    depth = (int)(fr.y >> 24);
    while (l - f > 16) {
      if (l - f <= 32) {
        isb_warp_small(e, f, l, depth, LOCAL ? nullptr : wheap, wpos);
        break;
      }
      if (!LOCAL && wbuf && l - f <= ISB_REG) {
        // everything below this size happens in the warp's shared-memory copy of the range: one read and one write of global
        // memory for all the levels that are left
        for (int t = lane; t < l - f; t += 32) wbuf[t] = e[f + t];
        __syncwarp();
        isb_warp_finish_t<true>(wbuf - f, nullptr, f, l, depth, wpos, nullptr, nullptr);
        __syncwarp();
        for (int t = lane; t < l - f; t += 32) e[f + t] = wbuf[t];
        __syncwarp();
        break;
      }
      if (depth == 0) {
        isw_heap(e, f, l, LOCAL ? nullptr : wheap);
        break;
      }
      --depth;
      const int cut = isb_warp_partition(e, pos, f, l, wpos);
      if (l - cut > 16 && sp < 40) st[sp++] = make_uint2((unsigned)cut, (unsigned)l | ((unsigned)depth << 24));
      l = cut;
    }
  }
}
__device__ __forceinline__ void isb_warp_finish(unsigned long long *e, int *pos, int f, int l, int depth, unsigned short *wpos,
                                                unsigned long long *wheap = nullptr, unsigned long long *wbuf = nullptr) {
  isb_warp_finish_t<false>(e, pos, f, l, depth, wpos, wheap, wbuf);
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
// wpos_all: optional shared memory, NW x ISB_REG 16-bit entries (per-warp swap positions of short ranges), or nullptr.
template <int NW>
__device__ void block_introsort_partitions(unsigned long long *e, int *pos, int n, uint2 *lists, int list_cap, IsbShared *sh,
                                           unsigned short *wpos_all = nullptr) {
  const int lane = threadIdx.x & 31;
  unsigned short *wpos = wpos_all ? wpos_all + (threadIdx.x >> 5) * ISB_REG : nullptr;
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
  const int w = threadIdx.x >> 5;
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
    if (nb == 0 && (ns >= NW || n < 64 * NW)) {
      // enough independent ranges (or a short list): every warp finishes its share without further block-wide steps
      for (int r = w; r < ns; r += NW) {
        const uint2 fl = small_[cur][r];
        isb_warp_finish(e, pos, (int)fl.x, (int)fl.y, depth, wpos);
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
      const int cut = isb_warp_partition(e, pos, (int)fl.x, (int)fl.y, wpos);
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

// ---------------------------------------------------------------------------------------------------------------------------
// Work-sharing variant for one list per CTA without any block-wide step: a task is a range (f, l, depth) that one warp
// partitions; the warp keeps the left part, hands the right part to a shared-memory queue when it is long enough to be worth
// another warp (ISW_SPLIT) and otherwise to its private stack.  The 10-20 dependent levels of a list thus overlap between the
// warps, and the top levels fan out over 1, 2, 4 ... warps.  The list stays with one SM, so its records are served from L1.
// Queue discipline: indices only grow.  push: outstanding += 1, slot = tail++, write the item, fence, ready[slot] = 1.
// pop: slot = head++ (every warp claims its own slot), wait until ready[slot] — or until outstanding == 0, which can only happen
// when every pushed task has completed, i.e. the claimed slot will never be filled: the warp leaves.  A task is completed
// (outstanding -= 1) after everything it kept privately is finished.  A watchdog turns an impossible wait into an error flag.
// ---------------------------------------------------------------------------------------------------------------------------
#define ISW_SPLIT 256  // = ISB_REG: what a warp finishes in its shared-memory copy is not worth handing over
#define ISW_CAP 512
#define ISW_HEAP 128  // depth-limit ranges up to this long are heap-sorted in shared memory

struct IswShared {
  int head, tail, outstanding, err;
  // one word per task, written once by its producer with atomicExch and read by its consumer with an atomic read: bit 63 = valid,
  // depth limit left << 48, l << 24, f (ranges of lists below 2^24 records).  A single-word message needs no separate "ready" flag.
  unsigned long long task[ISW_CAP];
};

__device__ __forceinline__ bool isw_push(IswShared *q, int f, int l, int depth) {
  const int slot = atomicAdd(&q->tail, 1);
  if (slot >= ISW_CAP) return false;  // queue exhausted: the caller keeps the range
  atomicAdd(&q->outstanding, 1);
  __threadfence_block();  // release: the records of [f, l) this warp has just moved are visible before the task is
  atomicExch(&q->task[slot], (1ull << 63) | ((unsigned long long)depth << 48) | ((unsigned long long)l << 24) | (unsigned long long)f);
  return true;
}

// depth limit reached: std::__partial_sort(first, last, last) == heapsort, sequential by nature; short ranges in shared memory
__device__ __forceinline__ void isw_heap(unsigned long long *e, int f, int l, unsigned long long *wheap) {
  const int lane = threadIdx.x & 31, len = l - f;
  if (wheap && len <= ISW_HEAP) {
    for (int t = lane; t < len; t += 32) wheap[t] = e[f + t];
    __syncwarp();
    if (lane == 0) ssc::heap_sort(wheap, (long)len);
    __syncwarp();
    for (int t = lane; t < len; t += 32) e[f + t] = wheap[t];
  } else if (lane == 0) {
    ssc::heap_sort(e + f, (long)len);
  }
  __syncwarp();
}

// every warp of the CTA calls; q must have been reset and the root task pushed (block_introsort_ws does both)
__device__ __forceinline__ void isw_worker(IswShared *q, unsigned long long *e, int *pos, unsigned short *wpos, unsigned long long *wheap,
                                           unsigned long long *wbuf) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&q->head, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= ISW_CAP) return;
    unsigned long long task = 0ull;
    if (lane == 0) {
      unsigned spins = 0;
      while (true) {
        task = atomicOr(&q->task[slot], 0ull);  // atomic read
        if (task) break;
        if (atomicAdd(&q->outstanding, 0) <= 0) break;  // nothing left that could fill this slot
        __nanosleep(100);
        if (++spins > (1u << 23)) { q->err = 1; break; }  // watchdog (~1 s): give up instead of hanging
      }
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    if (!task) return;
    __threadfence_block();  // acquire: the producer's records
    const int4 it = make_int4((int)(task & 0xffffffull), (int)((task >> 24) & 0xffffffull), (int)((task >> 48) & 0xffull), 0);
    uint2 st[40];  // private stack: y = l | depth << 24
    int sp = 0;
    st[sp++] = make_uint2((unsigned)it.x, (unsigned)it.y | ((unsigned)it.z << 24));
    while (sp > 0) {
      const uint2 fr = st[--sp];
      int f = (int)fr.x, l = (int)(fr.y & 0xffffffu), depth = (int)(fr.y >> 24);
      while (l - f > 16) {
        if (l - f <= ISB_REG) {  // short enough for the warp's shared-memory copy: finish it there
          isb_warp_finish(e, pos, f, l, depth, wpos, wheap, wbuf);
          break;
        }
        if (depth == 0) {
          isw_heap(e, f, l, wheap);
          break;
        }
        --depth;
        const int cut = isb_warp_partition(e, pos, f, l, wpos);
        const int rlen = l - cut;
        if (rlen > 16) {
          int pushed = 0;
          if (rlen > ISW_SPLIT) {
            if (lane == 0) pushed = isw_push(q, cut, l, depth) ? 1 : 0;
            pushed = __shfl_sync(0xffffffffu, pushed, 0);
          }
          if (!pushed && sp < 40) st[sp++] = make_uint2((unsigned)cut, (unsigned)l | ((unsigned)depth << 24));
        }
        l = cut;
      }
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicSub(&q->outstanding, 1); }
  }
}

// The partition phase of std::sort on e[0..n) by the NW warps of the CTA.  Long ranges (>= ISW_BIG) first, level by level, each
// partitioned by the whole CTA (isb_block_partition: the top levels of a long list would otherwise be one warp's serial,
// latency-bound chain); what they leave below ISW_BIG becomes the initial tasks of the work-sharing phase (isw_worker).
// pos: n ints of scratch.  wpos_all / wbuf_all: per-warp shared scratch (NW x ISB_REG 16-bit positions; NW x ISB_REG records: the
// warp's copy of a short range, also the heapsort buffer).  Every thread of the CTA must call; ends with a barrier.
#define ISW_BIG 2048
#define ISW_MAX_BIG 256  // lists of up to ISW_BIG * ISW_MAX_BIG = 524288 records start with cooperative partitions
struct IswBig {
  IsbShared coop;
  int2 range[2][ISW_MAX_BIG];
  int cnt[2];
};

template <int NW>
__device__ void block_introsort_ws(unsigned long long *e, int *pos, int n, IswShared *q, IswBig *big, unsigned short *wpos_all,
                                   unsigned long long *wbuf_all) {
  if (n <= 16) return;  // uniform
  if (n >= (1 << 24)) __trap();  // task words and stack entries carry 24-bit positions
  for (int t = threadIdx.x; t < ISW_CAP; t += NW * 32) q->task[t] = 0ull;
  int depth = 2 * (31 - __clz(n));
  if (threadIdx.x == 0) {
    q->head = 0; q->tail = 0; q->outstanding = 0; q->err = 0;
    big->cnt[0] = big->cnt[1] = 0;
    if (n >= ISW_BIG && n <= ISW_BIG * ISW_MAX_BIG) { big->range[0][0] = make_int2(0, n); big->cnt[0] = 1; }
  }
  __syncthreads();
  if (threadIdx.x == 0 && big->cnt[0] == 0) isw_push(q, 0, n, depth);
  int cur = 0;
  while (true) {
    __syncthreads();
    const int nb = big->cnt[cur];
    if (nb == 0) break;
    const int nxt = cur ^ 1;
    if (depth == 0) {  // depth limit on a long range: hand it to the workers, which heap-sort it
      if (threadIdx.x == 0)
        for (int r = 0; r < nb; ++r) isw_push(q, big->range[cur][r].x, big->range[cur][r].y, 0);
      break;
    }
    --depth;
    for (int r = 0; r < nb; ++r) {
      const int2 fl = big->range[cur][r];
      const int cut = isb_block_partition<NW>(e, pos, fl.x, fl.y, &big->coop);
      if (threadIdx.x == 0) {
        const int part[2][2] = {{fl.x, cut}, {cut, fl.y}};
        for (int c = 0; c < 2; ++c) {
          const int len = part[c][1] - part[c][0];
          if (len >= ISW_BIG) big->range[nxt][big->cnt[nxt]++] = make_int2(part[c][0], part[c][1]);
          else if (len > 16) isw_push(q, part[c][0], part[c][1], depth);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) big->cnt[cur] = 0;
    cur = nxt;
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  isw_worker(q, e, pos, wpos_all + w * ISB_REG, wbuf_all + (size_t)w * ISB_REG, wbuf_all + (size_t)w * ISB_REG);
  __syncthreads();
  if (q->err) __trap();  // the watchdog fired: fail the launch loudly rather than leave a half-ordered list behind
}
