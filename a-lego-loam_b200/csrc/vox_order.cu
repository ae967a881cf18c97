// vox_order.cu — the partition phase of libstdc++'s std::sort for a BATCH of independent record lists (introsort_block.cuh).  It is
// what turns the device's VoxelGrid into pcl::VoxelGrid's exact record order (laserOdometry.cpp:288-293, laserMapping.cpp:325-342):
// the kernels that build the (voxel, point) record lists leave a VoxState per list, the kernels here partition, the finish kernels
// apply the stable radix sort.  A list stays with one SM from start to finish (its few KB are then served from L1):
//   vox_order_warp_kernel  one WARP per list, private stack  — the 16 k per-ring lists of LaserOdometry (a few hundred to 2 k records
//                          each: the batch itself supplies the parallelism, nothing ever waits on a barrier)
//   vox_order_wide_kernel  one CTA of 16 warps per list, the warps share the ranges through a shared-memory queue — LaserMapping's
//                          clouds (hundreds to ~20 k records: the top levels fan out over 1, 2, 4 ... warps), and the ring lists too
//                          when a launch holds too few lists to fill the GPU with one warp each (latency regime)
//                          and ring lists above 2048 records (sweeps wider than 2048 columns) in a large batch
#include <cstdlib>

#include "common.cuh"
#include "sort_voxel.cuh"
#include "vox_order.cuh"

namespace {
#define VOW_WARPS 8
#define VO_WARP_MAX 512
#define VO_RING_WARP_MAX 2048  // 640 sends the dense near-range rings to the CTA kernel: 0.25 + 0.57 ms instead of 0.52 (issue-bound either way)
__global__ void __launch_bounds__(VOW_WARPS * 32)
vox_order_warp_kernel(const VoxState *__restrict__ state, int n_lists, u64 *__restrict__ buf_a, u64 *__restrict__ buf_b, int max_n, int group) {
  __shared__ unsigned short s_wpos[VOW_WARPS * ISB_REG];
  __shared__ u64 s_buf[VOW_WARPS * ISB_REG];  // per warp: its copy of a short range / heapsort buffer
  const int warp = threadIdx.x >> 5;
  int c = blockIdx.x * VOW_WARPS + warp;
  if (c >= n_lists) return;
  // group = lists per sequence, a multiple of the warps per CTA (the rings of a sweep): issue the CTAs ring block by ring block
  // across all sequences instead of sequence by sequence.  The low rings (ground, near range) hold the longest lists; started
  // first, they are not what the launch ends on.
  if (group > 0) {
    const int per = group / VOW_WARPS, n_seq = n_lists / group;
    const int rb = blockIdx.x / n_seq, b = blockIdx.x - rb * n_seq;
    if (rb < per) c = b * group + rb * VOW_WARPS + warp;
  }
  const int n = state[c].n;
  if (state[c].done || state[c].nv != n || n <= 16 || n > max_n) return;
  isb_warp_finish(buf_a + state[c].off, reinterpret_cast<int *>(buf_b + state[c].off_b), 0, n, 2 * (31 - __clz(n)), s_wpos + warp * ISB_REG,
                  s_buf + (size_t)warp * ISB_REG, s_buf + (size_t)warp * ISB_REG);
}

// LaserMapping's batch: `group` lists per sequence (corner, surf, outlier, union), the surf list an order of magnitude longer than
// the others.  A launch lasts as long as its longest list, and a list is as fast as the warps that share it: 16 warps per CTA (two
// CTAs per SM, so the 256 long lists of a full batch are one wave), and the long kind is issued FIRST (first_kind) so that the short
// lists fill the slots it frees instead of the other way round.  Dynamic shared memory (52 KB: above the static limit).
#define VOX_WIDE_WARPS 16
#define VOX_WIDE_SMEM (((sizeof(IswShared) + 15) & ~(size_t)15) + ((sizeof(IswBig) + 15) & ~(size_t)15) + \
                       (size_t)VOX_WIDE_WARPS * ISB_REG * (sizeof(u64) + sizeof(unsigned short)))
__global__ void __launch_bounds__(VOX_WIDE_WARPS * 32, 2)
vox_order_wide_kernel(const VoxState *__restrict__ state, u64 *__restrict__ buf_a, u64 *__restrict__ buf_b, int min_n, int group, int n_groups,
                      int first_kind) {
  extern __shared__ __align__(16) unsigned char vow_smem[];
  IswShared *q = reinterpret_cast<IswShared *>(vow_smem);
  IswBig *big = reinterpret_cast<IswBig *>(vow_smem + ((sizeof(IswShared) + 15) & ~(size_t)15));
  u64 *wbuf = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(big) + ((sizeof(IswBig) + 15) & ~(size_t)15));
  unsigned short *wpos = reinterpret_cast<unsigned short *>(wbuf + (size_t)VOX_WIDE_WARPS * ISB_REG);
  // block i -> (kind, sequence): kinds in the order first_kind, first_kind + 1, ... (mod group), sequences innermost
  const int k = blockIdx.x / n_groups, b = blockIdx.x - k * n_groups;
  int kind = first_kind + k;
  if (kind >= group) kind -= group;
  const VoxState &st = state[b * group + kind];
  const int n = st.n;
  if (st.done || st.nv != n || n <= 16 || n < min_n) return;  // uniform
  block_introsort_ws<VOX_WIDE_WARPS>(buf_a + st.off, reinterpret_cast<int *>(buf_b + st.off_b), n, q, big, wpos, wbuf);
}
}  // namespace

static int vox_wide_attr(AlegoHandle *h) {
  static bool attr_set[ALEGO_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!attr_set[h->dev]) {
    CUDA_TRY(h, cudaFuncSetAttribute(vox_order_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOX_WIDE_SMEM));
    attr_set[h->dev] = true;
  }
  return ALEGO_OK;
}

// below this many lists in a launch the batch does not fill the GPU with one warp per list (148 SMs x 32 resident warps): the
// launch then lasts as long as ONE warp needs for the longest list, and sharing a list among the warps of a CTA pays
#define VO_SMALL_BATCH_LISTS 1024
#define VO_SMALL_BATCH_SPLIT 256

int vox_order_lists_by_warp(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag,
                            int group, int max_len) {
  static const int small_batch = [] {  // tuning override (tools/latency.py compares both routings)
    const char *e = getenv("ALEGO_ORDER_SMALL_BATCH_LISTS");
    return e ? atoi(e) : VO_SMALL_BATCH_LISTS;
  }();
  if (n_lists <= small_batch) {
    // latency regime (a few sequences): 16 warps per list for everything a warp would not finish in its shared-memory copy
    const int rc = vox_wide_attr(h);
    if (rc != ALEGO_OK) return rc;
    { LAUNCH(h, tag);
      vox_order_wide_kernel<<<n_lists, VOX_WIDE_WARPS * 32, VOX_WIDE_SMEM, s>>>(state, buf_a, buf_b, VO_SMALL_BATCH_SPLIT + 1, 1, n_lists, 0); }
    { std::string t2 = std::string(tag) + "_short"; LAUNCH(h, t2.c_str());
      vox_order_warp_kernel<<<div_up(n_lists, VOW_WARPS), VOW_WARPS * 32, 0, s>>>(state, n_lists, buf_a, buf_b, VO_SMALL_BATCH_SPLIT, 0); }
    CUDA_TRY(h, cudaGetLastError());
    return ALEGO_OK;
  }
  // a list is one warp's serial work, and a launch of 16 k lists is about one wave: its duration is the LONGEST list's.  Lists
  // above VO_RING_WARP_MAX records (possible only for sweeps wider than that many columns) therefore go to work-sharing CTAs
  // (the two kernels touch disjoint lists)
  { LAUNCH(h, tag);
    vox_order_warp_kernel<<<div_up(n_lists, VOW_WARPS), VOW_WARPS * 32, 0, s>>>(state, n_lists, buf_a, buf_b, VO_RING_WARP_MAX,
                                                                               group % VOW_WARPS == 0 && n_lists % group == 0 ? group : 0); }
  if (max_len > VO_RING_WARP_MAX) {
    const int rc = vox_wide_attr(h);
    if (rc != ALEGO_OK) return rc;
    std::string t2 = std::string(tag) + "_long";
    LAUNCH(h, t2.c_str());
    vox_order_wide_kernel<<<n_lists, VOX_WIDE_WARPS * 32, VOX_WIDE_SMEM, s>>>(state, buf_a, buf_b, VO_RING_WARP_MAX + 1, 1, n_lists, 0);
  }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

int vox_order_lists_by_cta(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag,
                           int group, int first_kind) {
  // same routing by length (a batch whose lists range from a few hundred to ten thousand records: LaserMapping's clouds)
  const int rc = vox_wide_attr(h);
  if (rc != ALEGO_OK) return rc;
  { LAUNCH(h, tag);
    vox_order_wide_kernel<<<n_lists, VOX_WIDE_WARPS * 32, VOX_WIDE_SMEM, s>>>(state, buf_a, buf_b, VO_WARP_MAX + 1, group, n_lists / group,
                                                                             first_kind); }
  { std::string t2 = std::string(tag) + "_short"; LAUNCH(h, t2.c_str());
    vox_order_warp_kernel<<<div_up(n_lists, VOW_WARPS), VOW_WARPS * 32, 0, s>>>(state, n_lists, buf_a, buf_b, VO_WARP_MAX, 0); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
