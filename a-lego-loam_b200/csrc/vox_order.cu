// vox_order.cu — the partition phase of libstdc++'s std::sort for a BATCH of independent record lists (introsort_block.cuh).  It is
// what turns the device's VoxelGrid into pcl::VoxelGrid's exact record order (laserOdometry.cpp:288-293, laserMapping.cpp:325-342):
// the kernels that build the (voxel, point) record lists leave a VoxState per list, the kernels here partition, the finish kernels
// apply the stable radix sort.  A list stays with one SM from start to finish (its few KB are then served from L1):
//   vox_order_warp_kernel  one WARP per list, private stack  — the 16 k per-ring lists of LaserOdometry (a few hundred records each:
//                          the batch itself supplies the parallelism, nothing ever waits on a barrier)
//   vox_order_cta_kernel   one CTA per list, the warps share the ranges through a shared-memory queue — LaserMapping's clouds
//                          (hundreds to ~15 k records: the top levels fan out over 1, 2, 4 ... warps)
#include "common.cuh"
#include "sort_voxel.cuh"
#include "vox_order.cuh"

namespace {
#define VOW_WARPS 8
__global__ void __launch_bounds__(VOW_WARPS * 32)
vox_order_warp_kernel(const VoxState *__restrict__ state, int n_lists, u64 *__restrict__ buf_a, u64 *__restrict__ buf_b) {
  __shared__ unsigned short s_wpos[VOW_WARPS * ISB_REG];
  __shared__ u64 s_heap[VOW_WARPS * ISW_HEAP];
  const int warp = threadIdx.x >> 5;
  const int c = blockIdx.x * VOW_WARPS + warp;
  if (c >= n_lists) return;
  const int n = state[c].n;
  if (state[c].done || state[c].nv != n || n <= 16) return;
  isb_warp_finish(buf_a + state[c].off, reinterpret_cast<int *>(buf_b + state[c].off_b), 0, n, 2 * (31 - __clz(n)), s_wpos + warp * ISB_REG,
                  s_heap + (size_t)warp * ISW_HEAP);
}

#define VOC_WARPS 8
__global__ void __launch_bounds__(VOC_WARPS * 32)
vox_order_cta_kernel(const VoxState *__restrict__ state, u64 *__restrict__ buf_a, u64 *__restrict__ buf_b) {
  __shared__ IswShared s_q;
  __shared__ IswBig s_big;
  __shared__ unsigned short s_wpos[VOC_WARPS * ISB_REG];
  __shared__ u64 s_heap[VOC_WARPS * ISW_HEAP];
  const VoxState &st = state[blockIdx.x];
  const int n = st.n;
  if (st.done || st.nv != n || n <= 16) return;  // uniform
  block_introsort_ws<VOC_WARPS>(buf_a + st.off, reinterpret_cast<int *>(buf_b + st.off_b), n, &s_q, &s_big, s_wpos, s_heap);
}
}  // namespace

int vox_order_lists_by_warp(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag) {
  { LAUNCH(h, tag); vox_order_warp_kernel<<<div_up(n_lists, VOW_WARPS), VOW_WARPS * 32, 0, s>>>(state, n_lists, buf_a, buf_b); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

int vox_order_lists_by_cta(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag) {
  { LAUNCH(h, tag); vox_order_cta_kernel<<<n_lists, VOC_WARPS * 32, 0, s>>>(state, buf_a, buf_b); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
