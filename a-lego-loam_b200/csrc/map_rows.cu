// map_rows.cu — build of the voxel-row index (map_rows.cuh)
#include "map_rows.cuh"

namespace {

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) mr_bbox_kernel(const float4 *__restrict__ pts, size_t pts_stride, const int *__restrict__ n_ptr,
                                                      int *__restrict__ bbox) {
  const int b = blockIdx.y;
  const int n = n_ptr[b];
  const float4 *src = pts + (size_t)b * pts_stride;
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {  // four independent 16-byte loads in flight
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = min(i0 + u * stride, n - 1);  // past the end: re-read the last point (changes neither min nor max)
      p[u] = ldg_f4(src + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      mn[0] = fminf(mn[0], p[u].x); mn[1] = fminf(mn[1], p[u].y); mn[2] = fminf(mn[2], p[u].z);
      mx[0] = fmaxf(mx[0], p[u].x); mx[1] = fmaxf(mx[1], p[u].y); mx[2] = fmaxf(mx[2], p[u].z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  // one set of atomics per CTA (the six words of a sequence are a serialisation point in L2: per-warp atomics dominated the kernel)
  __shared__ float s_red[8][6];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
    for (int a = 0; a < 3; ++a) { s_red[wid][a] = mn[a]; s_red[wid][3 + a] = mx[a]; }
  __syncthreads();
  if (threadIdx.x < 6 && blockIdx.x * blockDim.x < n) {
    const int a = threadIdx.x;
    float v = s_red[0][a];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = a < 3 ? fminf(v, s_red[w][a]) : fmaxf(v, s_red[w][a]);
    if (a < 3) atomicMin(bbox + b * 6 + a, f2ord(v));
    else atomicMax(bbox + b * 6 + a, f2ord(v));
  }
}

__global__ void mr_bbox_init_kernel(int *bbox, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int a = 0; a < 3; ++a) { bbox[b * 6 + a] = 0x7fffffff; bbox[b * 6 + 3 + a] = (int)0x80000000; }
}

__global__ void mr_frame_kernel(const int *__restrict__ bbox, const int *__restrict__ n_ptr, MapFrame *__restrict__ frame, float leaf, int cap,
                                int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  MapFrame f;
  f.inv = 1.0f / leaf;
  f.n = n_ptr[b];
  f.valid = 0;
  f.W = 0;
  for (int a = 0; a < 3; ++a) { f.min_b[a] = 0; f.dim[a] = 0; }
  if (f.n > 0) {
    long long entries = 1;
    bool ok = true;
    for (int a = 0; a < 3; ++a) {
      const float mn = ord2f(bbox[b * 6 + a]), mx = ord2f(bbox[b * 6 + 3 + a]);
      ok = ok && isfinite(mn) && isfinite(mx) && fabsf(mn) < 1e6f && fabsf(mx) < 1e6f;
      f.min_b[a] = (int)floorf(mn * f.inv);
      f.dim[a] = (int)floorf(mx * f.inv) - f.min_b[a] + 1;
      ok = ok && f.dim[a] > 0 && f.dim[a] < (1 << 20);
    }
    if (ok) {
      f.W = (f.dim[0] + 31) >> 5;
      entries = (long long)f.W * f.dim[1] * f.dim[2];
      f.valid = (cap <= 0 || entries <= cap) ? 1 : 0;  // cap <= 0: sizing pass, the table does not exist yet
    }
  }
  frame[b] = f;
}

// clears the entries a sequence's frame actually spans (the table is sized for the largest one)
__global__ void __launch_bounds__(256) mr_clear_kernel(const MapFrame *__restrict__ frame, uint2 *__restrict__ tab, int cap) {
  const int b = blockIdx.y;
  const MapFrame f = frame[b];
  if (!f.valid) return;
  const int entries = f.W * f.dim[1] * f.dim[2];
  uint4 *t = reinterpret_cast<uint4 *>(tab + (size_t)b * cap);  // cap is a multiple of 1024 entries
  const int n4 = (entries + 1) >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) t[i] = make_uint4(0u, 0u, 0u, 0u);
}

// One thread per point.  pcl::VoxelGrid output lists the voxels in increasing key order, so the points of one 32-voxel word are
// consecutive in the cloud — a run of lanes.  The lane that opens a run sums its lanes' bits (distinct bits: a segment sum of the
// warp's modular prefix sum) and writes {mask, start} with one plain 8-byte store; only the runs cut by the warp's edges (the first
// when it continues the previous warp's word, and the last) go through atomicOr on the cleared table.  Issue-bound: the loop body
// is kept to ~80 instructions (32-bit keys, uniform look-behind load, no per-run masked reductions).
__global__ void __launch_bounds__(256) mr_fill_kernel(const float4 *__restrict__ pts, size_t pts_stride, MapFrame *__restrict__ frame,
                                                      uint2 *__restrict__ tab, int cap) {
  const int b = blockIdx.y;
  const MapFrame f = frame[b];
  if (!f.valid) return;
  const float4 *src = pts + (size_t)b * pts_stride;
  uint2 *t = tab + (size_t)b * cap;
  const int lane = threadIdx.x & 31;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < f.n; i0 += gridDim.x * blockDim.x) {  // whole warps
    const int i = i0 + lane;
    const bool live = i < f.n;
    // dead lanes re-read the last point: they extend the last run with bits it already has and are never openers
    const float4 p = ldg_f4(src + min(i, f.n - 1));
    const int ix = mr_voxel(p.x, f.inv, f.min_b[0]);
    const int row = mr_voxel(p.y, f.inv, f.min_b[1]) + f.dim[1] * mr_voxel(p.z, f.inv, f.min_b[2]);
    const int entry = row * f.W + (ix >> 5);
    const unsigned bit = live ? 1u << (ix & 31) : 0u;
    // the point before mine: lane - 1, or (lane 0) the last point of the previous 32 — one broadcast load for the warp
    const float4 q = ldg_f4(src + max(i0 - 1, 0));
    const int q_ix = mr_voxel(q.x, f.inv, f.min_b[0]);
    const int q_row = i0 == 0 ? -1 : mr_voxel(q.y, f.inv, f.min_b[1]) + f.dim[1] * mr_voxel(q.z, f.inv, f.min_b[2]);
    const int s_ix = __shfl_up_sync(0xffffffffu, ix, 1), s_row = __shfl_up_sync(0xffffffffu, row, 1);
    const int prev_ix = lane == 0 ? q_ix : s_ix, prev_row = lane == 0 ? q_row : s_row;
    const int prev_entry = prev_row * f.W + (prev_ix >> 5);  // negative for the cloud's first point
    // not pcl::VoxelGrid output order (keys must rise strictly), or not finite: the index does not apply
    const bool bad = live && (prev_row > row || (prev_row == row && prev_ix >= ix) || !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z)));
    if (bad) frame[b].valid = 0;
    const bool opens = live && entry != prev_entry;
    // modular inclusive prefix sum of the bits: a run's mask is the difference of two prefixes (its bits are distinct)
    unsigned pre = bit;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    const unsigned heads = __ballot_sync(0xffffffffu, opens);
    const unsigned above = heads & (0xfffffffeu << lane);
    const int last = above ? __ffs(above) - 2 : 31;  // last lane of my run
    const unsigned run_sum = __shfl_sync(0xffffffffu, pre, last) - (pre - bit);
    const bool inside = entry >= 0 && entry < cap;   // (a NaN coordinate lands anywhere; the frame is flagged invalid above)
    if (opens && !bad && inside) {
      if (above) t[entry] = make_uint2(run_sum, (unsigned)i);               // opened and closed in this warp: all mine
      else { atomicOr(&t[entry].x, run_sum); t[entry].y = (unsigned)i; }    // the warp's last run may go on in the next 32 points
    }
    // lanes before the first opener continue a word an earlier warp opened
    const int n_cont = heads ? __ffs(heads) - 1 : 32;
    if (n_cont > 0) {
      const unsigned cont = __shfl_sync(0xffffffffu, pre, n_cont - 1);
      if (lane == 0 && !bad && inside) atomicOr(&t[entry].x, cont);
    }
  }
}

}  // namespace

void map_rows_free(MapRows *m) {
  if (m->tab) cudaFree(m->tab);
  if (m->frame) cudaFree(m->frame);
  if (m->bbox) cudaFree(m->bbox);
  m->tab = nullptr;
  m->frame = nullptr;
  m->bbox = nullptr;
  m->cap = 0;
  m->usable = false;
}

static int mr_frames(AlegoHandle *h, MapRows *m, const float4 *pts, size_t pts_stride, const int *n_ptr, int cap, cudaStream_t s,
                     const char *tag, int points_cap) {
  const int B = h->B;
  std::string t0 = std::string("maprows_bbox_") + tag;
  mr_bbox_init_kernel<<<div_up(B, 128), 128, 0, s>>>(m->bbox, B);
  { LAUNCH(h, t0.c_str()); mr_bbox_kernel<<<dim3(std::min(div_up(points_cap, 256 * 8), 48), B), 256, 0, s>>>(pts, pts_stride, n_ptr, m->bbox); }
  mr_frame_kernel<<<div_up(B, 128), 128, 0, s>>>(m->bbox, n_ptr, m->frame, m->leaf, cap, B);
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

int map_rows_build(AlegoHandle *h, MapRows *m, const float4 *pts, size_t pts_stride, const int *n_ptr, const char *tag) {
  cudaStream_t s = h->launch_stream ? h->launch_stream : h->stream;
  const int B = h->B;
  int rc = mr_frames(h, m, pts, pts_stride, n_ptr, m->cap, s, tag, (int)pts_stride);
  if (rc != ALEGO_OK) return rc;
  std::string tc = std::string("maprows_clear_") + tag;
  { LAUNCH(h, tc.c_str()); mr_clear_kernel<<<dim3(std::min(div_up(m->cap / 2, 256), 64), B), 256, 0, s>>>(m->frame, m->tab, m->cap); }
  std::string t1 = std::string("maprows_fill_") + tag;
  { LAUNCH(h, t1.c_str()); mr_fill_kernel<<<dim3(std::min(div_up((int)pts_stride, 256), 128), B), 256, 0, s>>>(pts, pts_stride, m->frame, m->tab, m->cap); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}

int map_rows_validate(AlegoHandle *h, MapRows *m, const float4 *pts, size_t pts_stride, const int *n_ptr, float leaf, const char *tag) {
  const int B = h->B;
  cudaStream_t s = h->stream;
  m->usable = false;
  m->leaf = leaf;
  if (!m->frame) {
    CUDA_TRY(h, cudaMalloc(&m->frame, (size_t)B * sizeof(MapFrame)));
    CUDA_TRY(h, cudaMalloc(&m->bbox, (size_t)B * 6 * sizeof(int)));
  }
  // sizing pass: extents of every sequence's cloud
  int rc = mr_frames(h, m, pts, pts_stride, n_ptr, 0, s, tag, (int)pts_stride);
  if (rc != ALEGO_OK) return rc;
  std::vector<MapFrame> fr(B);
  CUDA_TRY(h, cudaMemcpyAsync(fr.data(), m->frame, (size_t)B * sizeof(MapFrame), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  long long need = 0;
  int with_points = 0;
  for (const MapFrame &f : fr) {
    if (f.n <= 0) continue;
    ++with_points;
    if (!f.valid) return ALEGO_OK;  // degenerate extents somewhere: hashed grid
    need = std::max(need, (long long)f.W * f.dim[1] * f.dim[2]);
  }
  if (with_points == 0 || need > (1ll << 19)) return ALEGO_OK;  // nothing to index / table larger than 4 MB per sequence: hashed grid
  const int cap = (int)((need + 1023) & ~1023ll);
  if (cap > m->cap) {
    ++h->graph_epoch;  // captured graphs hold the old table pointer
    if (m->tab) cudaFree(m->tab);
    m->tab = nullptr;
    CUDA_TRY(h, cudaMalloc(&m->tab, (size_t)B * cap * sizeof(uint2)));
    m->cap = cap;
  }
  h->launch_stream = nullptr;
  rc = map_rows_build(h, m, pts, pts_stride, n_ptr, tag);
  if (rc != ALEGO_OK) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(fr.data(), m->frame, (size_t)B * sizeof(MapFrame), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  bool ok = true;
  for (const MapFrame &f : fr) ok = ok && (f.n <= 0 || f.valid);
  m->usable = ok;
  return ALEGO_OK;
}
