// grid.cu — hashed uniform cell grid over a point cloud, one per sequence: the device replacement for
// pcl::KdTreeFLANN::setInputCloud (laserOdometry.cpp:321-322,533-534; laserMapping.cpp:356-357).
// Build = counting sort of the points by hashed cell: count (atomics on a per-sequence table), exclusive
// scan of the table, fill (the table is count, fill cursor and final bucket index in turn, see grid_scan_kernel).  Points are copied next to each other per bucket (xyz + original index), so a
// query reads whole buckets with coalesced 16-byte loads.  Bucket order is arbitrary; every consumer ranks
// candidates by (distance, original index), so results do not depend on it.
#include "common.cuh"
#include "grid.cuh"

namespace {

__global__ void __launch_bounds__(256) grid_count_kernel(const float4 *__restrict__ pts, size_t pts_stride, const int *__restrict__ n_ptr,
                                                         int n_stride, int *__restrict__ counts, int T, float inv_cell, int cap) {
  const int b = blockIdx.y;
  const int n = min(n_ptr[(size_t)b * n_stride], cap);
  const float4 *src = pts + (size_t)b * pts_stride;
  int *cnt = counts + (size_t)b * GRID_TABLE_STRIDE(T);
  // populations of the 4096-entry chunks of the table (they let the scan run chunk-parallel): accumulated per CTA in shared
  // memory, one global atomic per touched chunk at the end
  __shared__ int s_chunk[GRID_MAX_CHUNKS];
  const int n_chunks = (T + 4 + GRID_CHUNK - 1) >> GRID_CHUNK_SHIFT;
  if (blockIdx.x * blockDim.x >= n) return;  // block-uniform
  for (int t = threadIdx.x; t < n_chunks; t += blockDim.x) s_chunk[t] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = ldg_f4(src + i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const int e = 1 + grid_hash(grid_coord(p.x, inv_cell), grid_coord(p.y, inv_cell), grid_coord(p.z, inv_cell), T);
    atomicAdd(cnt + e, 1);
    atomicAdd(&s_chunk[e >> GRID_CHUNK_SHIFT], 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_chunks; t += blockDim.x)
    if (s_chunk[t]) atomicAdd(cnt + T + 4 + t, s_chunk[t]);
}

// In-place exclusive scan of the T+1 entries of the bucket table (entry 0 is always 0, entry 1+h holds the count of
// bucket h), one CTA per 4096-entry chunk: the chunk's base is the sum of the chunk populations before it (accumulated by
// grid_count next to the table), so the chunks of a sequence are scanned in parallel.  Afterwards entry 1+h = start of
// bucket h = the fill cursor; the fill kernel advances it to the END of bucket h = start of bucket h+1, so that after the
// fill entry h = start(h) and entry T = number of points: one array is count, cursor and index.
__global__ void __launch_bounds__(1024) grid_scan_kernel(int *__restrict__ table, int T) {
  const int n_chunks = (T + 4 + GRID_CHUNK - 1) >> GRID_CHUNK_SHIFT;
  int *arr = table + (size_t)blockIdx.y * GRID_TABLE_STRIDE(T);
  const int *chunk_tot = arr + T + 4;
  const int c = blockIdx.x;
  __shared__ int s_scan[34];
  __shared__ int s_base;
  if (threadIdx.x < 32) {
    int v = 0;
    for (int t = threadIdx.x; t < c; t += 32) v += chunk_tot[t];
    v = warp_sum_i(v);
    if (threadIdx.x == 0) s_base = v;
  }
  __syncthreads();
  const int i = c * GRID_CHUNK + threadIdx.x * 4;
  int4 v = make_int4(0, 0, 0, 0);
  if (i < T + 4) v = *reinterpret_cast<const int4 *>(arr + i);
  int total;
  const int ex = block_excl_scan(v.x + v.y + v.z + v.w, s_scan, &total) + s_base;
  if (i < T + 4) *reinterpret_cast<int4 *>(arr + i) = make_int4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
  (void)n_chunks;
}

__global__ void __launch_bounds__(256) grid_fill_kernel(const float4 *__restrict__ pts, size_t pts_stride, const int *__restrict__ n_ptr,
                                                        int n_stride, int *__restrict__ cursor, float4 *__restrict__ sorted, int T,
                                                        float inv_cell, int cap, int pack_ring) {
  const int b = blockIdx.y;
  const int n = min(n_ptr[(size_t)b * n_stride], cap);
  const float4 *src = pts + (size_t)b * pts_stride;
  int *cur = cursor + (size_t)b * GRID_TABLE_STRIDE(T) + 1;
  float4 *dst = sorted + (size_t)b * cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = ldg_f4(src + i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const int slot = atomicAdd(cur + grid_hash(grid_coord(p.x, inv_cell), grid_coord(p.y, inv_cell), grid_coord(p.z, inv_cell), T), 1);
    // pack_ring: the ring id int(intensity) (laserOdometry.cpp:347,436) rides in bits 24..30 next to the index
    const int tag = pack_ring ? (i | ((int)p.w << GRID_RING_SHIFT)) : i;
    dst[slot] = make_float4(p.x, p.y, p.z, __int_as_float(tag));
  }
}

}  // namespace

int grid_alloc(AlegoHandle *h, GridIndex *g, int cap, float cell, int table_factor_log2, int batch) {
  grid_free(g);
  const int nb = batch > 0 ? batch : h->B;
  ++h->graph_epoch;  // captured graphs hold the old table / array pointers
  g->cap = cap;
  g->cell = cell;
  // buckets: next_pow2(cap / 2) << table_factor_log2.  A query reads whole buckets, so every point that merely collides
  // with a neighbour cell is a wasted candidate: the big local-map grids use more buckets than points.
  int T = next_pow2(cap > 2048 ? cap / 2 : 1024);
  if (T < 1024) T = 1024;
  T <<= table_factor_log2;
  g->table_size = T;
  CUDA_TRY(h, cudaMalloc(&g->cell_start, (size_t)nb * GRID_TABLE_STRIDE(T) * sizeof(int)));
  CUDA_TRY(h, cudaMalloc(&g->sorted, (size_t)nb * cap * sizeof(float4)));
  CUDA_TRY(h, cudaMemsetAsync(g->cell_start, 0, (size_t)nb * GRID_TABLE_STRIDE(T) * sizeof(int), h->stream));
  return ALEGO_OK;
}

void grid_free(GridIndex *g) {
  if (g->cell_start) cudaFree(g->cell_start);
  if (g->sorted) cudaFree(g->sorted);
  g->cell_start = nullptr;
  g->sorted = nullptr;
  g->cap = g->table_size = 0;
}

int grid_build(AlegoHandle *h, GridIndex *g, const float4 *pts, size_t pts_stride, const int *n_ptr, int n_stride, const char *tag,
               bool pack_ring, int batch) {
  const int B = batch > 0 ? batch : h->B, T = g->table_size;
  cudaStream_t s = h->launch_stream ? h->launch_stream : h->stream;
  const float inv = 1.0f / g->cell;
  const int blocks = min(div_up(g->cap, 256), 128);  // capacity-sized clouds are mostly far from full: bounded grid, stride loops
  std::string t0 = std::string("grid_count_") + tag, t1 = std::string("grid_scan_") + tag, t2 = std::string("grid_fill_") + tag;
  CUDA_TRY(h, cudaMemsetAsync(g->cell_start, 0, (size_t)B * GRID_TABLE_STRIDE(T) * sizeof(int), s));
  { LAUNCH(h, t0.c_str());
    grid_count_kernel<<<dim3(blocks, B), 256, 0, s>>>(pts, pts_stride, n_ptr, n_stride, g->cell_start, T, inv, g->cap); }
  { LAUNCH(h, t1.c_str()); grid_scan_kernel<<<dim3((T + 4 + GRID_CHUNK - 1) >> GRID_CHUNK_SHIFT, B), 1024, 0, s>>>(g->cell_start, T); }
  { LAUNCH(h, t2.c_str());
    grid_fill_kernel<<<dim3(blocks, B), 256, 0, s>>>(pts, pts_stride, n_ptr, n_stride, g->cell_start, g->sorted, T, inv, g->cap, pack_ring ? 1 : 0); }
  CUDA_TRY(h, cudaGetLastError());
  return ALEGO_OK;
}
