// grid.cuh — device-side lookup helpers of the hashed cell grid (see grid.cu) and the host build entry.
#pragma once
#include "common.cuh"

__host__ __device__ __forceinline__ int grid_coord(float v, float inv_cell) { return (int)floorf(v * inv_cell); }
// Bucket of a cell.  Blocks of 2^GRID_XB consecutive x cells hash together and keep their x order inside the block, so
// the buckets of x-neighbours — and, after the counting sort, their points — are contiguous: a 3-cell x run is one or
// two contiguous ranges instead of three scattered ones.  T is a power of two >= 1024.
// bucket tables: per sequence T + 4 entries followed by GRID_MAX_CHUNKS chunk populations (see grid_scan_kernel)
#define GRID_CHUNK_SHIFT 12
#define GRID_CHUNK (1 << GRID_CHUNK_SHIFT)
#define GRID_MAX_CHUNKS 1028  // tables of up to 4 Mi buckets
#define GRID_TABLE_STRIDE(T) ((size_t)(T) + 4 + GRID_MAX_CHUNKS)
#define GRID_XB 3
__host__ __device__ __forceinline__ int grid_hash(int ix, int iy, int iz, int T) {
  const unsigned h = ((unsigned)(ix >> GRID_XB) * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u);
  return (int)(((h << GRID_XB) & (unsigned)(T - 1)) | ((unsigned)ix & ((1u << GRID_XB) - 1u)));
}

// squared distance accumulated in float exactly like ::flann::L2_Simple<float> (diff = a-b; result += diff*diff)
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, const float4 &p) {
  float r = 0.f, d;
  d = qx - p.x; r += d * d;
  d = qy - p.y; r += d * d;
  d = qz - p.z; r += d * d;
  return r;
}

// batch: number of clouds the index holds (default: one per sequence of the handle)
int grid_alloc(AlegoHandle *h, GridIndex *g, int cap, float cell, int table_factor_log2 = 0, int batch = -1);
void grid_free(GridIndex *g);
// pts: [B] clouds `pts_stride` points apart; point count of sequence b = n_ptr[b * n_stride]
// pack_ring: store int(intensity) (the ring id of LaserOdometry's feature clouds) in bits 24..30 of the index word
#define GRID_RING_SHIFT 24
#define GRID_INDEX_MASK 0x00ffffff
int grid_build(AlegoHandle *h, GridIndex *g, const float4 *pts, size_t pts_stride, const int *n_ptr, int n_stride, const char *tag,
               bool pack_ring = false, int batch = -1);
