// map_rows.cuh — "voxel row" index over a local-map cloud that is pcl::VoxelGrid OUTPUT (what corner_from_map_ds_ / surf_from_map_ds_
// are, laserMapping.cpp:316-319): such a cloud holds one point per occupied voxel, in ascending voxel index (x fastest, then y,
// then z).  The device replacement of pcl::KdTreeFLANN::setInputCloud (laserMapping.cpp:356-357) then needs NO copy and NO sort of
// the points: per (row = (y, z) voxel pair, 32-voxel word along x) one entry {occupancy mask, index of the word's first point}; the
// point of voxel x of a row is map[start + popc(mask below x)].  A query reads the entries of the <= 4 x 4 rows within the 1 m gate
// and exactly the occupied voxels in range.  Build = bounding box (read 16 B / point) + one streaming pass that sets mask bits
// and word starts (read 16 B / point, write 8 B per touched word) — against count + scan + scatter of the hashed grid.
// Clouds that are not in voxel order (checked, see map_rows_validate) keep using the hashed grid (grid.cuh).
#pragma once
#include "common.cuh"

// struct MapFrame / struct MapRows: common.cuh (the handle holds one)

void map_rows_free(MapRows *m);
// (Re)decides `usable` for the current clouds: sizes the table from the clouds' extents, builds it once and reads the per-sequence
// verdicts back (synchronises; call when the map changed, outside any stream capture).
int map_rows_validate(AlegoHandle *h, MapRows *m, const float4 *pts, size_t pts_stride, const int *n_ptr, float leaf, const char *tag);
// Stream-ordered rebuild (no host synchronisation): bounding box, frame, table.
int map_rows_build(AlegoHandle *h, MapRows *m, const float4 *pts, size_t pts_stride, const int *n_ptr, const char *tag);

__device__ __forceinline__ int mr_voxel(float v, float inv, int min_b) { return (int)floorf(v * inv) - min_b; }
