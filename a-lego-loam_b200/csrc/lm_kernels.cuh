// lm_kernels.cuh — host entries of the LaserMapping stage.
#pragma once
#include "common.cuh"

int lm_ensure_buffers(AlegoHandle *h, int need_c, int need_s, int need_o);
int lm_build_map_index(AlegoHandle *h);                                     // laserMapping.cpp:356-357
// map_index_event: when non-null the map index was (re)built on another stream; the main stream waits on it before the associations
int lm_scan2map_device(AlegoHandle *h, int *guard_dev, bool write_pose, cudaEvent_t map_index_event = nullptr);  // laserMapping.cpp:325-489
int voxel_grid_host(AlegoHandle *h, const float *xyzi, int n, float leaf, float *out_xyzi, int *n_out);
