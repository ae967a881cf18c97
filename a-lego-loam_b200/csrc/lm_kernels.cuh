// lm_kernels.cuh — host entries of the LaserMapping stage.
#pragma once
#include "common.cuh"

int lm_ensure_buffers(AlegoHandle *h, int need_c, int need_s, int need_o);
int lm_build_map_index(AlegoHandle *h);                                     // laserMapping.cpp:356-357
// once per map change (synchronises): can the surf map use the voxel-row index instead of the hashed grid?
int lm_validate_map_rows(AlegoHandle *h);
// index_ready: the caller has (re)built the local-map index already; launches go to h->launch_stream when it is set
int lm_scan2map_device(AlegoHandle *h, int *guard_dev, bool write_pose, bool index_ready);  // laserMapping.cpp:325-489
int voxel_grid_host(AlegoHandle *h, const float *xyzi, int n, float leaf, float *out_xyzi, int *n_out);
// N1 (laserMapping.cpp:194-323): concatenate host clouds, transform each by its keyframe matrix, VoxelGrid into dst (device)
int lm_assemble_cloud(AlegoHandle *h, const float *const *seg_ptr, const int *seg_n, int n_seg, const float *M_host, int n_mat,
                      int mat_shift, float leaf, float4 *dst, int *n_dst);
// N4 (laserMapping.cpp:652-711): pcl::IterativeClosestPoint of performLoopClosure, host clouds in, result out
int lc_icp_device(AlegoHandle *h, const float *src_host, int n_src, const float *tgt_host, int n_tgt, double max_corr_dist,
                  int max_iterations, double transformation_epsilon, double fitness_epsilon, AlegoIcpResult *out, double *trace_host);
