// vox_order.cuh — host entry points of the batch record-ordering kernels (vox_order.cu)
#pragma once
#include "common.cuh"
#include "sort_voxel.cuh"

// partition every list that needs it (state[i].done == 0, no pad keys, more than 16 records): records of list i start at
// buf_a + state[i].off, its scratch (as many ints) at buf_b + state[i].off_b
// group = lists per sequence (0: unknown); max_len = upper bound of a list's length (the launch for longer lists is skipped when none can exist)
int vox_order_lists_by_warp(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag,
                            int group, int max_len);
// group = lists per sequence (state index = sequence * group + kind); the CTAs of kind first_kind are issued first
int vox_order_lists_by_cta(AlegoHandle *h, const VoxState *state, int n_lists, u64 *buf_a, u64 *buf_b, cudaStream_t s, const char *tag,
                           int group, int first_kind);
