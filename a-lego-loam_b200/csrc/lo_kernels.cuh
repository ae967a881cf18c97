// lo_kernels.cuh — host entries of the LaserOdometry stages.
#pragma once
#include "common.cuh"

int lo_extract_device(AlegoHandle *h);    // laserOdometry.cpp:118-297
int lo_scan2scan_device(AlegoHandle *h);  // laserOdometry.cpp:316-535
// N2 (laserOdometry.cpp:557-726, IMU branch): in place on seg_cloud; queues [B][10][len] doubles
int lo_adjust_distortion_device(AlegoHandle *h, const double *queues_dev, int len, const int *ptr_last_dev, int *ptr_last_iter_dev,
                                const double *scan_time_dev, double scan_period, int *n_done_dev, float *start_pose_dev);
