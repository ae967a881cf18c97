// lo_kernels.cuh — host entries of the LaserOdometry stages.
#pragma once
#include "common.cuh"

int lo_extract_device(AlegoHandle *h);    // laserOdometry.cpp:118-297
int lo_scan2scan_device(AlegoHandle *h);  // laserOdometry.cpp:316-535
