// ip_kernels.cuh — device-side parameter block and host entry of the ImageProjection stage.
#pragma once
#include "common.cuh"

struct IpDev {
  int R, C, RC, ground_scan_id, seg_valid_point_num, seg_valid_line_num, seg_min_cluster;
  double ang_res_x, ang_res_y, ang_bottom, sensor_mount_ang, seg_theta;
  double row_scale, row_off, col_scale;
  float row_scale_f, row_off_f, col_scale_f, eps_row, eps_col;  // fast-path forms of the row / column scaling (exact forms near cell boundaries)
  double sin_x, cos_x, sin_y, cos_y;  // sin/cos of seg_alpha_x / seg_alpha_y (utility.h:60-61), host libm
};

IpDev make_ip_dev(const AlegoHandle *h);
int ip_run_device(AlegoHandle *h, bool want_labels);
int ip_label_device(AlegoHandle *h);  // label_mat_ on demand
