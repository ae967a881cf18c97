// Stand-in for pcl/io/pcd_io.h (TEST INFRASTRUCTURE, oracle/_ref build only): nothing is written.
#ifndef ALEGO_REF_SHIM_PCL_PCD_IO_H
#define ALEGO_REF_SHIM_PCL_PCD_IO_H
#include <string>
#include <pcl/point_cloud.h>
namespace pcl { namespace io {
template <typename PointT>
int savePCDFile(const std::string &, const PointCloud<PointT> &) { return 0; }
} }
#endif
