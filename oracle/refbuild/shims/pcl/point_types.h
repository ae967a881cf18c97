// Stand-in for pcl/point_types.h (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_PCL_POINT_TYPES_H
#define ALEGO_REF_SHIM_PCL_POINT_TYPES_H
#include <Eigen/Core>

#define PCL_ADD_POINT4D float x; float y; float z; float data_w;
#define PCL_ADD_INTENSITY float intensity
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)

namespace pcl {
// pcl::PointXYZI: 16-byte xyz1 block + intensity, 32 bytes; the default constructor zeroes x, y, z, intensity (PCL >= 1.7)
struct PointXYZI {
  float x, y, z, data_w;
  float intensity;
  float pad_[3];
  PointXYZI() : x(0.f), y(0.f), z(0.f), data_w(1.f), intensity(0.f), pad_{0.f, 0.f, 0.f} {}
} __attribute__((aligned(16)));
}  // namespace pcl
#endif
