// Stand-in for pcl::IterativeClosestPoint (TEST INFRASTRUCTURE, oracle/_ref build only).  Loop closure is off the pinned path:
// this class only lets LaserMapping::performLoopClosure compile; align() reports "not converged".
#ifndef ALEGO_REF_SHIM_PCL_ICP_H
#define ALEGO_REF_SHIM_PCL_ICP_H
#include <pcl/point_cloud.h>
namespace pcl {
template <typename S, typename T>
class IterativeClosestPoint {
 public:
  void setMaxCorrespondenceDistance(double) {}
  void setMaximumIterations(int) {}
  void setTransformationEpsilon(double) {}
  void setEuclideanFitnessEpsilon(double) {}
  void setRANSACIterations(int) {}
  void setInputSource(const typename PointCloud<S>::ConstPtr &) {}
  void setInputTarget(const typename PointCloud<T>::ConstPtr &) {}
  void align(PointCloud<S> &) {}
  bool hasConverged() const { return false; }
  double getFitnessScore() const { return 1e30; }
  Eigen::Matrix4f getFinalTransformation() const { return Eigen::Matrix4f::Identity(); }
};
}  // namespace pcl
#endif
