// Stand-in nav_msgs/Path (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_NAV_PATH_H
#define ALEGO_REF_SHIM_NAV_PATH_H
#include <vector>
#include <geometry_msgs/PoseWithCovarianceStamped.h>
namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
}  // namespace nav_msgs
#endif
