// Stand-in (TEST INFRASTRUCTURE, oracle/_ref build only): everything lives in gtsam/geometry/Pose3.h.
#include <gtsam/geometry/Pose3.h>
