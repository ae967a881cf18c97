// Stand-in tf::TransformBroadcaster (TEST INFRASTRUCTURE, oracle/_ref build only): transforms go nowhere.
#ifndef ALEGO_REF_SHIM_TF_BROADCASTER_H
#define ALEGO_REF_SHIM_TF_BROADCASTER_H
#include <tf/tf.h>
namespace tf {
struct TransformBroadcaster {
  void sendTransform(const StampedTransform &) {}
};
}  // namespace tf
#endif
