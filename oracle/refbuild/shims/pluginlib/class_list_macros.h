// Stand-in: the plugin registration is a no-op in the oracle/_ref build (TEST INFRASTRUCTURE).
#ifndef ALEGO_REF_SHIM_PLUGINLIB_H
#define ALEGO_REF_SHIM_PLUGINLIB_H
#define PLUGINLIB_EXPORT_CLASS(cls, base)
#endif
