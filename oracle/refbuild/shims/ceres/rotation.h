// Stand-in (TEST INFRASTRUCTURE, oracle/_ref build only): the reference includes ceres/rotation.h but uses nothing from it.
#ifndef ALEGO_REF_SHIM_CERES_ROTATION_H
#define ALEGO_REF_SHIM_CERES_ROTATION_H
#endif
