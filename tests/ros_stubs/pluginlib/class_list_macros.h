// stand-in for <pluginlib/class_list_macros.h>: instantiates the class once so that it is fully type-checked
#pragma once
#define ALEGO_STUB_CAT2(a, b) a##b
#define ALEGO_STUB_CAT(a, b) ALEGO_STUB_CAT2(a, b)
#define PLUGINLIB_EXPORT_CLASS(cls, base) \
  extern "C" base *ALEGO_STUB_CAT(alego_stub_make_, __LINE__)() { return new cls(); }
