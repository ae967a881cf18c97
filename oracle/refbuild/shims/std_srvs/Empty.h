// Stand-in std_srvs/Empty (TEST INFRASTRUCTURE, oracle/_ref build only).
#ifndef ALEGO_REF_SHIM_STD_SRVS_EMPTY_H
#define ALEGO_REF_SHIM_STD_SRVS_EMPTY_H
namespace std_srvs {
struct Empty { struct Request {}; struct Response {}; };
}  // namespace std_srvs
#endif
