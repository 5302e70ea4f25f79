// Library identification for the C ABI (include/jpb200.h).
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

#define JPB_STR2(x) #x
#define JPB_STR(x) JPB_STR2(x)

extern "C" int jpb_abi_version(void) { return 1; }

extern "C" const char* jpb_build_info(void) {
#ifdef JPB_HOST_EMU
  return "host-emulation (tests only) " __DATE__;
#else
  return "sm_100a nvcc " JPB_STR(__CUDACC_VER_MAJOR__) "." JPB_STR(__CUDACC_VER_MINOR__) " " __DATE__;
#endif
}
