// Small utility kernels shared by the loss chain: accumulator finalisation.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {
__global__ void finalize_kernel(const double* acc, const double* den, float scale, float* out, int n) {
  for (int i = blockIdx.x * JPB_NT + JPB_TID; i < n; i += gridDim.x * JPB_NT) {
    const double d = den ? den[i] : 1.0;
    out[i] = (float)(acc[i] / d * (double)scale);
  }
}
}  // namespace

extern "C" int jpb_finalize(const double* acc, const double* den, float scale, float* out, int n, void* stream) {
  if (!acc || !out || n < 1) return JPB_ERR_ARG;
  JPB_LAUNCH(finalize_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, acc, den, scale, out, n);
  return jpb_status();
}
