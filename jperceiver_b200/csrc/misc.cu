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

// Batched weight re-layout for the data-gradient GEMMs: for every registered convolution weight W [N][taps][Cin]
// (channels-last nn.Conv2d parameter) write WT [Cin][taps][N] with the taps reversed (= flip(2,3) + transpose(0,1)),
// all layers in ONE launch.  Block = one 32x32 (n, ci) tile of one tap of one layer, transposed through shared memory.
__global__ void __launch_bounds__(256) weight_flipT_kernel(const JpbWeightT* ent, int nent) {
  __shared__ float tile[32][33];
  // locate the entry that owns this block (block_start is an exclusive prefix sum, ascending)
  int lo = 0, hi = nent - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ent[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const JpbWeightT e = ent[lo];
  const int local = (int)blockIdx.x - e.block_start;
  const int tn = (e.N + 31) / 32, tc = (e.Cin + 31) / 32;
  const int tap = local / (tn * tc), rem = local - tap * (tn * tc);
  const int n0 = (rem / tc) * 32, c0 = (rem % tc) * 32;
  for (int i = JPB_TID; i < 32 * 32; i += JPB_NT) {
    const int r = i >> 5, c = i & 31;              // r: n offset, c: ci offset (contiguous in W)
    const int n = n0 + r, ci = c0 + c;
    tile[r][c] = (n < e.N && ci < e.Cin) ? e.src[((size_t)n * e.taps + tap) * e.Cin + ci] : 0.f;
  }
  __syncthreads();
  const int tdst = e.taps - 1 - tap;
  for (int i = JPB_TID; i < 32 * 32; i += JPB_NT) {
    const int r = i >> 5, c = i & 31;              // r: ci offset, c: n offset (contiguous in WT)
    const int ci = c0 + r, n = n0 + c;
    if (n < e.N && ci < e.Cin) e.dst[((size_t)ci * e.taps + tdst) * e.N + n] = tile[c][r];
  }
}
}  // namespace

extern "C" int jpb_weight_flipT(const JpbWeightT* entries_dev, int nent, int nblocks, void* stream) {
  if (!entries_dev || nent < 1 || nblocks < 1) return JPB_ERR_ARG;
  JPB_LAUNCH(weight_flipT_kernel, dim3((unsigned)nblocks), dim3(256), 0, (cudaStream_t)stream, entries_dev, nent);
  return jpb_status();
}

extern "C" int jpb_finalize(const double* acc, const double* den, float scale, float* out, int n, void* stream) {
  if (!acc || !out || n < 1) return JPB_ERR_ARG;
  JPB_LAUNCH(finalize_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, acc, den, scale, out, n);
  return jpb_status();
}
