// Memory-bound companions of the tensor-core convolutions (NHWC fp32).
//   jpb_act_bwd : dz = dy * act'(y) and the bias gradient (column sums) in one pass — the backward of the
//                 bias / LeakyReLU / ReLU / sigmoid epilogue fused into the forward convolution
//                 (reference: F.leaky_relu depth_decoder.py:60, nn.Sigmoid depth_decoder.py:35-38, nn.ReLU pose_decoder.py:17-20).
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

__device__ __forceinline__ float act_grad(float dy, float y, int act) {
  if (act == 1) return y > 0.f ? dy : 0.f;
  if (act == 2) return y > 0.f ? dy : 0.01f * dy;
  if (act == 3) return dy * y * (1.f - y);
  return dy;
}

// rows x C matrix; thread owns channels tid, tid+nt, ...; block owns a contiguous slab of rows
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* dy, const float* y, float* dz, long long rows, int C, int act, float* dbias) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  long long r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  if (C >= 32) {
    for (int c = JPB_TID; c < C; c += JPB_NT) {
      float s = 0.f;
#pragma unroll 4
      for (long long r = r0; r < r1; ++r) {
        const long long i = r * C + c;
        const float g = act_grad(dy[i], y ? y[i] : 0.f, act);
        if (dz) dz[i] = g;
        s += g;
      }
      if (dbias && r1 > r0) atomicAdd(&dbias[c], s);
    }
  } else {
    // narrow matrices (C = 1, 2, 6, 16): threads stride over the flat slab; per-channel sums through shared bins
    __shared__ float bins[32];
    for (int c = JPB_TID; c < 32; c += JPB_NT) bins[c] = 0.f;
    __syncthreads();
    float s = 0.f;
    int myc = -1;
    const long long e0 = r0 * C, e1 = r1 * C;
    // stride is a multiple of C when blockDim*C... keep it simple: one atomic per element group of equal channel
    for (long long i = e0 + JPB_TID; i < e1; i += JPB_NT) {
      const float g = act_grad(dy[i], y ? y[i] : 0.f, act);
      if (dz) dz[i] = g;
      const int c = (int)(i % C);
      if (c != myc) {
        if (myc >= 0 && dbias) atomicAdd(&bins[myc], s);
        myc = c; s = 0.f;
      }
      s += g;
    }
    if (myc >= 0 && dbias) atomicAdd(&bins[myc], s);
    __syncthreads();
    if (dbias)
      for (int c = JPB_TID; c < C; c += JPB_NT)
        if (bins[c] != 0.f) atomicAdd(&dbias[c], bins[c]);
  }
}

}  // namespace

extern "C" int jpb_act_bwd(const float* dy, const float* y, float* dz, long long rows, int C, int act, float* dbias, void* stream) {
  if (!dy || rows < 1 || C < 1 || (act != 0 && !y) || (!dz && !dbias)) return JPB_ERR_ARG;
  long long blocks = rows / 16 + 1;   // small-extent layers (rows = 480..5120) need many short slabs, not 8 blocks of 64 serial rows
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(act_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, dy, y, dz, rows, C, act, dbias);
  return jpb_status();
}
