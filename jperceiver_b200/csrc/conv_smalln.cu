// 3x3 convolutions with 1..4 output channels (the four disparity heads `dispL`, Cin = 256 read through a nearest 2x
// up-sampling, sigmoid — depth_decoder.py:35-38,68; the four BEV `topview` heads, Cin = 16 -> 2 — layout_model.py:158)
// on the CUDA cores.  As GEMMs these layers waste >= 94 % of the tensor-core N dimension and are bound by re-reading
// the im2col matrix; here one warp owns an output pixel (forward) / one thread owns a channel (weight gradient), the
// input is read through L1 once per tap and the weights sit in shared memory.
//   forward : y[p][n] = act(bias[n] + sum_{tap,c} x[src(p,tap)][c] * w[n][tap][c])
//   wgrad   : dw[n][tap][c] += sum_p dz[p][n] * x[src(p,tap)][c]
// src() folds ReflectionPad2d(1) / zero padding and the nearest 2x up-sampling, as the tensor-core gather does.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

struct SmallGeom {
  int B, Hs, Ws, C, up, Ho, Wo, N, reflect, act;
};

__device__ __forceinline__ bool small_src(const SmallGeom& g, int oy, int ox, int ky, int kx, int& sy, int& sx) {
  int iy = oy - 1 + ky, ix = ox - 1 + kx;
  if (g.reflect) { iy = jpb_reflect(iy, g.Ho); ix = jpb_reflect(ix, g.Wo); }
  else if (iy < 0 || iy >= g.Ho || ix < 0 || ix >= g.Wo) return false;
  sy = g.up ? iy >> 1 : iy;
  sx = g.up ? ix >> 1 : ix;
  return true;
}

__device__ __forceinline__ float small_act(float v, int act) {
  if (act == 1) return v > 0.f ? v : 0.f;
  if (act == 2) return v > 0.f ? v : 0.01f * v;
  if (act == 3) return 1.f / (1.f + expf(-v));
  return v;
}

// one "warp" (32 consecutive threads; a single emulated thread covers all lanes) per output pixel
__global__ void __launch_bounds__(256) smalln_fwd_kernel(const float* x, const float* w, const float* bias, float* y, SmallGeom g) {
  JPB_DYN_SMEM(float, sw);   // [N][9][C]
  const int KW = 9 * g.C;
  for (int i = JPB_TID; i < g.N * KW; i += JPB_NT) sw[i] = w[i];
  __syncthreads();
#ifdef JPB_HOST_EMU
  const int lanes = 1, lane = 0, warps_per_block = 1, warp = 0;
#else
  const int lanes = 32, lane = threadIdx.x & 31, warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5;
#endif
  const long long P = (long long)g.B * g.Ho * g.Wo;
  for (long long p = (long long)blockIdx.x * warps_per_block + warp; p < P; p += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(p / (g.Ho * g.Wo));
    const int rem = (int)(p - (long long)b * g.Ho * g.Wo);
    const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int sy, sx;
        if (!small_src(g, oy, ox, ky, kx, sy, sx)) continue;
        const float* xp = x + ((size_t)(b * g.Hs + sy) * g.Ws + sx) * g.C;
        const float* wp = sw + (ky * 3 + kx) * g.C;
        for (int c = lane * 4; c < g.C; c += lanes * 4) {
          const float4 v = *reinterpret_cast<const float4*>(xp + c);
          for (int n = 0; n < g.N; ++n) {
            const float* wn = wp + n * KW + c;
            acc[n] += v.x * wn[0] + v.y * wn[1] + v.z * wn[2] + v.w * wn[3];
          }
        }
      }
#ifndef JPB_HOST_EMU
    for (int n = 0; n < g.N; ++n)
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
#endif
    if (lane == 0)
      for (int n = 0; n < g.N; ++n) y[p * g.N + n] = small_act(acc[n] + (bias ? bias[n] : 0.f), g.act);
  }
}

// one thread per channel (blockDim.x == C rounded up to 32); each block reduces a slab of pixels
__global__ void __launch_bounds__(256) smalln_wgrad_kernel(const float* x, const float* dz, float* dw, SmallGeom g) {
  const long long P = (long long)g.B * g.Ho * g.Wo;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  long long p1 = p0 + per;
  if (p1 > P) p1 = P;
  for (int c = JPB_TID; c < g.C; c += JPB_NT) {
    float acc[4][9];
    for (int n = 0; n < 4; ++n)
      for (int t = 0; t < 9; ++t) acc[n][t] = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const int b = (int)(p / (g.Ho * g.Wo));
      const int rem = (int)(p - (long long)b * g.Ho * g.Wo);
      const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
      float d[4];
      for (int n = 0; n < g.N; ++n) d[n] = dz[p * g.N + n];
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          int sy, sx;
          if (!small_src(g, oy, ox, ky, kx, sy, sx)) continue;
          const float xv = x[((size_t)(b * g.Hs + sy) * g.Ws + sx) * g.C + c];
          for (int n = 0; n < g.N; ++n) acc[n][ky * 3 + kx] += d[n] * xv;
        }
    }
    if (p1 > p0)
      for (int n = 0; n < g.N; ++n)
        for (int t = 0; t < 9; ++t) atomicAdd(&dw[(size_t)(n * 9 + t) * g.C + c], acc[n][t]);
  }
}

}  // namespace

extern "C" int jpb_conv3x3_smalln_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Hs, int Ws, int C, int up,
                                      int N, int reflect, int act, void* stream) {
  if (!x || !w || !y || N < 1 || N > 4 || (C & 3) || C < 4) return JPB_ERR_ARG;
  SmallGeom g{B, Hs, Ws, C, up, up ? 2 * Hs : Hs, up ? 2 * Ws : Ws, N, reflect, act};
  const size_t smem = (size_t)N * 9 * C * sizeof(float);
  if (smem > 48 * 1024) return JPB_ERR_UNSUPPORTED;
  const long long P = (long long)B * g.Ho * g.Wo;
  long long blocks = (P + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(smalln_fwd_kernel, dim3((unsigned)blocks), dim3(256), smem, (cudaStream_t)stream, x, w, bias, y, g);
  return jpb_status();
}

extern "C" int jpb_conv3x3_smalln_wgrad(const float* x, const float* dz, float* dw, int B, int Hs, int Ws, int C, int up, int N, int reflect,
                                        void* stream) {
  if (!x || !dz || !dw || N < 1 || N > 4 || C < 1) return JPB_ERR_ARG;
  SmallGeom g{B, Hs, Ws, C, up, up ? 2 * Hs : Hs, up ? 2 * Ws : Ws, N, reflect, 0};
  const long long P = (long long)B * g.Ho * g.Wo;
  long long blocks = P / 256 + 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  int threads = ((C + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  JPB_LAUNCH(smalln_wgrad_kernel, dim3((unsigned)blocks), dim3(threads), 0, (cudaStream_t)stream, x, dz, dw, g);
  return jpb_status();
}
