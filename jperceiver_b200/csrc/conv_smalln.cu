// 3x3 stride-1 pad-1 convolutions with 1..2 output channels (the four disparity heads `dispL`: Cin = 256 read
// through a nearest 2x up-sampling, sigmoid — depth_decoder.py:35-38,68; the four BEV `topview` heads: Cin = 16 -> 2 —
// layout_model.py:158) on the CUDA cores, in "project, then shift-and-add" form.
//
// As GEMMs these layers waste >= 94 % of the tensor-core N dimension and are bound by re-reading the im2col matrix
// (9 x the input, at 4 x the pixel count when the input is up-sampled).  Because N is tiny, the 3x3 convolution is
// re-associated as nine 1x1 projections of the SOURCE pixels followed by a gather over taps:
//   forward : D[s][n,t] = sum_c x[s][c] w[n][t][c]              (reads x exactly once)
//             y[p][n]   = act(bias[n] + sum_t D[src(p,t)][n,t])  (tiny)
//   backward: G[s][n,t] = sum_{p : src(p,t) = s} dz[p][n]        (tiny, adjoint of the gather)
//             dw[n][t][c] += sum_s x[s][c] G[s][n,t]             (reads x exactly once)
//             dx[s][c]   = sum_{n,t} G[s][n,t] w[n][t][c]        (writes dx exactly once)
// src(p,t) folds ReflectionPad2d(1) / zero padding and the nearest 2x up-sampling.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

constexpr int MAXNT = 18;   // N * 9 with N <= 2

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

#if defined(JPB_HOST_EMU) && !defined(JPB_HOST_EMU_MT)
#define SM_LANES 1
#define SM_LANE 0
#define SM_WARPS 1
#define SM_WARP 0
#else
#define SM_LANES 32
#define SM_LANE (threadIdx.x & 31)
#define SM_WARPS (blockDim.x >> 5)
#define SM_WARP (threadIdx.x >> 5)
#endif

// D[s][n*9+t] = sum_c x[s][c] * w[n][t][c]; one warp per source pixel, lanes stride the channels in float4s
__global__ void __launch_bounds__(256) smalln_project_kernel(const float* x, const float* w, float* D, long long S, int C, int NT) {
  JPB_DYN_SMEM(float, sw);   // [NT][C]
  for (int i = JPB_TID; i < NT * C; i += JPB_NT) sw[i] = w[i];
  __syncthreads();
  if (C <= 32) {   // narrow inputs (topview heads, C = 16): one thread per source pixel, weights broadcast from shared memory
    for (long long s = (long long)blockIdx.x * JPB_NT + JPB_TID; s < S; s += (long long)gridDim.x * JPB_NT) {
      float acc[MAXNT];
      for (int j = 0; j < MAXNT; ++j) acc[j] = 0.f;
      const float* xp = x + (size_t)s * C;
      for (int c = 0; c < C; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xp + c);
        for (int j = 0; j < NT; ++j) {
          const float4 wj = *reinterpret_cast<const float4*>(sw + j * C + c);
          acc[j] += v.x * wj.x + v.y * wj.y + v.z * wj.z + v.w * wj.w;
        }
      }
      for (int j = 0; j < NT; ++j) D[s * NT + j] = acc[j];
    }
    return;
  }
#if !defined(JPB_HOST_EMU) || defined(JPB_HOST_EMU_MT)
  if (C == 256 && NT == 9) {
    // the disparity heads (256 channels -> 1): a lane owns channels [4 lane, +4) and [128 + 4 lane, +4) and keeps their 9 x 2 weight
    // vectors in registers — the first version re-read them from shared memory with 36 scalar loads per input vector (ncu: 487
    // warp instructions per pixel, MIO-throttled at 30 % issue)
    float4 wr[9][2];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      wr[j][0] = *reinterpret_cast<const float4*>(sw + j * 256 + SM_LANE * 4);
      wr[j][1] = *reinterpret_cast<const float4*>(sw + j * 256 + 128 + SM_LANE * 4);
    }
    for (long long s = (long long)blockIdx.x * SM_WARPS + SM_WARP; s < S; s += (long long)gridDim.x * SM_WARPS) {
      const float* xp = x + (size_t)s * 256;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(xp + SM_LANE * 4));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(xp + 128 + SM_LANE * 4));
      float acc[9];
#pragma unroll
      for (int j = 0; j < 9; ++j)
        acc[j] = (v0.x * wr[j][0].x + v0.y * wr[j][0].y + v0.z * wr[j][0].z + v0.w * wr[j][0].w) +
                 (v1.x * wr[j][1].x + v1.y * wr[j][1].y + v1.z * wr[j][1].z + v1.w * wr[j][1].w);
#pragma unroll
      for (int j = 0; j < 9; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      if (SM_LANE < 9) {
        float outv = acc[0];
#pragma unroll
        for (int j = 1; j < 9; ++j) outv = SM_LANE == j ? acc[j] : outv;
        D[s * 9 + SM_LANE] = outv;          // nine lanes store the nine taps: one coalesced 36-byte write per pixel
      }
    }
    return;
  }
#endif
  for (long long s = (long long)blockIdx.x * SM_WARPS + SM_WARP; s < S; s += (long long)gridDim.x * SM_WARPS) {
    float acc[MAXNT];
    for (int j = 0; j < MAXNT; ++j) acc[j] = 0.f;
    const float* xp = x + (size_t)s * C;
    for (int c = SM_LANE * 4; c < C; c += SM_LANES * 4) {
      const float4 v = *reinterpret_cast<const float4*>(xp + c);
      for (int j = 0; j < NT; ++j) {
        const float4 wj = *reinterpret_cast<const float4*>(sw + j * C + c);
        acc[j] += v.x * wj.x + v.y * wj.y + v.z * wj.z + v.w * wj.w;
      }
    }
#if !defined(JPB_HOST_EMU) || defined(JPB_HOST_EMU_MT)
    for (int j = 0; j < NT; ++j)
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#endif
    if (SM_LANE == 0)
      for (int j = 0; j < NT; ++j) D[s * NT + j] = acc[j];
  }
}

// y[p][n] = act(bias[n] + sum_t D[src(p,t)][n*9+t])
__global__ void __launch_bounds__(256) smalln_gather_kernel(const float* D, const float* bias, float* y, SmallGeom g) {
  const long long P = (long long)g.B * g.Ho * g.Wo;
  const int NT = g.N * 9;
  for (long long p = (long long)blockIdx.x * JPB_NT + JPB_TID; p < P; p += (long long)gridDim.x * JPB_NT) {
    const int b = (int)(p / (g.Ho * g.Wo));
    const int rem = (int)(p - (long long)b * g.Ho * g.Wo);
    const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
    float acc[2] = {0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int sy, sx;
        if (!small_src(g, oy, ox, ky, kx, sy, sx)) continue;
        const float* d = D + ((size_t)(b * g.Hs + sy) * g.Ws + sx) * NT + ky * 3 + kx;
        for (int n = 0; n < g.N; ++n) acc[n] += d[n * 9];
      }
    for (int n = 0; n < g.N; ++n) y[p * g.N + n] = small_act(acc[n] + (bias ? bias[n] : 0.f), g.act);
  }
}

// G[src(p,t)][n*9+t] += dz[p][n]   (G zero-filled by the caller)
__global__ void __launch_bounds__(256) smalln_adjoint_kernel(const float* dz, float* G, SmallGeom g) {
  const long long P = (long long)g.B * g.Ho * g.Wo;
  const int NT = g.N * 9;
  for (long long p = (long long)blockIdx.x * JPB_NT + JPB_TID; p < P; p += (long long)gridDim.x * JPB_NT) {
    const int b = (int)(p / (g.Ho * g.Wo));
    const int rem = (int)(p - (long long)b * g.Ho * g.Wo);
    const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
    float d[2];
    for (int n = 0; n < g.N; ++n) d[n] = dz[p * g.N + n];
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int sy, sx;
        if (!small_src(g, oy, ox, ky, kx, sy, sx)) continue;
        float* t = G + ((size_t)(b * g.Hs + sy) * g.Ws + sx) * NT + ky * 3 + kx;
        for (int n = 0; n < g.N; ++n) atomicAdd(t + n * 9, d[n]);
      }
  }
}

// dw[j][c] += sum_s x[s][c] * G[s][j]; a thread owns 4 channels and one of the block's pixel lanes; partial sums are
// reduced across pixel lanes in shared memory so each block issues one atomic per (tap, channel)
__global__ void __launch_bounds__(256) smalln_wgrad_kernel(const float* x, const float* G, float* dw, long long S, int C, int NT) {
  JPB_DYN_SMEM(float, part);   // [36][U], U = C4 * lanes_s <= 256
  const int C4 = C >> 2;
  const int lanes_s = 256 / C4 > 0 ? 256 / C4 : 1;   // pixel lanes (the launch uses 256 threads; emulation loops over u)
  const int U = C4 * lanes_s;
  const long long per = (S + gridDim.x - 1) / gridDim.x;
  const long long s0 = (long long)blockIdx.x * per;
  long long s1 = s0 + per;
  if (s1 > S) s1 = S;
  for (int j0 = 0; j0 < NT; j0 += 9) {            // one output channel (9 taps) at a time: 36 accumulators per thread
    for (int u = JPB_TID; u < U; u += JPB_NT) {
      const int cg = u % C4, ls = u / C4;
      float acc[9][4];
      for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
      for (long long s = s0 + ls; s < s1; s += lanes_s) {
        const float4 v = *reinterpret_cast<const float4*>(x + (size_t)s * C + cg * 4);
        const float* gp = G + s * NT + j0;
        for (int t = 0; t < 9; ++t) {
          const float gv = gp[t];
          acc[t][0] += gv * v.x; acc[t][1] += gv * v.y; acc[t][2] += gv * v.z; acc[t][3] += gv * v.w;
        }
      }
      for (int t = 0; t < 9; ++t)
        for (int q = 0; q < 4; ++q) part[(t * 4 + q) * U + u] = acc[t][q];
    }
    __syncthreads();
    for (int i = JPB_TID; i < 9 * C; i += JPB_NT) {
      const int t = i / C, c = i - t * C;
      const float* pp = part + (t * 4 + (c & 3)) * U + (c >> 2);
      float sum = 0.f;
      for (int ls = 0; ls < lanes_s; ++ls) sum += pp[ls * C4];
      if (s1 > s0) atomicAdd(&dw[(size_t)(j0 + t) * C + c], sum);
    }
    __syncthreads();
  }
}

// dx[s][c] = sum_j G[s][j] * w[j][c]
__global__ void __launch_bounds__(256) smalln_dgrad_kernel(const float* G, const float* w, float* dx, long long S, int C, int NT) {
  JPB_DYN_SMEM(float, sw);   // [NT][C]
  for (int i = JPB_TID; i < NT * C; i += JPB_NT) sw[i] = w[i];
  __syncthreads();
  const int C4 = C >> 2;
  const long long total = S * C4;
  for (long long e = (long long)blockIdx.x * JPB_NT + JPB_TID; e < total; e += (long long)gridDim.x * JPB_NT) {
    const long long s = e / C4;
    const int c = (int)(e - s * C4) * 4;
    const float* gp = G + s * NT;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < NT; ++j) {
      const float gv = gp[j];
      const float4 wj = *reinterpret_cast<const float4*>(sw + j * C + c);
      o.x += gv * wj.x; o.y += gv * wj.y; o.z += gv * wj.z; o.w += gv * wj.w;
    }
    *reinterpret_cast<float4*>(dx + (size_t)s * C + c) = o;
  }
}

inline unsigned sm_grid(long long work, int per_block, int cap = 148 * 8) {
  long long g = (work + per_block - 1) / per_block;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

inline bool small_ok(int N, int C) { return N >= 1 && N <= 2 && C >= 4 && (C & 3) == 0 && (size_t)N * 9 * C * 4 <= 48 * 1024; }

}  // namespace

extern "C" int jpb_conv3x3_smalln_fwd(const float* x, const float* w, const float* bias, float* y, float* work, int B, int Hs, int Ws,
                                      int C, int up, int N, int reflect, int act, void* stream) {
  if (!x || !w || !y || !work || !small_ok(N, C)) return JPB_ERR_ARG;
  SmallGeom g{B, Hs, Ws, C, up, up ? 2 * Hs : Hs, up ? 2 * Ws : Ws, N, reflect, act};
  const long long S = (long long)B * Hs * Ws;
  JPB_LAUNCH(smalln_project_kernel, dim3(sm_grid(S, C <= 32 ? 256 : 8)), dim3(256), (size_t)N * 9 * C * 4, (cudaStream_t)stream, x, w, work, S, C, N * 9);
  JPB_LAUNCH(smalln_gather_kernel, dim3(sm_grid((long long)B * g.Ho * g.Wo, 256)), dim3(256), 0, (cudaStream_t)stream, work, bias, y, g);
  return jpb_status();
}

extern "C" int jpb_conv3x3_smalln_bwd(const float* x, const float* w, const float* dz, float* work, float* dw, float* dx, int B, int Hs,
                                      int Ws, int C, int up, int N, int reflect, void* stream) {
  if (!x || !w || !dz || !work || !small_ok(N, C) || (!dw && !dx)) return JPB_ERR_ARG;
  SmallGeom g{B, Hs, Ws, C, up, up ? 2 * Hs : Hs, up ? 2 * Ws : Ws, N, reflect, 0};
  const long long S = (long long)B * Hs * Ws;
  JPB_LAUNCH(smalln_adjoint_kernel, dim3(sm_grid((long long)B * g.Ho * g.Wo, 256)), dim3(256), 0, (cudaStream_t)stream, dz, work, g);
  if (dw) {
    if (C / 4 > 256) return JPB_ERR_UNSUPPORTED;
    const int C4 = C / 4, lanes_s = 256 / C4 > 0 ? 256 / C4 : 1;
    JPB_LAUNCH(smalln_wgrad_kernel, dim3(sm_grid(S, 256, 148 * 2)), dim3(256), (size_t)36 * C4 * lanes_s * sizeof(float), (cudaStream_t)stream,
               x, work, dw, S, C, N * 9);
  }
  if (dx)
    JPB_LAUNCH(smalln_dgrad_kernel, dim3(sm_grid(S * (C / 4), 256)), dim3(256), (size_t)N * 9 * C * 4, (cudaStream_t)stream, work, w, dx, S, C, N * 9);
  return jpb_status();
}
