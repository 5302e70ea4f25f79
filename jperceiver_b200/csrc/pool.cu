// NHWC max pooling, forward (+ window-relative arg-max byte) and backward (deterministic gather).
// Reference call sites: nn.MaxPool2d(5,1,2) of the four chained-residual-pooling stages per decoder level
// (layers.py:184-199 — 16 pools per step on 256-channel maps), nn.MaxPool2d(3,2,1) of the ResNet stems
// (resnet.py:91), nn.MaxPool2d(2) of the layout encoder / CVT (layout_model.py:84, CrossViewTransformer.py:44).
// Ties resolve to the first maximum in (ky, kx) scan order, as ATen does, so gradients route identically.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

struct PoolGeom {
  int B, H, W, C, Ho, Wo, k, s, p;
};

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* x, float* y, unsigned char* idx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.B * g.Ho * g.Wo * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long r = t / C4;
    const int ox = (int)(r % g.Wo); r /= g.Wo;
    const int oy = (int)(r % g.Ho);
    const int b = (int)(r / g.Ho);
    const int y0 = oy * g.s - g.p, x0 = ox * g.s - g.p;
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    bool first = true;
    for (int ky = 0; ky < g.k; ++ky) {
      const int iy = y0 + ky;
      if (iy < 0 || iy >= g.H) continue;
      for (int kx = 0; kx < g.k; ++kx) {
        const int ix = x0 + kx;
        if (ix < 0 || ix >= g.W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)(b * g.H + iy) * g.W + ix) * g.C + c4 * 4);
        const int code = ky * g.k + kx;
        if (first) { m0 = v.x; m1 = v.y; m2 = v.z; m3 = v.w; i0 = i1 = i2 = i3 = code; first = false; }
        else {
          if (v.x > m0 || v.x != v.x) { m0 = v.x; i0 = code; }
          if (v.y > m1 || v.y != v.y) { m1 = v.y; i1 = code; }
          if (v.z > m2 || v.z != v.z) { m2 = v.z; i2 = code; }
          if (v.w > m3 || v.w != v.w) { m3 = v.w; i3 = code; }
        }
      }
    }
    const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
    *reinterpret_cast<float4*>(y + o) = make_float4(m0, m1, m2, m3);
    if (idx) {
      idx[o] = (unsigned char)i0; idx[o + 1] = (unsigned char)i1; idx[o + 2] = (unsigned char)i2; idx[o + 3] = (unsigned char)i3;
    }
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* gy, const unsigned char* idx, float* gx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.B * g.H * g.W * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long r = t / C4;
    const int ix = (int)(r % g.W); r /= g.W;
    const int iy = (int)(r % g.H);
    const int b = (int)(r / g.H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    // outputs whose window contains (iy, ix): oy*s - p <= iy <= oy*s - p + k - 1
    int oy_lo = iy + g.p - (g.k - 1);
    oy_lo = oy_lo <= 0 ? 0 : (oy_lo + g.s - 1) / g.s;
    int oy_hi = (iy + g.p) / g.s;
    if (oy_hi > g.Ho - 1) oy_hi = g.Ho - 1;
    int ox_lo = ix + g.p - (g.k - 1);
    ox_lo = ox_lo <= 0 ? 0 : (ox_lo + g.s - 1) / g.s;
    int ox_hi = (ix + g.p) / g.s;
    if (ox_hi > g.Wo - 1) ox_hi = g.Wo - 1;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const int ky = iy - (oy * g.s - g.p);
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const int code = ky * g.k + (ix - (ox * g.s - g.p));
        const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
        const unsigned char* id = idx + o;
        if (id[0] == code || id[1] == code || id[2] == code || id[3] == code) {
          const float4 v = *reinterpret_cast<const float4*>(gy + o);
          if (id[0] == code) a0 += v.x;
          if (id[1] == code) a1 += v.y;
          if (id[2] == code) a2 += v.z;
          if (id[3] == code) a3 += v.w;
        }
      }
    }
    *reinterpret_cast<float4*>(gx + ((size_t)(b * g.H + iy) * g.W + ix) * g.C + c4 * 4) = make_float4(a0, a1, a2, a3);
  }
}

inline unsigned pool_grid(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int jpb_maxpool_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, int k, int s, int p, void* stream) {
  if (!x || !y || (C & 3) || k < 1 || k > 15 || s < 1) return JPB_ERR_ARG;
  PoolGeom g{B, H, W, C, (H + 2 * p - k) / s + 1, (W + 2 * p - k) / s + 1, k, s, p};
  JPB_LAUNCH(maxpool_fwd_kernel, dim3(pool_grid((long long)B * g.Ho * g.Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, y, idx, g);
  return jpb_status();
}

extern "C" int jpb_maxpool_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, int k, int s, int p, void* stream) {
  if (!gy || !idx || !gx || (C & 3) || k < 1 || k > 15 || s < 1) return JPB_ERR_ARG;
  PoolGeom g{B, H, W, C, (H + 2 * p - k) / s + 1, (W + 2 * p - k) / s + 1, k, s, p};
  JPB_LAUNCH(maxpool_bwd_kernel, dim3(pool_grid((long long)B * H * W * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  return jpb_status();
}
