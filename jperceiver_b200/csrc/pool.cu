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

// ---- 5x5 / stride 1 / pad 2 (the 16 chained-residual-pooling pools per step): separable running maximum.
// A thread owns one output column x 4 channels and walks down POOL5_R output rows, keeping the last five row maxima
// (value + arg column) in registers: 5 loads per output row instead of 25.
constexpr int POOL5_R = 16;

struct Row5 {
  float v[4];
  int kx[4];
  bool valid;
};

__device__ __forceinline__ void pool5_row(const float* x, const PoolGeom& g, int b, int iy, int ox, int c4, Row5& r) {
  r.valid = iy >= 0 && iy < g.H;
  if (!r.valid) return;
  bool first = true;
  for (int kx = 0; kx < 5; ++kx) {
    const int ix = ox - 2 + kx;
    if (ix < 0 || ix >= g.W) continue;
    const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)(b * g.H + iy) * g.W + ix) * g.C + c4 * 4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    for (int q = 0; q < 4; ++q)
      if (first || vv[q] > r.v[q] || vv[q] != vv[q]) { r.v[q] = vv[q]; r.kx[q] = kx; }
    first = false;
  }
}

__global__ void __launch_bounds__(256) maxpool5_fwd_kernel(const float* x, float* y, unsigned char* idx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const int runs = (g.Ho + POOL5_R - 1) / POOL5_R;
  const long long total = (long long)g.B * runs * g.Wo * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long r = t / C4;
    const int ox = (int)(r % g.Wo); r /= g.Wo;
    const int run = (int)(r % runs);
    const int b = (int)(r / runs);
    const int oy0 = run * POOL5_R;
    Row5 w0, w1, w2, w3, w4;   // rows oy-2 .. oy+2 of the current output row
    pool5_row(x, g, b, oy0 - 2, ox, c4, w0);
    pool5_row(x, g, b, oy0 - 1, ox, c4, w1);
    pool5_row(x, g, b, oy0, ox, c4, w2);
    pool5_row(x, g, b, oy0 + 1, ox, c4, w3);
    for (int oy = oy0; oy < oy0 + POOL5_R && oy < g.Ho; ++oy) {
      pool5_row(x, g, b, oy + 2, ox, c4, w4);
      float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int code[4] = {0, 0, 0, 0};
      bool first = true;
      const Row5* rows[5] = {&w0, &w1, &w2, &w3, &w4};
      for (int ky = 0; ky < 5; ++ky) {
        const Row5& rr = *rows[ky];
        if (!rr.valid) continue;
        for (int q = 0; q < 4; ++q)
          if (first || rr.v[q] > m[q] || rr.v[q] != rr.v[q]) { m[q] = rr.v[q]; code[q] = ky * 5 + rr.kx[q]; }
        first = false;
      }
      const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
      *reinterpret_cast<float4*>(y + o) = make_float4(m[0], m[1], m[2], m[3]);
      if (idx) { idx[o] = (unsigned char)code[0]; idx[o + 1] = (unsigned char)code[1]; idx[o + 2] = (unsigned char)code[2]; idx[o + 3] = (unsigned char)code[3]; }
      w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
  }
}

// backward by scatter: every output element adds its gradient to its arg-max input (gx zero-filled by the caller).
// One scalar red.global.add per element; targets are spread over the map, so contention is low.
__global__ void __launch_bounds__(256) maxpool_bwd_scatter_kernel(const float* gy, const unsigned char* idx, float* gx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.B * g.Ho * g.Wo * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long r = t / C4;
    const int ox = (int)(r % g.Wo); r /= g.Wo;
    const int oy = (int)(r % g.Ho);
    const int b = (int)(r / g.Ho);
    const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
    const float4 v = *reinterpret_cast<const float4*>(gy + o);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    for (int q = 0; q < 4; ++q) {
      const int code = idx[o + q];
      const int iy = oy * g.s - g.p + code / g.k, ix = ox * g.s - g.p + code % g.k;
      atomicAdd(gx + ((size_t)(b * g.H + iy) * g.W + ix) * g.C + c4 * 4 + q, vv[q]);
    }
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
  if (k == 5 && s == 1 && p == 2) {
    const long long runs = (g.Ho + POOL5_R - 1) / POOL5_R;
    JPB_LAUNCH(maxpool5_fwd_kernel, dim3(pool_grid((long long)B * runs * g.Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, y, idx, g);
  } else {
    JPB_LAUNCH(maxpool_fwd_kernel, dim3(pool_grid((long long)B * g.Ho * g.Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, y, idx, g);
  }
  return jpb_status();
}

extern "C" int jpb_maxpool_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, int k, int s, int p, void* stream) {
  if (!gy || !idx || !gx || (C & 3) || k < 1 || k > 15 || s < 1) return JPB_ERR_ARG;
  PoolGeom g{B, H, W, C, (H + 2 * p - k) / s + 1, (W + 2 * p - k) / s + 1, k, s, p};
  if (k * k > 4 * s * s) {
    // overlapping windows (5x5/1, 3x3/2): scatter with atomics; gx must be zero-filled by the caller
    JPB_LAUNCH(maxpool_bwd_scatter_kernel, dim3(pool_grid((long long)B * g.Ho * g.Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  } else {
    JPB_LAUNCH(maxpool_bwd_kernel, dim3(pool_grid((long long)B * H * W * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  }
  return jpb_status();
}
