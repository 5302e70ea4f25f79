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
    if (idx) *reinterpret_cast<unsigned*>(idx + o) = (unsigned)i0 | ((unsigned)i1 << 8) | ((unsigned)i2 << 16) | ((unsigned)i3 << 24);
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
// A thread owns one output column x 4 channels and walks down POOL5_R output rows, keeping the last five row maxima (value +
// arg column) in registers: 5 loads per output row instead of 25.  Round-2 rewrite (the first version cost ~340 thread
// instructions per output and ran at 20 % of the HBM roofline, instruction bound): one NaN-aware compare per candidate
// (!(v <= m) == v > m || isnan(v)), arg columns / codes packed one byte per channel (one select and one 32-bit store for four
// channels), out-of-range taps as -inf behind predicated loads (no branches), the five-row window rotated by full unrolling
// (no register moves), the next row's loads issued before the current row is reduced.
constexpr int POOL5_R = 20;     // multiple of 5: the row window is indexed by (row mod 5) at compile time

struct Row5 {
  float v[4];
  unsigned kx;      // arg column of each channel, one byte each
};

#define JPB_NINF (-INFINITY)

// candidate `c` (value, byte code) against the running maximum, per channel q: first maximum in scan order wins, NaN propagates
#define JPB_POOL_TAKE(m, codes, q, val, code)                                 \
  if (!((val) <= (m))) { (m) = (val); (codes) = ((codes) & ~(0xffu << (8 * (q)))) | ((unsigned)(code) << (8 * (q))); }

__device__ __forceinline__ void pool5_load(const float* x, const PoolGeom& g, int b, int iy, int ox, int c4, float4 t[5], bool& rowok) {
  rowok = iy >= 0 && iy < g.H;
  const float* row = x + ((size_t)(b * g.H + (rowok ? iy : 0)) * g.W) * g.C + c4 * 4;
#pragma unroll
  for (int kx = 0; kx < 5; ++kx) {
    const int ix = ox - 2 + kx;
    const bool ok = rowok && ix >= 0 && ix < g.W;
    t[kx] = ok ? __ldg(reinterpret_cast<const float4*>(row + (size_t)ix * g.C)) : make_float4(JPB_NINF, JPB_NINF, JPB_NINF, JPB_NINF);
  }
}

__device__ __forceinline__ void pool5_hmax(const float4 t[5], Row5& r) {
  // start from the centre tap's code (always inside the image) at -inf: an out-of-range tap (-inf) never wins
  r.v[0] = r.v[1] = r.v[2] = r.v[3] = JPB_NINF;
  r.kx = 0x02020202u;
#pragma unroll
  for (int kx = 0; kx < 5; ++kx) {
    JPB_POOL_TAKE(r.v[0], r.kx, 0, t[kx].x, kx)
    JPB_POOL_TAKE(r.v[1], r.kx, 1, t[kx].y, kx)
    JPB_POOL_TAKE(r.v[2], r.kx, 2, t[kx].z, kx)
    JPB_POOL_TAKE(r.v[3], r.kx, 3, t[kx].w, kx)
  }
}

__global__ void __launch_bounds__(256) maxpool5_fwd_kernel(const float* x, float* y, unsigned char* idx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const int runs = (g.Ho + POOL5_R - 1) / POOL5_R;
  const long long total = (long long)g.B * runs * g.Wo * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long rr = t / C4;
    const int ox = (int)(rr % g.Wo); rr /= g.Wo;
    const int run = (int)(rr % runs);
    const int b = (int)(rr / runs);
    const int oy0 = run * POOL5_R;
    Row5 w[5];          // w[(iy + 2) mod 5 relative to oy0]: row maxima of input rows oy-2 .. oy+2
    float4 tap[5];
    bool ok;
#pragma unroll
    for (int j = 0; j < 4; ++j) {      // rows oy0-2 .. oy0+1
      pool5_load(x, g, b, oy0 - 2 + j, ox, c4, tap, ok);
      pool5_hmax(tap, w[j]);
    }
    pool5_load(x, g, b, oy0 + 2, ox, c4, tap, ok);
#pragma unroll
    for (int j = 0; j < POOL5_R; ++j) {
      const int oy = oy0 + j;
      if (oy >= g.Ho) break;
      pool5_hmax(tap, w[(j + 4) % 5]);                       // input row oy + 2
      if (j + 1 < POOL5_R) pool5_load(x, g, b, oy + 3, ox, c4, tap, ok);   // next row's taps are in flight during the reduction
      float m[4] = {JPB_NINF, JPB_NINF, JPB_NINF, JPB_NINF};
      unsigned code = 0x0c0c0c0cu;                           // centre (ky = 2, kx = 2) = 12
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const Row5& r = w[(j + ky) % 5];
        const unsigned rc = r.kx + 0x05050505u * (unsigned)ky;   // byte-wise ky * 5 + kx (no carries: < 25)
        JPB_POOL_TAKE(m[0], code, 0, r.v[0], (rc >> 0) & 0xffu)
        JPB_POOL_TAKE(m[1], code, 1, r.v[1], (rc >> 8) & 0xffu)
        JPB_POOL_TAKE(m[2], code, 2, r.v[2], (rc >> 16) & 0xffu)
        JPB_POOL_TAKE(m[3], code, 3, r.v[3], (rc >> 24) & 0xffu)
      }
      const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
      *reinterpret_cast<float4*>(y + o) = make_float4(m[0], m[1], m[2], m[3]);
      if (idx) *reinterpret_cast<unsigned*>(idx + o) = code;
    }
  }
}

// backward of the 5x5 / stride 1 / pad 2 pool as a GATHER: every input element sums the gradients of the (up to) 25 outputs whose
// window contains it and whose arg-max code points at it.  No atomics, no zero fill of gx, deterministic; the scatter it replaces
// issued 21 M scalar reductions per launch at the largest level.
__global__ void __launch_bounds__(256) maxpool5_bwd_kernel(const float* gy, const unsigned char* idx, float* gx, PoolGeom g) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.B * g.H * g.W * C4;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int c4 = (int)(t % C4);
    long long r = t / C4;
    const int ix = (int)(r % g.W); r /= g.W;
    const int iy = (int)(r % g.H);
    const int b = (int)(r / g.H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    // all 25 arg-max words first (independent loads, out-of-range windows read as "no match"), then the few matching gradients
    unsigned eqs[25];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int oy = iy + 2 - ky;            // the output row whose window row ky is iy
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ox = ix + 2 - kx;
        const bool ok = oy >= 0 && oy < g.Ho && ox >= 0 && ox < g.Wo;
        const size_t o = ((size_t)(b * g.Ho + (ok ? oy : 0)) * g.Wo + (ok ? ox : 0)) * g.C + c4 * 4;
        const unsigned codes = __ldg(reinterpret_cast<const unsigned*>(idx + o));
        eqs[ky * 5 + kx] = ok ? codes ^ ((unsigned)(ky * 5 + kx) * 0x01010101u) : 0xffffffffu;   // a zero byte = this channel's arg-max is (ky, kx)
      }
    }
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int oy = iy + 2 - ky;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const unsigned eq = eqs[ky * 5 + kx];
        if (((eq - 0x01010101u) & ~eq & 0x80808080u) == 0u) continue;  // no zero byte: nothing to add (the common case)
        const int ox = ix + 2 - kx;
        const size_t o = ((size_t)(b * g.Ho + oy) * g.Wo + ox) * g.C + c4 * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(gy + o));
        if (!(eq & 0x000000ffu)) a0 += v.x;
        if (!(eq & 0x0000ff00u)) a1 += v.y;
        if (!(eq & 0x00ff0000u)) a2 += v.z;
        if (!(eq & 0xff000000u)) a3 += v.w;
      }
    }
    *reinterpret_cast<float4*>(gx + ((size_t)(b * g.H + iy) * g.W + ix) * g.C + c4 * 4) = make_float4(a0, a1, a2, a3);
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

int g_pool_bwd = 0;   // 0: default (5x5: scatter, the rest: gather); 1: scatter for every overlapping window (round 1); 2: 5x5 gather

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

extern "C" int jpb_maxpool_set_bwd_variant(int v) {
  if (v < 0 || v > 2) return JPB_ERR_ARG;
  g_pool_bwd = v;
  return JPB_OK;
}

extern "C" int jpb_maxpool_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, int k, int s, int p, void* stream) {
  if (!gy || !idx || !gx || (C & 3) || k < 1 || k > 15 || s < 1) return JPB_ERR_ARG;
  PoolGeom g{B, H, W, C, (H + 2 * p - k) / s + 1, (W + 2 * p - k) / s + 1, k, s, p};
  // the caller does not zero-fill gx: the gather kernels write every element once, the scatter path clears gx itself
  const bool five = k == 5 && s == 1 && p == 2;
  if (five && g_pool_bwd == 2) {
    // A/B partner: 5x5 as a gather over the 25 covering windows — deterministic, but measured SLOWER on B200 (246 us vs 149 us for
    // 256 channels @80x256, B = 4: 25 dependent index loads per thread), so the default for 5x5 stays the scatter
    JPB_LAUNCH(maxpool5_bwd_kernel, dim3(pool_grid((long long)B * H * W * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  } else if (five || (g_pool_bwd == 1 && k * k > 4 * s * s)) {
    // scatter with red.global.add over a cleared gx (5x5 / stride 1: the default; other overlapping windows: variant 1)
#ifndef JPB_HOST_EMU
    if (cudaMemsetAsync(gx, 0, (size_t)B * H * W * C * sizeof(float), (cudaStream_t)stream) != cudaSuccess) return jpb_status();
#else
    for (size_t i = 0; i < (size_t)B * H * W * C; ++i) gx[i] = 0.f;
#endif
    JPB_LAUNCH(maxpool_bwd_scatter_kernel, dim3(pool_grid((long long)B * g.Ho * g.Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  } else {
    // gather over the (few) windows that contain the input element: 3x3 / stride 2 (stems), 2x2 / stride 2
    JPB_LAUNCH(maxpool_bwd_kernel, dim3(pool_grid((long long)B * H * W * (C / 4))), dim3(256), 0, (cudaStream_t)stream, gy, idx, gx, g);
  }
  return jpb_status();
}
