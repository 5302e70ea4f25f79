// Batched image preprocessing on the device: what MonoDataset.preprocess does per sample with PIL on
// dataloader workers (/root/reference/mono/datasets/mono_dataset.py:126-171, 202-203, 337-343, 417-431):
//   horizontal flip -> transforms.Resize(ANTIALIAS) -> transforms.ColorJitter -> ToTensor, and the nearest-resize +
//   "== 255" binarisation of the bird's-eye-view labels.
// The arithmetic is Pillow's, bit for bit (this is byte / fixed-point work): Resample.c's two-pass Lanczos with 22-bit
// integer coefficients (the coefficient tables themselves are built on the host in double precision, exactly as
// precompute_coeffs / normalize_coeffs_8bpc do), Blend.c's float blend with truncation, Convert.c's rgb2l / rgb2hsv /
// hsv2rgb.  Pillow is a dependency of the reference, absent from /root/reference: parity is pinned in the tests against
// Pillow itself (12.2, the version in this image) run on the same bytes.
//
// Layouts: decoded frames are uint8 HWC (what a JPEG/PNG decoder emits); float outputs are NCHW in [0,1] (ToTensor).
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

constexpr int PBITS = 22;   // Resample.c PRECISION_BITS = 32 - 8 - 2

__device__ __forceinline__ unsigned char clip8_fixed(int v) {   // clip8(): lookup of (v >> PRECISION_BITS) clamped to 0..255
  v >>= PBITS;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// non-fused float multiply-add: Pillow is compiled for baseline x86-64 (no FMA contraction)
__device__ __forceinline__ float mul_rn(float a, float b) {
#ifdef JPB_HOST_EMU
  volatile float r = a * b;
  return r;
#else
  return __fmul_rn(a, b);
#endif
}
__device__ __forceinline__ float add_rn(float a, float b) {
#ifdef JPB_HOST_EMU
  volatile float r = a + b;
  return r;
#else
  return __fadd_rn(a, b);
#endif
}

__device__ __forceinline__ double dmul_rn(double a, double b) {
#ifdef JPB_HOST_EMU
  volatile double r = a * b;
  return r;
#else
  return __dmul_rn(a, b);
#endif
}
__device__ __forceinline__ double dsub_rn(double a, double b) {
#ifdef JPB_HOST_EMU
  volatile double r = a - b;
  return r;
#else
  return __dsub_rn(a, b);
#endif
}

// ---- Resample.c ImagingResampleHorizontal_8bpc: one thread per output pixel (3 channels)
__global__ void __launch_bounds__(256) resize_h_kernel(const unsigned char* src, unsigned char* dst, int B, int H, int Win, int Wout,
                                                       const int* kx, const int* bx, int ks, const unsigned char* flip) {
  const long long total = (long long)B * H * Wout;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < total; i += (long long)gridDim.x * JPB_NT) {
    const int xx = (int)(i % Wout);
    const long long row = i / Wout;           // b * H + y
    const int b = (int)(row / H);
    const unsigned char* line = src + row * (long long)Win * 3;
    const bool fl = flip && flip[b];
    unsigned char* o = dst + i * 3;
    if (ks == 0) {                            // same width: Pillow skips the horizontal pass
      const int xs = fl ? Win - 1 - xx : xx;
      o[0] = line[xs * 3]; o[1] = line[xs * 3 + 1]; o[2] = line[xs * 3 + 2];
      continue;
    }
    const int xmin = bx[xx * 2], n = bx[xx * 2 + 1];
    const int* k = kx + (long long)xx * ks;
    int s0 = 1 << (PBITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
      const int xs = fl ? Win - 1 - (x + xmin) : x + xmin;
      const int w = k[x];
      s0 += (int)line[xs * 3] * w; s1 += (int)line[xs * 3 + 1] * w; s2 += (int)line[xs * 3 + 2] * w;
    }
    o[0] = clip8_fixed(s0); o[1] = clip8_fixed(s1); o[2] = clip8_fixed(s2);
  }
}

// ---- ImagingResampleVertical_8bpc (+ optional ToTensor output)
__global__ void __launch_bounds__(256) resize_v_kernel(const unsigned char* src, unsigned char* dst, float* dst_f, int B, int Hin, int Hout, int W,
                                                       const int* ky, const int* by, int ks) {
  const long long total = (long long)B * Hout * W;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < total; i += (long long)gridDim.x * JPB_NT) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int yy = (int)(r % Hout), b = (int)(r / Hout);
    const unsigned char* img = src + (long long)b * Hin * W * 3;
    unsigned char c0, c1, c2;
    if (ks == 0) {
      const unsigned char* p = img + ((long long)yy * W + x) * 3;
      c0 = p[0]; c1 = p[1]; c2 = p[2];
    } else {
      const int ymin = by[yy * 2], n = by[yy * 2 + 1];
      const int* k = ky + (long long)yy * ks;
      int s0 = 1 << (PBITS - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < n; ++y) {
        const unsigned char* p = img + ((long long)(y + ymin) * W + x) * 3;
        const int w = k[y];
        s0 += (int)p[0] * w; s1 += (int)p[1] * w; s2 += (int)p[2] * w;
      }
      c0 = clip8_fixed(s0); c1 = clip8_fixed(s1); c2 = clip8_fixed(s2);
    }
    if (dst) { unsigned char* o = dst + i * 3; o[0] = c0; o[1] = c1; o[2] = c2; }
    if (dst_f) {                                // ToTensor: HWC uint8 -> CHW float32 / 255
      const long long pl = (long long)Hout * W;
      float* o = dst_f + (long long)b * 3 * pl + (long long)yy * W + x;
      o[0] = (float)c0 / 255.f; o[pl] = (float)c1 / 255.f; o[2 * pl] = (float)c2 / 255.f;
    }
  }
}

// ---- Pillow pixel operators
__device__ __forceinline__ int rgb2l(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }   // Convert.c L24

// Blend.c ImagingBlend(in1 = degenerate, in2 = image, alpha)
__device__ __forceinline__ int blend8(int in1, int in2, float alpha) {
  const float t = add_rn((float)in1, mul_rn(alpha, (float)(in2 - in1)));
  if (alpha >= 0.f && alpha <= 1.0f) return (int)(unsigned char)(int)t;
  if (t <= 0.0f) return 0;
  if (t >= 255.0f) return 255;
  return (int)t;
}

// Convert.c rgb2hsv_row / hsv2rgb (float / double mix as in the C source)
__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) { uh = 0; us = 0; return; }
  const float cr = (float)(maxc - minc);
  const float s = cr / (float)maxc;
  const float rc = ((float)(maxc - r)) / cr, gc = ((float)(maxc - g)) / cr, bc = ((float)(maxc - b)) / cr;
  float h;
  if (r == maxc) h = bc - gc;
  else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
  else h = (float)(4.0 + (double)gc - (double)rc);
  h = (float)fmod(((double)h / 6.0 + 1.0), 1.0);
  int ih = (int)((double)h * 255.0), is = (int)((double)s * 255.0);
  uh = ih < 0 ? 0 : (ih > 255 ? 255 : ih);
  us = is < 0 ? 0 : (is > 255 ? 255 : is);
}

__device__ __forceinline__ int clip8i(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) { r = g = b = v; return; }
  const double h6 = (double)(float)h * 6.0 / 255.0;
  const int i = (int)floor(h6);
  const float f = (float)(h6 - (double)(float)i);
  const float fs = (float)((double)(float)s / 255.0);
  // explicit round-to-nearest products and differences: nvcc would otherwise contract 1 - a*b into one fused operation
  const double dv = (double)(float)v, dfs = (double)fs, df = (double)f;
  const int p = clip8i((int)round(dmul_rn(dv, dsub_rn(1.0, dfs))));
  const int q = clip8i((int)round(dmul_rn(dv, dsub_rn(1.0, dmul_rn(dfs, df)))));
  const int t = clip8i((int)round(dmul_rn(dv, dsub_rn(1.0, dmul_rn(dfs, dsub_rn(1.0, df))))));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// apply ops order[first..last) of one sample to a pixel; `mean` is the contrast operator's grey level
__device__ __forceinline__ void jitter_pixel(int& r, int& g, int& b, const int* order, const float* fac, int hue_shift, int first, int last, int mean) {
  for (int j = first; j < last; ++j) {
    const int op = order[j];
    if (op == 0) {                                    // ImageEnhance.Brightness: blend(black, image, f)
      r = blend8(0, r, fac[0]); g = blend8(0, g, fac[0]); b = blend8(0, b, fac[0]);
    } else if (op == 1) {                             // ImageEnhance.Contrast: blend(mean grey, image, f)
      r = blend8(mean, r, fac[1]); g = blend8(mean, g, fac[1]); b = blend8(mean, b, fac[1]);
    } else if (op == 2) {                             // ImageEnhance.Color: blend(L(image), image, f)
      const int l = rgb2l(r, g, b);
      r = blend8(l, r, fac[2]); g = blend8(l, g, fac[2]); b = blend8(l, b, fac[2]);
    } else {                                          // F.adjust_hue: HSV round trip with a wrapping uint8 shift of H
      int h, s, v;
      rgb2hsv(r, g, b, h, s, v);
      h = (h + hue_shift) & 255;
      hsv2rgb(h, s, v, r, g, b);
    }
  }
}

__device__ __forceinline__ int contrast_pos(const int* order) {
  for (int j = 0; j < 4; ++j)
    if (order[j] == 1) return j;
  return 4;
}

// pass 1: sum of L over the image as it looks when the contrast operator runs (after the operators that precede it)
__global__ void __launch_bounds__(256) jitter_mean_kernel(JpbJitterArgs a) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  if (a.enable && !a.enable[b]) return;
  const int* order = a.order + b * 4;
  const float* fac = a.factor + b * 4;
  const int cpos = contrast_pos(order);
  const long long npix = (long long)a.H * a.W;
  const unsigned char* img = a.src + (long long)b * npix * 3;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < npix; i += (long long)gridDim.x * JPB_NT) {
    int r = img[i * 3], g = img[i * 3 + 1], bl = img[i * 3 + 2];
    jitter_pixel(r, g, bl, order, fac, a.hue_shift[b], 0, cpos, 0);
    acc += (double)rgb2l(r, g, bl);
  }
  const double tot = jpb_block_sum<double>(acc, red);   // integers below 2^53: exact
  if (JPB_TID == 0) atomicAdd(&a.lsum[b], (unsigned long long)tot);
}

// pass 2: the four operators in the sample's order (+ optional ToTensor output)
__global__ void __launch_bounds__(256) jitter_apply_kernel(JpbJitterArgs a) {
  const int b = blockIdx.y;
  const long long npix = (long long)a.H * a.W;
  const unsigned char* img = a.src + (long long)b * npix * 3;
  const bool on = !a.enable || a.enable[b];
  const int* order = a.order + b * 4;
  const float* fac = a.factor + b * 4;
  // ImageStat.Stat(L).mean[0] = sum / count in double; int(mean + 0.5)
  const int mean = on ? (int)((double)a.lsum[b] / (double)npix + 0.5) : 0;
  const int hs = on ? a.hue_shift[b] : 0;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < npix; i += (long long)gridDim.x * JPB_NT) {
    int r = img[i * 3], g = img[i * 3 + 1], bl = img[i * 3 + 2];
    if (on) jitter_pixel(r, g, bl, order, fac, hs, 0, 4, mean);
    if (a.dst) {
      unsigned char* o = a.dst + ((long long)b * npix + i) * 3;
      o[0] = (unsigned char)r; o[1] = (unsigned char)g; o[2] = (unsigned char)bl;
    }
    if (a.dst_f) {
      float* o = a.dst_f + (long long)b * 3 * npix + i;
      o[0] = (float)r / 255.f; o[npix] = (float)g / 255.f; o[2 * npix] = (float)bl / 255.f;
    }
  }
}

// ---- BEV labels: Image.resize((size, size), NEAREST) + "== 255 -> 1.0" (mono_dataset.py:417-431), optional flip.
// xtab / ytab: Geometry.c ImagingScaleAffine's pre-tabulated source positions (accumulated in double on the host).
__global__ void __launch_bounds__(256) label_kernel(const unsigned char* src, float* dst, int B, int Hin, int Win, int size, const int* xtab,
                                                    const int* ytab, const unsigned char* flip) {
  const long long total = (long long)B * size * size;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < total; i += (long long)gridDim.x * JPB_NT) {
    const int x = (int)(i % size), y = (int)((i / size) % size), b = (int)(i / ((long long)size * size));
    int sx = xtab[x];
    const int sy = ytab[y];
    if (flip && flip[b]) sx = Win - 1 - sx;
    dst[i] = src[((long long)b * Hin + sy) * Win + sx] == 255 ? 1.f : 0.f;
  }
}

int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = 148 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int jpb_resize_lanczos_u8(const JpbResizeArgs* a, void* stream) {
  if (!a || !a->src || !a->tmp || (!a->dst && !a->dst_f) || a->B < 1 || a->Hin < 1 || a->Win < 1 || a->Hout < 1 || a->Wout < 1) return JPB_ERR_ARG;
  if ((a->ksx && (!a->kx || !a->bx)) || (a->ksy && (!a->ky || !a->by))) return JPB_ERR_ARG;
  if ((a->ksx == 0 && a->Win != a->Wout) || (a->ksy == 0 && a->Hin != a->Hout)) return JPB_ERR_ARG;
  JPB_LAUNCH(resize_h_kernel, dim3(grid_for((long long)a->B * a->Hin * a->Wout)), dim3(256), 0, (cudaStream_t)stream, a->src, a->tmp, a->B, a->Hin,
             a->Win, a->Wout, a->kx, a->bx, a->ksx, a->flip);
  JPB_LAUNCH(resize_v_kernel, dim3(grid_for((long long)a->B * a->Hout * a->Wout)), dim3(256), 0, (cudaStream_t)stream, a->tmp, a->dst, a->dst_f, a->B,
             a->Hin, a->Hout, a->Wout, a->ky, a->by, a->ksy);
  return jpb_status();
}

extern "C" int jpb_color_jitter_u8(const JpbJitterArgs* a, void* stream) {
  if (!a || !a->src || (!a->dst && !a->dst_f) || !a->order || !a->factor || !a->hue_shift || !a->lsum || a->B < 1 || a->H < 1 || a->W < 1)
    return JPB_ERR_ARG;
  const long long npix = (long long)a->H * a->W;
  int g = grid_for(npix);
  if (g > 148) g = 148;
  JPB_LAUNCH(jitter_mean_kernel, dim3(g, a->B), dim3(256), 0, (cudaStream_t)stream, *a);
  JPB_LAUNCH(jitter_apply_kernel, dim3(grid_for(npix), a->B), dim3(256), 0, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_bev_label_u8(const unsigned char* src, float* dst, int B, int Hin, int Win, int size, const int* xtab, const int* ytab,
                                const unsigned char* flip, void* stream) {
  if (!src || !dst || !xtab || !ytab || B < 1 || Hin < 1 || Win < 1 || size < 1) return JPB_ERR_ARG;
  JPB_LAUNCH(label_kernel, dim3(grid_for((long long)B * size * size)), dim3(256), 0, (cudaStream_t)stream, src, dst, B, Hin, Win, size, xtab, ytab, flip);
  return jpb_status();
}
