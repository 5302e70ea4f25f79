// Evaluation metrics on the device.
//
// Replaces the per-sample host round trip of the reference's validation hook
// (/root/reference/mono/core/evaluation/eval_hooks.py:149-197 and pixel_error.py:27-118): disp -> .cpu() -> cv2.resize
// -> boolean-mask gather -> np.median x2 -> compute_errors, and argmax -> .cpu() -> np.unique / mask stacks for the
// bird's-eye-view IoU / precision.  Here: one compaction pass over the ground-truth frame, one CTA per sample that finds
// the two medians with an exact radix select over the fp32 bit patterns and accumulates the seven depth errors, and one
// counting pass over the BEV logits.  All integer results (counts, medians as order statistics) are exact; the error means
// are fp32 terms summed in double.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

// cv2.resize(INTER_LINEAR) source index and weight of one axis (OpenCV resize.cpp: fx = (dx + 0.5) * scale - 0.5 evaluated in
// double, cast to float; taps clamped at both ends with the weight of the clamped tap forced to 0)
__device__ __forceinline__ void cv_axis(int d, double scale, int n, int& i0, int& i1, float& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= n - 1) { s = n - 1; f = 0.f; }
  i0 = s;
  i1 = min(s + 1, n - 1);
  w1 = f;
}

constexpr int CP_T = 256, CP_PER = 4;   // compaction: threads per block, pixels per thread

__global__ void __launch_bounds__(CP_T) depth_eval_compact_kernel(JpbDepthEvalArgs a) {
  __shared__ int s_n, s_base;
  __shared__ float s_g[CP_T * CP_PER], s_p[CP_T * CP_PER];   // the block's valid (gt, prediction) pairs, in arrival order
  const int b = blockIdx.y;
  const int npix = a.gh * a.gw;
  const float* gt = a.gt + (size_t)b * npix;
  const float* disp = a.disp + (size_t)b * a.h * a.w;
  float* wg = a.work + (size_t)b * 2 * npix;
  float* wp = wg + npix;
  const double sx = (double)a.w / (double)a.gw, sy = (double)a.h / (double)a.gh;
  const float m = a.min_disp, r = a.max_disp - a.min_disp;
  if (JPB_TID == 0) s_n = 0;
  __syncthreads();
  const int first = blockIdx.x * (CP_T * CP_PER);
  for (int i = JPB_TID; i < CP_T * CP_PER; i += JPB_NT) {
    const int p = first + i;
    if (p >= npix) continue;
    const int y = p / a.gw, x = p - y * a.gw;
    const float g = gt[p];
    if (!(g > a.min_depth && g < a.max_depth && y >= a.crop[0] && y < a.crop[1] && x >= a.crop[2] && x < a.crop[3])) continue;
    int x0, x1, y0, y1;
    float wx, wy;
    cv_axis(x, sx, a.w, x0, x1, wx);
    cv_axis(y, sy, a.h, y0, y1, wy);
    // scaled disparity (pixel_error.py:43-48) of the four taps; horizontal pass first, then vertical, as OpenCV
    const float d00 = m + r * disp[y0 * a.w + x0], d01 = m + r * disp[y0 * a.w + x1];
    const float d10 = m + r * disp[y1 * a.w + x0], d11 = m + r * disp[y1 * a.w + x1];
    const float r0 = d00 * (1.f - wx) + d01 * wx, r1 = d10 * (1.f - wx) + d11 * wx;
    const float pd = r0 * (1.f - wy) + r1 * wy;
    const int slot = atomicAdd(&s_n, 1);
    s_g[slot] = g;
    s_p[slot] = 1.f / pd;                                    // pred_depth = 1 / resized disparity (eval_hooks.py:165)
  }
  __syncthreads();
  if (JPB_TID == 0) s_base = s_n ? atomicAdd(&a.count[b], s_n) : 0;
  __syncthreads();
  for (int i = JPB_TID; i < s_n; i += JPB_NT) {
    wg[s_base + i] = s_g[i];
    wp[s_base + i] = s_p[i];
  }
}

// k-th smallest (0-based) of n positive floats: radix select over the bit patterns, 8 bits per pass.  All threads of the
// block call it; the result is returned to every thread.
__device__ float select_kth(const float* v, int n, int k, unsigned* hist, unsigned* bcast) {
  unsigned prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = JPB_TID; i < 256; i += JPB_NT) hist[i] = 0;
    __syncthreads();
    for (int i = JPB_TID; i < n; i += JPB_NT) {
      const unsigned u = __float_as_uint(v[i]);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (JPB_TID == 0) {
      unsigned cum = 0;
      int bin = 0;
      for (; bin < 255; ++bin) {
        if (cum + hist[bin] > (unsigned)k) break;
        cum += hist[bin];
      }
      bcast[0] = (unsigned)bin;
      bcast[1] = cum;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    k -= (int)bcast[1];
    __syncthreads();
  }
  return __uint_as_float(prefix);
}

__device__ __forceinline__ float median_of(const float* v, int n, unsigned* hist, unsigned* bcast) {   // np.median
  const float lo = select_kth(v, n, (n - 1) / 2, hist, bcast);
  if (n & 1) return lo;
  const float hi = select_kth(v, n, n / 2, hist, bcast);
  return (lo + hi) * 0.5f;
}

__global__ void __launch_bounds__(1024) depth_eval_reduce_kernel(JpbDepthEvalArgs a) {
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[2];
  __shared__ double red[32];
  const int b = blockIdx.x;
  const int npix = a.gh * a.gw;
  const int n = a.count[b];
  const float* wg = a.work + (size_t)b * 2 * npix;
  const float* wp = wg + npix;
  double* out = a.out + b * 8;
  if (n <= 0) {   // numpy: mean / median of an empty selection is NaN
    for (int i = JPB_TID; i < 8; i += JPB_NT) out[i] = (double)NAN;
    return;
  }
  const float med_g = median_of(wg, n, hist, bcast), med_p = median_of(wp, n, hist, bcast);
  const float ratio = med_g / med_p;                                   // eval_hooks.py:180
  const float scale = a.fixed_scale > 0.f ? a.fixed_scale : ratio;     // :181-184 (stereo_scale: x36)
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = JPB_TID; i < n; i += JPB_NT) {
    const float g = wg[i];
    float p = wp[i] * scale;
    p = p < a.min_depth ? a.min_depth : (p > a.max_depth ? a.max_depth : p);   // :186-187
    const float th = fmaxf(g / p, p / g);                                      // pixel_error.py:27-40
    const float df = g - p, dl = logf(g) - logf(p);
    s[0] += (double)(fabsf(df) / g);
    s[1] += (double)((df * df) / g);
    s[2] += (double)(df * df);
    s[3] += (double)(dl * dl);
    s[4] += th < 1.25f ? 1.0 : 0.0;
    s[5] += th < 1.5625f ? 1.0 : 0.0;
    s[6] += th < 1.953125f ? 1.0 : 0.0;
  }
  for (int j = 0; j < 7; ++j) {
    const double t = jpb_block_sum<double>(s[j], red);
    if (JPB_TID == 0) {
      const double m = t / (double)n;
      out[j] = (j == 2 || j == 3) ? sqrt(m) : m;
    }
  }
  if (JPB_TID == 0) out[7] = (double)ratio;
}

// prediction = argmax over the two logits (ties -> class 0, as torch.argmax returns the first maximum)
__global__ void __launch_bounds__(256) bev_confusion_kernel(const float* logits, long long sb, long long sc, long long sp, const float* label,
                                                            int npix, long long* counts) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float* l0 = logits + (size_t)b * sb;
  const float* lab = label + (size_t)b * npix;
  float n11 = 0.f, np1 = 0.f, ng1 = 0.f;   // per-thread counts stay far below 2^24: exact in fp32
  for (int p = blockIdx.x * JPB_NT + JPB_TID; p < npix; p += gridDim.x * JPB_NT) {
    const bool pr = l0[(size_t)p * sp + sc] > l0[(size_t)p * sp];
    const bool g = lab[p] != 0.f;
    n11 += (pr && g) ? 1.f : 0.f;
    np1 += pr ? 1.f : 0.f;
    ng1 += g ? 1.f : 0.f;
  }
  const float t0 = jpb_block_sum<float>(n11, red), t1 = jpb_block_sum<float>(np1, red), t2 = jpb_block_sum<float>(ng1, red);
  if (JPB_TID == 0) {
    atomicAdd((unsigned long long*)&counts[b * 3 + 0], (unsigned long long)t0);
    atomicAdd((unsigned long long*)&counts[b * 3 + 1], (unsigned long long)t1);
    atomicAdd((unsigned long long*)&counts[b * 3 + 2], (unsigned long long)t2);
  }
}

}  // namespace

extern "C" int jpb_depth_eval(const JpbDepthEvalArgs* a, void* stream) {
  if (!a || !a->disp || !a->gt || !a->work || !a->count || !a->out || a->B < 1 || a->h < 1 || a->w < 1 || a->gh < 1 || a->gw < 1)
    return JPB_ERR_ARG;
  if ((long long)a->gh * a->gw >= (1ll << 30)) return JPB_ERR_UNSUPPORTED;
  const int npix = a->gh * a->gw;
  dim3 grid((npix + CP_T * CP_PER - 1) / (CP_T * CP_PER), a->B);
  JPB_LAUNCH(depth_eval_compact_kernel, grid, dim3(CP_T), 0, (cudaStream_t)stream, *a);
  JPB_LAUNCH(depth_eval_reduce_kernel, dim3(a->B), dim3(1024), 0, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_bev_confusion(const float* logits, long long stride_b, long long stride_c, long long stride_p, const float* label,
                                 int B, int occ, long long* counts, void* stream) {
  if (!logits || !label || !counts || B < 1 || occ < 1) return JPB_ERR_ARG;
  const int npix = occ * occ;
  int blocks = (npix + 255) / 256;
  if (blocks > 64) blocks = 64;
  JPB_LAUNCH(bev_confusion_kernel, dim3(blocks, B), dim3(256), 0, (cudaStream_t)stream, logits, stride_b, stride_c, stride_p, label, npix, counts);
  return jpb_status();
}
