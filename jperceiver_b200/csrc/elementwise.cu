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
  if (C >= 32 && (C & 3) == 0 && ((((uintptr_t)dy) | ((uintptr_t)y) | ((uintptr_t)dz)) & 15) == 0) {
    // 16-byte accesses: a thread owns one float4 of channels and one of the block's row lanes (the scalar form below moved 4 bytes
    // per access with four loads in flight per thread); the row lanes' bias sums meet in shared memory: one atomic per channel and block
    JPB_DYN_SMEM(float, part);                          // [4][256]
    const int C4 = C >> 2;
    const int Ct = C4 < 256 ? C4 : 256;                 // float4 channel groups per pass
    const int lanes_r = 256 / Ct > 0 ? 256 / Ct : 1;
    const int U = Ct * lanes_r;
    for (int cbase = 0; cbase < C4; cbase += Ct) {
      for (int u = JPB_TID; u < U; u += JPB_NT) {
        const int c4 = cbase + u % Ct, lr = u / Ct;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (c4 < C4) {
#pragma unroll 4
          for (long long r = r0 + lr; r < r1; r += lanes_r) {
            const long long i = r * C + c4 * 4;
            const float4 g = *reinterpret_cast<const float4*>(dy + i);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y) v = *reinterpret_cast<const float4*>(y + i);
            const float4 o = make_float4(act_grad(g.x, v.x, act), act_grad(g.y, v.y, act), act_grad(g.z, v.z, act), act_grad(g.w, v.w, act));
            if (dz) *reinterpret_cast<float4*>(dz + i) = o;
            s0 += o.x; s1 += o.y; s2 += o.z; s3 += o.w;
          }
        }
        part[u] = s0; part[256 + u] = s1; part[512 + u] = s2; part[768 + u] = s3;
      }
      __syncthreads();
      if (dbias && r1 > r0)
        for (int i = JPB_TID; i < Ct * 4; i += JPB_NT) {
          const int cl = i >> 2, k = i & 3;
          if (cbase + cl < C4) {
            float a1 = 0.f;
            for (int lr = 0; lr < lanes_r; ++lr) a1 += part[k * 256 + lr * Ct + cl];
            atomicAdd(&dbias[(cbase + cl) * 4 + k], a1);
          }
        }
      __syncthreads();
    }
  } else if (C >= 32) {
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


// epilogue of a split-K convolution: the partial tiles were summed into `z` without bias / residual / activation (partial
// sums cannot run the epilogue); finish in place: z = act(z + bias[c] + residual)
__global__ void __launch_bounds__(256) bias_act_kernel(float* z, const float* bias, const float* residual, long long n4, int C, int act) {
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += (long long)gridDim.x * JPB_NT) {
    const int c = (int)((i * 4) % C);
    float4 v = *reinterpret_cast<float4*>(z + i * 4);
    if (bias) { v.x += bias[c]; v.y += bias[c + 1]; v.z += bias[c + 2]; v.w += bias[c + 3]; }
    if (residual) { const float4 r = *reinterpret_cast<const float4*>(residual + i * 4); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    float o[4] = {v.x, v.y, v.z, v.w};
    for (int k = 0; k < 4; ++k) {
      float t = o[k];
      if (act == 1) t = t > 0.f ? t : 0.f;
      else if (act == 2) t = t > 0.f ? t : 0.01f * t;
      else if (act == 3) t = 1.f / (1.f + __expf(-t));
      o[k] = t;
    }
    *reinterpret_cast<float4*>(z + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}


// 3xTF32 operand split (precision mode "3xtf32" of the tensor-core convolutions): every fp32 value x becomes
//   hi = x rounded to the nearest TF32 (10 explicit mantissa bits, ties to even), lo = x - hi (exact in fp32).
// rows x C in, rows x 2*Cp out (Cp = C rounded up to 4): hi in columns [0, Cp), lo in [Cp, 2*Cp), padding columns zero.
// The convolution then runs  hi*Whi + hi*Wlo + lo*Whi  as ONE GEMM over a 3x longer K (the split halves are extra K chunks of
// the same gather table), all three products accumulating in the same fp32 TMEM tile.
__global__ void __launch_bounds__(256) tf32_split_kernel(const float* x, float* out, long long rows, int C, int Cp) {
  const long long n = rows * Cp;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += (long long)gridDim.x * JPB_NT) {
    const long long r = i / Cp;
    const int c = (int)(i - r * Cp);
    float hi = 0.f, lo = 0.f;
    if (c < C) {
      const float v = x[r * C + c];
      uint32_t b = __float_as_uint(v);
      if ((b & 0x7f800000u) != 0x7f800000u) {       // finite: round to nearest even at bit 13
        b += 0xfffu + ((b >> 13) & 1u);
        b &= 0xffffe000u;
      }
      hi = __uint_as_float(b);
      lo = v - hi;
    }
    out[r * 2 * Cp + c] = hi;
    out[r * 2 * Cp + Cp + c] = lo;
  }
}


// Space-to-depth form of the stem input (7x7 / stride-2 convolutions, ResnetEncoder.py / resnet.py conv1): x [B,H,W,Cp] (Cp = 4 or
// 8: the 3 / 6 image channels zero-padded) -> x3 [B, H/2, W/2 + 1, 8*Cp]: position (oy, p) holds the 2 x 4 input pixels of rows
// 2*oy + dy (dy < 2) and columns 2*(p - 1) + dx (dx < 4), channel ((dy*4 + dx)*Cp + c), zero outside the image.  The 4-wide column
// windows overlap (stride 2) and the one at p = 0 reaches the valid columns 0 and 1 — hence the extra leading position.  The
// stride-2 7x7 window of output (oy, ox) becomes 4 x 2 stride-1 taps over x3: rows oy-2 .. oy+1, positions ox - 1 and ox + 1,
// which the TMA-patch convolution reads without any gather.
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* x, float* x3, int B, int H, int W, int Cp) {
  const int H2 = H >> 1, W3 = (W >> 1) + 1, v4 = Cp >> 2;       // float4 per (dy, dx) group
  const long long n = (long long)B * H2 * W3 * 8 * v4;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += (long long)gridDim.x * JPB_NT) {
    const int q = (int)(i % v4);
    long long r = i / v4;
    const int g = (int)(r % 8); r /= 8;
    const int p = (int)(r % W3); r /= W3;
    const int oy = (int)(r % H2);
    const int b = (int)(r / H2);
    const int dy = g >> 2, dx = g & 3;
    const int iy = 2 * oy + dy, ix = 2 * (p - 1) + dx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ix >= 0 && ix < W) v = *reinterpret_cast<const float4*>(x + (((long long)b * H + iy) * W + ix) * Cp + q * 4);
    *reinterpret_cast<float4*>(x3 + i * 4) = v;
  }
}

// nearest 2x up-sampling of an NHWC map (layers.py:16-19 upsample = F.interpolate(scale_factor=2, mode="nearest")), materialised
// for the TMA-row weight gradient, whose operand boxes need a dense full-resolution source (the other convolution kernels read
// up-sampled sources through their gather)
__global__ void __launch_bounds__(256) upsample2x_kernel(const float4* x, float4* y, int B, int H, int W, int C4) {
  const long long n = (long long)B * (2 * H) * (2 * W) * C4;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += (long long)gridDim.x * JPB_NT) {
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int ox = (int)(r % (2 * W)); r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    y[i] = x[(((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * C4 + c];
  }
}

// zero-padded channel copy of an NHWC map, C -> Cp (Cp % 4 == 0): the 1-channel disparity that the iconv layers concatenate becomes
// a source of whole 16-byte chunks (forward / data gradient gather: cp.async instead of synchronous scalar loads) or of a whole
// 32-channel block (TMA-row weight gradient)
__global__ void __launch_bounds__(256) pad_channels_kernel(const float* x, float4* y, long long rows, int C, int Cp4) {
  const long long n = rows * Cp4;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += (long long)gridDim.x * JPB_NT) {
    const long long r = i / Cp4;
    const int c = (int)(i - r * Cp4) * 4;
    const float* px = x + r * C;
    y[i] = make_float4(c < C ? px[c] : 0.f, c + 1 < C ? px[c + 1] : 0.f, c + 2 < C ? px[c + 2] : 0.f, c + 3 < C ? px[c + 3] : 0.f);
  }
}

// left-to-right sum of up to 8 equally shaped tensors: the residual chain of a chained-residual-pooling block
// (layers.py:186-199: x = top_i + x after every stage) as ONE pass — (((x + t1) + t2) + t3) + t4 is bit-identical to the four
// binary adds and reads / writes 6 n instead of 12 n floats
struct SumSrc {
  const float4* p[8];
};
__global__ void __launch_bounds__(256) sum_n_kernel(SumSrc src, int n, float4* y, long long n4) {
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += (long long)gridDim.x * JPB_NT) {
    float4 a = src.p[0][i];
    for (int k = 1; k < n; ++k) {
      const float4 b = src.p[k][i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    y[i] = a;
  }
}

}  // namespace

extern "C" int jpb_sum_n(const float* const* xs, int n, float* y, long long count, void* stream) {
  if (!xs || !y || n < 1 || n > 8 || count < 4 || (count & 3) || ((uintptr_t)y & 15)) return JPB_ERR_ARG;
  SumSrc src;
  for (int k = 0; k < 8; ++k) {
    src.p[k] = reinterpret_cast<const float4*>(xs[k < n ? k : 0]);
    if (k < n && (!xs[k] || ((uintptr_t)xs[k] & 15))) return JPB_ERR_ARG;
  }
  long long blocks = (count / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(sum_n_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, n, reinterpret_cast<float4*>(y), count / 4);
  return jpb_status();
}

extern "C" int jpb_pad_channels(const float* x, float* y, long long rows, int C, int Cp, void* stream) {
  if (!x || !y || rows < 1 || C < 1 || Cp < C || (Cp & 3) || ((uintptr_t)y & 15)) return JPB_ERR_ARG;
  const long long n = rows * (Cp / 4);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(pad_channels_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, reinterpret_cast<float4*>(y), rows, C, Cp / 4);
  return jpb_status();
}

extern "C" int jpb_upsample2x(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y || B < 1 || H < 1 || W < 1 || C < 4 || (C & 3) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return JPB_ERR_ARG;
  const long long n = (long long)B * 4 * H * W * (C / 4);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(upsample2x_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(x),
             reinterpret_cast<float4*>(y), B, H, W, C / 4);
  return jpb_status();
}

extern "C" int jpb_bias_act(float* z, const float* bias, const float* residual, long long rows, int C, int act, void* stream) {
  if (!z || rows < 1 || C < 4 || (C & 3)) return JPB_ERR_ARG;
  const long long n4 = rows * C / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(bias_act_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, z, bias, residual, n4, C, act);
  return jpb_status();
}

extern "C" int jpb_act_bwd(const float* dy, const float* y, float* dz, long long rows, int C, int act, float* dbias, void* stream) {
  if (!dy || rows < 1 || C < 1 || (act != 0 && !y) || (!dz && !dbias)) return JPB_ERR_ARG;
  long long blocks = rows / 16 + 1;   // small-extent layers (rows = 480..5120) need many short slabs, not 8 blocks of 64 serial rows
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(act_bwd_kernel, dim3((unsigned)blocks), dim3(256), 4 * 256 * sizeof(float), (cudaStream_t)stream, dy, y, dz, rows, C, act, dbias);
  return jpb_status();
}

extern "C" int jpb_tf32_split(const float* x, float* out, long long rows, int C, void* stream) {
  if (!x || !out || rows < 1 || C < 1) return JPB_ERR_ARG;
  const int Cp = (C + 3) & ~3;
  long long blocks = (rows * Cp + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(tf32_split_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, out, rows, C, Cp);
  return jpb_status();
}

extern "C" int jpb_stem_s2d(const float* x, float* x3, int B, int H, int W, int Cp, void* stream) {
  if (!x || !x3 || B < 1 || H < 2 || W < 4 || (H & 1) || (W & 1) || (Cp != 4 && Cp != 8)) return JPB_ERR_ARG;
  const long long n = (long long)B * (H / 2) * (W / 2 + 1) * 8 * (Cp / 4);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(stem_s2d_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, x3, B, H, W, Cp);
  return jpb_status();
}
