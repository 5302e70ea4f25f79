// Training-mode BatchNorm2d on NHWC fp32, fused with the residual add and ReLU that follow it in every ResNet block
// (resnet.py:28-45: out = relu(bn2(conv2(.)) + residual)) and layout-decoder stage (layout_model.py:146-158).
// Per-GPU batch statistics (no SyncBN, as the reference), running statistics updated in place with the unbiased variance.
//   forward : stats   acc[c] += (sum x, sum x^2)                       one pass over x
//             apply   y = relu?((x-mean)*rstd*gamma + beta (+ res))     one pass; block 0 also updates running stats
//   backward: reduce  acc[c] += (sum g, sum g*xhat),  g = dy * [y > 0]  one pass over (dy, x, y)
//             apply   dx = gamma*rstd*(g - s1/n - xhat*s2/n), dres = g   one pass
// The library path needs separate kernels for BN, the add and the ReLU in each direction.
#include <stdlib.h>
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

constexpr int BN_MAX_BLOCKS = 148 * 2;      // cap of the column-sum grids.  Measured on B200 (JPB_BN_BLOCKS): 592 blocks run the big layers 10-18 % faster in
                                            // isolation (268 MB layout stem backward 388 -> 327 us) but the multi-stream step 1 % SLOWER (18.9 -> 19.1 ms): they
                                            // take SMs from the convolutions running beside them
inline int bn_max_blocks() {
  static int v = 0;
  if (!v) { const char* e = getenv("JPB_BN_BLOCKS"); v = e ? atoi(e) : BN_MAX_BLOCKS; if (v < 1) v = BN_MAX_BLOCKS; }
  return v;
}

constexpr int BN_MAX_C = 2048;
// Workspace layout (jpb_bn_workspace_doubles): one int ticket counter in the first 16 bytes | final[2 * BN_MAX_C] column sums
// of the last launch (read by the backward apply pass) | running[2 * BN_MAX_C] accumulators.  Ticket and accumulators are zero
// at allocation and every launch leaves them at zero again, so launches of any C can share one workspace on a stream.
__device__ __forceinline__ int* bn_ticket(double* ws) { return reinterpret_cast<int*>(ws); }
__device__ __forceinline__ double* bn_sums(double* ws) { return ws + 2; }
__device__ __forceinline__ double* bn_running(double* ws) { return ws + 2 + 2 * BN_MAX_C; }

// y = (x - mean) * rstd * gamma + beta with a FIXED rounding sequence (sub, mul, fma): the backward pass re-derives the ReLU mask of
// a BatchNorm without residual from x instead of reading y back (two of its seven passes), which only works if forward and backward
// round identically whatever the compiler would contract.
__device__ __forceinline__ float bn_affine(float x, float mean, float rstd, float g, float b) {
#ifdef JPB_HOST_EMU
  // host emulation: the unfused sequence of the CPU oracle (sub, mul, mul, add), so that the hard arg-max ties downstream
  // (CCT attention) resolve as in the oracle; what matters for the mask is only that forward and backward share THIS function
  volatile float t = (x - mean) * rstd;
  volatile float u = t * g;
  return u + b;
#else
  return __fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), g, b);
#endif
}

struct BnTail {   // what the last block to finish does after folding the partials (one launch instead of three)
  long long rows;
  float eps, momentum;
  float* stat;                 // MODE 0: (mean, rstd) out
  float* running_mean;         // MODE 0: updated in place (may be NULL)
  float* running_var;
  long long* num_batches_tracked;  // MODE 0: += nbt_inc (may be NULL)
  int nbt_inc;
  float* dgamma;               // MODE 1: (+)= s2
  float* dbeta;                // MODE 1: (+)= s1
  int accumulate;              // MODE 1: 1 = add into dgamma/dbeta (they are views of the flat gradient buffer), 0 = overwrite
};

// column sums of two per-element quantities over a [rows][C] matrix.  Each block reduces a slab of rows and writes its
// partial sums (no atomics: hundreds of blocks hammering 2C addresses serialise in L2); the LAST block to finish (ticket
// counter) folds the partials in double precision into ws[0 .. 2C) and finalises: batch statistics + running statistics
// (MODE 0) or the affine-parameter gradients (MODE 1).
// MODE 0: (x, x*x)    MODE 1: (g, g*xhat) with g = dy*[y>0 or no relu], xhat = (x-mean)*rstd
template <int MODE>
__global__ void __launch_bounds__(256) bn_colsum_kernel(const float* x, const float* dy, const float* y, const float* stat, long long rows,
                                                       int C, int relu, double* ws, BnTail tail, const float* gamma, const float* beta) {
  JPB_DYN_SMEM(float, part);   // [8][256]
  __shared__ int s_last;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  long long r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  const int C4 = C >> 2;
  const int Ct = C4 < 256 ? C4 : 256;               // float4 channel groups covered per pass
  const int lanes_r = 256 / Ct > 0 ? 256 / Ct : 1;  // row lanes
  const int U = Ct * lanes_r;
  double* run = bn_running(ws);
  for (int cbase = 0; cbase < C4; cbase += Ct) {
    for (int u = JPB_TID; u < U; u += JPB_NT) {
      const int c4 = cbase + u % Ct, lr = u / Ct;
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      if (c4 < C4) {
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
        float gam[4] = {1.f, 1.f, 1.f, 1.f}, bet[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == 1) {
          for (int k = 0; k < 4; ++k) { mean[k] = stat[c4 * 4 + k]; rstd[k] = stat[C + c4 * 4 + k]; }
          if (relu && !y)
            for (int k = 0; k < 4; ++k) { gam[k] = gamma[c4 * 4 + k]; bet[k] = beta[c4 * 4 + k]; }
        }
#pragma unroll 8
        for (long long r = r0 + lr; r < r1; r += lanes_r) {
          const long long i = r * C + c4 * 4;
          if (MODE == 0) {
            const float4 v = *reinterpret_cast<const float4*>(x + i);
            s1[0] += v.x; s2[0] += v.x * v.x; s1[1] += v.y; s2[1] += v.y * v.y;
            s1[2] += v.z; s2[2] += v.z * v.z; s1[3] += v.w; s2[3] += v.w * v.w;
          } else {
            const float4 gv = *reinterpret_cast<const float4*>(dy + i);
            const float4 xv = *reinterpret_cast<const float4*>(x + i);
            float g[4] = {gv.x, gv.y, gv.z, gv.w};
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
            if (relu) {
              if (y) {
                const float4 yv = *reinterpret_cast<const float4*>(y + i);
                if (!(yv.x > 0.f)) g[0] = 0.f;
                if (!(yv.y > 0.f)) g[1] = 0.f;
                if (!(yv.z > 0.f)) g[2] = 0.f;
                if (!(yv.w > 0.f)) g[3] = 0.f;
              } else {            // no residual: the forward's y > 0 is bn_affine(x) > 0, bit for bit
                for (int k = 0; k < 4; ++k)
                  if (!(bn_affine(xs[k], mean[k], rstd[k], gam[k], bet[k]) > 0.f)) g[k] = 0.f;
              }
            }
            for (int k = 0; k < 4; ++k) { s1[k] += g[k]; s2[k] += g[k] * ((xs[k] - mean[k]) * rstd[k]); }
          }
        }
      }
      for (int k = 0; k < 4; ++k) { part[k * 256 + u] = s1[k]; part[(4 + k) * 256 + u] = s2[k]; }
    }
    __syncthreads();
    for (int i = JPB_TID; i < Ct * 4; i += JPB_NT) {
      const int cl = i >> 2, k = i & 3;
      const int c = (cbase + cl) * 4 + k;
      if (c < C) {
        float a1 = 0.f, a2 = 0.f;
        for (int lr = 0; lr < lanes_r; ++lr) { a1 += part[k * 256 + lr * Ct + cl]; a2 += part[(4 + k) * 256 + lr * Ct + cl]; }
        // <= 296 blocks per address: the double-precision reductions in L2 cost less than a pass over per-block partials
        if (r1 > r0) { atomicAdd(&run[c], (double)a1); atomicAdd(&run[C + c], (double)a2); }
      }
    }
    __syncthreads();
  }
  // ---- last block to finish: publish the sums, re-zero the accumulators, finalise
  __threadfence();
  __syncthreads();
  if (JPB_TID == 0) s_last = (atomicAdd(bn_ticket(ws), 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    double* sums = bn_sums(ws);
    for (int j = JPB_TID; j < 2 * C; j += JPB_NT) {
      sums[j] = atomicAdd(&run[j], 0.0);   // read through L2 (the other blocks' reductions never touched this SM's L1)
      run[j] = 0.0;
    }
    __syncthreads();
    for (int c = JPB_TID; c < C; c += JPB_NT) {
      if (MODE == 0) {
        const double mean = sums[c] / (double)tail.rows;
        double var = sums[C + c] / (double)tail.rows - mean * mean;
        if (var < 0.0) var = 0.0;
        tail.stat[c] = (float)mean;
        tail.stat[C + c] = (float)(1.0 / sqrt(var + (double)tail.eps));
        if (tail.running_mean) {
          const double unb = tail.rows > 1 ? var * (double)tail.rows / (double)(tail.rows - 1) : var;
          tail.running_mean[c] = (float)((1.0 - tail.momentum) * tail.running_mean[c] + tail.momentum * mean);
          tail.running_var[c] = (float)((1.0 - tail.momentum) * tail.running_var[c] + tail.momentum * unb);
        }
      } else {
        if (tail.accumulate) { tail.dbeta[c] += (float)sums[c]; tail.dgamma[c] += (float)sums[C + c]; }
        else { tail.dbeta[c] = (float)sums[c]; tail.dgamma[c] = (float)sums[C + c]; }
      }
    }
    if (JPB_TID == 0) {
      if (MODE == 0 && tail.num_batches_tracked) tail.num_batches_tracked[0] += tail.nbt_inc;
      *bn_ticket(ws) = 0;
    }
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* x, const float* res, const float* stat, const float* gamma, const float* beta,
                                                      float* y, long long n4, int C, int relu) {
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += (long long)gridDim.x * JPB_NT) {
    const int c = (int)((i * 4) % C);
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    float o[4] = {v.x, v.y, v.z, v.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (res) { const float4 q = *reinterpret_cast<const float4*>(res + i * 4); r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
    for (int k = 0; k < 4; ++k) {
      float t = bn_affine(o[k], stat[c + k], stat[C + c + k], gamma[c + k], beta[c + k]) + r[k];
      if (relu && !(t > 0.f)) t = 0.f;
      o[k] = t;
    }
    *reinterpret_cast<float4*>(y + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Forward apply when the statistics were accumulated by the producing convolution's epilogue (JpbConvArgs.stats): every block
// derives (mean, rstd) of all channels from the double accumulators, normalises its slab, and the last block to finish
// (ticket) writes stat / running statistics / num_batches_tracked and re-zeroes the accumulators for the next layer.
__global__ void __launch_bounds__(256) bn_apply_from_sums_kernel(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                                                long long n4, int C, int relu, double* ws, BnTail tail) {
  JPB_DYN_SMEM(float, sstat);   // [2][C]
  __shared__ int s_last;
  const double* run = bn_running(ws);
  for (int c = JPB_TID; c < C; c += JPB_NT) {
    const double mean = run[c] / (double)tail.rows;
    double var = run[C + c] / (double)tail.rows - mean * mean;
    if (var < 0.0) var = 0.0;
    sstat[c] = (float)mean;
    sstat[C + c] = (float)(1.0 / sqrt(var + (double)tail.eps));
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += (long long)gridDim.x * JPB_NT) {
    const int c = (int)((i * 4) % C);
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    float o[4] = {v.x, v.y, v.z, v.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (res) { const float4 q = *reinterpret_cast<const float4*>(res + i * 4); r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
    for (int k = 0; k < 4; ++k) {
      float t = bn_affine(o[k], sstat[c + k], sstat[C + c + k], gamma[c + k], beta[c + k]) + r[k];
      if (relu && !(t > 0.f)) t = 0.f;
      o[k] = t;
    }
    *reinterpret_cast<float4*>(y + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __threadfence();
  __syncthreads();
  if (JPB_TID == 0) s_last = (atomicAdd(bn_ticket(ws), 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  double* runw = bn_running(ws);
  for (int c = JPB_TID; c < C; c += JPB_NT) {
    const double mean = runw[c] / (double)tail.rows;
    double var = runw[C + c] / (double)tail.rows - mean * mean;
    if (var < 0.0) var = 0.0;
    tail.stat[c] = sstat[c];
    tail.stat[C + c] = sstat[C + c];
    if (tail.running_mean) {
      const double unb = tail.rows > 1 ? var * (double)tail.rows / (double)(tail.rows - 1) : var;
      tail.running_mean[c] = (float)((1.0 - tail.momentum) * tail.running_mean[c] + tail.momentum * mean);
      tail.running_var[c] = (float)((1.0 - tail.momentum) * tail.running_var[c] + tail.momentum * unb);
    }
  }
  __syncthreads();
  for (int j = JPB_TID; j < 2 * C; j += JPB_NT) runw[j] = 0.0;
  if (JPB_TID == 0) {
    if (tail.num_batches_tracked) tail.num_batches_tracked[0] += tail.nbt_inc;
    *bn_ticket(ws) = 0;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* x, const float* dy, const float* y, const float* stat, const float* gamma,
                                                          const double* acc, float* dx, float* dres, long long rows, long long n4, int C, int relu,
                                                          const float* beta) {
  const double inv_n = 1.0 / (double)rows;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += (long long)gridDim.x * JPB_NT) {
    const int c = (int)((i * 4) % C);
    const float4 xv = *reinterpret_cast<const float4*>(x + i * 4);
    const float4 gv = *reinterpret_cast<const float4*>(dy + i * 4);
    float xs[4] = {xv.x, xv.y, xv.z, xv.w}, g[4] = {gv.x, gv.y, gv.z, gv.w}, o[4];
    if (relu) {
      if (y) {
        const float4 yv = *reinterpret_cast<const float4*>(y + i * 4);
        if (!(yv.x > 0.f)) g[0] = 0.f;
        if (!(yv.y > 0.f)) g[1] = 0.f;
        if (!(yv.z > 0.f)) g[2] = 0.f;
        if (!(yv.w > 0.f)) g[3] = 0.f;
      } else {
        for (int k = 0; k < 4; ++k)
          if (!(bn_affine(xs[k], stat[c + k], stat[C + c + k], gamma[c + k], beta[c + k]) > 0.f)) g[k] = 0.f;
      }
    }
    for (int k = 0; k < 4; ++k) {
      const float mean = stat[c + k], rstd = stat[C + c + k];
      const float xh = (xs[k] - mean) * rstd;
      const float m1 = (float)(acc[c + k] * inv_n), m2 = (float)(acc[C + c + k] * inv_n);
      o[k] = gamma[c + k] * rstd * (g[k] - m1 - xh * m2);
    }
    *reinterpret_cast<float4*>(dx + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
    if (dres) *reinterpret_cast<float4*>(dres + i * 4) = make_float4(g[0], g[1], g[2], g[3]);
  }
}

// blocks of the column-sum kernels: every thread should own >= 4 rows, at most BN_MAX_BLOCKS blocks
inline unsigned bn_colsum_grid(long long rows, int C) {
  const int C4 = C >> 2, Ct = C4 < 256 ? C4 : 256;
  const int lanes_r = 256 / Ct > 0 ? 256 / Ct : 1;
  long long g = (rows + 4LL * lanes_r - 1) / (4LL * lanes_r);
  if (g > bn_max_blocks()) g = bn_max_blocks();
  return (unsigned)(g < 1 ? 1 : g);
}

inline unsigned bn_grid(long long work, int per_block, int cap) {
  long long g = (work + per_block - 1) / per_block;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" double* jpb_bn_stats_accumulator(double* ws) { return ws ? ws + 2 + 2 * BN_MAX_C : nullptr; }

extern "C" long long jpb_bn_workspace_doubles(int C) { (void)C; return 2 + (long long)4 * BN_MAX_C; }

extern "C" int jpb_bn_train_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                long long* num_batches_tracked, int nbt_inc, float momentum, float eps, int relu, float* y, float* stat,
                                double* ws, long long rows, int C, int stats_ready, void* stream) {
  if (!x || !gamma || !beta || !y || !stat || !ws || rows < 1 || C < 4 || (C & 3) || C > BN_MAX_C) return JPB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = bn_colsum_grid(rows, C);
  BnTail t = {};
  t.rows = rows; t.eps = eps; t.momentum = momentum; t.stat = stat; t.running_mean = running_mean; t.running_var = running_var;
  t.num_batches_tracked = num_batches_tracked; t.nbt_inc = nbt_inc;
  if (stats_ready) {   // the producing convolution already accumulated (sum, sum of squares): one launch
    const long long n4s = rows * C / 4;
    JPB_LAUNCH(bn_apply_from_sums_kernel, dim3(bn_grid(n4s, 256 * 4, 148 * 8)), dim3(256), 2 * C * sizeof(float), st, x, res, gamma, beta, y, n4s, C,
               relu, ws, t);
    return jpb_status();
  }
  JPB_LAUNCH(bn_colsum_kernel<0>, dim3(nb), dim3(256), 8 * 256 * sizeof(float), st, x, nullptr, nullptr, nullptr, rows, C, 0, ws, t, nullptr, nullptr);
  const long long n4 = rows * C / 4;
  JPB_LAUNCH(bn_apply_kernel, dim3(bn_grid(n4, 256 * 4, 148 * 8)), dim3(256), 0, st, x, res, stat, gamma, beta, y, n4, C, relu);
  return jpb_status();
}

extern "C" int jpb_bn_eval_fwd(const float* x, const float* res, const float* gamma, const float* beta, const float* stat, int relu, float* y,
                               long long rows, int C, void* stream) {
  if (!x || !gamma || !beta || !y || !stat || rows < 1 || (C & 3)) return JPB_ERR_ARG;
  const long long n4 = rows * C / 4;
  JPB_LAUNCH(bn_apply_kernel, dim3(bn_grid(n4, 256 * 4, 148 * 8)), dim3(256), 0, (cudaStream_t)stream, x, res, stat, gamma, beta, y, n4, C, relu);
  return jpb_status();
}

extern "C" int jpb_bn_train_bwd(const float* x, const float* dy, const float* y, const float* stat, const float* gamma, int relu, float* dx,
                                float* dres, float* dgamma, float* dbeta, int accumulate, double* ws, long long rows, int C, void* stream,
                                const float* beta) {
  // relu without y: the mask is re-derived from x (BatchNorm + ReLU WITHOUT residual only; needs beta)
  if (!x || !dy || !stat || !gamma || !dx || !dgamma || !dbeta || !ws || (relu && !y && (!beta || dres)) || (C & 3) || C > BN_MAX_C) return JPB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = bn_colsum_grid(rows, C);
  BnTail t = {};
  t.rows = rows; t.dgamma = dgamma; t.dbeta = dbeta; t.accumulate = accumulate;
  JPB_LAUNCH(bn_colsum_kernel<1>, dim3(nb), dim3(256), 8 * 256 * sizeof(float), st, x, dy, y, stat, rows, C, relu, ws, t, gamma, beta);
  const long long n4 = rows * C / 4;
  JPB_LAUNCH(bn_bwd_apply_kernel, dim3(bn_grid(n4, 256 * 4, 148 * 8)), dim3(256), 0, st, x, dy, y, stat, gamma, ws + 2, dx, dres, rows, n4, C, relu, beta);
  return jpb_status();
}
