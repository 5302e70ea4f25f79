// Small fused operators around the network trunks (CUDA cores; all of them also build under JPB_HOST_EMU):
//   jpb_image_prep    : ResnetEncoder.forward's (x - 0.45) / 0.225 (ResnetEncoder.py:99, depth_encoder.py:37, pose_encoder.py:84)
//                       fused with the bilinear resize in front of the pose / layout trunks (net.py:633, F.interpolate
//                       align_corners=False), the channel concatenation of a frame pair (net.py:636-638) and the
//                       NCHW -> NHWC(+zero channel padding to 4/8) re-layout the tensor-core gather wants.
//   jpb_dropout       : nn.Dropout(0.5) on the two deepest encoder features (depth_decoder.py:47-48), counter-based mask.
//   jpb_pose_head_*   : PoseDecoder's spatial mean x 0.01 (pose_decoder.py:22-26) + transformation_from_parameters
//                       (net.py:704-756: Rodrigues with axis = v / (|v| + 1e-7), translation, optional inversion), and its
//                       hand-derived backward.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

// ATen upsample_bilinear2d(align_corners=False) source index + weight for one axis
__device__ __forceinline__ void prep_axis(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(256) image_prep_kernel(const float* im0, const float* im1, float* out, int B, int Hs, int Ws, int Ho, int Wo,
                                                        int Cpad) {
  const long long total = (long long)B * Ho * Wo;
  const bool resize = Ho != Hs || Wo != Ws;
  const float sy = (float)Hs / (float)Ho, sx = (float)Ws / (float)Wo;
  const size_t plane = (size_t)Hs * Ws;
  for (long long t = (long long)blockIdx.x * JPB_NT + JPB_TID; t < total; t += (long long)gridDim.x * JPB_NT) {
    const int ox = (int)(t % Wo);
    const long long r = t / Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int y0 = oy, y1 = oy, x0 = ox, x1 = ox;
    float ly = 0.f, lx = 0.f;
    if (resize) { prep_axis(oy, sy, Hs, y0, y1, ly); prep_axis(ox, sx, Ws, x0, x1, lx); }
    for (int im = 0; im < 2; ++im) {
      const float* p = im == 0 ? im0 : im1;
      if (!p) continue;
      p += (size_t)b * 3 * plane;
      for (int c = 0; c < 3; ++c) {
        const float* q = p + c * plane;
        float val;
        if (resize) {
          const float a = q[(size_t)y0 * Ws + x0], bb = q[(size_t)y0 * Ws + x1], cc = q[(size_t)y1 * Ws + x0], d = q[(size_t)y1 * Ws + x1];
          val = (1.f - ly) * ((1.f - lx) * a + lx * bb) + ly * ((1.f - lx) * cc + lx * d);
        } else {
          val = q[(size_t)oy * Ws + ox];
        }
        v[im * 3 + c] = (val - 0.45f) / 0.225f;
      }
    }
    float* o = out + (size_t)t * Cpad;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    if (Cpad == 8) *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// keep mask of element i: Philox word > p (4 elements per counter)
__device__ __forceinline__ float drop_keep(uint64_t seed, uint64_t stream, long long i, float p) {
  uint32_t r[4];
  jpb_philox4(seed, stream, (uint64_t)(i >> 2), r);
  return jpb_u01(r[i & 3]) >= p ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) dropout_kernel(const float* x, const float* mask, float* y, long long n, float p, uint64_t seed,
                                                     uint64_t stream, const long long* step) {
  const float scale = 1.f / (1.f - p);
  const uint64_t st = stream + (step ? 4096ull * (uint64_t)step[0] : 0ull);
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += (long long)gridDim.x * JPB_NT) {
    const float k = mask ? mask[i] : drop_keep(seed, st, i, p);
    y[i] = x[i] * k * scale;
  }
}

struct PoseFwd {   // everything the backward needs, recomputed from the 6 means
  double aa[3], t[3], ang, ax[3], ca, sa, R[9];
};

__device__ __forceinline__ void pose_forward(const double v[6], PoseFwd& f) {
  for (int i = 0; i < 3; ++i) { f.aa[i] = v[i]; f.t[i] = v[3 + i]; }
  f.ang = sqrt(f.aa[0] * f.aa[0] + f.aa[1] * f.aa[1] + f.aa[2] * f.aa[2]);
  for (int i = 0; i < 3; ++i) f.ax[i] = f.aa[i] / (f.ang + 1e-7);
  f.ca = cos(f.ang); f.sa = sin(f.ang);
  const double C = 1.0 - f.ca, x = f.ax[0], y = f.ax[1], z = f.ax[2];
  f.R[0] = x * x * C + f.ca;      f.R[1] = x * y * C - z * f.sa;  f.R[2] = z * x * C + y * f.sa;
  f.R[3] = x * y * C + z * f.sa;  f.R[4] = y * y * C + f.ca;      f.R[5] = y * z * C - x * f.sa;
  f.R[6] = z * x * C - y * f.sa;  f.R[7] = y * z * C + x * f.sa;  f.R[8] = z * z * C + f.ca;
}

// x: [B][hw][C] (C >= 6, first six channels used); T: [B][4][4]
__global__ void pose_head_fwd_kernel(const float* x, float* T, float* mean6, int B, int hw, int C, int invert) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  double v[6];
  for (int c = 0; c < 6; ++c) {
    double s = 0.0;
    for (int i = JPB_TID; i < hw; i += JPB_NT) s += (double)x[((size_t)b * hw + i) * C + c];
    s = jpb_block_sum<double>(s, red);
    v[c] = s;
  }
  if (JPB_TID == 0) {
    // the reference takes mean(3) then mean(2) in fp32 and scales by 0.01; the means here are double sums rounded once
    for (int c = 0; c < 6; ++c) { v[c] = 0.01 * (double)(float)(v[c] / (double)hw); if (mean6) mean6[b * 6 + c] = (float)v[c]; }
    PoseFwd f;
    pose_forward(v, f);
    float* M = T + b * 16;
    for (int i = 0; i < 16; ++i) M[i] = 0.f;
    M[15] = 1.f;
    if (!invert) {
      for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M[i * 4 + j] = (float)f.R[i * 3 + j]; M[i * 4 + 3] = (float)f.t[i]; }
    } else {
      for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int j = 0; j < 3; ++j) { M[i * 4 + j] = (float)f.R[j * 3 + i]; s -= f.R[j * 3 + i] * f.t[j]; }
        M[i * 4 + 3] = (float)s;
      }
    }
  }
}

// gT: [B][4][4]; mean6: the six scaled means of the forward; gx: [B][hw][C] (every pixel of a sample gets the same gradient)
__global__ void pose_head_bwd_kernel(const float* gT, const float* mean6, float* gx, int B, int hw, int C, int invert) {
  __shared__ float g6[8];
  const int b = blockIdx.x;
  if (JPB_TID == 0) {
    double v[6];
    for (int c = 0; c < 6; ++c) v[c] = (double)mean6[b * 6 + c];
    PoseFwd f;
    pose_forward(v, f);
    const float* G = gT + b * 16;
    double dR[9], dt[3];
    if (!invert) {
      for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) dR[i * 3 + j] = (double)G[i * 4 + j]; dt[i] = (double)G[i * 4 + 3]; }
    } else {
      // M33 = R^T, Mt_i = -sum_j R[j][i] t_j
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int i = 0; i < 3; ++i) {
          dR[j * 3 + i] = (double)G[i * 4 + j] - f.t[j] * (double)G[i * 4 + 3];
          s -= f.R[j * 3 + i] * (double)G[i * 4 + 3];
        }
        dt[j] = s;
      }
    }
    const double C1 = 1.0 - f.ca, x = f.ax[0], y = f.ax[1], z = f.ax[2], sa = f.sa;
    // R entries as functions of (x, y, z, ca, sa, C1)
    double dx = 0, dy = 0, dz = 0, dca = 0, dsa = 0, dC = 0;
    dx += dR[0] * 2 * x * C1; dC += dR[0] * x * x; dca += dR[0];
    dx += dR[1] * y * C1; dy += dR[1] * x * C1; dC += dR[1] * x * y; dz -= dR[1] * sa; dsa -= dR[1] * z;
    dz += dR[2] * x * C1; dx += dR[2] * z * C1; dC += dR[2] * z * x; dy += dR[2] * sa; dsa += dR[2] * y;
    dx += dR[3] * y * C1; dy += dR[3] * x * C1; dC += dR[3] * x * y; dz += dR[3] * sa; dsa += dR[3] * z;
    dy += dR[4] * 2 * y * C1; dC += dR[4] * y * y; dca += dR[4];
    dy += dR[5] * z * C1; dz += dR[5] * y * C1; dC += dR[5] * y * z; dx -= dR[5] * sa; dsa -= dR[5] * x;
    dz += dR[6] * x * C1; dx += dR[6] * z * C1; dC += dR[6] * z * x; dy -= dR[6] * sa; dsa -= dR[6] * y;
    dy += dR[7] * z * C1; dz += dR[7] * y * C1; dC += dR[7] * y * z; dx += dR[7] * sa; dsa += dR[7] * x;
    dz += dR[8] * 2 * z * C1; dC += dR[8] * z * z; dca += dR[8];
    dca -= dC;                                       // C1 = 1 - ca
    double dang = -sa * dca + f.ca * dsa;            // ca = cos(ang), sa = sin(ang)
    // ax_j = aa_j / (ang + eps)
    const double den = f.ang + 1e-7;
    const double dax[3] = {dx, dy, dz};
    double daa[3] = {0, 0, 0};
    for (int j = 0; j < 3; ++j) { daa[j] += dax[j] / den; dang -= dax[j] * f.aa[j] / (den * den); }
    if (f.ang > 0.0) for (int i = 0; i < 3; ++i) daa[i] += dang * f.aa[i] / f.ang;   // torch.norm backward (0 at the origin)
    const double k = 0.01 / (double)hw;
    for (int c = 0; c < 3; ++c) { g6[c] = (float)(daa[c] * k); g6[3 + c] = (float)(dt[c] * k); }
  }
  __syncthreads();
  for (long long i = JPB_TID; i < (long long)hw * C; i += JPB_NT) {
    const int c = (int)(i % C);
    gx[(size_t)b * hw * C + i] = c < 6 ? g6[c] : 0.f;
  }
}

}  // namespace

extern "C" int jpb_image_prep(const float* im0, const float* im1, float* out, int B, int Hs, int Ws, int Ho, int Wo, int Cpad, void* stream) {
  if (!im0 || !out || B < 1 || (Cpad != 4 && Cpad != 8) || (im1 && Cpad != 8)) return JPB_ERR_ARG;
  long long blocks = ((long long)B * Ho * Wo + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(image_prep_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, im0, im1, out, B, Hs, Ws, Ho, Wo, Cpad);
  return jpb_status();
}

extern "C" int jpb_dropout(const float* x, const float* mask, float* y, long long n, float p, uint64_t seed, uint64_t stream_id,
                           const long long* step, void* stream) {
  if (!x || !y || n < 1 || !(p >= 0.f && p < 1.f)) return JPB_ERR_ARG;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  JPB_LAUNCH(dropout_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, mask, y, n, p, seed, stream_id, step);
  return jpb_status();
}

extern "C" int jpb_pose_head_fwd(const float* x, float* T, float* mean6, int B, int hw, int C, int invert, void* stream) {
  if (!x || !T || !mean6 || B < 1 || hw < 1 || C < 6) return JPB_ERR_ARG;
  JPB_LAUNCH(pose_head_fwd_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, x, T, mean6, B, hw, C, invert);
  return jpb_status();
}

extern "C" int jpb_pose_head_bwd(const float* gT, const float* mean6, float* gx, int B, int hw, int C, int invert, void* stream) {
  if (!gT || !mean6 || !gx || B < 1 || hw < 1 || C < 6) return JPB_ERR_ARG;
  JPB_LAUNCH(pose_head_bwd_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, gT, mean6, gx, B, hw, C, invert);
  return jpb_status();
}

// =====================================================================================================================
// Cross-view transformer core and cycled view projection (CrossViewTransformer.py:45-92, CycledViewProjection.py:11-67).
// The 1x1 / 3x3 convolutions around them stay tensor-core launches; everything between them — the n x n energies, the hard
// max / arg-max over front positions, the gather of the projected values, the S-weighted residual and the broadcast
// (h x w)·(h x w) matrix product with the depth values — is one launch per stage instead of ~20 library kernels per head.
// All tensors are NHWC: x[b][j][c] with j = row * w + col the position ("token") index, n = h * w <= JPB_CCT_MAX_N.
namespace {

constexpr int JPB_CCT_MAX_N = 256;

constexpr int CCT_JT = 8;   // positions per block: grid = (B, ceil(n / CCT_JT)[, 2]) so the 4-sample batch still fills tens of SMs

// S[j] = max_i <k_i, q_j>, arg[j] = first maximising i; T[j][:] = v[arg[j]][:]; blockIdx.z = 1: the same max for the depth
// pair (attn, argd)
__global__ void __launch_bounds__(128) cct_select_fwd_kernel(const float* q, const float* k, const float* v, const float* qd, const float* kd,
                                                            float* T, float* S, int* arg, float* attn, int* argd, int n, int Cq, int C) {
  __shared__ int s_arg[CCT_JT];
  const int b = blockIdx.x, j0 = blockIdx.y * CCT_JT, depth = blockIdx.z;
  const float* Q = (depth ? qd : q) + (size_t)b * n * Cq;
  const float* K = (depth ? kd : k) + (size_t)b * n * Cq;
  for (int t = JPB_TID; t < CCT_JT; t += JPB_NT) {
    const int j = j0 + t;
    if (j >= n) continue;
    float best = -INFINITY;
    int bi = 0;
    for (int i = 0; i < n; ++i) {
      float e = 0.f;
      for (int c = 0; c < Cq; ++c) e += K[(size_t)i * Cq + c] * Q[(size_t)j * Cq + c];
      if (e > best || i == 0) { best = e; bi = i; }
    }
    if (depth) { attn[b * n + j] = best; argd[b * n + j] = bi; }
    else { S[b * n + j] = best; arg[b * n + j] = bi; s_arg[t] = bi; }
  }
  if (depth) return;
  __syncthreads();
  for (int t = JPB_TID; t < CCT_JT * C; t += JPB_NT) {
    const int jj = t / C, c = t - jj * C;
    if (j0 + jj < n) T[((size_t)b * n + j0 + jj) * C + c] = v[((size_t)b * n + s_arg[jj]) * C + c];
  }
}

// gv[i][c] = sum_{j: arg[j]=i} gT[j][c];  gq[j][:] = gS[j] k[arg[j]][:];  gk[i][:] = sum_{j: arg[j]=i} gS[j] q[j][:]  (+ depth pair)
__global__ void __launch_bounds__(128) cct_select_bwd_kernel(const float* q, const float* k, const float* qd, const float* kd, const int* arg,
                                                            const int* argd, const float* gT, const float* gS, const float* gattn, float* gq,
                                                            float* gk, float* gv, float* gqd, float* gkd, int n, int Cq, int C) {
  const int b = blockIdx.x, i0 = blockIdx.y * CCT_JT;
  const int* ab = arg + b * n;
  const int* adb = argd + b * n;
  for (int t = JPB_TID; t < CCT_JT * C; t += JPB_NT) {
    const int i = i0 + t / C, c = t % C;
    if (i >= n) continue;
    float s = 0.f;
    for (int j = 0; j < n; ++j)
      if (ab[j] == i) s += gT[((size_t)b * n + j) * C + c];
    gv[((size_t)b * n + i) * C + c] = s;
  }
  for (int t = JPB_TID; t < 2 * CCT_JT * Cq; t += JPB_NT) {
    const int depth = t / (CCT_JT * Cq), r = t - depth * CCT_JT * Cq;
    const int j = i0 + r / Cq, c = r % Cq;
    if (j >= n) continue;
    const float* Q = (depth ? qd : q) + (size_t)b * n * Cq;
    const float* K = (depth ? kd : k) + (size_t)b * n * Cq;
    const float* g = (depth ? gattn : gS) + b * n;
    const int* a = depth ? adb : ab;
    float* GQ = (depth ? gqd : gq) + (size_t)b * n * Cq;
    float* GK = (depth ? gkd : gk) + (size_t)b * n * Cq;
    GQ[(size_t)j * Cq + c] = g[j] * K[(size_t)a[j] * Cq + c];
    float s = 0.f;                       // here j plays the role of the front position i
    for (int jj = 0; jj < n; ++jj)
      if (a[jj] == j) s += g[jj] * Q[(size_t)jj * Cq + c];
    GK[(size_t)j * Cq + c] = s;
  }
}

// out[(r,s)][c] = front + fused * S[(r,s)] + sum_t attn[(r,t)] * vd[(t,s)][c]
__global__ void __launch_bounds__(128) cct_combine_fwd_kernel(const float* front, const float* fused, const float* S, const float* attn,
                                                             const float* vd, float* out, int h, int w, int C) {
  const int b = blockIdx.x, n = h * w, j0 = blockIdx.y * CCT_JT;
  for (int t = JPB_TID; t < CCT_JT * C; t += JPB_NT) {
    const int j = j0 + t / C, c = t % C;
    if (j >= n) continue;
    const int r = j / w, s = j - r * w;
    const size_t o = ((size_t)b * n + j) * C + c;
    float m = 0.f;
    for (int tt = 0; tt < w; ++tt) m += attn[b * n + r * w + tt] * vd[((size_t)b * n + tt * w + s) * C + c];
    out[o] = front[o] + fused[o] * S[b * n + j] + m;
  }
}

// backward of the combine stage: gfused = g*S, gS = <g, fused>_c, gattn[(r,t)] = sum_{s,c} g[(r,s)][c] vd[(t,s)][c],
// gvd[(t,s)][c] = sum_r attn[(r,t)] g[(r,s)][c]     (the gradient w.r.t. front is g itself)
__global__ void __launch_bounds__(128) cct_combine_bwd_kernel(const float* g, const float* fused, const float* S, const float* attn, const float* vd,
                                                             float* gfused, float* gS, float* gattn, float* gvd, int h, int w, int C) {
  const int b = blockIdx.x, n = h * w, j0 = blockIdx.y * CCT_JT;
  for (int t = JPB_TID; t < CCT_JT * C; t += JPB_NT) {
    const int j = j0 + t / C, c = t % C;
    if (j >= n) continue;
    const int tt = j / w, s = j - tt * w;          // j = (t, s) as an index of vd
    const size_t o = ((size_t)b * n + j) * C + c;
    gfused[o] = g[o] * S[b * n + j];
    float m = 0.f;
    for (int r = 0; r < h; ++r) m += attn[b * n + r * w + tt] * g[((size_t)b * n + r * w + s) * C + c];
    gvd[o] = m;
  }
  for (int t = JPB_TID; t < 2 * CCT_JT; t += JPB_NT) {
    const int which = t / CCT_JT, j = j0 + t % CCT_JT;
    if (j >= n) continue;
    if (which == 0) {
      float s1 = 0.f;
      for (int c = 0; c < C; ++c) s1 += g[((size_t)b * n + j) * C + c] * fused[((size_t)b * n + j) * C + c];
      gS[b * n + j] = s1;
    } else {
      const int r = j / w, tt = j - r * w;         // j = (r, t) as an index of attn
      float s2 = 0.f;
      for (int s = 0; s < w; ++s)
        for (int c = 0; c < C; ++c) s2 += g[((size_t)b * n + r * w + s) * C + c] * vd[((size_t)b * n + tt * w + s) * C + c];
      gattn[b * n + j] = s2;
    }
  }
}

// CycledViewProjection's transform module: per (sample, channel) a 2-layer MLP over the n positions; x, y: [B][n][C].
// One generic stage, block = (sample, output position), thread = channel:
//   TRANS = 0:  y[b][o][c] = relu(sum_i W[o][i] x[b][i][c] + bias[o])                         (forward layer)
//   TRANS = 1:  y[b][i][c] = [gate[b][i][c] > 0] * sum_o W[o][i] x[b][o][c]                   (backward through a layer)
template <int TRANS>
__global__ void __launch_bounds__(128) cvp_stage_kernel(const float* x, const float* W, const float* bias, const float* gate, float* y, int n, int C) {
  const int b = blockIdx.x, o = blockIdx.y;
  for (int c = JPB_TID; c < C; c += JPB_NT) {
    float s = (!TRANS && bias) ? bias[o] : 0.f;
    for (int i = 0; i < n; ++i) s += (TRANS ? W[i * n + o] : W[o * n + i]) * x[((size_t)b * n + i) * C + c];
    const size_t idx = ((size_t)b * n + o) * C + c;
    if (TRANS) y[idx] = (!gate || gate[idx] > 0.f) ? s : 0.f;
    else y[idx] = s > 0.f ? s : 0.f;
  }
}

// dz = g * [y > 0]
__global__ void __launch_bounds__(256) relu_gate_kernel(const float* g, const float* y, float* dz, long long nel) {
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < nel; i += (long long)gridDim.x * JPB_NT) dz[i] = y[i] > 0.f ? g[i] : 0.f;
}

// dW[o][i] += sum_{b,c} dz[b][o][c] * xin[b][i][c];  db[o] += sum_{b,c} dz[b][o][c]      block = output position o, thread = i
__global__ void __launch_bounds__(128) cvp_wgrad_kernel(const float* dz, const float* xin, float* dW, float* db, int B, int n, int C) {
  const int o = blockIdx.x;
  for (int i = JPB_TID; i <= n; i += JPB_NT) {      // i == n: the bias column
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* zr = dz + ((size_t)b * n + o) * C;
      const float* xr = xin + ((size_t)b * n + (i < n ? i : 0)) * C;
      if (i < n) for (int c = 0; c < C; ++c) s += zr[c] * xr[c];
      else for (int c = 0; c < C; ++c) s += zr[c];
    }
    if (i < n) dW[o * n + i] += s; else db[o] += s;
  }
}

}  // namespace

extern "C" int jpb_cct_select_fwd(const float* q, const float* k, const float* v, const float* qd, const float* kd, float* T, float* S, int* arg,
                                  float* attn, int* argd, int B, int n, int Cq, int C, void* stream) {
  if (!q || !k || !v || !qd || !kd || !T || !S || !arg || !attn || !argd || B < 1 || n < 1 || n > JPB_CCT_MAX_N) return JPB_ERR_ARG;
  JPB_LAUNCH(cct_select_fwd_kernel, dim3(B, (n + CCT_JT - 1) / CCT_JT, 2), dim3(128), 0, (cudaStream_t)stream, q, k, v, qd, kd, T, S, arg, attn, argd, n, Cq, C);
  return jpb_status();
}

extern "C" int jpb_cct_select_bwd(const float* q, const float* k, const float* qd, const float* kd, const int* arg, const int* argd, const float* gT,
                                  const float* gS, const float* gattn, float* gq, float* gk, float* gv, float* gqd, float* gkd, int B, int n, int Cq,
                                  int C, void* stream) {
  if (!q || !k || !qd || !kd || !arg || !argd || !gT || !gS || !gattn || !gq || !gk || !gv || !gqd || !gkd || n > JPB_CCT_MAX_N) return JPB_ERR_ARG;
  JPB_LAUNCH(cct_select_bwd_kernel, dim3(B, (n + CCT_JT - 1) / CCT_JT), dim3(128), 0, (cudaStream_t)stream, q, k, qd, kd, arg, argd, gT, gS, gattn, gq, gk, gv, gqd, gkd, n, Cq, C);
  return jpb_status();
}

extern "C" int jpb_cct_combine_fwd(const float* front, const float* fused, const float* S, const float* attn, const float* vd, float* out, int B,
                                   int h, int w, int C, void* stream) {
  if (!front || !fused || !S || !attn || !vd || !out || B < 1 || h != w) return JPB_ERR_ARG;   // attn @ vd needs square maps (as the reference)
  JPB_LAUNCH(cct_combine_fwd_kernel, dim3(B, (h * w + CCT_JT - 1) / CCT_JT), dim3(128), 0, (cudaStream_t)stream, front, fused, S, attn, vd, out, h, w, C);
  return jpb_status();
}

extern "C" int jpb_cct_combine_bwd(const float* g, const float* fused, const float* S, const float* attn, const float* vd, float* gfused, float* gS,
                                   float* gattn, float* gvd, int B, int h, int w, int C, void* stream) {
  if (!g || !fused || !S || !attn || !vd || !gfused || !gS || !gattn || !gvd || B < 1 || h != w) return JPB_ERR_ARG;
  JPB_LAUNCH(cct_combine_bwd_kernel, dim3(B, (h * w + CCT_JT - 1) / CCT_JT), dim3(128), 0, (cudaStream_t)stream, g, fused, S, attn, vd, gfused, gS, gattn, gvd, h, w, C);
  return jpb_status();
}

extern "C" int jpb_cvp_mlp_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y1, float* y2, int B, int n,
                               int C, void* stream) {
  if (!x || !W1 || !b1 || !W2 || !b2 || !y1 || !y2 || B < 1 || n < 1) return JPB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  JPB_LAUNCH(cvp_stage_kernel<0>, dim3(B, n), dim3(128), 0, st, x, W1, b1, nullptr, y1, n, C);
  JPB_LAUNCH(cvp_stage_kernel<0>, dim3(B, n), dim3(128), 0, st, y1, W2, b2, nullptr, y2, n, C);
  return jpb_status();
}

extern "C" int jpb_cvp_mlp_bwd(const float* x, const float* W1, const float* W2, const float* y1, const float* y2, const float* g, float* dz2, float* dz1,
                               float* dx, float* dW1, float* db1, float* dW2, float* db2, int B, int n, int C, void* stream) {
  if (!x || !W1 || !W2 || !y1 || !y2 || !g || !dz2 || !dz1 || !dx || !dW1 || !db1 || !dW2 || !db2 || B < 1) return JPB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nel = (long long)B * n * C;
  JPB_LAUNCH(relu_gate_kernel, dim3((unsigned)((nel + 255) / 256)), dim3(256), 0, st, g, y2, dz2, nel);
  JPB_LAUNCH(cvp_stage_kernel<1>, dim3(B, n), dim3(128), 0, st, dz2, W2, nullptr, y1, dz1, n, C);
  JPB_LAUNCH(cvp_stage_kernel<1>, dim3(B, n), dim3(128), 0, st, dz1, W1, nullptr, nullptr, dx, n, C);
  JPB_LAUNCH(cvp_wgrad_kernel, dim3(n), dim3(128), 0, st, dz2, y1, dW2, db2, B, n, C);
  JPB_LAUNCH(cvp_wgrad_kernel, dim3(n), dim3(128), 0, st, dz1, x, dW1, db1, B, n, C);
  return jpb_status();
}
