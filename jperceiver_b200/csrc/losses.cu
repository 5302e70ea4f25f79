// The rest of the self-supervised / BEV loss chain as fused CUDA-core kernels (HBM-bound, no host
// round-trips).  Reference code replaced (paths under /root/reference/mono/model/mono_baseline):
//   area pyramid + smoothness : net.py:182-190, 758-786 (get_smooth_loss, gradient)
//   CGT scale label           : net.py:212-310, 403-476, 529-543; layers.py:214-252 (SE3)
//   CGT scale loss            : net.py:193-211
//   signed distance map       : boundary_loss.py:121-147 (scipy EDT + skimage boundaries on the CPU)
//   BEV head loss             : net.py:554-617, dice_loss.py:31-81,293-331, boundary_loss.py:150-192
//   transform (L1) loss       : net.py:619-622
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

// =============================================================================== area pyramid
// J_s = exact (2^(s+1))^2 box mean of the target frame, s = 0..nlev-1, one pass over the frame:
// each CTA owns a 16x16-pixel block of the finest requested factor (up to 16x16 input pixels).
__global__ void __launch_bounds__(256) area_pyramid_kernel(const float* img, int BC, int H, int W, JpbPyramid out) {
  __shared__ float s[256 + 64 + 16 + 4 + 1];  // 16x16 input pixels, then 8x8, 4x4, 2x2, 1x1 means
  const int bc = blockIdx.z;
  const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
  for (int e = JPB_TID; e < 256; e += JPB_NT) {
    const int y = y0 + (e >> 4), x = x0 + (e & 15);
    s[e] = (y < H && x < W) ? img[((size_t)bc * H + y) * W + x] : 0.f;
  }
  __syncthreads();
  const float* prev = s;
  float* cur = s + 256;
  int pn = 16;
  for (int lv = 0; lv < out.nlev; ++lv) {
    const int n = pn >> 1, f = 16 / n;   // level lv is the 2x2 mean of the previous level (box means compose exactly)
    const int h = H / f, w = W / f;
    for (int e = JPB_TID; e < n * n; e += JPB_NT) {
      const int ty = e / n, tx = e - ty * n;
      const float* p = prev + (2 * ty) * pn + 2 * tx;
      const float r = 0.25f * (p[0] + p[1] + p[pn] + p[pn + 1]);
      cur[e] = r;
      const int oy = y0 / f + ty, ox = x0 / f + tx;
      if (oy < h && ox < w) out.level[lv][((size_t)bc * h + oy) * w + ox] = r;
    }
    __syncthreads();
    prev = cur;
    cur += n * n;
    pn = n;
  }
}

// =============================================================================== smoothness
// acc[b][0..5] += sum |d_k disp_b| * exp(-0.5 mean_c |d_k J_b|) for the six derivative stencils,
// acc[b][6] += sum disp_b.   k: 0 dx, 1 dy, 2 dxx, 3 dxy(=dy of dx), 4 dyx(=dx of dy), 5 dyy.
struct Stencil {
  float d[6];  // derivative of the disparity
  float w[6];  // edge-aware weight
  bool ok[6];
};

__device__ __forceinline__ float fdx(const float* a, int w, int y, int x) { return a[y * w + x + 1] - a[y * w + x]; }
__device__ __forceinline__ float fdy(const float* a, int w, int y, int x) { return a[(y + 1) * w + x] - a[y * w + x]; }

__device__ __forceinline__ void stencils(const float* a, int h, int w, int y, int x, float out[6], bool ok[6]) {
  ok[0] = x + 1 < w; ok[1] = y + 1 < h; ok[2] = x + 2 < w; ok[3] = ok[0] && ok[1]; ok[4] = ok[3]; ok[5] = y + 2 < h;
  out[0] = ok[0] ? fdx(a, w, y, x) : 0.f;
  out[1] = ok[1] ? fdy(a, w, y, x) : 0.f;
  out[2] = ok[2] ? fdx(a, w, y, x + 1) - fdx(a, w, y, x) : 0.f;
  out[3] = ok[3] ? fdx(a, w, y + 1, x) - fdx(a, w, y, x) : 0.f;
  out[4] = ok[4] ? fdy(a, w, y, x + 1) - fdy(a, w, y, x) : 0.f;
  out[5] = ok[5] ? fdy(a, w, y + 1, x) - fdy(a, w, y, x) : 0.f;
}

__device__ __forceinline__ void smooth_weights(const float* J, int h, int w, int y, int x, float wt[6]) {
  float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  bool ok[6];
  for (int c = 0; c < 3; ++c) {
    float g[6];
    stencils(J + (size_t)c * h * w, h, w, y, x, g, ok);
    for (int k = 0; k < 6; ++k) m[k] += fabsf(g[k]);
  }
  for (int k = 0; k < 6; ++k) wt[k] = expf(-0.5f * (m[k] * (1.f / 3.f)));
}

__global__ void __launch_bounds__(256) smooth_fwd_kernel(const float* disp, const float* J, int h, int w, double* acc) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const float* d = disp + (size_t)b * h * w;
  const float* Jb = J + (size_t)b * 3 * h * w;
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < h * w; e += gridDim.x * JPB_NT) {
    const int y = e / w, x = e - y * w;
    float g[6], wt[6];
    bool ok[6];
    stencils(d, h, w, y, x, g, ok);
    smooth_weights(Jb, h, w, y, x, wt);
    for (int k = 0; k < 6; ++k)
      if (ok[k]) s[k] += fabsf(g[k]) * wt[k];
    s[6] += d[e];
  }
  for (int k = 0; k < 7; ++k) {
    const double t = jpb_block_sum<double>((double)s[k], red);
    if (JPB_TID == 0) atomicAdd(&acc[b * 7 + k], t);
  }
}

// loss = sw * sum_k (1/(B N_k)) sum_b S_bk / (m_b + 1e-7)      (disp_norm)   m_b = mean disp_b
__device__ __forceinline__ void smooth_counts(int h, int w, double N[6]) {
  N[0] = (double)h * (w - 1); N[1] = (double)(h - 1) * w; N[2] = (double)h * (w - 2);
  N[3] = (double)(h - 1) * (w - 1); N[4] = N[3]; N[5] = (double)(h - 2) * w;
}

__global__ void smooth_finalize_kernel(const double* acc, int B, int h, int w, int disp_norm, float weight, float* out) {
  if (blockIdx.x != 0 || JPB_TID != 0) return;
  double N[6];
  smooth_counts(h, w, N);
  double tot = 0.0;
  for (int b = 0; b < B; ++b) {
    const double den = disp_norm ? (double)((float)(acc[b * 7 + 6] / ((double)h * w)) + 1e-7f) : 1.0;
    for (int k = 0; k < 6; ++k) tot += acc[b * 7 + k] / den / ((double)B * N[k]);
  }
  out[0] = (float)(tot * (double)weight);
}

__global__ void __launch_bounds__(256) smooth_bwd_kernel(const float* disp, const float* J, int B, int h, int w, const double* acc,
                                                         int disp_norm, float weight, const float* gout, float* gdisp) {
  const int b = blockIdx.y;
  const float* d = disp + (size_t)b * h * w;
  const float* Jb = J + (size_t)b * 3 * h * w;
  float* gd = gdisp + (size_t)b * h * w;
  double N[6];
  smooth_counts(h, w, N);
  const double den = disp_norm ? (double)((float)(acc[b * 7 + 6] / ((double)h * w)) + 1e-7f) : 1.0;
  const float g0 = gout[0] * weight;
  float ck[6];
  double mean_term = 0.0;
  for (int k = 0; k < 6; ++k) {
    ck[k] = (float)((double)g0 / ((double)B * N[k]) / den);
    mean_term += (double)g0 * acc[b * 7 + k] / ((double)B * N[k]) / (den * den);
  }
  const float mt = disp_norm ? (float)(-mean_term / ((double)h * w)) : 0.f;
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < h * w; e += gridDim.x * JPB_NT) {
    const int y = e / w, x = e - y * w;
    float g[6], wt[6];
    bool ok[6];
    stencils(d, h, w, y, x, g, ok);
    smooth_weights(Jb, h, w, y, x, wt);
    float t[6];
    for (int k = 0; k < 6; ++k) t[k] = ok[k] ? ck[k] * wt[k] * (g[k] > 0.f ? 1.f : (g[k] < 0.f ? -1.f : 0.f)) : 0.f;
    // adjoint of each stencil, scattered
    float c00 = mt - t[0] - t[1] + t[2] + t[3] + t[4] + t[5];
    atomicAdd(&gd[e], c00);
    if (ok[0]) atomicAdd(&gd[e + 1], t[0] - 2.f * t[2] - t[3] - t[4]);
    if (ok[1]) atomicAdd(&gd[e + w], t[1] - t[3] - t[4] - 2.f * t[5]);
    if (ok[2]) atomicAdd(&gd[e + 2], t[2]);
    if (ok[3]) atomicAdd(&gd[e + w + 1], t[3] + t[4]);
    if (ok[5]) atomicAdd(&gd[e + 2 * w], t[5]);
  }
}

// =============================================================================== CGT scale label
struct LabelGeom {
  float S[9];  // src_norm <- dst_norm homography (torchgeometry's warp grid)
};

__device__ __forceinline__ void inv3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, Bq = -(d * i - f * g), Cq = d * h - e * g;
  const double det = a * A + b * Bq + c * Cq, id = 1.0 / det;
  o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
  o[3] = Bq * id; o[4] = (a * i - c * g) * id; o[5] = -(a * f - c * d) * id;
  o[6] = Cq * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}
__device__ __forceinline__ void mul3(const double* a, const double* b, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// bilinear sample with zero padding of a virtual occ x occ map given by functor
template <typename Fn>
__device__ __forceinline__ float sample_zeros(Fn fn, int n, float ix, float iy) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  float v = 0.f;
  const bool xa = x0 >= 0 && x0 < n, xb = x0 + 1 >= 0 && x0 + 1 < n, ya = y0 >= 0 && y0 < n, yb = y0 + 1 >= 0 && y0 + 1 < n;
  if (ya && xa) v += fn(y0, x0) * (1.f - tx) * (1.f - ty);
  if (ya && xb) v += fn(y0, x0 + 1) * tx * (1.f - ty);
  if (yb && xa) v += fn(y0 + 1, x0) * (1.f - tx) * ty;
  if (yb && xb) v += fn(y0 + 1, x0 + 1) * tx * ty;
  return v;
}

__global__ void __launch_bounds__(256) scale_label_kernel(JpbScaleLabelArgs a) {
  __shared__ LabelGeom geo;
  const int b = blockIdx.y;
  const int occ = a.occ, Hf = a.Hf, Wf = a.Wf;
  if (JPB_TID == 0) {
    // cam_T_ground = Tr * inverse(SE3(I,[0,0,h]))  ->  columns r1, r2 and t' = t - h*r3
    const float* K = a.K3 + (size_t)b * a.k_stride;
    const float* T = a.Tr + (size_t)b * 16;
    double Kd[9], R[9], Hm[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Kd[i * 3 + j] = K[i * a.k_row + j];
    for (int i = 0; i < 3; ++i) {
      // the reference builds these in fp32 (torch.bmm); keep fp32 products for the translation column
      const float tcol = T[i * 4 + 2] * (-a.cam_height) + T[i * 4 + 3];
      R[i * 3 + 0] = T[i * 4 + 0]; R[i * 3 + 1] = T[i * 4 + 1]; R[i * 3 + 2] = tcol;
    }
    mul3(Kd, R, Hm);                       // img_H_ground
    double Hi[9], A[9];
    inv3(Hm, Hi);                          // ground_H_img
    const double s = (double)occ / 40.0;
    const double Sh[9] = {s, 0, 0, 0, s, (double)(occ / 2), 0, 0, 1};
    mul3(Sh, Hi, A);                       // shiftedground_H_img : image px -> BEV px  (= inverse of the warp's M)
    // src_norm <- dst_norm = Ns * A * Nd^-1, Ns/Nd = torchgeometry's normal_transform_pixel
    const double Ns[9] = {2.0 / (occ - 1), 0, -1, 0, 2.0 / (occ - 1), -1, 0, 0, 1};
    const double Ndi[9] = {(Wf - 1) / 2.0, 0, (Wf - 1) / 2.0, 0, (Hf - 1) / 2.0, (Hf - 1) / 2.0, 0, 0, 1};
    double t1[9], t2[9];
    mul3(Ns, A, t1);
    mul3(t1, Ndi, t2);
    for (int i = 0; i < 9; ++i) geo.S[i] = (float)t2[i];
  }
  __syncthreads();
  const float* L = a.label + (size_t)b * occ * occ;
  const float zs = 40.f / (float)occ;
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < Hf * Wf; e += gridDim.x * JPB_NT) {
    const int y = e / Wf, x = e - y * Wf;
    const float xn = -1.f + 2.f * (float)x / (float)(Wf - 1), yn = -1.f + 2.f * (float)y / (float)(Hf - 1);
    const float q0 = geo.S[0] * xn + geo.S[1] * yn + geo.S[2];
    const float q1 = geo.S[3] * xn + geo.S[4] * yn + geo.S[5];
    const float q2 = geo.S[6] * xn + geo.S[7] * yn + geo.S[8];
    const float gx = q0 / q2, gy = q1 / q2;
    float ix, iy;
    if (a.align_corners) { ix = (gx + 1.f) * 0.5f * (float)(occ - 1); iy = (gy + 1.f) * 0.5f * (float)(occ - 1); }
    else { ix = ((gx + 1.f) * (float)occ - 1.f) * 0.5f; iy = ((gy + 1.f) * (float)occ - 1.f) * 0.5f; }
    float out = 0.f;
    if (ix > -1.f && ix < (float)occ && iy > -1.f && iy < (float)occ) {
      // rot90(k=3): map_rot[i][j] = map[occ-1-j][i];  z_rot[i][j] = (j+1)*40/occ - delta
      const float wz = sample_zeros([&](int i, int j) { return (float)(j + 1) * zs - a.z_offset; }, occ, ix, iy);
      if (a.mode == 2) {
        out = a.quad[e] != 0 ? wz : 0.f;                                    // dynamic: cv2 quad only, the label is never warped
      } else {
        const float wl = sample_zeros([&](int i, int j) { return L[(occ - 1 - j) * occ + i]; }, occ, ix, iy);
        if (a.mode == 0) out = wz * wl;                                     // Argo_both: product of the two warps
        else out = (wl >= 0.99999905f && a.quad[e] != 0) ? wz : 0.f;        // static: exact-1 mask AND cv2 quad
      }
    }
    a.out[(size_t)b * Hf * Wf + e] = out;
  }
}

// =============================================================================== CGT scale loss
__device__ __forceinline__ void up_axis_sl(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

template <bool BWD>
__global__ void __launch_bounds__(256) scale_loss_kernel(JpbScaleLossArgs a) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const int Hf = a.Hf, Wf = a.Wf, hs = a.hs, ws = a.ws;
  const float* d = a.disp + (size_t)b * hs * ws;
  const float* lab = a.label + (size_t)b * Hf * Wf;
  const float sy = (float)hs / (float)Hf, sx = (float)ws / (float)Wf;
  const float rng = a.max_disp - a.min_disp;
  float gscale = 0.f;
  if (BWD) gscale = a.grad_out[0] * a.weight / (float)a.acc[1];
  double s = 0.0, n = 0.0;
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < Hf * Wf; e += gridDim.x * JPB_NT) {
    const float g = lab[e];
    if (!(g > 0.f)) continue;
    const int y = e / Wf, x = e - y * Wf;
    if (a.crop && !(y >= 153 && y < 371 && x >= 44 && x < 1197)) continue;
    int y0, y1, x0, x1;
    float ly, lx;
    up_axis_sl(y, sy, hs, y0, y1, ly);
    up_axis_sl(x, sx, ws, x0, x1, lx);
    const float z00 = 1.f / (a.min_disp + rng * d[y0 * ws + x0]), z01 = 1.f / (a.min_disp + rng * d[y0 * ws + x1]);
    const float z10 = 1.f / (a.min_disp + rng * d[y1 * ws + x0]), z11 = 1.f / (a.min_disp + rng * d[y1 * ws + x1]);
    const float raw = (1.f - ly) * ((1.f - lx) * z00 + lx * z01) + ly * ((1.f - lx) * z10 + lx * z11);
    const float p = fminf(fmaxf(raw, 1e-3f), 80.f);
    if (!BWD) {
      s += (double)(fabsf(g - p) / g);
      n += 1.0;
    } else {
      if (raw < 1e-3f || raw > 80.f) continue;
      const float df = g - p;
      const float gp = gscale * (df > 0.f ? -1.f : (df < 0.f ? 1.f : 0.f)) / g;
      float* gd = a.grad_disp + (size_t)b * hs * ws;
      atomicAdd(&gd[y0 * ws + x0], gp * (1.f - ly) * (1.f - lx) * (-rng * z00 * z00));
      atomicAdd(&gd[y0 * ws + x1], gp * (1.f - ly) * lx * (-rng * z01 * z01));
      atomicAdd(&gd[y1 * ws + x0], gp * ly * (1.f - lx) * (-rng * z10 * z10));
      atomicAdd(&gd[y1 * ws + x1], gp * ly * lx * (-rng * z11 * z11));
    }
  }
  if (!BWD) {
    const double ts = jpb_block_sum<double>(s, red);
    const double tn = jpb_block_sum<double>(n, red);
    if (JPB_TID == 0 && tn > 0.0) {
      atomicAdd(&a.acc[0], ts);
      atomicAdd(&a.acc[1], tn);
    }
  }
}

// =============================================================================== signed distance map
constexpr int SDF_INF = 20000;

// pass 1: per column, vertical distance to the nearest background (gfg) / foreground (gbg) pixel
__global__ void __launch_bounds__(128) sdf_columns_kernel(const float* label, int n, int* gfg, int* gbg) {
  const int b = blockIdx.y;
  const float* L = label + (size_t)b * n * n;
  int* gf = gfg + (size_t)b * n * n;
  int* gb = gbg + (size_t)b * n * n;
  for (int x = blockIdx.x * JPB_NT + JPB_TID; x < n; x += gridDim.x * JPB_NT) {
    int df = SDF_INF, db = SDF_INF;
    for (int y = 0; y < n; ++y) {
      const bool fg = L[y * n + x] > 0.5f;
      df = fg ? min(df + 1, SDF_INF) : 0;   // distance to nearest background above
      db = fg ? 0 : min(db + 1, SDF_INF);   // distance to nearest foreground above
      gf[y * n + x] = df;
      gb[y * n + x] = db;
    }
    df = SDF_INF; db = SDF_INF;
    for (int y = n - 1; y >= 0; --y) {
      const bool fg = L[y * n + x] > 0.5f;
      df = fg ? min(df + 1, SDF_INF) : 0;
      db = fg ? 0 : min(db + 1, SDF_INF);
      gf[y * n + x] = min(gf[y * n + x], df);
      gb[y * n + x] = min(gb[y * n + x], db);
    }
  }
}

// pass 2: per row, exact lower envelope by brute force over the row (n <= 1024), then sign/boundary
__global__ void __launch_bounds__(256) sdf_rows_kernel(const float* label, int n, const int* gfg, const int* gbg, float* sdf) {
  JPB_DYN_SMEM(int, sm);  // [2][n]
  const int b = blockIdx.y, y = blockIdx.x;
  const float* L = label + (size_t)b * n * n;
  int* sf = sm;
  int* sb = sm + n;
  for (int x = JPB_TID; x < n; x += JPB_NT) {
    sf[x] = gfg[((size_t)b * n + y) * n + x];
    sb[x] = gbg[((size_t)b * n + y) * n + x];
  }
  __syncthreads();
  for (int x = JPB_TID; x < n; x += JPB_NT) {
    const bool fg = L[y * n + x] > 0.5f;
    const int* g = fg ? sf : sb;
    long long best = (long long)SDF_INF * SDF_INF;
    for (int xp = 0; xp < n; ++xp) {
      const long long dx = x - xp, gy = g[xp];
      const long long v = dx * dx + gy * gy;
      best = v < best ? v : best;
    }
    float out;
    if (best >= (long long)SDF_INF * SDF_INF) out = 0.f;  // no foreground (or no background) in the map
    else {
      const double dist = sqrt((double)best);
      out = fg ? (float)(-dist) : (float)dist;
      if (fg) {  // inner boundary (4-connectivity, image edge replicated) -> 0
        const bool bnd = (x > 0 && !(L[y * n + x - 1] > 0.5f)) || (x + 1 < n && !(L[y * n + x + 1] > 0.5f)) ||
                         (y > 0 && !(L[(y - 1) * n + x] > 0.5f)) || (y + 1 < n && !(L[(y + 1) * n + x] > 0.5f));
        if (bnd) out = 0.f;
      }
    }
    sdf[((size_t)b * n + y) * n + x] = out;
  }
}

// a map without any foreground must give an all-zero SDF even though its background EDT is finite
// (boundary_loss.py:137 `if posmask.any()`): gbg == INF everywhere handles it above.

// =============================================================================== BEV head loss
// out = loss_weight * region + [loss_sum >= 2] loss2_weight * BD + [loss_sum == 3] CE     (net.py:554-617)
// region (a.region): 0 soft IoU, 1 soft Dice, 2 Tversky(alpha=.3, beta=.7) — all of the form
//   -mean_{b,c} (k*TP + 1) / (k*TP + al*FP + be*FN + 1)   (dice_loss.py:255-372; k = 2 for Dice) —
// or 3 focal: mean over pixels of -alpha_y (1-pt)^2 log(pt), pt = (1-s) p_y + s p_other + s, s = 1e-5, alpha = (.25, .75)
// (focal_loss.py:7-92).
// acc layout: [b*4 + {A,Bq,Cq,Dq}] per-sample soft confusion sums (A = sum p0[y=0], Bq = sum p0[y=1], Cq = sum p1[y=0],
// Dq = sum p1[y=1]), then [4B + {ce_num, ce_den, bd, focal}]
struct RegionCoef { double k, al, be; };
__device__ __forceinline__ RegionCoef region_coef(int region) {
  RegionCoef r;
  r.k = region == 1 ? 2.0 : 1.0;
  r.al = region == 2 ? 0.3 : 1.0;
  r.be = region == 2 ? 0.7 : 1.0;
  return r;
}
constexpr float FOCAL_S = 1e-5f;

__global__ void __launch_bounds__(256) bev_fwd_kernel(JpbBevArgs a) {
  __shared__ double red[32];
  const int b = blockIdx.y, n2 = a.occ * a.occ;
  const float* l0 = a.logits + (size_t)b * a.stride_b;
  const float* lab = a.label + (size_t)b * n2;
  const float* phi = a.sdf + (size_t)b * n2;
  float A = 0.f, Bq = 0.f, Cq = 0.f, Dq = 0.f;
  double cen = 0.0, ced = 0.0, bd = 0.0, foc = 0.0;
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < n2; e += gridDim.x * JPB_NT) {
    const float u = l0[(size_t)e * a.stride_p], v = l0[(size_t)e * a.stride_p + a.stride_c];
    const float m = fmaxf(u, v);
    const float eu = expf(u - m), ev = expf(v - m);
    const float den = eu + ev;
    const float p0 = eu / den, p1 = ev / den;
    const bool fg = lab[e] > 0.5f;
    if (fg) { Bq += p0; Dq += p1; } else { A += p0; Cq += p1; }
    const float wy = fg ? a.w_fg : 1.f;
    const float nll = (m + logf(den)) - (fg ? v : u);
    cen += (double)(wy * nll);
    ced += (double)wy;
    bd += (double)(p1 * phi[e]);
    if (a.region == 3) {
      const float pt = (1.f - FOCAL_S) * (fg ? p1 : p0) + FOCAL_S * (fg ? p0 : p1) + FOCAL_S;
      const float om = 1.f - pt;
      foc += (double)(-(fg ? 0.75f : 0.25f) * om * om * logf(pt));
    }
  }
  const double vals[8] = {(double)A, (double)Bq, (double)Cq, (double)Dq, cen, ced, bd, foc};
  for (int k = 0; k < 8; ++k) {
    const double t = jpb_block_sum<double>(vals[k], red);
    if (JPB_TID == 0) atomicAdd(k < 4 ? &a.acc[b * 4 + k] : &a.acc[4 * a.B + (k - 4)], t);
  }
}

__global__ void bev_finalize_kernel(const double* acc, int B, int occ, float lw, float l2w, int region, int loss_sum, float* out) {
  if (blockIdx.x != 0 || JPB_TID != 0) return;
  double reg = 0.0;
  if (region == 3) {
    reg = acc[4 * B + 3] / ((double)B * occ * occ);
  } else {
    const RegionCoef rc = region_coef(region);
    for (int b = 0; b < B; ++b) {
      const double A = acc[b * 4], Bq = acc[b * 4 + 1], Cq = acc[b * 4 + 2], Dq = acc[b * 4 + 3];
      // class 0: TP = A, FP = Bq, FN = Cq;  class 1: TP = Dq, FP = Cq, FN = Bq
      reg += (rc.k * A + 1.0) / (rc.k * A + rc.al * Bq + rc.be * Cq + 1.0) + (rc.k * Dq + 1.0) / (rc.k * Dq + rc.al * Cq + rc.be * Bq + 1.0);
    }
    reg = -reg / (2.0 * B);
  }
  double v = (double)lw * reg;
  if (loss_sum >= 2) v += (double)l2w * acc[4 * B + 2] / ((double)B * occ * occ);
  if (loss_sum == 3) v += acc[4 * B] / acc[4 * B + 1];
  out[0] = (float)v;
}

__global__ void __launch_bounds__(256) bev_bwd_kernel(JpbBevArgs a, const float* gout, float* glogits) {
  const int b = blockIdx.y, n2 = a.occ * a.occ;
  const float* l0 = a.logits + (size_t)b * a.stride_b;
  float* gl = glogits + (size_t)b * a.stride_b;
  const float* lab = a.label + (size_t)b * n2;
  const float* phi = a.sdf + (size_t)b * n2;
  const float g = gout[0];
  // d loss / d{A, Bq, Cq, Dq} of the ratio-type region terms
  float cA = 0.f, cB = 0.f, cC = 0.f, cD = 0.f;
  if (a.region != 3) {
    const RegionCoef rc = region_coef(a.region);
    const double A = a.acc[b * 4], Bq = a.acc[b * 4 + 1], Cq = a.acc[b * 4 + 2], Dq = a.acc[b * 4 + 3];
    const double U0 = rc.k * A + rc.al * Bq + rc.be * Cq + 1.0, U1 = rc.k * Dq + rc.al * Cq + rc.be * Bq + 1.0;
    const double N0 = rc.k * A + 1.0, N1 = rc.k * Dq + 1.0;
    const double k = -(double)a.loss_weight / (2.0 * a.B) * g;
    cA = (float)(k * rc.k * (rc.al * Bq + rc.be * Cq) / (U0 * U0));
    cD = (float)(k * rc.k * (rc.al * Cq + rc.be * Bq) / (U1 * U1));
    cB = (float)(k * (-rc.al * N0 / (U0 * U0) - rc.be * N1 / (U1 * U1)));
    cC = (float)(k * (-rc.be * N0 / (U0 * U0) - rc.al * N1 / (U1 * U1)));
  }
  const float cew = a.loss_sum == 3 ? g / (float)a.acc[4 * a.B + 1] : 0.f;
  const float bdk = a.loss_sum >= 2 ? g * a.loss2_weight / ((float)a.B * (float)n2) : 0.f;
  const float fok = g * a.loss_weight / ((float)a.B * (float)n2);
  for (int e = blockIdx.x * JPB_NT + JPB_TID; e < n2; e += gridDim.x * JPB_NT) {
    const float u = l0[(size_t)e * a.stride_p], v = l0[(size_t)e * a.stride_p + a.stride_c];
    const float m = fmaxf(u, v);
    const float eu = expf(u - m), ev = expf(v - m);
    const float den = eu + ev;
    const float p0 = eu / den, p1 = ev / den;
    const bool fg = lab[e] > 0.5f;
    // d loss / d p_c
    float g0 = fg ? cB : cA;
    float g1 = (fg ? cD : cC) + bdk * phi[e];
    if (a.region == 3) {
      const float pt = (1.f - FOCAL_S) * (fg ? p1 : p0) + FOCAL_S * (fg ? p0 : p1) + FOCAL_S;
      const float om = 1.f - pt;
      // d/dpt [-al (1-pt)^2 log pt] = al (2 (1-pt) log pt - (1-pt)^2 / pt)
      const float dpt = fok * (fg ? 0.75f : 0.25f) * (2.f * om * logf(pt) - om * om / pt);
      g0 += dpt * (fg ? FOCAL_S : 1.f - FOCAL_S);
      g1 += dpt * (fg ? 1.f - FOCAL_S : FOCAL_S);
    }
    const float dot = g0 * p0 + g1 * p1;
    const float wy = fg ? a.w_fg : 1.f;
    gl[(size_t)e * a.stride_p] = p0 * (g0 - dot) + cew * wy * (p0 - (fg ? 0.f : 1.f));
    gl[(size_t)e * a.stride_p + a.stride_c] = p1 * (g1 - dot) + cew * wy * (p1 - (fg ? 1.f : 0.f));
  }
}

// =============================================================================== mean |a - b|
__global__ void __launch_bounds__(256) l1_mean_fwd_kernel(const float* x, const float* y, long long n, double* acc) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * JPB_NT + JPB_TID; e < n; e += (long long)gridDim.x * JPB_NT) s += (double)fabsf(x[e] - y[e]);
  const double t = jpb_block_sum<double>(s, red);
  if (JPB_TID == 0) atomicAdd(acc, t);
}
__global__ void __launch_bounds__(256) l1_mean_bwd_kernel(const float* x, const float* y, long long n, const float* gout, float* gx, float* gy) {
  const float k = gout[0] / (float)n;
  for (long long e = (long long)blockIdx.x * JPB_NT + JPB_TID; e < n; e += (long long)gridDim.x * JPB_NT) {
    const float d = x[e] - y[e];
    const float s = d > 0.f ? k : (d < 0.f ? -k : 0.f);
    gx[e] = s;
    gy[e] = -s;
  }
}

inline int grid_for(long long n, int block, int cap = 148 * 8) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int jpb_area_pyramid(const float* img, int BC, int H, int W, const JpbPyramid* out, void* stream) {
  if (!img || !out || out->nlev < 1 || out->nlev > 4 || (H % (1 << out->nlev)) || (W % (1 << out->nlev))) return JPB_ERR_ARG;
  dim3 grid((W + 15) / 16, (H + 15) / 16, BC);
  JPB_LAUNCH(area_pyramid_kernel, grid, dim3(256), 0, (cudaStream_t)stream, img, BC, H, W, *out);
  return jpb_status();
}

extern "C" int jpb_smooth_fwd(const float* disp, const float* J, int B, int h, int w, int disp_norm, float weight,
                              double* acc, float* out, void* stream) {
  if (!disp || !J || !acc || !out || h < 3 || w < 3) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)h * w, 256, 148 * 4), B);
  JPB_LAUNCH(smooth_fwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, disp, J, h, w, acc);
  JPB_LAUNCH(smooth_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, acc, B, h, w, disp_norm, weight, out);
  return jpb_status();
}

extern "C" int jpb_smooth_bwd(const float* disp, const float* J, int B, int h, int w, int disp_norm, float weight,
                              const double* acc, const float* grad_out, float* grad_disp, void* stream) {
  if (!disp || !J || !acc || !grad_out || !grad_disp) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)h * w, 256, 148 * 4), B);
  JPB_LAUNCH(smooth_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, disp, J, B, h, w, acc, disp_norm, weight, grad_out, grad_disp);
  return jpb_status();
}

extern "C" int jpb_scale_label(const JpbScaleLabelArgs* a, void* stream) {
  if (!a || !a->K3 || !a->Tr || !a->out || a->mode < 0 || a->mode > 2) return JPB_ERR_ARG;
  if ((a->mode != 2 && !a->label) || (a->mode != 0 && !a->quad)) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)a->Hf * a->Wf, 256, 148 * 4), a->B);
  JPB_LAUNCH(scale_label_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_scale_loss_fwd(const JpbScaleLossArgs* a, void* stream) {
  if (!a || !a->disp || !a->label || !a->acc) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)a->Hf * a->Wf, 256, 148 * 4), a->B);
  JPB_LAUNCH(scale_loss_kernel<false>, grid, dim3(256), 0, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_scale_loss_bwd(const JpbScaleLossArgs* a, void* stream) {
  if (!a || !a->disp || !a->label || !a->acc || !a->grad_out || !a->grad_disp) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)a->Hf * a->Wf, 256, 148 * 4), a->B);
  JPB_LAUNCH(scale_loss_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_signed_distance(const float* label, int B, int n, int* work, float* sdf, void* stream) {
  if (!label || !work || !sdf || n < 2 || n > 4096) return JPB_ERR_ARG;
  int* gfg = work;
  int* gbg = work + (size_t)B * n * n;
  JPB_LAUNCH(sdf_columns_kernel, dim3((n + 127) / 128, B), dim3(128), 0, (cudaStream_t)stream, label, n, gfg, gbg);
  JPB_LAUNCH(sdf_rows_kernel, dim3(n, B), dim3(256), 2 * n * sizeof(int), (cudaStream_t)stream, label, n, gfg, gbg, sdf);
  return jpb_status();
}

extern "C" int jpb_bev_loss_fwd(const JpbBevArgs* a, float* out, void* stream) {
  if (!a || !a->logits || !a->label || !a->sdf || !a->acc || !out) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)a->occ * a->occ, 256, 64), a->B);
  JPB_LAUNCH(bev_fwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *a);
  JPB_LAUNCH(bev_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, a->acc, a->B, a->occ, a->loss_weight, a->loss2_weight, a->region, a->loss_sum, out);
  return jpb_status();
}

extern "C" int jpb_bev_loss_bwd(const JpbBevArgs* a, const float* grad_out, float* grad_logits, void* stream) {
  if (!a || !a->logits || !a->label || !a->sdf || !a->acc || !grad_out || !grad_logits) return JPB_ERR_ARG;
  dim3 grid(grid_for((long long)a->occ * a->occ, 256, 64), a->B);
  JPB_LAUNCH(bev_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *a, grad_out, grad_logits);
  return jpb_status();
}

extern "C" int jpb_l1_mean_fwd(const float* x, const float* y, long long n, double* acc, void* stream) {
  if (!x || !y || !acc || n < 1) return JPB_ERR_ARG;
  JPB_LAUNCH(l1_mean_fwd_kernel, dim3(grid_for(n, 256, 148)), dim3(256), 0, (cudaStream_t)stream, x, y, n, acc);
  return jpb_status();
}

extern "C" int jpb_l1_mean_bwd(const float* x, const float* y, long long n, const float* grad_out, float* gx, float* gy, void* stream) {
  if (!x || !y || !grad_out || !gx || !gy) return JPB_ERR_ARG;
  JPB_LAUNCH(l1_mean_bwd_kernel, dim3(grid_for(n, 256, 148)), dim3(256), 0, (cudaStream_t)stream, x, y, n, grad_out, gx, gy);
  return jpb_status();
}
