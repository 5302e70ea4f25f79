// Fused photometric-reprojection loss, one launch per scale (forward) and one per scale (backward).
//
// Replaces, per scale s, the reference chain (all paths under /root/reference/mono/model/mono_baseline):
//   net.py:690-702  generate_images_pred : bilinear up-sample of disp_s -> disp_to_depth -> Backproject
//                                          -> Project -> F.grid_sample(border)
//   layers.py:41-82 Backproject / Project ; layers.py:85-107 SSIM ; net.py:84-92 robust_l1 + mix
//   net.py:159-175  automask identity terms (+N(0,1)*1e-5), min over candidates, argmin, mean
// which in eager PyTorch is ~10 kernels x F x 4 scales with ~30 full-resolution temporaries per call.
// Here each CTA owns a 32x16 pixel tile of one sample: it stages target / identity / warped pixels
// for the tile plus the 1-pixel SSIM apron in shared memory (reflect-indexed at the image border),
// evaluates all 2F candidates from shared memory, takes min/argmin and reduces the tile's sum with
// warp shuffles; HBM sees each input pixel once (+apron) and only scalar / 1-byte-per-pixel outputs.
//
// Algorithmic bytes per sample per scale: 4*H*W*(3+3F) + 4*hs*ws  (forward).
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

constexpr int PT_W = 32, PT_H = 16;           // output tile
constexpr int P1_W = PT_W + 2, P1_H = PT_H + 2, P1_N = P1_W * P1_H;  // +1 apron (SSIM window)
constexpr int P2_W = PT_W + 4, P2_H = PT_H + 4, P2_N = P2_W * P2_H;  // +2 apron (backward)
constexpr float SSIM_C1 = 1e-4f, SSIM_C2 = 9e-4f;

struct SrcGeom {  // P = (K T)[:3,:]  (layers.py:74); p = P * (z * invK3x3 * (x,y,1), 1) in the reference's order
  float P[12];
};

__device__ __forceinline__ void make_geom(const float* K, const float* T, SrcGeom& g) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += K[i * 4 + k] * T[k * 4 + j];
      g.P[i * 4 + j] = s;
    }
}

// camera ray invK[:3,:3] * (x, y, 1)   (layers.py:58)
__device__ __forceinline__ void pixel_ray(const float* iK, int x, int y, float rc[3]) {
  const float fx = (float)x, fy = (float)y;
  rc[0] = iK[0] * fx + iK[1] * fy + iK[2];
  rc[1] = iK[4] * fx + iK[5] * fy + iK[6];
  rc[2] = iK[8] * fx + iK[9] * fy + iK[10];
}

// ATen upsample_bilinear2d(align_corners=False) source index + weights for one axis
__device__ __forceinline__ void up_axis(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

struct DispTap {
  int y0, y1, x0, x1;
  float ly, lx;
};

__device__ __forceinline__ float disp_upsample(const float* d, int hs, int ws, float sy, float sx, int y, int x, DispTap& tp) {
  up_axis(y, sy, hs, tp.y0, tp.y1, tp.ly);
  up_axis(x, sx, ws, tp.x0, tp.x1, tp.lx);
  const float a = d[tp.y0 * ws + tp.x0], b = d[tp.y0 * ws + tp.x1];
  const float c = d[tp.y1 * ws + tp.x0], e = d[tp.y1 * ws + tp.x1];
  return (1.f - tp.ly) * ((1.f - tp.lx) * a + tp.lx * b) + tp.ly * ((1.f - tp.lx) * c + tp.lx * e);
}

struct Sample {      // grid_sample(bilinear, border, align_corners=False) footprint
  int x0, y0;        // north-west tap
  float tx, ty;      // fractional offsets
  float mx, my;      // d(ix)/d(unclipped ix): 0 where the coordinate was clipped
  float p[3];        // projected homogeneous point
};

__device__ __forceinline__ void project(const SrcGeom& g, float z, const float rc[3], int W, int H, Sample& s) {
  const float X0 = z * rc[0], X1 = z * rc[1], X2 = z * rc[2];
  s.p[0] = g.P[0] * X0 + g.P[1] * X1 + g.P[2] * X2 + g.P[3];
  s.p[1] = g.P[4] * X0 + g.P[5] * X1 + g.P[6] * X2 + g.P[7];
  s.p[2] = g.P[8] * X0 + g.P[9] * X1 + g.P[10] * X2 + g.P[11];
  const float den = s.p[2] + 1e-7f;
  const float u = s.p[0] / den, v = s.p[1] / den;
  const float gx = (u / (float)(W - 1) - 0.5f) * 2.f, gy = (v / (float)(H - 1) - 0.5f) * 2.f;
  float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
  s.mx = 1.f; s.my = 1.f;
  // clip_coordinates: NaN-safe min/max order as ATen (min(max(x,0),size-1))
  if (!(ix > 0.f)) { ix = 0.f; s.mx = 0.f; }
  if (ix > (float)(W - 1)) { ix = (float)(W - 1); s.mx = 0.f; }
  if (!(iy > 0.f)) { iy = 0.f; s.my = 0.f; }
  if (iy > (float)(H - 1)) { iy = (float)(H - 1); s.my = 0.f; }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  s.x0 = (int)fx0; s.y0 = (int)fy0;
  s.tx = ix - fx0; s.ty = iy - fy0;
}

__device__ __forceinline__ void gather3(const float* img, int H, int W, const Sample& s, float out[3]) {
  const int x1 = s.x0 + 1, y1 = s.y0 + 1;
  const bool xin = x1 < W, yin = y1 < H;   // north-west tap is always inside after clipping
  const float wnw = (1.f - s.tx) * (1.f - s.ty), wne = s.tx * (1.f - s.ty);
  const float wsw = (1.f - s.tx) * s.ty, wse = s.tx * s.ty;
  const size_t plane = (size_t)H * W;
  const size_t o00 = (size_t)s.y0 * W + s.x0;
  for (int c = 0; c < 3; ++c) {
    const float* p = img + c * plane;
    float v = p[o00] * wnw;
    if (xin) v += p[o00 + 1] * wne;
    if (yin) v += p[o00 + W] * wsw;
    if (xin && yin) v += p[o00 + W + 1] * wse;
    out[c] = v;
  }
}

// 0.85*mean_c SSIM(x,y) + 0.15*mean_c sqrt((x-y)^2+1e-6) from 3x3 windows in shared memory
__device__ __forceinline__ float reproj_error(const float* xs, const float* ys, int plane, int stride, int ctr,
                                              const float* my, const float* syy) {
  float ssim = 0.f, l1 = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float* x = xs + c * plane + ctr;
    const float* y = ys + c * plane + ctr;
    float sx = 0.f, sxx = 0.f, sxy = 0.f;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const float xv = x[dy * stride + dx], yv = y[dy * stride + dx];
        sx += xv; sxx += xv * xv; sxy += xv * yv;
      }
    const float mx = sx * (1.f / 9.f), m_y = my[c];
    const float vx = sxx * (1.f / 9.f) - mx * mx, vxy = sxy * (1.f / 9.f) - mx * m_y;
    const float n = (2.f * mx * m_y + SSIM_C1) * (2.f * vxy + SSIM_C2);
    const float d = (mx * mx + m_y * m_y + SSIM_C1) * (vx + syy[c] + SSIM_C2);
    ssim += __saturatef((1.f - n / d) * 0.5f);
    const float df = y[0] - x[0];
    l1 += sqrtf(df * df + 1e-6f);
  }
  return 0.85f * (ssim * (1.f / 3.f)) + 0.15f * (l1 * (1.f / 3.f));
}

// ------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) photometric_fwd_generic_kernel(JpbPhotoArgs a) {
  JPB_DYN_SMEM(float, sm);
  __shared__ SrcGeom geom[JPB_MAX_SRC];
  __shared__ double red[32];
  const int b = blockIdx.z, x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const int F = a.F, H = a.H, W = a.W;
  const int nid = a.automask ? F : 0;
  float* s_tgt = sm;                       // [3][P1_N]
  float* s_id = s_tgt + 3 * P1_N;          // [nid][3][P1_N]
  float* s_wp = s_id + nid * 3 * P1_N;     // [F][3][P1_N]
  const size_t plane = (size_t)H * W;

  for (int f = JPB_TID; f < F; f += JPB_NT) make_geom(a.K + b * 16, a.T[f] + b * 16, geom[f]);
  __syncthreads();

  const float* disp = a.disp + (size_t)b * a.hs * a.ws;
  const float sy = (float)a.hs / (float)H, sx = (float)a.ws / (float)W;
  for (int e = JPB_TID; e < P1_N; e += JPB_NT) {
    const int hy = e / P1_W, hx = e - hy * P1_W;
    const int y = jpb_reflect(min(y0 + hy - 1, H), H), x = jpb_reflect(min(x0 + hx - 1, W), W);
    const size_t o = (size_t)y * W + x;
    const float* tg = a.target + (size_t)b * 3 * plane + o;
    s_tgt[e] = tg[0]; s_tgt[P1_N + e] = tg[plane]; s_tgt[2 * P1_N + e] = tg[2 * plane];
    DispTap tp;
    const float D = disp_upsample(disp, a.hs, a.ws, sy, sx, y, x, tp);
    const float z = 1.f / (a.min_disp + (a.max_disp - a.min_disp) * D);
    const bool interior = hy >= 1 && hy <= PT_H && hx >= 1 && hx <= PT_W && (y0 + hy - 1) < H && (x0 + hx - 1) < W;
    float rc[3];
    pixel_ray(a.invK + b * 16, x, y, rc);
    for (int f = 0; f < F; ++f) {
      const float* sp = a.src[f] + (size_t)b * 3 * plane;
      if (nid) {
        float* d = s_id + f * 3 * P1_N + e;
        d[0] = sp[o]; d[P1_N] = sp[plane + o]; d[2 * P1_N] = sp[2 * plane + o];
      }
      Sample s;
      project(geom[f], z, rc, W, H, s);
      float v[3];
      gather3(sp, H, W, s, v);
      float* d = s_wp + f * 3 * P1_N + e;
      d[0] = v[0]; d[P1_N] = v[1]; d[2 * P1_N] = v[2];
      if (interior && a.warped[f]) {
        float* wo = a.warped[f] + (size_t)b * 3 * plane + o;
        wo[0] = v[0]; wo[plane] = v[1]; wo[2 * plane] = v[2];
      }
    }
  }
  __syncthreads();

  float local = 0.f;
  for (int e = JPB_TID; e < PT_W * PT_H; e += JPB_NT) {
    const int ty = e / PT_W, tx = e - ty * PT_W;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int ctr = (ty + 1) * P1_W + tx + 1;
    float my[3], vyy[3];
    for (int c = 0; c < 3; ++c) {
      const float* yp = s_tgt + c * P1_N + ctr;
      float s1 = 0.f, s2 = 0.f;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const float v = yp[dy * P1_W + dx];
          s1 += v; s2 += v * v;
        }
      my[c] = s1 * (1.f / 9.f);
      vyy[c] = s2 * (1.f / 9.f) - my[c] * my[c];
    }
    float best = 3.0e38f;
    int besti = 0;
    const size_t po = (size_t)b * plane + (size_t)y * W + x;
    for (int f = 0; f < nid; ++f) {
      float err = reproj_error(s_id + f * 3 * P1_N, s_tgt, P1_N, P1_W, ctr, my, vyy);
      if (a.noise[f]) err += a.noise[f][po];
      else if (a.noise_scale != 0.f)
        err += a.noise_scale * jpb_randn(a.seed, a.stream + (uint64_t)f + (a.step ? 64ull * (uint64_t)a.step[0] : 0ull), (uint64_t)po);
      if (err < best) { best = err; besti = f; }
    }
    for (int f = 0; f < F; ++f) {
      const float err = reproj_error(s_wp + f * 3 * P1_N, s_tgt, P1_N, P1_W, ctr, my, vyy);
      if (err < best) { best = err; besti = nid + f; }
    }
    if (a.min_index) a.min_index[po] = (long long)besti;
    if (a.winner) a.winner[po] = (unsigned char)besti;
    local += best;
  }
  const double tot = jpb_block_sum<double>((double)local, red);
  if (JPB_TID == 0) atomicAdd(a.loss_sum, tot);
}


// ------------------------------------------------------------------------------------ forward, fast path (F <= 2)
// Same staging as the generic kernel (phase 1), but on 32x32 tiles, and the 3x3 window statistics are separable running
// sums: a thread owns one column of an 8-row strip, walks down it once per (channel, candidate) keeping the horizontal
// 3-sums of x, x^2 and x*y of the previous two rows in registers, so one output costs 3 shared loads + ~40 flops per
// (candidate, channel) instead of 18 loads + ~70 flops; the target's own window sums are computed once per channel and
// shared by all candidates; one Philox call per pixel feeds the noise of both identity candidates.
constexpr int QT_W = 32, QT_H = 32, QSR = 4;                   // tile and strip height
constexpr int Q1_W = QT_W + 2, Q1_H = QT_H + 2, Q1_N = Q1_W * Q1_H;
constexpr int QNC = 4;                                          // candidates handled by the fast path (2 identity + 2 warped)

__device__ __forceinline__ float fast_sqrt(float x) {
#ifdef JPB_HOST_EMU
  return sqrtf(x);
#else
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

// two independent standard normals from one Philox call (fast-math Box-Muller; the draw only breaks ties at 1e-5)
__device__ __forceinline__ void randn2(uint64_t seed, uint64_t stream, uint64_t ctr, float& n0, float& n1) {
  uint32_t r[4];
  jpb_philox4(seed, stream, ctr, r);
  const float u1 = jpb_u01(r[0]), u2 = jpb_u01(r[1]);
  const float rad = fast_sqrt(-2.0f * __logf(u1));
  const float ang = 6.28318530717958647692f * u2;
  n0 = rad * __cosf(ang);
  n1 = rad * __sinf(ang);
}

// one candidate, one channel, one strip: running 3x3 sums down the column.  X(r, j) returns the candidate's value in strip
// row r (0 = apron row above the strip) and horizontal neighbour j (0 left, 1 centre, 2 right).
#define JPB_STRIP_CANDIDATE(X, ERR)                                                                        \
  {                                                                                                        \
    float h1a = 0.f, h1b = 0.f, h2a = 0.f, h2b = 0.f, h3a = 0.f, h3b = 0.f, xc_prev = 0.f;                 \
    _Pragma("unroll") for (int r = 0; r < QSR + 2; ++r) {                                                  \
      const float xa = X(r, 0), xb = X(r, 1), xc = X(r, 2);                                                \
      const float h1 = xa + xb + xc;                                                                       \
      const float h2 = xa * xa + xb * xb + xc * xc;                                                        \
      const float h3 = xa * yw[r][0] + xb * yw[r][1] + xc * yw[r][2];                                      \
      if (r >= 2) {                                                                                        \
        const int o = r - 2; /* output row; its window centre is strip row r-1 */                          \
        const float mx = (h1a + h1b + h1) * (1.f / 9.f), m_y = my[o];                                      \
        const float vx = (h2a + h2b + h2) * (1.f / 9.f) - mx * mx;                                         \
        const float vxy = (h3a + h3b + h3) * (1.f / 9.f) - mx * m_y;                                       \
        const float n = (2.f * mx * m_y + SSIM_C1) * (2.f * vxy + SSIM_C2);                                \
        const float d = (mx * mx + cy1[o]) * (vx + cy2[o]);                                                \
        const float ssim = __saturatef(0.5f - 0.5f * __fdividef(n, d));                                    \
        const float df = yw[r - 1][1] - xc_prev;                                                           \
        ERR[o] += (0.85f / 3.f) * ssim + (0.15f / 3.f) * fast_sqrt(df * df + 1e-6f);                       \
      }                                                                                                    \
      h1a = h1b; h1b = h1; h2a = h2b; h2b = h2; h3a = h3b; h3b = h3; xc_prev = xb;                         \
    }                                                                                                      \
  }

__global__ void __launch_bounds__(256, 3) photometric_fwd_kernel(JpbPhotoArgs a) {
  JPB_DYN_SMEM(float, sm);
  __shared__ SrcGeom geom[JPB_MAX_SRC];
  __shared__ double red[32];
  const int b = blockIdx.z, x0 = blockIdx.x * QT_W, y0 = blockIdx.y * QT_H;
  const int F = a.F, H = a.H, W = a.W;
  const int nid = a.automask ? F : 0;
  float* s_tgt = sm;                       // [3][Q1_N]
  float* s_wp = s_tgt + 3 * Q1_N;          // [F][3][Q1_N] warped sources (the identity candidates are read from global / L1)
  const size_t plane = (size_t)H * W;

  for (int f = JPB_TID; f < F; f += JPB_NT) make_geom(a.K + b * 16, a.T[f] + b * 16, geom[f]);
  __syncthreads();

  // ---- phase 1: stage target and warped pixels of the tile + 1-pixel apron (reflect-indexed at the border)
  const float* disp = a.disp + (size_t)b * a.hs * a.ws;
  const float sy = (float)a.hs / (float)H, sx = (float)a.ws / (float)W;
  const float* iK = a.invK + b * 16;
#pragma unroll 2
  for (int e = JPB_TID; e < Q1_N; e += JPB_NT) {
    const int hy = e / Q1_W, hx = e - hy * Q1_W;
    const int y = jpb_reflect(min(y0 + hy - 1, H), H), x = jpb_reflect(min(x0 + hx - 1, W), W);
    const size_t o = (size_t)y * W + x;
    const float* tg = a.target + (size_t)b * 3 * plane + o;
    s_tgt[e] = tg[0]; s_tgt[Q1_N + e] = tg[plane]; s_tgt[2 * Q1_N + e] = tg[2 * plane];
    DispTap tp;
    const float D = disp_upsample(disp, a.hs, a.ws, sy, sx, y, x, tp);
    const float z = 1.f / (a.min_disp + (a.max_disp - a.min_disp) * D);
    const bool interior = hy >= 1 && hy <= QT_H && hx >= 1 && hx <= QT_W && (y0 + hy - 1) < H && (x0 + hx - 1) < W;
    float rc[3];
    pixel_ray(iK, x, y, rc);
    for (int f = 0; f < F; ++f) {
      const float* sp = a.src[f] + (size_t)b * 3 * plane;
      Sample s;
      project(geom[f], z, rc, W, H, s);
      float v[3];
      gather3(sp, H, W, s, v);
      float* d = s_wp + f * 3 * Q1_N + e;
      d[0] = v[0]; d[Q1_N] = v[1]; d[2 * Q1_N] = v[2];
      if (interior && a.warped[f]) {
        float* wo = a.warped[f] + (size_t)b * 3 * plane + o;
        wo[0] = v[0]; wo[plane] = v[1]; wo[2 * plane] = v[2];
      }
    }
  }
  __syncthreads();

  // ---- phase 2: one (column, QSR-row strip) per thread
  float local = 0.f;
  for (int e = JPB_TID; e < QT_W * (QT_H / QSR); e += JPB_NT) {
    const int tx = e % QT_W, ry0 = (e / QT_W) * QSR;
    const int x = x0 + tx;
    if (x >= W || y0 + ry0 >= H) continue;
    float err[QNC][QSR];
#pragma unroll
    for (int k = 0; k < QNC; ++k)
#pragma unroll
      for (int r = 0; r < QSR; ++r) err[k][r] = 0.f;
    const int col = tx + 1;   // apron column of the window centre
    // global (reflected) coordinates of the strip's window rows / columns, for the identity candidates
    const int gxa = jpb_reflect(x - 1, W), gxc = jpb_reflect(min(x + 1, W), W);
    int grow[QSR + 2];
#pragma unroll
    for (int r = 0; r < QSR + 2; ++r) grow[r] = jpb_reflect(min(y0 + ry0 + r - 1, H), H) * W;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      // target window: the three horizontal neighbours of every row of the strip (+apron rows), and per output row
      // my = mean, cy1 = my^2 + C1, cy2 = var_y + C2
      float yw[QSR + 2][3], my[QSR], cy1[QSR], cy2[QSR];
      {
        const float* yp = s_tgt + c * Q1_N + ry0 * Q1_W + col;
        float h1a = 0.f, h1b = 0.f, h2a = 0.f, h2b = 0.f;
#pragma unroll
        for (int r = 0; r < QSR + 2; ++r) {
          const float ya = yp[r * Q1_W - 1], yb = yp[r * Q1_W], yc = yp[r * Q1_W + 1];
          yw[r][0] = ya; yw[r][1] = yb; yw[r][2] = yc;
          const float h1 = ya + yb + yc, h2 = ya * ya + yb * yb + yc * yc;
          if (r >= 2) {
            const float m = (h1a + h1b + h1) * (1.f / 9.f);
            my[r - 2] = m;
            cy1[r - 2] = m * m + SSIM_C1;
            cy2[r - 2] = (h2a + h2b + h2) * (1.f / 9.f) - m * m + SSIM_C2;
          }
          h1a = h1b; h1b = h1; h2a = h2b; h2b = h2;
        }
      }
      if (nid) {
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          if (f < F) {
            const float* gp = a.src[f] + ((size_t)b * 3 + c) * plane;
#define JPB_XG(r, j) __ldg(gp + grow[r] + ((j) == 0 ? gxa : ((j) == 1 ? x : gxc)))
            JPB_STRIP_CANDIDATE(JPB_XG, err[f])
#undef JPB_XG
          }
        }
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        if (f < F) {
          const float* xp = s_wp + (f * 3 + c) * Q1_N + ry0 * Q1_W + col;
#define JPB_XS(r, j) xp[(r) * Q1_W + (j) - 1]
          JPB_STRIP_CANDIDATE(JPB_XS, err[2 + f])
#undef JPB_XS
        }
      }
    }
    // ---- candidates -> min / argmin (identity terms first, with their tie-breaking noise)
#pragma unroll
    for (int r = 0; r < QSR; ++r) {
      const int y = y0 + ry0 + r;
      if (y >= H) continue;
      const size_t po = (size_t)b * plane + (size_t)y * W + x;
      float nz[2] = {0.f, 0.f};
      if (nid && !a.noise[0] && a.noise_scale != 0.f) {
        randn2(a.seed, a.stream + (a.step ? 64ull * (uint64_t)a.step[0] : 0ull), (uint64_t)po, nz[0], nz[1]);
        nz[0] *= a.noise_scale; nz[1] *= a.noise_scale;
      }
      float best = 3.0e38f;
      int besti = 0;
      if (nid) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
          if (f < F) {
            const float v = err[f][r] + (a.noise[f] ? a.noise[f][po] : nz[f]);
            if (v < best) { best = v; besti = f; }
          }
      }
#pragma unroll
      for (int f = 0; f < 2; ++f)
        if (f < F) {
          const float v = err[2 + f][r];
          if (v < best) { best = v; besti = nid + f; }
        }
      if (a.min_index) a.min_index[po] = (long long)besti;
      if (a.winner) a.winner[po] = (unsigned char)besti;
      local += best;
    }
  }
  const double tot = jpb_block_sum<double>((double)local, red);
  if (JPB_TID == 0) atomicAdd(a.loss_sum, tot);
}
#undef JPB_STRIP_CANDIDATE

// ------------------------------------------------------------------------------------ forward, packed variant (F <= 2)
// Same tile (32x32 + 1-pixel apron) and the same two phases as photometric_fwd_kernel, re-organised around what the ncu
// capture of that kernel showed (profiles/README.md: 1800 thread instructions per pixel, FMA pipe and gather latency bound):
//  * phase 1 issues every load of a staged pixel unconditionally (taps clamped instead of predicated: a clamped tap always
//    carries weight 0 because the sampling coordinate was clipped first), uses 32-bit offsets, one MUFU.RCP for the
//    perspective divide and a folded pixel->grid->pixel scale (ix = u*W/(W-1) - 0.5), and stages the identity sources too;
//  * shared memory holds the two source frames of a candidate kind interleaved as float2 (f0, f1), and phase 2 evaluates both
//    frames with one packed FADD2 / FMUL2 / FFMA2 per operation (jpb_*2 wrappers): half the FMA-pipe issue slots;
//  * phase 2 streams down the strip once per channel with all window sums in registers (no local arrays).
namespace v3 {
constexpr int TW = 32, TH = 32, SR = 4;                 // tile and strip height
constexpr int AW = TW + 2, AH = TH + 2, AN = AW * AH;   // tile + apron

// MUFU.RCP / MUFU.SQRT without the denormal pre-scaling of div.approx / sqrt.approx (arguments here are >= 1e-8)
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef JPB_HOST_EMU
  return 1.f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
__device__ __forceinline__ float fast_sqrt_ftz(float x) {
#ifdef JPB_HOST_EMU
  return sqrtf(x);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
// keep a per-sample base pointer in a 64-bit register pair so that every access is one IMAD.WIDE + load
#ifdef JPB_HOST_EMU
#define JPB_PIN_PTR(p) ((void)0)
#else
#define JPB_PIN_PTR(p) asm volatile("" : "+l"(p))
#endif

// warped sample of one staged pixel for one source frame: Project (layers.py:73-82) + grid_sample(bilinear, border);
// c0/c1/c2 are the three colour planes of the source frame
__device__ __forceinline__ void warp3(const SrcGeom& g, float X0, float X1, float X2, const float* c0, const float* c1, const float* c2,
                                      int H, int W, float cw, float ch, float v[3]) {
  const float p0 = g.P[0] * X0 + g.P[1] * X1 + g.P[2] * X2 + g.P[3];
  const float p1 = g.P[4] * X0 + g.P[5] * X1 + g.P[6] * X2 + g.P[7];
  const float p2 = g.P[8] * X0 + g.P[9] * X1 + g.P[10] * X2 + g.P[11];
  const float inv = fast_rcp(p2 + 1e-7f);
  float ix = (p0 * inv) * cw - 0.5f, iy = (p1 * inv) * ch - 0.5f;
  // clip_coordinates, NaN-safe order as ATen (min(max(x, 0), size-1))
  if (!(ix > 0.f)) ix = 0.f;
  if (ix > (float)(W - 1)) ix = (float)(W - 1);
  if (!(iy > 0.f)) iy = 0.f;
  if (iy > (float)(H - 1)) iy = (float)(H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int xa = (int)fx0, ya = (int)fy0;
  const float tx = ix - fx0, ty = iy - fy0;
  const int xb = min(xa + 1, W - 1), yb = min(ya + 1, H - 1);   // clamped taps have tx (ty) == 0
  const int o00 = ya * W + xa, o01 = ya * W + xb, o10 = yb * W + xa, o11 = yb * W + xb;
  const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
  const float a0 = __ldg(c0 + o00), a1 = __ldg(c0 + o01), a2 = __ldg(c0 + o10), a3 = __ldg(c0 + o11);
  const float b0 = __ldg(c1 + o00), b1 = __ldg(c1 + o01), b2 = __ldg(c1 + o10), b3 = __ldg(c1 + o11);
  const float d0 = __ldg(c2 + o00), d1 = __ldg(c2 + o01), d2 = __ldg(c2 + o10), d3 = __ldg(c2 + o11);
  v[0] = ((a0 * wnw + a1 * wne) + a2 * wsw) + a3 * wse;
  v[1] = ((b0 * wnw + b1 * wne) + b2 * wsw) + b3 * wse;
  v[2] = ((d0 * wnw + d1 * wne) + d2 * wsw) + d3 * wse;
}

struct PairState {   // horizontal 3-sums of the previous two rows (x, x^2, x*y) and the previous centre value, two frames each
  float2 h1a, h1b, h2a, h2b, h3a, h3b, xbp;
};
struct TargetOut {   // per output row, broadcast to both lanes
  float2 MY, TWO_MY, CY1, CY2, YBP;
};

__device__ __forceinline__ void pair_row(PairState& S, const float2 xa, const float2 xb, const float2 xc, const float2 Ya, const float2 Yb,
                                         const float2 Yc, const bool emit, const TargetOut& t, float2& err) {
  const float2 h1 = jpb_add2(jpb_add2(xa, xb), xc);
  const float2 h2 = jpb_fma2(xc, xc, jpb_fma2(xb, xb, jpb_mul2(xa, xa)));
  const float2 h3 = jpb_fma2(xc, Yc, jpb_fma2(xb, Yb, jpb_mul2(xa, Ya)));
  if (emit) {
    const float2 NINTH = jpb_dup2(1.f / 9.f), NNINTH = jpb_dup2(-1.f / 9.f);
    const float2 s1 = jpb_add2(jpb_add2(S.h1a, S.h1b), h1);
    const float2 s2 = jpb_add2(jpb_add2(S.h2a, S.h2b), h2);
    const float2 s3 = jpb_add2(jpb_add2(S.h3a, S.h3b), h3);
    const float2 mx = jpb_mul2(s1, NINTH), nmx = jpb_mul2(s1, NNINTH);
    const float2 d1 = jpb_fma2(mx, mx, t.CY1);                                   // mu_x^2 + mu_y^2 + C1
    const float2 d2 = jpb_fma2(nmx, mx, jpb_fma2(s2, NINTH, t.CY2));             // sigma_x + sigma_y + C2
    const float2 vxy = jpb_fma2(nmx, t.MY, jpb_mul2(s3, NINTH));                 // sigma_xy
    const float2 n1 = jpb_fma2(mx, t.TWO_MY, jpb_dup2(SSIM_C1));
    const float2 n2 = jpb_fma2(vxy, jpb_dup2(2.f), jpb_dup2(SSIM_C2));
    const float2 n = jpb_mul2(n1, n2), d = jpb_mul2(d1, d2);
    const float2 ssim = make_float2(__saturatef(0.5f - 0.5f * (n.x * fast_rcp(d.x))), __saturatef(0.5f - 0.5f * (n.y * fast_rcp(d.y))));
    const float2 df = jpb_fma2(S.xbp, jpb_dup2(-1.f), t.YBP);
    const float2 e2 = jpb_fma2(df, df, jpb_dup2(1e-6f));
    const float2 l1 = make_float2(fast_sqrt_ftz(e2.x), fast_sqrt_ftz(e2.y));
    err = jpb_fma2(ssim, jpb_dup2(0.85f / 3.f), err);
    err = jpb_fma2(l1, jpb_dup2(0.15f / 3.f), err);
  }
  S.h1a = S.h1b; S.h1b = h1; S.h2a = S.h2b; S.h2b = h2; S.h3a = S.h3b; S.h3b = h3; S.xbp = xb;
}

// IDENT: 0 = no identity candidates (automask off); 1 = computed here (and stored to a.ident_err when a.ident_mode == 1);
// 2 = read from a.ident_err (computed by another scale's launch of the same step: the identity terms compare the target with
// the un-warped full-resolution sources, net.py:159-166, and do not depend on the scale).  Modes 0 / 2 stage and evaluate only
// the warped pair: fewer instructions, no identity tile in shared memory, three resident CTAs per SM instead of two.
template <int IDENT>
__global__ void __launch_bounds__(256, IDENT == 1 ? 2 : 3) photometric_fwd_kernel(JpbPhotoArgs a) {
  JPB_DYN_SMEM(float, sm);
  __shared__ SrcGeom geom[2];
  __shared__ double red[32];
  const int b = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int F = a.F, H = a.H, W = a.W;
  constexpr bool ident = IDENT == 1;
  const int pl = H * W;
  float* s_tgt = sm;                                                  // [3][AN]
  float2* s_wp = reinterpret_cast<float2*>(sm + 3 * AN);              // [3][AN] warped (f0, f1)
  float2* s_id = s_wp + 3 * AN;                                       // [3][AN] identity (f0, f1), automask only
  const float* tb0 = a.target + (size_t)b * 3 * pl;
  const float* sa0 = a.src[0] + (size_t)b * 3 * pl;
  const float* sb0 = a.src[F > 1 ? 1 : 0] + (size_t)b * 3 * pl;
  const float *tb1 = tb0 + pl, *tb2 = tb1 + pl, *sa1 = sa0 + pl, *sa2 = sa1 + pl, *sb1 = sb0 + pl, *sb2 = sb1 + pl;

  for (int f = JPB_TID; f < F; f += JPB_NT) make_geom(a.K + b * 16, a.T[f] + b * 16, geom[f]);
  __syncthreads();

  // ---- phase 1: stage target, identity and warped pixels of the tile + apron (reflect-indexed at the border)
  const float* disp = a.disp + (size_t)b * a.hs * a.ws;
  const int hs = a.hs, ws = a.ws;
  const float sy = (float)hs / (float)H, sx = (float)ws / (float)W;
  const float cw = (float)W / (float)(W - 1), ch = (float)H / (float)(H - 1);
  const float drange = a.max_disp - a.min_disp, dmin = a.min_disp;
  float ik[9];
  {
    const float* iK = a.invK + b * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) { ik[3 * i] = __ldg(iK + 4 * i); ik[3 * i + 1] = __ldg(iK + 4 * i + 1); ik[3 * i + 2] = __ldg(iK + 4 * i + 2); }
  }
  float* wo0 = a.warped[0] ? a.warped[0] + (size_t)b * 3 * pl : nullptr;
  float* wo1 = (F > 1 && a.warped[1]) ? a.warped[1] + (size_t)b * 3 * pl : nullptr;
  // one register pair per colour plane: every access below is one IMAD.WIDE + LDG
  JPB_PIN_PTR(tb0); JPB_PIN_PTR(tb1); JPB_PIN_PTR(tb2); JPB_PIN_PTR(sa0); JPB_PIN_PTR(sa1); JPB_PIN_PTR(sa2);
  JPB_PIN_PTR(sb0); JPB_PIN_PTR(sb1); JPB_PIN_PTR(sb2); JPB_PIN_PTR(disp);
#pragma unroll 2
  for (int e = JPB_TID; e < AN; e += JPB_NT) {
    const int hy = e / AW, hx = e - hy * AW;
    const int y = jpb_reflect(min(y0 + hy - 1, H), H), x = jpb_reflect(min(x0 + hx - 1, W), W);
    const int o = y * W + x;
    const float t0 = __ldg(tb0 + o), t1 = __ldg(tb1 + o), t2 = __ldg(tb2 + o);
    float2 i0, i1, i2;
    if (ident) {
      i0.x = __ldg(sa0 + o); i1.x = __ldg(sa1 + o); i2.x = __ldg(sa2 + o);
      i0.y = __ldg(sb0 + o); i1.y = __ldg(sb1 + o); i2.y = __ldg(sb2 + o);
    }
    // bilinear up-sampling of disp_s (align_corners=False), disp_to_depth (layers.py:33-38)
    int uy0, uy1, ux0, ux1;
    float uly, ulx;
    up_axis(y, sy, hs, uy0, uy1, uly);
    up_axis(x, sx, ws, ux0, ux1, ulx);
    const int ur0 = uy0 * ws, ur1 = uy1 * ws;
    const float da = __ldg(disp + (ur0 + ux0)), db = __ldg(disp + (ur0 + ux1));
    const float dc = __ldg(disp + (ur1 + ux0)), de = __ldg(disp + (ur1 + ux1));
    const float D = (1.f - uly) * ((1.f - ulx) * da + ulx * db) + uly * ((1.f - ulx) * dc + ulx * de);
    const float z = fast_rcp(dmin + drange * D);
    const float fx = (float)x, fy = (float)y;
    const float X0 = z * (ik[0] * fx + ik[1] * fy + ik[2]);   // Backproject (layers.py:57-61)
    const float X1 = z * (ik[3] * fx + ik[4] * fy + ik[5]);
    const float X2 = z * (ik[6] * fx + ik[7] * fy + ik[8]);
    float v0[3], v1[3];
    warp3(geom[0], X0, X1, X2, sa0, sa1, sa2, H, W, cw, ch, v0);
    if (F > 1) warp3(geom[1], X0, X1, X2, sb0, sb1, sb2, H, W, cw, ch, v1);
    else { v1[0] = v0[0]; v1[1] = v0[1]; v1[2] = v0[2]; }
    s_tgt[e] = t0; s_tgt[AN + e] = t1; s_tgt[2 * AN + e] = t2;
    s_wp[e] = make_float2(v0[0], v1[0]); s_wp[AN + e] = make_float2(v0[1], v1[1]); s_wp[2 * AN + e] = make_float2(v0[2], v1[2]);
    if (ident) { s_id[e] = i0; s_id[AN + e] = i1; s_id[2 * AN + e] = i2; }
    const bool interior = hy >= 1 && hy <= TH && hx >= 1 && hx <= TW && (y0 + hy - 1) < H && (x0 + hx - 1) < W;
    if (interior) {
      if (wo0) { wo0[o] = v0[0]; wo0[pl + o] = v0[1]; wo0[2 * pl + o] = v0[2]; }
      if (wo1) { wo1[o] = v1[0]; wo1[pl + o] = v1[1]; wo1[2 * pl + o] = v1[2]; }
    }
  }
  __syncthreads();

  // ---- phase 2: one (column, SR-row strip) per thread, both frames of a candidate kind per packed instruction
  float local = 0.f;
  for (int e = JPB_TID; e < TW * (TH / SR); e += JPB_NT) {
    const int tx = e % TW, ry0 = (e / TW) * SR;
    const int x = x0 + tx;
    if (x >= W || y0 + ry0 >= H) continue;
    float2 errI[SR], errW[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) { errI[r] = jpb_dup2(0.f); errW[r] = jpb_dup2(0.f); }
    const int base = ry0 * AW + tx + 1;   // apron index of the strip's first window row (row above the strip), centre column
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      const float* yp = s_tgt + c * AN + base;
      const float2* wp = s_wp + c * AN + base;
      const float2* ip = s_id + c * AN + base;
      float t1a = 0.f, t1b = 0.f, t2a = 0.f, t2b = 0.f, ybp = 0.f;
      PairState SI, SW;
      SI.h1a = SI.h1b = SI.h2a = SI.h2b = SI.h3a = SI.h3b = SI.xbp = jpb_dup2(0.f);
      SW = SI;
#pragma unroll
      for (int r = 0; r < SR + 2; ++r) {
        const float ya = yp[r * AW - 1], yb = yp[r * AW], yc = yp[r * AW + 1];
        const float t1 = (ya + yb) + yc, t2 = yc * yc + (yb * yb + ya * ya);
        TargetOut t;
        if (r >= 2) {
          const float m = ((t1a + t1b) + t1) * (1.f / 9.f);
          const float mm = m * m;
          t.MY = jpb_dup2(m); t.TWO_MY = jpb_dup2(2.f * m);
          t.CY1 = jpb_dup2(mm + SSIM_C1);
          t.CY2 = jpb_dup2((((t2a + t2b) + t2) * (1.f / 9.f) - mm) + SSIM_C2);
          t.YBP = jpb_dup2(ybp);
        }
        const float2 Ya = jpb_dup2(ya), Yb = jpb_dup2(yb), Yc = jpb_dup2(yc);
        if (ident) pair_row(SI, ip[r * AW - 1], ip[r * AW], ip[r * AW + 1], Ya, Yb, Yc, r >= 2, t, errI[r >= 2 ? r - 2 : 0]);
        pair_row(SW, wp[r * AW - 1], wp[r * AW], wp[r * AW + 1], Ya, Yb, Yc, r >= 2, t, errW[r >= 2 ? r - 2 : 0]);
        t1a = t1b; t1b = t1; t2a = t2b; t2b = t2; ybp = yb;
      }
    }
    // ---- candidates -> min / argmin (identity terms first, with their tie-breaking noise)
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      const int y = y0 + ry0 + r;
      if (y >= H) continue;
      const size_t po = (size_t)b * pl + (size_t)y * W + x;
      float best = 3.0e38f;
      int besti = 0;
      int nid = 0;
      if (IDENT != 0) {
        nid = F;
        if (IDENT == 2) errI[r] = __ldg(reinterpret_cast<const float2*>(a.ident_err) + po);
        else if (a.ident_mode == 1) reinterpret_cast<float2*>(a.ident_err)[po] = errI[r];
        float nz0 = 0.f, nz1 = 0.f;
        if (!a.noise[0] && a.noise_scale != 0.f) {
          randn2(a.seed, a.stream + (a.step ? 64ull * (uint64_t)a.step[0] : 0ull), (uint64_t)po, nz0, nz1);
          nz0 *= a.noise_scale; nz1 *= a.noise_scale;
        }
        const float v0 = errI[r].x + (a.noise[0] ? a.noise[0][po] : nz0);
        if (v0 < best) { best = v0; besti = 0; }
        if (F > 1) {
          const float v1 = errI[r].y + (a.noise[1] ? a.noise[1][po] : nz1);
          if (v1 < best) { best = v1; besti = 1; }
        }
      }
      if (errW[r].x < best) { best = errW[r].x; besti = nid; }
      if (F > 1 && errW[r].y < best) { best = errW[r].y; besti = nid + 1; }
      if (a.min_index) a.min_index[po] = (long long)besti;
      if (a.winner) a.winner[po] = (unsigned char)besti;
      local += best;
    }
  }
  const double tot = jpb_block_sum<double>((double)local, red);
  if (JPB_TID == 0) atomicAdd(a.loss_sum, tot);
}
}  // namespace v3

// ------------------------------------------------------------------------------------ backward
// d(loss_s)/d(disp_s) and d(loss_s)/d(T_f).  Gradient reaches a warped candidate only where it is the
// arg-min; each warped pixel feeds the (up to) 9 SSIM windows around it, with multiplicity 2 where the
// reflect padding maps two taps of a border window onto the same pixel.
__global__ void __launch_bounds__(256) photometric_bwd_kernel(JpbPhotoArgs a, JpbPhotoGrad g) {
  JPB_DYN_SMEM(float, sm);
  __shared__ SrcGeom geom[JPB_MAX_SRC];
  __shared__ float red[32];
  __shared__ float s_dd[(PT_H + 4) * (PT_W + 4)];  // disparity-gradient footprint of the tile (factor >= 1... <=2 px/px)
  const int b = blockIdx.z, x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const int F = a.F, H = a.H, W = a.W;
  const int nid = a.automask ? F : 0;
  float* s_tgt = sm;                      // [3][P2_N]
  float* s_wp = s_tgt + 3 * P2_N;         // [F][3][P2_N]
  float* s_coef = s_wp + F * 3 * P2_N;    // [9][P1_N]  (alpha,beta,gamma) x 3 channels of the winning window
  int* s_sel = reinterpret_cast<int*>(s_coef + 9 * P1_N);  // [P1_N] winning warped source of window q, or -1
  const size_t plane = (size_t)H * W;
  const float gpix = g.grad_out[0] * g.inv_count;

  for (int f = JPB_TID; f < F; f += JPB_NT) make_geom(a.K + b * 16, a.T[f] + b * 16, geom[f]);
  const float* disp = a.disp + (size_t)b * a.hs * a.ws;
  const float sy = (float)a.hs / (float)H, sx = (float)a.ws / (float)W;
  // disparity footprint of this tile
  int fy0, fx0, tmp; float tl;
  up_axis(min(y0, H - 1), sy, a.hs, fy0, tmp, tl);
  up_axis(min(x0, W - 1), sx, a.ws, fx0, tmp, tl);
  const int FW = PT_W + 4;
  for (int e = JPB_TID; e < (PT_H + 4) * (PT_W + 4); e += JPB_NT) s_dd[e] = 0.f;
  __syncthreads();

  // phase A: target + warped sources on the 2-pixel apron
  for (int e = JPB_TID; e < P2_N; e += JPB_NT) {
    const int hy = e / P2_W, hx = e - hy * P2_W;
    const int y = jpb_reflect(jpb_clampi(y0 + hy - 2, -(H - 1), H), H), x = jpb_reflect(jpb_clampi(x0 + hx - 2, -(W - 1), W), W);
    const size_t o = (size_t)y * W + x;
    const float* tg = a.target + (size_t)b * 3 * plane + o;
    s_tgt[e] = tg[0]; s_tgt[P2_N + e] = tg[plane]; s_tgt[2 * P2_N + e] = tg[2 * plane];
    DispTap tp;
    const float D = disp_upsample(disp, a.hs, a.ws, sy, sx, y, x, tp);
    const float z = 1.f / (a.min_disp + (a.max_disp - a.min_disp) * D);
    float rc[3];
    pixel_ray(a.invK + b * 16, x, y, rc);
    for (int f = 0; f < F; ++f) {
      float v[3];
      if (a.warped[f]) {   // the forward launch of this scale kept its warped frame (outputs[("color",f,s)]): stage it as it is
        const float* wv = a.warped[f] + (size_t)b * 3 * plane + o;
        v[0] = wv[0]; v[1] = wv[plane]; v[2] = wv[2 * plane];
      } else {
        Sample s;
        project(geom[f], z, rc, W, H, s);
        gather3(a.src[f] + (size_t)b * 3 * plane, H, W, s, v);
      }
      float* d = s_wp + f * 3 * P2_N + e;
      d[0] = v[0]; d[P2_N] = v[1]; d[2 * P2_N] = v[2];
    }
  }
  __syncthreads();

  // phase B: per SSIM window q (tile + 1 apron): linear form of d(err_q)/d(x_i) = alpha + beta*x_i + gamma*y_i
  for (int e = JPB_TID; e < P1_N; e += JPB_NT) {
    const int hy = e / P1_W, hx = e - hy * P1_W;
    const int y = y0 + hy - 1, x = x0 + hx - 1;
    int sel = -1;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const int w = (int)g.winner[(size_t)b * plane + (size_t)y * W + x];
      if (w >= nid) sel = w - nid;
    }
    s_sel[e] = sel;
    if (sel < 0) continue;
    const int ctr = (hy + 1) * P2_W + hx + 1;
    for (int c = 0; c < 3; ++c) {
      const float* xp = s_wp + (sel * 3 + c) * P2_N + ctr;
      const float* yp = s_tgt + c * P2_N + ctr;
      float sxv = 0.f, syv = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const float xv = xp[dy * P2_W + dx], yv = yp[dy * P2_W + dx];
          sxv += xv; syv += yv; sxx += xv * xv; syy += yv * yv; sxy += xv * yv;
        }
      const float ma = sxv * (1.f / 9.f), mb = syv * (1.f / 9.f);
      const float va = sxx * (1.f / 9.f) - ma * ma, vb = syy * (1.f / 9.f) - mb * mb, vab = sxy * (1.f / 9.f) - ma * mb;
      const float n1 = 2.f * ma * mb + SSIM_C1, n2 = 2.f * vab + SSIM_C2;
      const float d1 = ma * ma + mb * mb + SSIM_C1, d2 = va + vb + SSIM_C2;
      const float inv = 1.f / (d1 * d2);
      const float S = (1.f - n1 * n2 * inv) * 0.5f;
      float al = 0.f, be = 0.f, ga = 0.f;
      if (S >= 0.f && S <= 1.f) {
        // dS/dx_i = -(1/2) * (2/9) * [ (mb*n2 + n1*(y_i-mb))*inv - n1*n2*inv^2*(ma*d2 + d1*(x_i-ma)) ]
        const float k = -(1.f / 9.f) * (0.85f / 3.f) * gpix;
        const float q = n1 * n2 * inv * inv;
        al = k * ((mb * n2 - n1 * mb) * inv - q * (ma * d2 - d1 * ma));
        be = k * (-q * d1);
        ga = k * (n1 * inv);
      }
      s_coef[(c * 3 + 0) * P1_N + e] = al;
      s_coef[(c * 3 + 1) * P1_N + e] = be;
      s_coef[(c * 3 + 2) * P1_N + e] = ga;
    }
  }
  __syncthreads();

  // phase C: gather window contributions per pixel, back through grid_sample / projection / up-sampling
  float G[JPB_MAX_SRC][12];
  for (int f = 0; f < JPB_MAX_SRC; ++f)
    for (int i = 0; i < 12; ++i) G[f][i] = 0.f;
  const float l1k = (0.15f / 3.f) * gpix;
  for (int e = JPB_TID; e < PT_W * PT_H; e += JPB_NT) {
    const int ty = e / PT_W, tx = e - ty * PT_W;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int c1 = (ty + 1) * P1_W + tx + 1, c2 = (ty + 2) * P2_W + tx + 2;
    float gw[JPB_MAX_SRC][3];
    bool any[JPB_MAX_SRC];
    for (int f = 0; f < F; ++f) { gw[f][0] = gw[f][1] = gw[f][2] = 0.f; any[f] = false; }
    for (int dy = -1; dy <= 1; ++dy) {
      const int qy = y + dy;
      if (qy < 0 || qy >= H) continue;
      const float mrow = ((qy == 0 && y == 1) || (qy == H - 1 && y == H - 2)) ? 2.f : 1.f;
      for (int dx = -1; dx <= 1; ++dx) {
        const int qx = x + dx;
        if (qx < 0 || qx >= W) continue;
        const int qe = c1 + dy * P1_W + dx;
        const int f = s_sel[qe];
        if (f < 0) continue;
        const float m = mrow * (((qx == 0 && x == 1) || (qx == W - 1 && x == W - 2)) ? 2.f : 1.f);
        any[f] = true;
        for (int c = 0; c < 3; ++c) {
          const float xv = s_wp[(f * 3 + c) * P2_N + c2], yv = s_tgt[c * P2_N + c2];
          gw[f][c] += m * (s_coef[(c * 3 + 0) * P1_N + qe] + s_coef[(c * 3 + 1) * P1_N + qe] * xv + s_coef[(c * 3 + 2) * P1_N + qe] * yv);
        }
      }
    }
    {
      const int f = s_sel[c1];
      if (f >= 0)
        for (int c = 0; c < 3; ++c) {
          const float df = s_wp[(f * 3 + c) * P2_N + c2] - s_tgt[c * P2_N + c2];
          gw[f][c] += l1k * df / sqrtf(df * df + 1e-6f);
        }
    }
    DispTap tp;
    const float D = disp_upsample(disp, a.hs, a.ws, sy, sx, y, x, tp);
    const float sd = a.min_disp + (a.max_disp - a.min_disp) * D;
    const float z = 1.f / sd;
    float gz = 0.f;
    float rc[3];
    pixel_ray(a.invK + b * 16, x, y, rc);
    const float X0 = z * rc[0], X1 = z * rc[1], X2 = z * rc[2];
    for (int f = 0; f < F; ++f) {
      if (!any[f]) continue;
      Sample s;
      project(geom[f], z, rc, W, H, s);
      // d warped_c / d ix, d iy
      const float* img = a.src[f] + (size_t)b * 3 * plane;
      const int x1 = s.x0 + 1, y1 = s.y0 + 1;
      const bool xin = x1 < W, yin = y1 < H;
      const size_t o00 = (size_t)s.y0 * W + s.x0;
      float gix = 0.f, giy = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float* p = img + c * plane;
        const float nw = p[o00], ne = xin ? p[o00 + 1] : 0.f, sw = yin ? p[o00 + W] : 0.f, se = (xin && yin) ? p[o00 + W + 1] : 0.f;
        gix += gw[f][c] * ((ne - nw) * (1.f - s.ty) + (se - sw) * s.ty);
        giy += gw[f][c] * ((sw - nw) * (1.f - s.tx) + (se - ne) * s.tx);
      }
      // ix = u*W/(W-1) - 0.5  (clip mask mx), u = px/(pz+eps)
      const float gu = gix * s.mx * ((float)W / (float)(W - 1)), gv = giy * s.my * ((float)H / (float)(H - 1));
      const float den = s.p[2] + 1e-7f, iden = 1.f / den;
      const float gp0 = gu * iden, gp1 = gv * iden;
      const float gp2 = -(gu * s.p[0] + gv * s.p[1]) * iden * iden;
      const float* P = geom[f].P;
      gz += gp0 * (P[0] * rc[0] + P[1] * rc[1] + P[2] * rc[2]) + gp1 * (P[4] * rc[0] + P[5] * rc[1] + P[6] * rc[2]) +
            gp2 * (P[8] * rc[0] + P[9] * rc[1] + P[10] * rc[2]);
      // G[i][j] += gp_i * Xh_j with Xh = (z*rc, 1)
      const float gp[3] = {gp0, gp1, gp2};
      for (int i = 0; i < 3; ++i) {
        G[f][i * 4 + 0] += gp[i] * X0; G[f][i * 4 + 1] += gp[i] * X1;
        G[f][i * 4 + 2] += gp[i] * X2; G[f][i * 4 + 3] += gp[i];
      }
    }
    if (gz != 0.f) {
      const float gD = gz * (-(a.max_disp - a.min_disp) * z * z);
      const int r0 = (tp.y0 - fy0) * FW - fx0, r1 = (tp.y1 - fy0) * FW - fx0;
      atomicAdd(&s_dd[r0 + tp.x0], gD * (1.f - tp.ly) * (1.f - tp.lx));
      atomicAdd(&s_dd[r0 + tp.x1], gD * (1.f - tp.ly) * tp.lx);
      atomicAdd(&s_dd[r1 + tp.x0], gD * tp.ly * (1.f - tp.lx));
      atomicAdd(&s_dd[r1 + tp.x1], gD * tp.ly * tp.lx);
    }
  }
  __syncthreads();
  float* gd = g.grad_disp + (size_t)b * a.hs * a.ws;
  for (int e = JPB_TID; e < (PT_H + 4) * (PT_W + 4); e += JPB_NT) {
    const float v = s_dd[e];
    if (v == 0.f) continue;
    const int fy = fy0 + e / FW, fx = fx0 + e % FW;
    if (fy < a.hs && fx < a.ws) atomicAdd(&gd[fy * a.ws + fx], v);
  }
  // pose gradient: dT[k][j] = sum_i K[i][k] * G[i][j]
  for (int f = 0; f < F; ++f) {
    if (!g.grad_T[f]) continue;
    float tot[12];
    for (int i = 0; i < 12; ++i) tot[i] = jpb_block_sum<float>(G[f][i], red);
    if (JPB_TID == 0) {
      const float* K = a.K + b * 16;
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) {
          const float v = K[0 * 4 + k] * tot[0 * 4 + j] + K[1 * 4 + k] * tot[1 * 4 + j] + K[2 * 4 + k] * tot[2 * 4 + j];
          if (v != 0.f) atomicAdd(&g.grad_T[f][b * 16 + k * 4 + j], v);
        }
    }
  }
}


// ------------------------------------------------------------------------------------ backward, F <= 2 (every reference configuration)
// Same mathematics as photometric_bwd_kernel, re-organised around its ncu capture (profiles/r2_ncu_photo_bwd.csv: 108 M warp
// instructions per launch at 17 of 32 lanes active, pose accumulators in local memory, 48 block barriers for the 24 pose sums,
// shared-memory float atomics that serialise 16-fold on the coarse scales):
//  * F is a template parameter: every per-frame array lives in registers and the frame loops unroll;
//  * the two warped frames are staged interleaved as float2, the window coefficients (alpha, beta, gamma) of a window's
//    winning frame as one float4 per channel (w of channel 0 = the winning frame or -1): phase C reads 3 LDS.128 per window;
//  * the transposed bilinear up-sampling of the disparity gradient is separable and gather-based — per-pixel d(loss)/d(D)
//    goes to shared memory, then (row, texel column) sums, then (texel row, texel column) sums: no shared-memory atomics and one
//    global atomic per texel of the tile's footprint;
//  * the 12 F pose sums are reduced with warp shuffles and ONE shared-memory transposition (two barriers in all).
namespace v4 {
constexpr int TW = 32, TH = 16;
constexpr int W2 = TW + 4, H2 = TH + 4, N2 = W2 * H2;   // tile + 2-pixel apron: staged pixels
constexpr int W1 = TW + 2, H1 = TH + 2, N1 = W1 * H1;   // tile + 1-pixel apron: SSIM windows
constexpr int FJ = TW + 2, FI = TH + 2;                 // largest disparity footprint of a tile (up-sampling factor >= 1)
constexpr int NWARP_MAX = 8;
constexpr int SMEM_FLOATS = 4 * 3 * N1 + 3 * N2 + 2 * 3 * N2 + TW * TH + TH * FJ + NWARP_MAX * 24 + 24 + 4;

// weight with which pixel coordinate `p` of the up-sampled axis reads texel `t` (both taps may coincide at the far border)
__device__ __forceinline__ float up_weight(int p, float scale, int in_size, int t) {
  int i0, i1;
  float l1;
  up_axis(p, scale, in_size, i0, i1, l1);
  return (i0 == t ? 1.f - l1 : 0.f) + (i1 == t ? l1 : 0.f);
}

template <int F>
__global__ void __launch_bounds__(256, 3) photometric_bwd_kernel(JpbPhotoArgs a, JpbPhotoGrad g) {
  JPB_DYN_SMEM(float, sm);
  __shared__ SrcGeom geom[2];
  const int b = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int H = a.H, W = a.W, hs = a.hs, ws = a.ws;
  const int nid = a.automask ? F : 0;
  float4* s_cf = reinterpret_cast<float4*>(sm);                 // [3][N1] (alpha, beta, gamma, sel) of the window's winning frame
  float* s_tgt = sm + 4 * 3 * N1;                               // [3][N2]
  float2* s_wp = reinterpret_cast<float2*>(s_tgt + 3 * N2);     // [3][N2] warped (f0, f1)
  float* s_gd = s_tgt + 3 * N2 + 2 * 3 * N2;                    // [TH][TW] d(loss)/d(up-sampled disparity)
  float* s_h = s_gd + TW * TH;                                  // [TH][FJ] horizontally gathered
  float* s_red = s_h + TH * FJ;                                 // [NWARP_MAX][24] + [24]
  const int pl = H * W;
  const float gpix = g.grad_out[0] * g.inv_count;

  for (int f = JPB_TID; f < F; f += JPB_NT) make_geom(a.K + b * 16, a.T[f] + b * 16, geom[f]);
  const float* disp = a.disp + (size_t)b * hs * ws;
  const float sy = (float)hs / (float)H, sx = (float)ws / (float)W;
  const float* tg0 = a.target + (size_t)b * 3 * pl;
  const float* sp0 = a.src[0] + (size_t)b * 3 * pl;
  const float* sp1 = a.src[F > 1 ? 1 : 0] + (size_t)b * 3 * pl;
  const float* wk0 = a.warped[0] ? a.warped[0] + (size_t)b * 3 * pl : nullptr;
  const float* wk1 = (F > 1 && a.warped[1]) ? a.warped[1] + (size_t)b * 3 * pl : nullptr;
  const bool kept = wk0 != nullptr && (F == 1 || wk1 != nullptr);
  if (!kept) __syncthreads();   // geom is read in phase A only when the apron is re-projected

  // ---- phase A: target + warped frames on the 2-pixel apron
  for (int e = JPB_TID; e < N2; e += JPB_NT) {
    const int hy = e / W2, hx = e - hy * W2;
    const int y = jpb_reflect(jpb_clampi(y0 + hy - 2, -(H - 1), H), H), x = jpb_reflect(jpb_clampi(x0 + hx - 2, -(W - 1), W), W);
    const int o = y * W + x;
    s_tgt[e] = __ldg(tg0 + o); s_tgt[N2 + e] = __ldg(tg0 + pl + o); s_tgt[2 * N2 + e] = __ldg(tg0 + 2 * pl + o);
    float v0[3], v1[3];
    if (kept) {   // the forward launch of this scale kept its warped frames (outputs[("color",f,s)]): stage them as they are
      v0[0] = __ldg(wk0 + o); v0[1] = __ldg(wk0 + pl + o); v0[2] = __ldg(wk0 + 2 * pl + o);
      if (F > 1) { v1[0] = __ldg(wk1 + o); v1[1] = __ldg(wk1 + pl + o); v1[2] = __ldg(wk1 + 2 * pl + o); }
    } else {
      DispTap tp;
      const float D = disp_upsample(disp, hs, ws, sy, sx, y, x, tp);
      const float z = 1.f / (a.min_disp + (a.max_disp - a.min_disp) * D);
      float rc[3];
      pixel_ray(a.invK + b * 16, x, y, rc);
      Sample s;
      project(geom[0], z, rc, W, H, s);
      gather3(sp0, H, W, s, v0);
      if (F > 1) {
        project(geom[1], z, rc, W, H, s);
        gather3(sp1, H, W, s, v1);
      }
    }
    if (F == 1) { v1[0] = v0[0]; v1[1] = v0[1]; v1[2] = v0[2]; }
    s_wp[e] = make_float2(v0[0], v1[0]); s_wp[N2 + e] = make_float2(v0[1], v1[1]); s_wp[2 * N2 + e] = make_float2(v0[2], v1[2]);
  }
  __syncthreads();

  // ---- phase B: per SSIM window q (tile + 1 apron), for its winning warped frame: d(err_q)/d(x_i) = alpha + beta*x_i + gamma*y_i.
  // Only windows whose arg-min is a warped frame do any work (a quarter to a half of them): they are first COMPACTED into a list
  // (shared-memory counter) so that the coefficient pass runs with full warps — the ncu source view of the uncompacted version
  // showed 7.5 of 32 lanes active here.
  int* s_cnt = reinterpret_cast<int*>(s_red + NWARP_MAX * 24 + 24);     // [4] counters: windows, frame-0 pixels, frame-1 pixels
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_gd);     // window list (<= N1 entries of 2 bytes) in the not yet used s_gd / s_h
  for (int i = JPB_TID; i < 4; i += JPB_NT) s_cnt[i] = 0;
  __syncthreads();
  for (int e = JPB_TID; e < N1; e += JPB_NT) {
    const int hy = e / W1, hx = e - hy * W1;
    const int y = y0 + hy - 1, x = x0 + hx - 1;
    int sel = -1;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const int w = (int)g.winner[(size_t)b * pl + (size_t)y * W + x];
      if (w >= nid && w - nid < F) sel = w - nid;
    }
    if (sel < 0) {
      s_cf[e] = make_float4(0.f, 0.f, 0.f, -1.f);
      s_cf[N1 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
      s_cf[2 * N1 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      s_list[atomicAdd(&s_cnt[0], 1)] = (unsigned short)(e | (sel << 15));
    }
  }
  __syncthreads();
  const int nwin = s_cnt[0];
  for (int li = JPB_TID; li < nwin; li += JPB_NT) {
    const int ent = (int)s_list[li];
    const int e = ent & 0x7fff, sel = ent >> 15;
    const int hy = e / W1, hx = e - hy * W1;
    const int ctr = (hy + 1) * W2 + hx + 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2* xp = s_wp + c * N2 + ctr;
      const float* yp = s_tgt + c * N2 + ctr;
      float sxv = 0.f, syv = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const float2 xw = xp[dy * W2 + dx];
          const float xv = sel ? xw.y : xw.x, yv = yp[dy * W2 + dx];
          sxv += xv; syv += yv; sxx += xv * xv; syy += yv * yv; sxy += xv * yv;
        }
      const float ma = sxv * (1.f / 9.f), mb = syv * (1.f / 9.f);
      const float va = sxx * (1.f / 9.f) - ma * ma, vb = syy * (1.f / 9.f) - mb * mb, vab = sxy * (1.f / 9.f) - ma * mb;
      const float n1 = 2.f * ma * mb + SSIM_C1, n2 = 2.f * vab + SSIM_C2;
      const float d1 = ma * ma + mb * mb + SSIM_C1, d2 = va + vb + SSIM_C2;
      const float inv = 1.f / (d1 * d2);
      const float S = (1.f - n1 * n2 * inv) * 0.5f;
      float al = 0.f, be = 0.f, ga = 0.f;
      if (S >= 0.f && S <= 1.f) {
        // dS/dx_i = -(1/2) * (2/9) * [ (mb*n2 + n1*(y_i-mb))*inv - n1*n2*inv^2*(ma*d2 + d1*(x_i-ma)) ]
        const float k = -(1.f / 9.f) * (0.85f / 3.f) * gpix;
        const float q = n1 * n2 * inv * inv;
        al = k * ((mb * n2 - n1 * mb) * inv - q * (ma * d2 - d1 * ma));
        be = k * (-q * d1);
        ga = k * (n1 * inv);
      }
      s_cf[c * N1 + e] = make_float4(al, be, ga, c == 0 ? (float)sel : 0.f);
    }
  }
  __syncthreads();

  // ---- phase C1: window contributions per pixel -> d(loss)/d(warped frame f, channel c) at the pixel, branch-free over the 3x3
  // windows (the coefficients of a window without a warped winner are zero).  Each thread keeps the result of its (at most two)
  // pixels in registers until every thread has finished reading the coefficient tile, which then becomes the work list of C2.
#if defined(JPB_HOST_EMU) && !defined(JPB_HOST_EMU_MT)
  constexpr int PPT = TW * TH;                    // host emulation with one "thread" per block: that thread owns every pixel
#else
  constexpr int PPT = (TW * TH + 255) / 256;      // pixels per thread at 256 threads
#endif
  float gwr[PPT][2][3];
  const float l1k = (0.15f / 3.f) * gpix;
  const float drange = a.max_disp - a.min_disp;
#pragma unroll
  for (int it = 0; it < PPT; ++it) {
    const int e = JPB_TID + it * JPB_NT;
#pragma unroll
    for (int f = 0; f < 2; ++f) gwr[it][f][0] = gwr[it][f][1] = gwr[it][f][2] = 0.f;
    if (e >= TW * TH) continue;
    const int ty = e / TW, tx = e - ty * TW;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int c1 = (ty + 1) * W1 + tx + 1, c2 = (ty + 2) * W2 + tx + 2;
    const float2 xw0 = s_wp[c2], xw1 = s_wp[N2 + c2], xw2 = s_wp[2 * N2 + c2];
    const float yv0 = s_tgt[c2], yv1 = s_tgt[N2 + c2], yv2 = s_tgt[2 * N2 + c2];
    float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int qy = y + dy;
      // a window outside the image has no entry in the loss: weight 0 (its coefficients are zero as well)
      const float mrow = (qy < 0 || qy >= H) ? 0.f : (((qy == 0 && y == 1) || (qy == H - 1 && y == H - 2)) ? 2.f : 1.f);
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int qx = x + dx;
        const float m = (qx < 0 || qx >= W) ? 0.f : mrow * (((qx == 0 && x == 1) || (qx == W - 1 && x == W - 2)) ? 2.f : 1.f);
        const int qe = c1 + dy * W1 + dx;
        const float4 k0 = s_cf[qe], k1 = s_cf[N1 + qe], k2 = s_cf[2 * N1 + qe];
        const bool second = k0.w > 0.5f;
        const float t0 = m * (k0.x + k0.y * (second ? xw0.y : xw0.x) + k0.z * yv0);
        const float t1 = m * (k1.x + k1.y * (second ? xw1.y : xw1.x) + k1.z * yv1);
        const float t2 = m * (k2.x + k2.y * (second ? xw2.y : xw2.x) + k2.z * yv2);
        g0[0] += second ? 0.f : t0; g0[1] += second ? 0.f : t1; g0[2] += second ? 0.f : t2;
        g1[0] += second ? t0 : 0.f; g1[1] += second ? t1 : 0.f; g1[2] += second ? t2 : 0.f;
      }
    }
    {
      const float selp = s_cf[c1].w;
      if (selp >= 0.f) {
        const bool second = selp > 0.5f;
        const float d0 = (second ? xw0.y : xw0.x) - yv0, d1 = (second ? xw1.y : xw1.x) - yv1, d2 = (second ? xw2.y : xw2.x) - yv2;
        const float a0 = l1k * d0 / sqrtf(d0 * d0 + 1e-6f), a1 = l1k * d1 / sqrtf(d1 * d1 + 1e-6f), a2 = l1k * d2 / sqrtf(d2 * d2 + 1e-6f);
        if (second) { g1[0] += a0; g1[1] += a1; g1[2] += a2; }
        else { g0[0] += a0; g0[1] += a1; g0[2] += a2; }
      }
    }
    gwr[it][0][0] = g0[0]; gwr[it][0][1] = g0[1]; gwr[it][0][2] = g0[2];
    gwr[it][1][0] = g1[0]; gwr[it][1][1] = g1[1]; gwr[it][1][2] = g1[2];
  }
  __syncthreads();            // every thread is done with the coefficient tile

  // ---- work lists of phase C2 in the coefficient tile's memory: per frame f, entries (pixel, gw0, gw1, gw2) of the pixels that
  // carry a gradient for that frame (compaction again: the projection / gather / pose part is ~250 instructions per entry and
  // ran at 12 of 32 lanes when every pixel walked through it under a branch)
  float4* s_work = s_cf;                                  // [2][TW * TH]
  for (int e = JPB_TID; e < TW * TH; e += JPB_NT) s_gd[e] = 0.f;
#pragma unroll
  for (int it = 0; it < PPT; ++it) {
    const int e = JPB_TID + it * JPB_NT;
#pragma unroll
    for (int f = 0; f < F; ++f)
      if (gwr[it][f][0] != 0.f || gwr[it][f][1] != 0.f || gwr[it][f][2] != 0.f)
        s_work[f * (TW * TH) + atomicAdd(&s_cnt[1 + f], 1)] = make_float4(__uint_as_float((unsigned)e), gwr[it][f][0], gwr[it][f][1], gwr[it][f][2]);
  }
  __syncthreads();

  // ---- phase C2: back through grid_sample / projection / disparity, one dense pass per frame
  float G[F][12];
#pragma unroll
  for (int f = 0; f < F; ++f)
#pragma unroll
    for (int i = 0; i < 12; ++i) G[f][i] = 0.f;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const int nwork = s_cnt[1 + f];
    const float* img = f ? sp1 : sp0;
    for (int li = JPB_TID; li < nwork; li += JPB_NT) {
      const float4 wk = s_work[f * (TW * TH) + li];
      const int e = (int)__float_as_uint(wk.x);
      const int ty = e / TW, tx = e - ty * TW;
      const int y = y0 + ty, x = x0 + tx;
      DispTap tp;
      const float D = disp_upsample(disp, hs, ws, sy, sx, y, x, tp);
      const float z = 1.f / (a.min_disp + drange * D);
      float rc[3];
      pixel_ray(a.invK + b * 16, x, y, rc);
      const float X0 = z * rc[0], X1 = z * rc[1], X2 = z * rc[2];
      Sample s;
      project(geom[f], z, rc, W, H, s);
      const int x1 = s.x0 + 1, y1 = s.y0 + 1;
      const bool xin = x1 < W, yin = y1 < H;
      const int o00 = s.y0 * W + s.x0;
      const float gwc[3] = {wk.y, wk.z, wk.w};
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p = img + c * pl;
        const float nw = __ldg(p + o00), ne = xin ? __ldg(p + o00 + 1) : 0.f, sw = yin ? __ldg(p + o00 + W) : 0.f,
                    se = (xin && yin) ? __ldg(p + o00 + W + 1) : 0.f;
        gix += gwc[c] * ((ne - nw) * (1.f - s.ty) + (se - sw) * s.ty);
        giy += gwc[c] * ((sw - nw) * (1.f - s.tx) + (se - ne) * s.tx);
      }
      // ix = u*W/(W-1) - 0.5  (clip mask mx), u = px/(pz+eps)
      const float gu = gix * s.mx * ((float)W / (float)(W - 1)), gv = giy * s.my * ((float)H / (float)(H - 1));
      const float iden = 1.f / (s.p[2] + 1e-7f);
      const float gp0 = gu * iden, gp1 = gv * iden;
      const float gp2 = -(gu * s.p[0] + gv * s.p[1]) * iden * iden;
      const float* P = geom[f].P;
      const float gz = gp0 * (P[0] * rc[0] + P[1] * rc[1] + P[2] * rc[2]) + gp1 * (P[4] * rc[0] + P[5] * rc[1] + P[6] * rc[2]) +
                       gp2 * (P[8] * rc[0] + P[9] * rc[1] + P[10] * rc[2]);
      // G[i][j] += gp_i * Xh_j with Xh = (z*rc, 1)
      G[f][0] += gp0 * X0; G[f][1] += gp0 * X1; G[f][2] += gp0 * X2; G[f][3] += gp0;
      G[f][4] += gp1 * X0; G[f][5] += gp1 * X1; G[f][6] += gp1 * X2; G[f][7] += gp1;
      G[f][8] += gp2 * X0; G[f][9] += gp2 * X1; G[f][10] += gp2 * X2; G[f][11] += gp2;
      const float gD = gz * (-drange * z * z);
      if (F == 1) s_gd[e] = gD;                 // one entry per pixel
      else atomicAdd(&s_gd[e], gD);             // at most two entries (one per frame) meet on a pixel
    }
  }
  __syncthreads();

  // ---- phase D: transposed bilinear up-sampling of gD, separable, gather form.  Footprint of the tile in disp_s:
  int fy0, fx0, fy1, fx1, tmp;
  float tl;
  up_axis(min(y0, H - 1), sy, hs, fy0, tmp, tl);
  up_axis(min(x0, W - 1), sx, ws, fx0, tmp, tl);
  up_axis(min(y0 + TH - 1, H - 1), sy, hs, tmp, fy1, tl);
  up_axis(min(x0 + TW - 1, W - 1), sx, ws, tmp, fx1, tl);
  const int nj = min(fx1 - fx0 + 1, FJ), ni = min(fy1 - fy0 + 1, FI);
  const float rx = (float)W / (float)ws, ry = (float)H / (float)hs;
  const int xe = min(TW, W - x0), ye = min(TH, H - y0);          // valid pixels of the tile
  // per tile column / row: the two taps and the weight of the second one, computed once (the first version recomputed them
  // for every (row, texel, pixel) candidate: 15 % of the kernel's instructions)
  int* s_tapx = reinterpret_cast<int*>(s_work);                    // [TW] i0 | i1 << 16
  float* s_lx = reinterpret_cast<float*>(s_tapx + TW);            // [TW]
  int* s_tapy = reinterpret_cast<int*>(s_lx + TW);                // [TH]
  float* s_ly = reinterpret_cast<float*>(s_tapy + TH);            // [TH]
  for (int e = JPB_TID; e < TW + TH; e += JPB_NT) {
    int i0, i1;
    float l1;
    if (e < TW) { up_axis(min(x0 + e, W - 1), sx, ws, i0, i1, l1); s_tapx[e] = i0 | (i1 << 16); s_lx[e] = l1; }
    else { up_axis(min(y0 + e - TW, H - 1), sy, hs, i0, i1, l1); s_tapy[e - TW] = i0 | (i1 << 16); s_ly[e - TW] = l1; }
  }
  __syncthreads();
  for (int e = JPB_TID; e < TH * nj; e += JPB_NT) {
    const int ty = e / nj, j = e - ty * nj, t = fx0 + j;
    // pixels whose taps can touch texel t: source coordinate in [t - 1, t + 1)  ->  x + 0.5 in [(t - 0.5) rx, (t + 1.5) rx)
    const int lo = t == 0 ? 0 : max((int)floorf(((float)t - 0.5f) * rx - 0.5f) - 1 - x0, 0);
    const int hi = t == ws - 1 ? xe - 1 : min((int)ceilf(((float)t + 1.5f) * rx - 0.5f) + 1 - x0, xe - 1);
    float acc = 0.f;
    if (ty < ye)
      for (int px = lo; px <= hi; ++px) {
        const int tp2 = s_tapx[px];
        const float l1 = s_lx[px];
        const float wgt = ((tp2 & 0xffff) == t ? 1.f - l1 : 0.f) + ((tp2 >> 16) == t ? l1 : 0.f);
        acc += wgt * s_gd[ty * TW + px];
      }
    s_h[ty * FJ + j] = acc;
  }
  __syncthreads();
  float* gd = g.grad_disp + (size_t)b * hs * ws;
  for (int e = JPB_TID; e < ni * nj; e += JPB_NT) {
    const int i = e / nj, j = e - i * nj, t = fy0 + i;
    const int lo = t == 0 ? 0 : max((int)floorf(((float)t - 0.5f) * ry - 0.5f) - 1 - y0, 0);
    const int hi = t == hs - 1 ? ye - 1 : min((int)ceilf(((float)t + 1.5f) * ry - 0.5f) + 1 - y0, ye - 1);
    float acc = 0.f;
    for (int py = lo; py <= hi; ++py) {
      const int tp2 = s_tapy[py];
      const float l1 = s_ly[py];
      const float wgt = ((tp2 & 0xffff) == t ? 1.f - l1 : 0.f) + ((tp2 >> 16) == t ? l1 : 0.f);
      acc += wgt * s_h[py * FJ + j];
    }
    if (acc != 0.f) atomicAdd(&gd[t * ws + fx0 + j], acc);
  }

  // ---- phase E: pose gradient, dT[k][j] = sum_i K[i][k] * G[i][j]: warp shuffles, one shared-memory transposition
  float* s_tot = s_red + NWARP_MAX * 24;
#if defined(JPB_HOST_EMU) && !defined(JPB_HOST_EMU_MT)
  for (int f = 0; f < F; ++f)
    for (int i = 0; i < 12; ++i) s_tot[f * 12 + i] = G[f][i];
#else
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float v = G[f][i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[wid * 24 + f * 12 + i] = v;
      }
    __syncthreads();
    if ((int)threadIdx.x < 12 * F) {
      float v = 0.f;
      for (int w = 0; w < nw; ++w) v += s_red[w * 24 + threadIdx.x];
      s_tot[threadIdx.x] = v;
    }
    __syncthreads();
  }
#endif
  for (int e = JPB_TID; e < 16 * F; e += JPB_NT) {
    const int f = e >> 4, k = (e >> 2) & 3, j = e & 3;
    if (!g.grad_T[f]) continue;
    const float* K = a.K + b * 16;
    const float* tot = s_tot + f * 12;
    const float v = K[0 * 4 + k] * tot[0 * 4 + j] + K[1 * 4 + k] * tot[1 * 4 + j] + K[2 * 4 + k] * tot[2 * 4 + j];
    if (v != 0.f) atomicAdd(&g.grad_T[f][b * 16 + k * 4 + j], v);
  }
}
}  // namespace v4

}  // namespace

static int g_fwd_variant = 3;   // 3: v3::photometric_fwd_kernel (packed fp32; measured 24 % faster on B200, profiles/r2_photometric_ab.jsonl); 2: photometric_fwd_kernel

static int g_bwd_variant = 4;   // 4: v4::photometric_bwd_kernel<F> (F <= 2); 1: photometric_bwd_kernel (any F)

extern "C" int jpb_photometric_set_bwd_variant(int bwd_variant) {
  if (bwd_variant != 1 && bwd_variant != 4) return JPB_ERR_ARG;
  g_bwd_variant = bwd_variant;
  return JPB_OK;
}

extern "C" int jpb_photometric_get_variant(void) { return g_fwd_variant; }

extern "C" int jpb_photometric_set_variant(int fwd_variant) {
  if (fwd_variant != 2 && fwd_variant != 3) return JPB_ERR_ARG;
  g_fwd_variant = fwd_variant;
  return JPB_OK;
}

extern "C" int jpb_photometric_fwd(const JpbPhotoArgs* a, void* stream) {
  if (!a || a->F < 1 || a->F > JPB_MAX_SRC || a->H < 3 || a->W < 3 || !a->loss_sum) return JPB_ERR_ARG;
  const int nid = a->automask ? a->F : 0;
  if (a->ident_mode < 0 || a->ident_mode > 2 || (a->ident_mode && (!a->ident_err || !nid))) return JPB_ERR_ARG;
  if (a->F <= 2 && g_fwd_variant == 3 && (long long)a->H * a->W * 3 < (1ll << 31)) {   // packed variant (default)
    const int mode = !nid ? 0 : (a->ident_mode == 2 ? 2 : 1);
    const size_t smem = (size_t)(3 + 6 + (mode == 1 ? 6 : 0)) * v3::AN * sizeof(float);
    dim3 grid((a->W + v3::TW - 1) / v3::TW, (a->H + v3::TH - 1) / v3::TH, a->B);
#ifndef JPB_HOST_EMU
    static bool configured = false;
    if (!configured) {
      const int big = (int)((3 + 6 + 6) * v3::AN * sizeof(float)), lean = (int)((3 + 6) * v3::AN * sizeof(float));
      if (cudaFuncSetAttribute(v3::photometric_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
          cudaFuncSetAttribute(v3::photometric_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lean) != cudaSuccess ||
          cudaFuncSetAttribute(v3::photometric_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lean) != cudaSuccess)
        return JPB_ERR_UNSUPPORTED;
      configured = true;
    }
#endif
    if (mode == 0) JPB_LAUNCH(v3::photometric_fwd_kernel<0>, grid, dim3(256), smem, (cudaStream_t)stream, *a);
    else if (mode == 1) JPB_LAUNCH(v3::photometric_fwd_kernel<1>, grid, dim3(256), smem, (cudaStream_t)stream, *a);
    else JPB_LAUNCH(v3::photometric_fwd_kernel<2>, grid, dim3(256), smem, (cudaStream_t)stream, *a);
    return jpb_status();
  }
  if (a->ident_mode) return JPB_ERR_UNSUPPORTED;   // the shared identity terms exist in the packed schedule only
  if (a->F <= 2) {   // fast path: 32x32 tiles, separable window sums (every reference configuration: F <= 2)
    const size_t smem = (size_t)(3 + 3 * a->F) * Q1_N * sizeof(float);
    dim3 grid((a->W + QT_W - 1) / QT_W, (a->H + QT_H - 1) / QT_H, a->B);
#ifndef JPB_HOST_EMU
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      if (cudaFuncSetAttribute(photometric_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
      configured = smem;
    }
#endif
    JPB_LAUNCH(photometric_fwd_kernel, grid, dim3(256), smem, (cudaStream_t)stream, *a);
    return jpb_status();
  }
  const size_t smem = (size_t)(3 + 3 * nid + 3 * a->F) * P1_N * sizeof(float);
  dim3 grid((a->W + PT_W - 1) / PT_W, (a->H + PT_H - 1) / PT_H, a->B);
#ifndef JPB_HOST_EMU
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(photometric_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = smem;
  }
#endif
  JPB_LAUNCH(photometric_fwd_generic_kernel, grid, dim3(256), smem, (cudaStream_t)stream, *a);
  return jpb_status();
}

extern "C" int jpb_photometric_bwd(const JpbPhotoArgs* a, const JpbPhotoGrad* g, void* stream) {
  if (!a || !g || a->F < 1 || a->F > JPB_MAX_SRC || !g->winner || !g->grad_disp || !g->grad_out) return JPB_ERR_ARG;
  if (a->hs > a->H || a->ws > a->W) return JPB_ERR_UNSUPPORTED;  // footprint buffer assumes up-sampling
  if (a->F <= 2 && g_bwd_variant == 4 && (long long)a->H * a->W * 3 < (1ll << 31)) {   // register-resident schedule (default)
    const size_t smem4 = (size_t)v4::SMEM_FLOATS * sizeof(float);
    dim3 grid4((a->W + v4::TW - 1) / v4::TW, (a->H + v4::TH - 1) / v4::TH, a->B);
#ifndef JPB_HOST_EMU
    static bool configured4 = false;
    if (!configured4) {
      if (cudaFuncSetAttribute(v4::photometric_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) != cudaSuccess ||
          cudaFuncSetAttribute(v4::photometric_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) != cudaSuccess)
        return JPB_ERR_UNSUPPORTED;
      configured4 = true;
    }
#endif
    if (a->F == 1) JPB_LAUNCH(v4::photometric_bwd_kernel<1>, grid4, dim3(256), smem4, (cudaStream_t)stream, *a, *g);
    else JPB_LAUNCH(v4::photometric_bwd_kernel<2>, grid4, dim3(256), smem4, (cudaStream_t)stream, *a, *g);
    return jpb_status();
  }
  const size_t smem = (size_t)(3 + 3 * a->F) * P2_N * sizeof(float) + (size_t)9 * P1_N * sizeof(float) + (size_t)P1_N * sizeof(int);
  dim3 grid((a->W + PT_W - 1) / PT_W, (a->H + PT_H - 1) / PT_H, a->B);
#ifndef JPB_HOST_EMU
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(photometric_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = smem;
  }
#endif
  JPB_LAUNCH(photometric_bwd_kernel, grid, dim3(256), smem, (cudaStream_t)stream, *a, *g);
  return jpb_status();
}
