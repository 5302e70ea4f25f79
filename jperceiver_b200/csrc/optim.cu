// Flat-buffer optimizer step: one pass for the squared gradient norm, one fused pass for
//   1/world scaling (post-allreduce) -> global-norm clipping -> Adam.
// Replaces (per training step) the reference's  mono/core/utils/dist_utils.py:34-60  chain
//   allreduce_grads (flatten, all_reduce, div_, unflatten+copy_) -> mmcv clip_grads
//   (torch.nn.utils.clip_grad_norm_, 466 tensors) -> torch.optim.Adam.step (foreach over 466 tensors)
// with two launches over one contiguous fp32 buffer (52 M elements).
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* g, long long n, double* acc) {
  __shared__ double red[32];
  double s = 0.0;
  const long long stride = (long long)gridDim.x * JPB_NT;
  const long long n4 = (((uintptr_t)g & 15) == 0) ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += stride) {
    const float4 q = reinterpret_cast<const float4*>(g)[i];
    s += ((double)q.x * (double)q.x + (double)q.y * (double)q.y) + ((double)q.z * (double)q.z + (double)q.w * (double)q.w);
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += stride) {
    const float v = g[i];
    s += (double)v * (double)v;
  }
  const double t = jpb_block_sum<double>(s, red);
  if (JPB_TID == 0) atomicAdd(acc, t);
}

__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* g, float* m, float* v, long long n, JpbAdamArgs a) {
  // step counter lives on the device so the launch is CUDA-graph replayable
  const long long t = a.step[0] + 1;
  const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
  const float step_size = (float)((double)a.lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  float gscale = a.grad_scale;
  if (a.max_norm > 0.f && a.normsq) {
    const float total = (float)(sqrt(a.normsq[0]) * (double)a.grad_scale);   // norm of the averaged gradient
    const float coef = a.max_norm / (total + 1e-6f);
    if (coef < 1.f) gscale *= coef;
  }
  const float b1 = a.beta1, b2 = a.beta2, wd = a.weight_decay, eps = a.eps;
  auto upd = [&](float gi, float& pi, float& mi, float& vi) {
    gi *= gscale;
    if (wd != 0.f) gi += wd * pi;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  };
  const long long stride = (long long)gridDim.x * JPB_NT;
  // 16-byte accesses (four parameters per thread and iteration: 4 x 16 B of loads in flight) when the buffers allow it
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * JPB_NT + JPB_TID; i < n4; i += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * JPB_NT + JPB_TID; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(g[i], pi, mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}

__global__ void bump_step_kernel(long long* step) {
  if (blockIdx.x == 0 && JPB_TID == 0) step[0] += 1;
}

}  // namespace

extern "C" int jpb_sumsq(const float* g, long long n, double* acc, void* stream) {
  if (!g || !acc || n < 1) return JPB_ERR_ARG;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(sumsq_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, g, n, acc);
  return jpb_status();
}

extern "C" int jpb_adam_step(float* p, const float* g, float* m, float* v, long long n, const JpbAdamArgs* a, void* stream) {
  if (!p || !g || !m || !v || !a || !a->step || n < 1) return JPB_ERR_ARG;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  JPB_LAUNCH(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, *a);
  JPB_LAUNCH(bump_step_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, a->step);
  return jpb_status();
}
