// Micro-benchmark: tcgen05.mma kind::tf32 issue ceiling on this part, cta_group::1 vs cta_group::2 (2-CTA clusters), SS operands
// (garbage data in shared memory; no loads, no epilogue).  Every CTA (pair) issues `iters` x 4 MMAs of M = 128 (256 per pair),
// K = 8, N = NT into one TMEM accumulator and commits once at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_ceiling tools/microbench/mma_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n .reg .pred p;\nW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\nD_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// mode 0: fixed 1024-B-aligned A tile (8-row group stride 1024 B) — the GEMM case;  mode 1: group stride 2048 B (rows of a 16-pixel
// patch), aligned start;  mode 2: as 1, start address cycling through the nine 3x3 taps (ky * 2048 + kx * 128: NOT 1024-aligned for
// kx != 0) — the TMA-patch convolution;  mode 3: as 2, B cycling through four stages
template <int NT, int CG>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (36864 + 4 * NT * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (float)(i & 7) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && rank == 0) {
    // K-major SW128 descriptors: A 128 rows x 128 B, B NT (per CTA: NT / CG) rows x 128 B
    const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = (((smem_u32(smem) + 36864) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t hia = mode >= 1 ? ((2048u >> 4) | (1u << 14) | (2u << 29)) : hi;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    int tap = 0, bs = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      if (mode >= 2) { a_lo += (uint32_t)(((tap / 3) * 2048 + (tap % 3) * 128) >> 4); if (++tap == 9) tap = 0; }
      if (mode >= 3) { b_lo += (uint32_t)((bs * NT * 128) >> 4); if (++bs == 4) bs = 0; }
      asm volatile(
          "{\n .reg .pred pe, pa, pt;\n .reg .b64 da, db;\n .reg .b32 al, bl;\n"
          " elect.sync _|pe, 0xffffffff;\n setp.ne.b32 pa, %5, 0;\n setp.eq.b32 pt, %5, %5;\n"
          " mov.b64 da, {%1, %7};\n mov.b64 db, {%2, %3};\n"
          " @pe tcgen05.mma.cta_group::%6.kind::tf32 [%0], da, db, %4, pa;\n"
          " add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %7};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::%6.kind::tf32 [%0], da, db, %4, pt;\n"
          " add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %7};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::%6.kind::tf32 [%0], da, db, %4, pt;\n"
          " add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %7};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::%6.kind::tf32 [%0], da, db, %4, pt;\n"
          "}" ::"r"(tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(it), "n"(CG), "r"(hia) : "memory");
    }
    if (CG == 1) asm volatile("{\n .reg .pred pe;\n elect.sync _|pe, 0xffffffff;\n @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(&done_bar)) : "memory");
    else asm volatile("{\n .reg .pred pe;\n .reg .b16 m;\n mov.b16 m, 3;\n elect.sync _|pe, 0xffffffff;\n @pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}" ::"r"(smem_u32(&done_bar)) : "memory");
  }
  mbar_wait(&done_bar, 0);     // both CTAs of a pair: the commit is multicast to both barriers
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n tcgen05.wait::ld.sync.aligned;" : "=r"(r0) : "r"(tmem));
    if (threadIdx.x == 0 && out) out[blockIdx.x] = __uint_as_float(r0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

template <int NT, int CG>
void run(int iters, int mode) {
  const int smem = 36864 + 4 * NT * 128 + 2048;
  cudaFuncSetAttribute(mma_kernel<NT, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 148;
  float* out;
  cudaMalloc(&out, 4 * sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, mma_kernel<NT, CG>, iters, out, mode);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaError_t e2 = cudaGetLastError();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 128 * NT * 8 * 4.0 * iters * sms;   // per CTA: M = 128 rows of the (pair's) tile
    if (rep == 2) printf("cta_group::%d N=%3d mode=%d iters=%d  %.3f ms  %.1f TFLOP/s  (%s / %s)\n", CG, NT, mode, iters, ms, flop / (ms * 1e-3) / 1e12, cudaGetErrorString(err), cudaGetErrorString(e2));
  }
  cudaFree(out);
}

int main() {
  const int iters = 20000;
  for (int mode = 0; mode < 4; ++mode) { run<64, 1>(iters, mode); run<128, 1>(iters, mode); run<256, 1>(iters, mode); }
  run<64, 2>(iters, 0); run<128, 2>(iters, 0); run<256, 2>(iters, 0);
  return 0;
}
