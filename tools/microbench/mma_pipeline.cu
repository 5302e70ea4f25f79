// Micro-benchmark: what separates the convolution kernels' main loop (300-350 TFLOP/s on the wide layers) from the 1114 TFLOP/s
// that tcgen05.mma.cta_group::1.kind::tf32 reaches in isolation (mma_ceiling.cu)?  The skeleton of conv_tc_fwd_kernel — 4 producer
// warps, one bulk-copy warp, one MMA-issuing warp, a ring of STAGES {A 128 x 128 B, B NT x 128 B} stages, full / empty mbarriers,
// tcgen05.commit per K block — with the pieces switched on one at a time (mode bits):
//   1  per-K-block hand-shake (commit -> empty barrier -> producers -> full barrier -> MMA warp), no data movement
//   2  producers fill the A stage with 16-byte cp.async gathers from an L2-resident array (1024 per stage) + fence.proxy.async
//   4  the B stage arrives by cp.async.bulk (global -> shared, complete_tx on the full barrier)
//   8  the gather rows are 9 x re-read scattered image rows (im2col-like) instead of one linear range
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_pipeline tools/microbench/mma_pipeline.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n .reg .pred p;\nW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\nD_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }

template <int NT, int STAGES>
__global__ void __launch_bounds__(192, 1) pipe_kernel(int nkb, const float* gsrc, size_t gfloats, float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int A_STAGE = 128 * 128, B_STAGE = NT * 128, STAGE = A_STAGE + B_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < STAGES * STAGE / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (float)(i & 7) * 0.125f;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128 + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const bool shake = mode & 1, gather = mode & 2, bulk = mode & 4, scat = mode & 8;
  if (warp < 4) {
    if (shake) {
      // row r of the A tile (128 rows x 128 B): thread owns row tid, 8 x 16-byte chunks, written 128B-swizzled
      const size_t rows = gfloats / 32;
      size_t row = ((size_t)blockIdx.x * 128 + tid) % rows;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], (((uint32_t)(kb / STAGES)) & 1u) ^ 1u);
        if (gather) {
          const uint32_t dst = smem_u32(smem + s * STAGE) + (uint32_t)tid * 128u;
          const float* g = gsrc + row * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async16(dst + (uint32_t)((c ^ (tid & 7)) << 4), g + c * 4);
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          row = scat ? (row * 9 + 1031 * (size_t)(kb % 9)) % rows : (row + (size_t)gridDim.x * 128) % rows;
        }
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == 4) {
    if (shake && lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], (((uint32_t)(kb / STAGES)) & 1u) ^ 1u);
        if (bulk) {
          mbar_expect_tx(&full_bar[s], (uint32_t)B_STAGE);
          const float* g = gsrc + ((size_t)kb * B_STAGE / 4) % (gfloats - B_STAGE / 4);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(smem + s * STAGE + A_STAGE)), "l"(g), "r"(B_STAGE), "r"(smem_u32(&full_bar[s])) : "memory");
        } else mbar_arrive(&full_bar[s]);
      }
    }
  } else {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      if (shake) {
        mbar_wait(&full_bar[s], ((uint32_t)(kb / STAGES)) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t sa = smem_u32(smem + s * STAGE);
      const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16), b_lo = (((sa + A_STAGE) >> 4) & 0x3FFFu) | (1u << 16);
      asm volatile(
          "{\n .reg .pred pe, pa, pt;\n .reg .b64 da, db;\n .reg .b32 al, bl;\n"
          " elect.sync _|pe, 0xffffffff;\n setp.ne.b32 pa, %5, 0;\n setp.eq.b32 pt, %5, %5;\n"
          " mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
          " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pa;\n"
          " add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n"
          " add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n"
          " add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
          " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n"
          "}" ::"r"(tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(kb) : "memory");
      if (shake) asm volatile("{\n .reg .pred pe;\n elect.sync _|pe, 0xffffffff;\n @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(&empty_bar[s])) : "memory");
    }
    asm volatile("{\n .reg .pred pe;\n elect.sync _|pe, 0xffffffff;\n @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(done_bar)) : "memory");
  }
  mbar_wait(done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n tcgen05.wait::ld.sync.aligned;" : "=r"(r0) : "r"(tmem));
    if (tid == 0 && out) out[blockIdx.x] = __uint_as_float(r0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int NT, int STAGES>
void run(int nkb, int mode, const float* gsrc, size_t gfloats) {
  const int smem = STAGES * (128 * 128 + NT * 128) + 1024 + 256;
  cudaFuncSetAttribute(pipe_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int sms = 148;
  float* out;
  cudaMalloc(&out, 4 * sms);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    pipe_kernel<NT, STAGES><<<sms, 192, smem>>>(nkb, gsrc, gfloats, out, mode);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const cudaError_t err = cudaGetLastError();
  const double flop = 2.0 * 128 * NT * 32.0 * nkb * sms;
  printf("N=%3d stages=%d mode=%2d  %.3f ms  %7.1f TFLOP/s  %.0f cycles/K-block  (%s)\n", NT, STAGES, mode, ms, flop / (ms * 1e-3) / 1e12,
         ms * 1e-3 * 1.965e9 / nkb, cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  const size_t gfloats = (size_t)16 << 20;   // 64 MB: L2-resident after the first pass
  float* g;
  cudaMalloc(&g, gfloats * 4);
  cudaMemset(g, 0, gfloats * 4);
  const int nkb = 20000;
  for (int mode : {0, 1, 3, 5, 7, 15}) {
    run<256, 4>(nkb, mode, g, gfloats);
    run<256, 2>(nkb, mode, g, gfloats);
    run<128, 4>(nkb, mode, g, gfloats);
    run<64, 6>(nkb, mode, g, gfloats);
  }
  cudaFree(g);
  return 0;
}
