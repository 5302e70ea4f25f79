// Micro-benchmark: the A operand of the implicit-GEMM convolution fetched by TMA instead of 16-byte cp.async gathers.
// Same skeleton as mma_pipeline.cu (ring of STAGES {A 128 x 128 B, B NT x 128 B}, full / empty mbarriers, tcgen05.commit per K
// block), but ONE elected thread issues, per K block, a 2-D tensor load {32 channels, 128 pixels} of an NHWC activation
// [P pixels, C = 256 channels] at an im2col-like position (tile base + tap shift, channel block) and a 2-D load {32, NT} of the
// weight [NT, 9 * C]; nobody else touches the operands.  Prints cycles per K block and the TFLOP/s that implies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_tma_gather tools/microbench/mma_tma_gather.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n .reg .pred p;\nW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\nD_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <int NT, int STAGES, int MINB>
__global__ void __launch_bounds__(192, MINB) k(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap, int nkb, int P, int W,
                                               float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int A_STAGE = 128 * 128, B_STAGE = NT * 128, STAGE = A_STAGE + B_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (warp == 4) {
    if (lane == 0) {
      const int tiles = P / 128;
      const int p0 = (int)((blockIdx.x * 37u) % (unsigned)tiles) * 128;       // this CTA's output tile (128 consecutive pixels)
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], (((uint32_t)(kb / STAGES)) & 1u) ^ 1u);
        const int kk = kb % 72, tap = kk / 8, cb = kk % 8;                     // 9 taps x 8 channel blocks of a 256-channel 3x3 layer
        const int shift = (tap / 3 - 1) * W + (tap % 3 - 1);
        int pp = p0 + shift + (kb / 72) * 128 * 41;
        pp = ((pp % (P - 128)) + (P - 128)) % (P - 128);
        mbar_expect_tx(&full_bar[s], (uint32_t)((mode & 1 ? A_STAGE : 0) + (mode & 2 ? B_STAGE : 0)));
        if (mode & 1) tma_load_2d(smem_u32(smem + s * STAGE), &xmap, &full_bar[s], cb * 32, pp);
        if (mode & 2) tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE), &wmap, &full_bar[s], kk * 32, 0);
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], ((uint32_t)(kb / STAGES)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
      asm volatile("{\n .reg .pred pe;\n elect.sync _|pe, 0xffffffff;\n @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(&empty_bar[s])) : "memory");
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
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NT) : "memory");
}

static CUtensorMap make_map(float* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  CUtensorMap m;
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {inner * 4};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return m;
}

template <int NT, int STAGES, int MINB>
void run(int nkb, int mode, float* x, int P, int W, float* w) {
  const int smem = STAGES * (128 * 128 + NT * 128) + 1024 + 256;
  cudaFuncSetAttribute(k<NT, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  CUtensorMap xmap = make_map(x, 256, (uint64_t)P, 32, 128), wmap = make_map(w, 9 * 256, NT, 32, NT);
  const int grid = 148 * MINB;
  float* out;
  cudaMalloc(&out, 4 * grid);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<NT, STAGES, MINB><<<grid, 192, smem>>>(xmap, wmap, nkb, P, W, out, mode);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const cudaError_t err = cudaGetLastError();
  const double flop = 2.0 * 128 * NT * 32.0 * nkb * grid;
  printf("N=%3d stages=%d ctas/SM=%d mode=%d (A %s, B %s)  %.3f ms  %7.1f TFLOP/s  %.0f cycles per K block per SM  (%s)\n", NT, STAGES, MINB, mode,
         mode & 1 ? "TMA" : "-", mode & 2 ? "TMA" : "-", ms, flop / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / nkb / MINB, cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  cuInit(0);
  const int W = 256, P = 4 * 80 * 256;        // iconv1 / merge1 extent: 81920 pixels x 256 channels = 84 MB
  float *x, *w;
  cudaMalloc(&x, (size_t)P * 256 * 4);
  cudaMalloc(&w, (size_t)256 * 9 * 256 * 4);
  cudaMemset(x, 0, (size_t)P * 256 * 4);
  cudaMemset(w, 0, (size_t)256 * 9 * 256 * 4);
  const int nkb = 7200;
  for (int mode : {0, 1, 2, 3}) {
    run<256, 4, 1>(nkb, mode, x, P, W, w);
    run<256, 2, 2>(nkb, mode, x, P, W, w);
    run<128, 4, 2>(nkb, mode, x, P, W, w);
    run<64, 6, 2>(nkb, mode, x, P, W, w);
  }
  return 0;
}
