// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, TF32 x TF32 -> FP32 in TMEM).
//
//   D[M = B*Ho*Wo, N = Cout] = A[M, K] * W[N, K]^T,   K = taps x (padded) input channels
//
// A is never materialised: 128 producer threads gather it 16 bytes (4 channels) at a time with
// cp.async straight into the 128-byte-swizzled K-major shared-memory layout that tcgen05.mma reads.
// The gather folds everything the reference does as separate full-tensor passes in front of a conv:
//   * ReflectionPad2d / zero padding                      (layers.py:156-167, layout_model.py:31-47)
//   * nearest 2x up-sampling  F.interpolate(scale_factor=2) (layers.py:110, layout_model.py:50-53)
//   * channel concatenation   torch.cat((reduce, up, disp),1) (depth_decoder.py:76,96,115)
//   * stride-2 sampling
// through a small per-convolution chunk table (one int4 per 16-byte chunk of K: source tensor, tap
// offset, channel offset, valid bytes).  W tiles arrive by TMA (cp.async.bulk.tensor, 128B swizzle)
// from the K-major weight matrix (the channels-last nn.Conv2d weight as it sits in the flat parameter
// buffer, or a packed copy when Cin is not a multiple of 4).  One elected thread issues
// tcgen05.mma.cta_group::1.kind::tf32 (M=128, N<=256, K=8 per instruction, 4 per 32-float K block) into
// a TMEM accumulator; stages are recycled with tcgen05.commit -> mbarrier.  The epilogue reads TMEM with
// tcgen05.ld (32 lanes x 32 columns per warp) and fuses bias, residual add and LeakyReLU/ReLU/sigmoid
// before the only global write of the layer.
//
// Warp roles (192 threads): warps 0-3 gather A, then run the epilogue (warp w owns TMEM lanes 32w..32w+31);
// warp 4 allocates TMEM and drives the weight TMA; warp 5 issues the MMAs.
#include "jpb_common.cuh"
#include "../../include/jpb200.h"

#ifndef JPB_HOST_EMU
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

namespace {

constexpr int BM = 128;          // output pixels per CTA (UMMA M)
constexpr int BK = 32;           // floats per K block = one 128-byte swizzle row
constexpr int A_STAGE = BM * BK * 4;   // 16 KB
constexpr int NPROD = 128;       // A producer threads (warps 0-3)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Explicit shared-space accesses.  All shared pointers of these kernels derive from one dynamically aligned base, so the
// compiler no longer knows their address space and emits GENERIC loads/stores (LD.E/ST.E): those sit on the same (long)
// scoreboard as in-flight global loads — an offset-table read then waits for an unrelated L2 round trip.
__device__ __forceinline__ int lds32(uint32_t a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned long long lds64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// L1-allocating variant: neighbouring filter taps of one channel block re-read (almost) the same pixels; issued back to
// back (host K-block order) they hit L1 instead of crossing the L2 fabric again
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// wait until the oldest of `pending` (1..4) committed groups has landed
__device__ __forceinline__ void cp_async_wait_oldest(int pending) {
  if (pending <= 1) cp_async_wait<0>();
  else if (pending == 2) cp_async_wait<1>();
  else if (pending == 3) cp_async_wait<2>();
  else cp_async_wait<3>();
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): [0,14) start address >> 4 | [16,30) leading byte offset >> 4 |
// [32,46) stride byte offset >> 4 | [46,48) version = 1 | [49,52) base offset (0: the swizzle follows absolute address bits) |
// [61,64) layout (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B).  Built as 32-bit halves by the issue helpers below.
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// `scale`: truncation-bias compensation of kind::tf32 (JpbConvArgs.acc_scale), applied to the raw accumulator
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v, float scale = 1.f) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * scale;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v, float scale = 1.f) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * scale;
}

// BatchNorm statistics fused into the convolution epilogue: lanes l, l+8, l+16, l+24 hold sums of the same four output
// channels over different rows of the warp's 32-row slab; fold them and add the warp's totals (sum, sum of squares) into the
// double-precision accumulators the BatchNorm apply pass finalises (csrc/bn.cu) — the statistics pass over the convolution
// output disappears.  Executed by all 32 lanes.
__device__ __forceinline__ void bn_stats_flush(double* stats, int N, int col, bool valid, int lane, float4 s1, float4 s2) {
  float v[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
  }
  if (lane < 8 && valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(stats + col + q, (double)v[q]);
      atomicAdd(stats + N + col + q, (double)v[4 + q]);
    }
  }
}

// same fold, into shared memory: dst[c] (+NT: sums of squares) for the 32 columns of this pass; lanes 0-7 write 4 columns each
__device__ __forceinline__ void bn_stats_to_smem(uint32_t dst, int NT, int lane, float4 s1, float4 s2) {
  float v[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
  }
  if (lane < 8 && lane * 4 < NT) {   // a 16-column tile has only four column groups: the others would spill into the next region
    sts128(dst + (uint32_t)lane * 16u, make_float4(v[0], v[1], v[2], v[3]));
    sts128(dst + (uint32_t)(NT + lane * 4) * 4u, make_float4(v[4], v[5], v[6], v[7]));
  }
}


// ---- lean MMA issue (measured: the issuing warp, not the tensor pipe, bounded every kernel of this file) ----
// `if (lane == 0) tcgen05.mma ...` makes the uniform-datapath instructions execute in divergent code: ptxas wraps EACH of them in
// an ELECT / BRA.U.ANY "waterfall" loop and moves the descriptors through R2UR, ~100 dependent scalar instructions (~700 cycles)
// per 32-float K block — more than the 128..512 cycles the four MMAs of that block take on the tensor pipe (ncu source view of
// conv_tc_patch_kernel<64>: the issuing warp never waits on a barrier, the TMA producer spins on the empty barriers).
// Here the whole warp stays converged and ONE elect.sync predicate guards the four MMAs of a K block inside a single asm
// statement; descriptors are 32-bit low words + constant high words, so the advance along K is one integer add.
__device__ __forceinline__ void umma_tf32_k4(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                             uint32_t acc_first, uint32_t enable) {
  asm volatile(
      "{\n"
      " .reg .pred pe, pa, pt, pv;\n"
      " .reg .b64 da, db;\n"
      " .reg .b32 al, bl;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " setp.ne.b32 pv, %7, 0;\n"
      " and.pred pe, pe, pv;\n"
      " setp.ne.b32 pa, %6, 0;\n"
      " setp.eq.b32 pt, %6, %6;\n"
      " mov.b64 da, {%1, %3};\n"
      " mov.b64 db, {%2, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pa;\n"
      " add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(enable) : "memory");
}
// same, with the descriptor advance per MMA as an operand (`step` in 16-byte units: 2 for K-major SW128 operands, 64 for the
// MN-major operands of the weight-gradient kernel)
__device__ __forceinline__ void umma_tf32_k4s(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                              uint32_t acc_first, uint32_t step) {
  asm volatile(
      "{\n"
      " .reg .pred pe, pa, pt;\n"
      " .reg .b64 da, db;\n"
      " .reg .b32 al, bl;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " setp.ne.b32 pa, %6, 0;\n"
      " setp.eq.b32 pt, %6, %6;\n"
      " mov.b64 da, {%1, %3};\n"
      " mov.b64 db, {%2, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pa;\n"
      " add.u32 al, %1, %7;\n add.u32 bl, %2, %7;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, al, %7;\n add.u32 bl, bl, %7;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, al, %7;\n add.u32 bl, bl, %7;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, pt;\n"
      "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(step) : "memory");
}
// tcgen05.commit -> mbarrier arrive, issued by one elected lane of a converged warp
__device__ __forceinline__ void umma_commit_elect(uint32_t bar_saddr) {
  asm volatile(
      "{\n"
      " .reg .pred pe;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}" ::"r"(bar_saddr) : "memory");
}
// ---- CTA pairs (tcgen05 cta_group::2): two CTAs of a 2-cluster on the two SMs of a TPC compute one 256-row tile; each holds its
// 128 rows of A, HALF of the B tile and its 128 lanes of the accumulator.  The leader (cluster rank 0) issues the MMAs for both.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster (one elected lane of a converged warp)
__device__ __forceinline__ void mbar_arrive_remote_elect(uint32_t bar_saddr, uint32_t rank) {
  asm volatile(
      "{\n"
      " .reg .pred pe;\n"
      " .reg .b32 ra;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " mapa.shared::cluster.u32 ra, %0, %1;\n"
      " @pe mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}" ::"r"(bar_saddr), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote arrivals)
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAITC_%=:\n"
      " mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONEC_%=;\n"
      " bra WAITC_%=;\n"
      "DONEC_%=:\n"
      "}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_tf32_k4_cg2(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                 uint32_t acc_first) {
  asm volatile(
      "{\n"
      " .reg .pred pe, pa, pt;\n"
      " .reg .b64 da, db;\n"
      " .reg .b32 al, bl;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " setp.ne.b32 pa, %6, 0;\n"
      " setp.eq.b32 pt, %6, %6;\n"
      " mov.b64 da, {%1, %3};\n"
      " mov.b64 db, {%2, %4};\n"
      " @pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, pa;\n"
      " add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, pt;\n"
      " add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %4};\n"
      " @pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, pt;\n"
      "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar_saddr) {
  asm volatile(
      "{\n"
      " .reg .pred pe;\n"
      " .reg .b16 m;\n"
      " mov.b16 m, 3;\n"
      " elect.sync _|pe, 0xffffffff;\n"
      " @pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}" ::"r"(bar_saddr) : "memory");
}
constexpr uint32_t DESC_HI_SW128(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }   // bits [32,64) of the descriptor
constexpr uint32_t DESC_LO_LBO1 = 1u << 16;

struct SrcDev {
  const float* ptr;
  int C, H, W, up;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return v > 0.f ? v : 0.f;
  if (act == 2) return v > 0.f ? v : 0.01f * v;
  if (act == 3) return 1.f / (1.f + __expf(-v));
  return v;
}
template <int ACT>
__device__ __forceinline__ float act_c(float v) {
  if (ACT == 1) return v > 0.f ? v : 0.f;
  if (ACT == 2) return v > 0.f ? v : 0.01f * v;
  if (ACT == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

// ---- epilogue inner loop, shared by the forward kernels ----
// Per epilogue warp: a padded 32 x 32 staging tile (36-float rows) and a row table of 32 x {output row pointer | atomic flag,
// residual row pointer}.  One call handles one 32-column pass: 8 lanes cover the 128-byte segment of a row, 4 rows per
// iteration.  The activation and the residual are TEMPLATE parameters: with a run-time `act` every element went through three
// compare-and-branch pairs, and those branches (ncu: stall_branch_resolving on the ISETPs), not memory, made the epilogue as
// long as the main loop (graph-timed: 128 -> 128 channels 51 us with, 28 us without the epilogue).
constexpr int EPI_WARP_FLOATS = 32 * 36 + 128;          // staging tile + row table (32 x 16 bytes)
template <int ACT, bool RES>
__device__ __forceinline__ void epi_rows(uint32_t sbuf, uint32_t rowtab, int r0, int c4, int col, float4 bq, float4& st1, float4& st2) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 4 * i;
    unsigned long long rp, rr;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(rp), "=l"(rr) : "r"(rowtab + (uint32_t)r * 16u));
    if (!rp) continue;
    float* op = reinterpret_cast<float*>((uintptr_t)(rp & ~1ull)) + col;
    float4 o = lds128(sbuf + (uint32_t)(r * 36 + c4) * 4u);
    o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
    if (RES) {
      const float4 rq = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>((uintptr_t)rr) + col);
      o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w;
    }
    o.x = act_c<ACT>(o.x); o.y = act_c<ACT>(o.y); o.z = act_c<ACT>(o.z); o.w = act_c<ACT>(o.w);
    if (rp & 1ull) asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(op), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    else *reinterpret_cast<float4*>(op) = o;
    st1.x += o.x; st1.y += o.y; st1.z += o.z; st1.w += o.w;
    st2.x += o.x * o.x; st2.y += o.y * o.y; st2.z += o.z * o.z; st2.w += o.w * o.w;
  }
}
__device__ __forceinline__ void epi_rows_dispatch(int act, bool res, uint32_t sbuf, uint32_t rowtab, int r0, int c4, int col, float4 bq, float4& st1,
                                                  float4& st2) {
  if (res) {
    switch (act) {
      case 1: epi_rows<1, true>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      case 2: epi_rows<2, true>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      case 3: epi_rows<3, true>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      default: epi_rows<0, true>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
    }
  } else {
    switch (act) {
      case 1: epi_rows<1, false>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      case 2: epi_rows<2, false>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      case 3: epi_rows<3, false>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
      default: epi_rows<0, false>(sbuf, rowtab, r0, c4, col, bq, st1, st2); break;
    }
  }
}
__device__ __forceinline__ void sts_row(uint32_t rowtab, int lane, unsigned long long out_ptr_flag, unsigned long long res_ptr) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(rowtab + (uint32_t)lane * 16u), "l"(out_ptr_flag), "l"(res_ptr) : "memory");
}

struct WgradRowMaps {
  CUtensorMap m[JPB_CONV_MAX_SRC][3];   // per source: boxes of 32 channels x {32, 31, 1} pixels
};

// NT: N tile (UMMA N), multiple of 16 in [16,256].  STAGES: smem pipeline depth.
// CG = 2: CTA-pair version (launched as 2-clusters along blockIdx.x).  CTA x owns output rows [128 x, 128 x + 128) exactly as in
// the single-CTA version — gather, epilogue and accumulator lanes are unchanged — but loads only rows [rank * NT/2, +NT/2) of
// the weight tile: the pair's tcgen05.mma.cta_group::2 (M = 256) reads both halves, so the weight stream per output row — two
// thirds of the L2 -> SM traffic that bounds the wide layers (tools/microbench/mma_pipeline.cu, profiles/README.md) — is halved.
// The peer's MMA warp forwards "my stage is full" to the leader's pfull barrier; the leader's commits arrive in both CTAs.
// ROWS: the A operand arrives by TMA (JpbConvArgs.rows; stride 1, dense sources of whole 32-channel blocks).  The tile's 128 rows are
// four 32-pixel segments of image rows (the output raster has a row length `rows_wv` that is a multiple of 32 — for the data
// gradient of a reflection-padded layer the padded width rounded up, whose surplus pixels are computed and dropped); per K block
// (one source, one tap, 32 channels) lanes 0-3 of the copy warp issue one {32 channels, 32 pixels} box each — out-of-range rows,
// pixels and images are the box's zero fill, reflection mirrors the row coordinate and splits the box at the row ends into
// 31 + 1 pixels — and lane 4 the weight tile.  No thread gathers: measured (tools/microbench/mma_tma_gather.cu) the operand
// path then sustains 650-680 cycles per K block and SM against ~1650 with 1024 cp.async per block.
template <int NT, int STAGES, int MINB, int CG = 1, bool ROWS = false>
__global__ void __launch_bounds__(192, MINB) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ WgradRowMaps xmaps,
                                                                JpbConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [STAGES][A 16KB][B (NT/CG)*128] | barriers | tmem ptr | src table
  constexpr int B_STAGE = (NT / CG) * BK * 4;
  constexpr int STAGE = A_STAGE + B_STAGE;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint64_t* pfull_bar = accum_bar + 1;                     // CG == 2, leader: the peer's stage s is full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull_bar + (CG == 2 ? STAGES : 0));
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  SrcDev* srcs = reinterpret_cast<SrcDev*>(tmem_slot + 2);
  int* s_off = reinterpret_cast<int*>(srcs + JPB_CONV_MAX_SRC);
  const uint32_t srcs_u32 = smem_u32(srcs), soff_u32 = smem_u32(s_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wv = ROWS ? a.rows_wv : a.Wo;               // row length of the tile raster (ROWS: a multiple of 32, >= Wo)
  const int M = a.B * a.Ho * Wv;
  const float acc_scale = a.acc_scale != 0.f ? a.acc_scale : 1.f;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * NT;
  // debug timeline (tools/conv_timeline.py): stamps[cta][warp][slot] = globaltimer ns at fixed points of each role
#define JPB_STAMP(slot)                                                                                         \
  do {                                                                                                          \
    if (a.dbg && lane == 0 && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x < 512) {                        \
      unsigned long long t_;                                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                    \
      a.dbg[((size_t)blockIdx.x * 6 + warp) * 8 + (slot)] = (long long)t_;                                      \
    }                                                                                                           \
  } while (0)
  JPB_STAMP(0);
  // split-K: gridDim.z CTAs share one output tile, each reduces a slice of the K blocks and adds its partial tile
  // atomically (only used for epilogue-free convolutions; the output is zero-filled by the caller)
  const int kb_per = (a.nkb + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb0 = (int)blockIdx.z * kb_per;
  int nkb = a.nkb - kb0;
  if (nkb > kb_per) nkb = kb_per;
  if (nkb < 0) nkb = 0;
  const bool split = gridDim.z > 1;
  // rotated K loop (see the persistent kernel): CTAs of one N column must not all stream the same weight tile at once
  // (timing experiment JPB_CONV_SKIP=16: no rotation)
  const int rot = (nkb > 1 && !(a.dbg_skip & 16)) ? (int)(((((uint32_t)blockIdx.x / (uint32_t)CG) * 0x9E3779B1u) >> 12) % (uint32_t)nkb) : 0;   // same for a pair
#define JPB_KROT(kb) ((kb) + rot >= nkb ? (kb) + rot - nkb : (kb) + rot)

  if (tid < a.nsrc) {
    srcs[tid].ptr = a.src[tid];
    srcs[tid].C = a.src_C[tid]; srcs[tid].H = a.src_H[tid]; srcs[tid].W = a.src_W[tid]; srcs[tid].up = a.src_up[tid];
  }
  if (tid == 160) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], ROWS ? 1 : NPROD + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    if (CG == 2)
      for (int s = 0; s < STAGES; ++s) mbar_init(&pfull_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {   // TMEM allocation (power of two >= 32 columns), owned by warp 4 (of each CTA of a pair)
    constexpr uint32_t cols = NT < 32 ? 32 : NT;
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  JPB_STAMP(1);

  // ---- per-tile gather offsets: s_off[(tap*nsrc + src)*BM + row] = element offset of the pixel that tile row `row` reads for
  // filter tap `tap` in source `src` (padding / reflection / up-sampling / stride folded), or -1.  Thread = tile row: three
  // integer divisions per row, then the taps are walked with counters (the per-entry form with six divisions per entry cost
  // 2.2 us of every tile's 4.7 us set-up).
  if (!ROWS && tid < BM) {
    const int m = m0 + tid;
    const bool rowok = m < M;
    const int HoWo = a.Ho * a.Wo;
    const int b = rowok ? m / HoWo : 0, rem = m - b * HoWo;
    const int oy = rem / a.Wo, ox = rem - oy * a.Wo;
    int tap = 0;
    for (int ky = 0; tap < a.ntaps; ++ky)
      for (int kx = 0; kx < a.kw && tap < a.ntaps; ++kx, ++tap) {
        int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
        bool ok = rowok;
        if (a.in_div == 2) { ok = ok && !((iy | ix) & 1); iy >>= 1; ix >>= 1; }   // dgrad of a stride-2 convolution
        if (a.reflect) { iy = jpb_reflect(iy, a.Hin); ix = jpb_reflect(ix, a.Win); }
        else ok = ok && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win;
        for (int si = 0; si < a.nsrc; ++si) {
          int off = -1;
          if (ok) {
            const int sy = a.src_up[si] ? (iy >> 1) : iy, sx = a.src_up[si] ? (ix >> 1) : ix;
            off = ((b * a.src_H[si] + sy) * a.src_W[si] + sx) * a.src_C[si];
          }
          sts32(soff_u32 + (uint32_t)((tap * a.nsrc + si) * BM + tid) * 4u, off);
        }
      }
  } else if (tid == 128) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");   // descriptor fetch off the first TMA's path
  }
  __syncthreads();
  JPB_STAMP(2);

  if (warp < 4) {
    // ===================================================== A gather (producers)
    // The element offset of (tile row, tap, source) does not depend on the channel block: all threads computed it once
    // into s_off[(tap*nsrc + src)*128 + row] (-1 = zero fill) before the role split, so the K loop below is one shared
    // load, one add and one cp.async per 16-byte chunk (the gather used to spend ~60 integer instructions per chunk,
    // which made the four producer warps — not L2, not the tensor pipe — the bottleneck of the whole kernel).
    const int c = tid & 7;           // 16-byte chunk column inside the 128-byte K row
    const int rbase = tid >> 3;      // rows rbase + 16*i
    const uint32_t swz = (uint32_t)((c ^ (rbase & 7)) << 4) + (uint32_t)(rbase & 7) * 128u + (uint32_t)(rbase >> 3) * 1024u;
    // Producer state machine: a K block is published (proxy fence + arrive on its full barrier) as soon as its copies have
    // landed, independently of whether the NEXT block's stage is free — an in-order "issue kb, then publish kb-2" loop
    // made every publish wait for an MMA two blocks back and left the tensor pipe idle 60 % of the time.
    constexpr int MAXFLY = STAGES < 4 ? STAGES : 4;
    const bool use_ca = a.l1_gather != 0;
    int issued = 0, published = 0;
    // the chunk-table row of the NEXT K block is fetched one block ahead: a dependent global load per block (an L2 round trip:
    // L1 is carved down to almost nothing by the pipeline stages) was the largest single stall of the producer warps
    int4 e_next = make_int4(-1, 0, 0, 0);
    if (nkb > 0) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(0)) * 8 + c);
    while (!ROWS && published < nkb) {
      bool can = false;
      if (issued < nkb && issued - published < MAXFLY) {
        const int s = issued % STAGES;
        const uint32_t ph = ((uint32_t)(issued / STAGES) & 1u) ^ 1u;
        if (issued == published) { mbar_wait(&empty_bar[s], ph); can = true; }
        else can = mbar_test(&empty_bar[s], ph);
      }
      if (can) {
        const int kb = issued, s = kb % STAGES;
        // x: source | (tap*nsrc + source) << 8, or -1; y: dy<<16 | dx (wgrad only); z: channel offset; w: valid bytes
        const int4 e = e_next;
        if (kb + 1 < nkb) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(kb + 1)) * 8 + c);
        const uint32_t sbase = smem_u32(smem + s * STAGE) + swz;
        const bool live = e.x >= 0;
        const float* base = reinterpret_cast<const float*>((uintptr_t)lds64(srcs_u32 + (uint32_t)(live ? (e.x & 0xff) : 0) * (uint32_t)sizeof(SrcDev))) + e.z;
        const uint32_t offs = soff_u32 + (uint32_t)(((live ? (e.x >> 8) : 0) * BM + rbase) * 4);
        if (a.dbg_skip & 1) {
          // timing experiment: no A traffic
        } else if (!live || e.w == 16) {
          if (use_ca) {
            for (int i = 0; i < 8; ++i) {
              const int off = lds32(offs + 64u * (uint32_t)i);
              const bool ok = live && off >= 0;
              cp_async16_ca(sbase + (uint32_t)i * 2048u, base + (ok ? off : 0), ok ? 16u : 0u);
            }
          } else {
            for (int i = 0; i < 8; ++i) {
              const int off = lds32(offs + 64u * (uint32_t)i);
              const bool ok = live && off >= 0;
              cp_async16(sbase + (uint32_t)i * 2048u, base + (ok ? off : 0), ok ? 16u : 0u);
            }
          }
        } else {
          // partial chunk (a source whose channel count is not a multiple of 4, e.g. the 1-channel disparity):
          // synchronous scalar loads, zero padded
          const int nval = e.w >> 2;
          for (int i = 0; i < 8; ++i) {
            const int off = lds32(offs + 64u * (uint32_t)i);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off >= 0) {
              const float* g = base + off;
              v.x = g[0];
              if (nval > 1) v.y = g[1];
              if (nval > 2) v.z = g[2];
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + (uint32_t)i * 2048u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
        }
        cp_async_commit();
        ++issued;
      } else {
        cp_async_wait_oldest(issued - published);
        fence_async_proxy();
        mbar_arrive(&full_bar[published % STAGES]);
        ++published;
      }
    }

    // ===================================================== epilogue
    JPB_STAMP(3);
    if (nkb > 0) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    JPB_STAMP(4);
    const int row = warp * 32 + lane;
    int m = m0 + row;
    bool mok = m < M;
    if (ROWS && Wv != a.Wo) {                 // tile raster with a padded row length: back to the dense output index
      const int b_ = m / (a.Ho * Wv), rem_ = m - b_ * (a.Ho * Wv), oy_ = rem_ / Wv, ox_ = rem_ - oy_ * Wv;
      mok = mok && ox_ < a.Wo;
      m = mok ? (b_ * a.Ho + oy_) * a.Wo + ox_ : 0;
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* out = nullptr;
    const float* res = nullptr;
    bool atomic = split, vec_ok = (a.N & 3) == 0;
    int nvalid = a.N - n0;                    // channels of this N tile that exist
    if (!a.scatter) {
      out = a.out + (size_t)m * a.N + n0;
      res = a.residual ? a.residual + (size_t)m * a.N + n0 : nullptr;
    } else {
      // dgrad scatter: this launch's output pixel (py,px) lives in the (padded) gradient domain; fold it back through
      // the reflection padding and the nearest up-sampling of the forward gather, into the source that owns channels n0..
      int j = 0, cbase = 0;
      while (j + 1 < a.ndst && n0 >= cbase + a.dst_C[j]) { cbase += a.dst_C[j]; ++j; }
      vec_ok = (a.dst_C[j] & 3) == 0;
      nvalid = cbase + a.dst_C[j] - n0;
      if (mok) {
        const int b = m / (a.Ho * a.Wo), rem = m - b * (a.Ho * a.Wo);
        int ty = rem / a.Wo - a.fold_pad, tx = rem % a.Wo - a.fold_pad;
        if (a.fold_reflect) { ty = jpb_reflect(ty, a.fold_H); tx = jpb_reflect(tx, a.fold_W); }
        atomic = split || a.dst_up[j] || (a.fold_reflect && (ty <= 1 || ty >= a.fold_H - 2 || tx <= 1 || tx >= a.fold_W - 2));
        if (a.dst_up[j]) { ty >>= 1; tx >>= 1; }
        if (a.dst_mul > 1) { ty = ty * a.dst_mul + a.dst_oy; tx = tx * a.dst_mul + a.dst_ox; }   // one parity class of a stride-2 data gradient
        out = a.dst[j] + ((size_t)(b * a.dst_H[j] + ty) * a.dst_W[j] + tx) * a.dst_C[j] + (n0 - cbase);
      }
    }
    const bool vec_ok_all = vec_ok;           // uniform over the CTA (depends on n0 only)
    const int nvalid_all = nvalid;
    if (vec_ok_all) {
      // Coalesced epilogue: 32 accumulator columns per pass go TMEM -> registers (lane = row) -> a padded per-warp tile in
      // the (now idle) pipeline memory -> registers (8 lanes = one 128-byte row segment), so every global access of the
      // bias / residual / output covers whole 128-byte lines instead of 32 rows x 16 bytes.
      const uint32_t sbuf = smem_u32(smem) + (uint32_t)warp * EPI_WARP_FLOATS * 4u;
      const uint32_t rowptr = sbuf + 32 * 36 * 4;
      const uint32_t sstats = smem_u32(smem) + 4u * EPI_WARP_FLOATS * 4u;   // [4 warps][2][NT] floats, behind the staging tiles
      sts_row(rowptr, lane, mok ? (unsigned long long)(uintptr_t)out | (atomic ? 1ull : 0ull) : 0ull,
              (unsigned long long)(uintptr_t)(a.residual ? a.residual + (size_t)(mok ? m : 0) * a.N + n0 : nullptr));
      __syncwarp();
      const int c4 = (lane & 7) * 4, r0 = lane >> 3;
      for (int j = 0; j < NT; j += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)j, v, acc_scale);
        for (int q = 0; q < 32; q += 4)
          sts128(sbuf + (uint32_t)(lane * 36 + q) * 4u, make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]));
        __syncwarp();
        const int col = j + c4;
        float4 st1 = make_float4(0.f, 0.f, 0.f, 0.f), st2 = make_float4(0.f, 0.f, 0.f, 0.f);   // BatchNorm statistics of this pass
        if (col < nvalid_all) {
          float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) bq = *reinterpret_cast<const float4*>(a.bias + n0 + col);
          epi_rows_dispatch(a.act, a.residual != nullptr, sbuf, rowptr, r0, c4, col, bq, st1, st2);
        }
        if (a.stats) {
          if (NT >= 32) bn_stats_to_smem(sstats + (uint32_t)(warp * 2 * NT + j) * 4u, NT, lane, st1, st2);
          else bn_stats_flush(a.stats, a.N, n0 + col, col < nvalid_all, lane, st1, st2);
        }
        __syncwarp();
      }
      if (a.stats && NT >= 32) {
        // fold the four epilogue warps' column sums in shared memory: 2*NT reductions per CTA instead of 8*NT
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int t = tid; t < 2 * NT; t += 128) {
          const int cc = t % NT, which = t / NT;
          if (cc < nvalid_all) {
            float v = 0.f;
            for (int ww = 0; ww < 4; ++ww) v += __int_as_float(lds32(sstats + (uint32_t)(ww * 2 * NT + which * NT + cc) * 4u));
            atomicAdd(a.stats + (size_t)which * a.N + n0 + cc, (double)v);
          }
        }
      }
    } else {
    for (int j = 0; j < NT; j += 16) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)j, v, acc_scale);   // warp-collective: every lane executes it, even for rows >= M
      if (mok) {
        const int nleft = nvalid - j;
        if (nleft >= 16 && vec_ok) {
          for (int q = 0; q < 16; q += 4) {
            float4 o = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            if (a.bias) { const float4 bq = *reinterpret_cast<const float4*>(a.bias + n0 + j + q); o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w; }
            if (res) { const float4 rq = *reinterpret_cast<const float4*>(res + j + q); o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w; }
            o.x = apply_act(o.x, a.act); o.y = apply_act(o.y, a.act); o.z = apply_act(o.z, a.act); o.w = apply_act(o.w, a.act);
            if (atomic) asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + j + q), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            else *reinterpret_cast<float4*>(out + j + q) = o;
          }
        } else {
          for (int q = 0; q < 16 && q < nleft; ++q) {
            float o = v[q];
            if (a.bias) o += a.bias[n0 + j + q];
            if (res) o += res[j + q];
            o = apply_act(o, a.act);
            if (atomic) atomicAdd(out + j + q, o); else out[j + q] = o;
          }
        }
      }
    }
    }
    }
    JPB_STAMP(5);
    tc_fence_before();
  } else if (warp == 4) {
    // ===================================================== weight TMA producer
    if (ROWS) {
      if (lane <= 4) {
        // lane j < 4: segment j of the tile = 32 consecutive pixels (ox0 ..) of output row oy of image b (b >= B: past the end,
        // the box is all zero fill)
        const int ms = m0 + 32 * lane;
        const int sb = ms / (a.Ho * Wv), srem = ms - sb * (a.Ho * Wv), soy = srem / Wv, sox = srem - soy * Wv;
        const uint32_t seg_off = (uint32_t)lane * 4096u;
        int4 e_next = make_int4(-1, 0, 0, 0);
        if (nkb > 0 && lane < 4) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(0)) * 8);
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
          const int4 e = e_next;
          if (kb + 1 < nkb && lane < 4) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(kb + 1)) * 8);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (lane == 4) {
            mbar_expect_tx(&full_bar[s], (uint32_t)(B_STAGE + A_STAGE));
            const int kr = kb0 + JPB_KROT(kb);
            tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE), &wmap, &full_bar[s], a.kcol ? a.kcol[kr] : kr * BK, n0 + (int)crank * (NT / CG));
          } else {
            const int si = e.x < 0 ? 0 : (e.x & 0xff), tap = e.x < 0 ? 0 : (e.x >> 8) / a.nsrc;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            int iy = soy - a.pad + ky;
            const int ix = sox - a.pad + kx;
            int bb = e.x < 0 ? a.B : sb;                      // padding K block: all zero
            if (a.reflect) iy = jpb_reflect(iy, a.Hin);
            const uint32_t dst = smem_u32(smem + s * STAGE) + seg_off;
            if (a.reflect && ix < 0) {                       // pixel -1 mirrors to pixel 1
              tma_load_4d(dst + 128u, &xmaps.m[si][1], &full_bar[s], e.z, 0, iy, bb);
              tma_load_4d(dst, &xmaps.m[si][2], &full_bar[s], e.z, 1, iy, bb);
            } else if (a.reflect && ix + 32 > a.Win) {       // pixel W mirrors to pixel W - 2
              tma_load_4d(dst, &xmaps.m[si][1], &full_bar[s], e.z, ix, iy, bb);
              tma_load_4d(dst + 31u * 128u, &xmaps.m[si][2], &full_bar[s], e.z, a.Win - 2, iy, bb);
            } else {
              tma_load_4d(dst, &xmaps.m[si][0], &full_bar[s], e.z, ix, iy, bb);
            }
          }
        }
      }
    } else
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (a.dbg_skip & 2) { mbar_arrive(&full_bar[s]); continue; }   // timing experiment: no B traffic
        mbar_expect_tx(&full_bar[s], (uint32_t)B_STAGE);
        const int kr = kb0 + JPB_KROT(kb);
        tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE), &wmap, &full_bar[s], a.kcol ? a.kcol[kr] : kr * BK, n0 + (int)crank * (NT / CG));
      }
    }
  } else if (CG == 2 && crank != 0) {
    // ===================================================== pair, peer CTA: tell the leader when this CTA's stage is full
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], (uint32_t)(kb / STAGES) & 1u);
      mbar_arrive_remote_elect(smem_u32(&pfull_bar[s]), 0u);
    }
  } else {
    // ===================================================== MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10),
    // both K-major, n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      if (CG == 2) mbar_wait_cluster(&pfull_bar[s], ph);
      tc_fence_after();
      if (kb == 0) JPB_STAMP(3);
      {   // converged warp, one elected lane issues (see umma_tf32_k4): 8 TF32 (32 bytes) per instruction, start address + 2 x 16 B each
        const uint32_t sa = smem_u32(smem + s * STAGE);
        if (CG == 2) {
          umma_tf32_k4_cg2(tmem_base, ((sa >> 4) & 0x3FFFu) | DESC_LO_LBO1, (((sa + A_STAGE) >> 4) & 0x3FFFu) | DESC_LO_LBO1, DESC_HI_SW128(1024),
                           DESC_HI_SW128(1024), idesc, kb ? 1u : 0u);
          umma_commit_pair_elect(smem_u32(&empty_bar[s]));
          if (kb == nkb - 1) umma_commit_pair_elect(smem_u32(accum_bar));
        } else {
          umma_tf32_k4(tmem_base, ((sa >> 4) & 0x3FFFu) | DESC_LO_LBO1, (((sa + A_STAGE) >> 4) & 0x3FFFu) | DESC_LO_LBO1, DESC_HI_SW128(1024),
                       DESC_HI_SW128(1024), idesc, kb ? 1u : 0u, 1u);
          umma_commit_elect(smem_u32(&empty_bar[s]));
          if (kb == nkb - 1) umma_commit_elect(smem_u32(accum_bar));
        }
      }
    }
    JPB_STAMP(4);
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // both CTAs are done with each other's shared memory and tensor memory
  JPB_STAMP(6);
  if (warp == 4) {
    tc_fence_after();
    constexpr uint32_t cols = NT < 32 ? 32 : NT;
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
#undef JPB_STAMP
#undef JPB_KROT
}


// ======================================================================================== persistent forward (v2)
// Same GEMM, different schedule.  The per-CTA timeline of the one-tile-per-CTA kernel above showed 20-45 % of every tile
// outside the main loop (4.7 us of set-up + first operands, 5-20 us of epilogue during which every SM writes at once and
// the tensor pipe idles).  Here one CTA per SM slot stays resident and walks over tiles:
//   warps 0-3  epilogue   (TMEM lanes 32w..32w+31 -> padded smem tile -> coalesced global, fused bias/residual/act/scatter)
//   warps 4-7  A gather   (cp.async through the per-tile offset table; the ring of STAGES buffers runs across tiles)
//   warp  8    weight TMA
//   warp  9    MMA issue  (accumulator double-buffered in TMEM: tile i+1 is computed while tile i is drained)
template <int NT, int STAGES, int MINB>
__global__ void __launch_bounds__(320, MINB) conv_tc_fwd2_kernel(const __grid_constant__ CUtensorMap wmap, JpbConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int B_STAGE = NT * BK * 4;
  constexpr int STAGE = A_STAGE + B_STAGE;
  constexpr int ACC_STRIDE = NT < 32 ? 32 : NT;          // TMEM columns per accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE);
  constexpr int EPI_STATS_FLOATS = NT <= 64 ? 4 * 2 * NT : 0;   // cross-warp fold of the fused BatchNorm statistics
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + 4 * EPI_WARP_FLOATS + EPI_STATS_FLOATS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accf_bar = empty_bar + STAGES;   // [2] accumulator buffer complete (MMA -> epilogue)
  uint64_t* acce_bar = accf_bar + 2;         // [2] accumulator buffer drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acce_bar + 2);
  SrcDev* srcs = reinterpret_cast<SrcDev*>(tmem_slot + 2);
  int* s_off = reinterpret_cast<int*>(srcs + JPB_CONV_MAX_SRC);
  const uint32_t srcs_u32 = smem_u32(srcs), soff_u32 = smem_u32(s_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.B * a.Ho * a.Wo;
  const float acc_scale = a.acc_scale != 0.f ? a.acc_scale : 1.f;
  const int mtiles = (M + BM - 1) / BM, ntiles = (a.N + NT - 1) / NT;
  const int ks = a.ksplit > 1 ? a.ksplit : 1;
  const int total = mtiles * ntiles * ks;
  const int kb_per = (a.nkb + ks - 1) / ks;
  const bool split = ks > 1;

  if (tid < a.nsrc) {
    srcs[tid].ptr = a.src[tid];
    srcs[tid].C = a.src_C[tid]; srcs[tid].H = a.src_H[tid]; srcs[tid].W = a.src_W[tid]; srcs[tid].up = a.src_up[tid];
  }
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], NPROD + 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accf_bar[i], 1); mbar_init(&acce_bar[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> (m tile, n tile, K slice); every role walks the same sequence
#define JPB_TILE_DECODE(t)                                              \
  const int mt_ = (t) % mtiles, nz_ = (t) / mtiles;                     \
  const int m0 = mt_ * BM, n0 = (nz_ % ntiles) * NT;                    \
  const int kb0 = (nz_ / ntiles) * kb_per;                              \
  int nkb = a.nkb - kb0;                                                \
  if (nkb > kb_per) nkb = kb_per;                                       \
  if (nkb < 0) nkb = 0;                                                 \
  const int rot = nkb > 1 ? (int)((((uint32_t)(t) * 0x9E3779B1u) >> 12) % (uint32_t)nkb) : 0;
  // `rot` rotates the K loop of each tile: K is a sum, so any order is valid, and CTAs that would otherwise stream the SAME
  // weight tile at the same moment (every CTA of an N column walks K in lock-step) now read different ones — without it the
  // few L2 slices holding the current 8-32 KB weight tile serve all 148 SMs at once and bound the whole main loop.
#define JPB_KROT(kb) ((kb) + rot >= nkb ? (kb) + rot - nkb : (kb) + rot)

  if (warp >= 4 && warp < 8) {
    // ===================================================== A gather (producers)
    const int ptid = tid - 128;
    const int c = ptid & 7;           // 16-byte chunk column inside the 128-byte K row
    const int rbase = ptid >> 3;      // rows rbase + 16*i
    const uint32_t swz = (uint32_t)((c ^ (rbase & 7)) << 4) + (uint32_t)(rbase & 7) * 128u + (uint32_t)(rbase >> 3) * 1024u;
    constexpr int MAXFLY = STAGES < 4 ? STAGES : 4;
    const bool use_ca = a.l1_gather != 0;
    const int HoWo = a.Ho * a.Wo;
    int ring = 0;                     // K blocks issued by this CTA so far (ring position)
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      JPB_TILE_DECODE(t)
      (void)n0;
      // ---- gather offsets of this tile: thread = tile row; s_off[(tap*nsrc + src)*BM + row] = element offset of the pixel the
      // row reads for filter tap `tap` in source `src` (padding / reflection / up-sampling / stride folded), or -1
      {
        const int m = m0 + ptid;
        const bool rowok = m < M;
        const int b = rowok ? m / HoWo : 0, rem = m - b * HoWo;
        const int oy = rem / a.Wo, ox = rem - oy * a.Wo;
        int tap = 0;
        for (int ky = 0; tap < a.ntaps; ++ky)
          for (int kx = 0; kx < a.kw && tap < a.ntaps; ++kx, ++tap) {
            int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
            bool ok = rowok;
            if (a.in_div == 2) { ok = ok && !((iy | ix) & 1); iy >>= 1; ix >>= 1; }   // dgrad of a stride-2 convolution
            if (a.reflect) { iy = jpb_reflect(iy, a.Hin); ix = jpb_reflect(ix, a.Win); }
            else ok = ok && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win;
            for (int si = 0; si < a.nsrc; ++si) {
              int off = -1;
              if (ok) {
                const int sy = a.src_up[si] ? (iy >> 1) : iy, sx = a.src_up[si] ? (ix >> 1) : ix;
                off = ((b * a.src_H[si] + sy) * a.src_W[si] + sx) * a.src_C[si];
              }
              sts32(soff_u32 + (uint32_t)((tap * a.nsrc + si) * BM + ptid) * 4u, off);
            }
          }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int issued = 0, published = 0;
      int4 e_next = make_int4(-1, 0, 0, 0);
      if (nkb > 0) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(0)) * 8 + c);
      while (published < nkb) {
        bool can = false;
        if (issued < nkb && issued - published < MAXFLY) {
          const int ri = ring + issued, s = ri % STAGES;
          const uint32_t ph = ((uint32_t)(ri / STAGES) & 1u) ^ 1u;
          if (issued == published) { mbar_wait(&empty_bar[s], ph); can = true; }
          else can = mbar_test(&empty_bar[s], ph);
        }
        if (can) {
          const int kb = issued, s = (ring + issued) % STAGES;
          const int4 e = e_next;   // x: source | (tap*nsrc + source) << 8, or -1; z: channel offset; w: valid bytes
          if (kb + 1 < nkb) e_next = __ldg(reinterpret_cast<const int4*>(a.table) + (size_t)(kb0 + JPB_KROT(kb + 1)) * 8 + c);
          const uint32_t sbase = smem_u32(smem + s * STAGE) + swz;
          const bool live = e.x >= 0;
          const float* base = reinterpret_cast<const float*>((uintptr_t)lds64(srcs_u32 + (uint32_t)(live ? (e.x & 0xff) : 0) * (uint32_t)sizeof(SrcDev))) + e.z;
          const uint32_t offs = soff_u32 + (uint32_t)(((live ? (e.x >> 8) : 0) * BM + rbase) * 4);
          if (!live || e.w == 16) {
            if (use_ca) {
              for (int i = 0; i < 8; ++i) {
                const int off = lds32(offs + 64u * (uint32_t)i);
                const bool ok = live && off >= 0;
                cp_async16_ca(sbase + (uint32_t)i * 2048u, base + (ok ? off : 0), ok ? 16u : 0u);
              }
            } else {
              for (int i = 0; i < 8; ++i) {
                const int off = lds32(offs + 64u * (uint32_t)i);
                const bool ok = live && off >= 0;
                cp_async16(sbase + (uint32_t)i * 2048u, base + (ok ? off : 0), ok ? 16u : 0u);
              }
            }
          } else {
            // partial chunk (a source whose channel count is not a multiple of 4, e.g. the 1-channel disparity)
            const int nval = e.w >> 2;
            for (int i = 0; i < 8; ++i) {
              const int off = lds32(offs + 64u * (uint32_t)i);
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (off >= 0) {
                const float* g = base + off;
                v.x = g[0];
                if (nval > 1) v.y = g[1];
                if (nval > 2) v.z = g[2];
              }
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + (uint32_t)i * 2048u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
          }
          cp_async_commit();
          ++issued;
        } else {
          cp_async_wait_oldest(issued - published);
          fence_async_proxy();
          mbar_arrive(&full_bar[(ring + published) % STAGES]);
          ++published;
        }
      }
      ring += nkb;
      asm volatile("bar.sync 1, 128;" ::: "memory");   // nobody still reads s_off when the next tile's offsets are written
    }
  } else if (warp == 8) {
    // ===================================================== weight TMA producer
    if (lane == 0) {
      int ring = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        JPB_TILE_DECODE(t)
        (void)m0;
        for (int kb = 0; kb < nkb; ++kb, ++ring) {
          const int s = ring % STAGES;
          const uint32_t ph = (uint32_t)(ring / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], (uint32_t)B_STAGE);
          const int kr = kb0 + JPB_KROT(kb);
          tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE), &wmap, &full_bar[s], a.kcol ? a.kcol[kr] : kr * BK, n0);
        }
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int ring = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      JPB_TILE_DECODE(t)
      (void)m0; (void)n0; (void)kb0;
      if (nkb == 0) continue;
      const int buf = it & 1;
      mbar_wait(&acce_bar[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_STRIDE);
      for (int kb = 0; kb < nkb; ++kb, ++ring) {
        const int s = ring % STAGES;
        const uint32_t ph = (uint32_t)(ring / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        {
          const uint32_t sa = smem_u32(smem + s * STAGE);
          umma_tf32_k4(tacc, ((sa >> 4) & 0x3FFFu) | DESC_LO_LBO1, (((sa + A_STAGE) >> 4) & 0x3FFFu) | DESC_LO_LBO1, DESC_HI_SW128(1024),
                       DESC_HI_SW128(1024), idesc, kb ? 1u : 0u, 1u);
          umma_commit_elect(smem_u32(&empty_bar[s]));
          if (kb == nkb - 1) umma_commit_elect(smem_u32(&accf_bar[buf]));
        }
      }
      ++it;
    }
  } else {
    // ===================================================== epilogue (warps 0-3)
    const uint32_t sbuf = smem_u32(epi) + (uint32_t)(warp * EPI_WARP_FLOATS) * 4u;
    const uint32_t rowptr = sbuf + 32 * 36 * 4;
    const int c4 = (lane & 7) * 4, r0 = lane >> 3;
    constexpr int RS_PASSES = NT <= 64 ? (NT + 31) / 32 : 1;
    float run_stats[RS_PASSES][8];            // BatchNorm column sums of all tiles of this CTA (fused statistics, N <= NT <= 64)
#pragma unroll
    for (int pz = 0; pz < RS_PASSES; ++pz)
#pragma unroll
      for (int qz = 0; qz < 8; ++qz) run_stats[pz][qz] = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      JPB_TILE_DECODE(t)
      (void)kb0;
      if (nkb == 0) continue;
      const int buf = it & 1;
      const int row = warp * 32 + lane;
      const int m = m0 + row;
      float* out = nullptr;
      bool atomic = split, vec_ok = (a.N & 3) == 0;
      int nvalid = a.N - n0;                    // channels of this N tile that exist
      if (!a.scatter) {
        out = a.out + (size_t)m * a.N + n0;
      } else {
        // dgrad scatter: this launch's output pixel (py,px) lives in the (padded) gradient domain; fold it back through
        // the reflection padding and the nearest up-sampling of the forward gather, into the source that owns channels n0..
        int j = 0, cbase = 0;
        while (j + 1 < a.ndst && n0 >= cbase + a.dst_C[j]) { cbase += a.dst_C[j]; ++j; }
        vec_ok = (a.dst_C[j] & 3) == 0;
        nvalid = cbase + a.dst_C[j] - n0;
        if (m < M) {
          const int b = m / (a.Ho * a.Wo), rem = m - b * (a.Ho * a.Wo);
          int ty = rem / a.Wo - a.fold_pad, tx = rem % a.Wo - a.fold_pad;
          if (a.fold_reflect) { ty = jpb_reflect(ty, a.fold_H); tx = jpb_reflect(tx, a.fold_W); }
          atomic = split || a.dst_up[j] || (a.fold_reflect && (ty <= 1 || ty >= a.fold_H - 2 || tx <= 1 || tx >= a.fold_W - 2));
          if (a.dst_up[j]) { ty >>= 1; tx >>= 1; }
          if (a.dst_mul > 1) { ty = ty * a.dst_mul + a.dst_oy; tx = tx * a.dst_mul + a.dst_ox; }   // one parity class of a stride-2 data gradient
          out = a.dst[j] + ((size_t)(b * a.dst_H[j] + ty) * a.dst_W[j] + tx) * a.dst_C[j] + (n0 - cbase);
        }
      }
      if (nvalid > NT) nvalid = NT;
      sts_row(rowptr, lane, (m < M) ? (unsigned long long)(uintptr_t)out | (atomic ? 1ull : 0ull) : 0ull,
              (unsigned long long)(uintptr_t)(a.residual ? a.residual + (size_t)(m < M ? m : 0) * a.N + n0 : nullptr));
      mbar_wait(&accf_bar[buf], ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * ACC_STRIDE);
      for (int j = 0; j < NT && j < nvalid; j += 32) {
        const int col = j + c4;
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias && vec_ok && col < nvalid) bq = *reinterpret_cast<const float4*>(a.bias + n0 + col);   // issued early: overlaps the TMEM load
        float v[32];
        tmem_ld32(taddr + (uint32_t)j, v, acc_scale);
        for (int q = 0; q < 32; q += 4)
          sts128(sbuf + (uint32_t)(lane * 36 + q) * 4u, make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]));
        __syncwarp();
        float4 st1 = make_float4(0.f, 0.f, 0.f, 0.f), st2 = make_float4(0.f, 0.f, 0.f, 0.f);   // BatchNorm statistics of this pass
        if (vec_ok) {
          if (col < nvalid) epi_rows_dispatch(a.act, a.residual != nullptr, sbuf, rowptr, r0, c4, col, bq, st1, st2);
          if (a.stats) {
            if (ntiles == 1 && NT <= 64) {       // every tile of this CTA has the same columns: keep running sums
              float* rs = run_stats[(NT <= 64) ? j / 32 : 0];
              rs[0] += st1.x; rs[1] += st1.y; rs[2] += st1.z; rs[3] += st1.w;
              rs[4] += st2.x; rs[5] += st2.y; rs[6] += st2.z; rs[7] += st2.w;
            } else {
              bn_stats_flush(a.stats, a.N, n0 + col, col < nvalid, lane, st1, st2);
            }
          }
        } else {
          // ragged channel counts (destination C % 4 != 0): scalar accesses, lane = column
          for (int r = 0; r < 32; ++r) {
            const unsigned long long rp = lds64(rowptr + (uint32_t)r * 16u);
            const int cc = j + lane;
            if (!rp || cc >= nvalid) continue;
            float* op = reinterpret_cast<float*>((uintptr_t)(rp & ~1ull)) + cc;
            float o = __int_as_float(lds32(sbuf + (uint32_t)(r * 36 + lane) * 4u));
            if (a.bias) o += a.bias[n0 + cc];
            if (a.residual) o += a.residual[(size_t)(m0 + warp * 32 + r) * a.N + n0 + cc];
            o = apply_act(o, a.act);
            if (rp & 1ull) atomicAdd(op, o); else *op = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acce_bar[buf]);
      ++it;
    }
    if (a.stats && ntiles == 1 && NT <= 64) {
      // one set of reductions per CTA: warp partials -> shared memory -> 2*NT double-precision adds
      const uint32_t sst = smem_u32(epi) + 4u * EPI_WARP_FLOATS * 4u;   // [4 warps][2][NT] floats
#pragma unroll
      for (int pz = 0; pz < RS_PASSES; ++pz) {
        const float4 a1 = make_float4(run_stats[pz][0], run_stats[pz][1], run_stats[pz][2], run_stats[pz][3]);
        const float4 a2 = make_float4(run_stats[pz][4], run_stats[pz][5], run_stats[pz][6], run_stats[pz][7]);
        bn_stats_to_smem(sst + (uint32_t)(warp * 2 * NT + pz * 32) * 4u, NT, lane, a1, a2);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      for (int tt = tid; tt < 2 * NT; tt += 128) {
        const int cc = tt % NT, which = tt / NT;
        if (cc < a.N) {
          float v = 0.f;
          for (int ww = 0; ww < 4; ++ww) v += __int_as_float(lds32(sst + (uint32_t)(ww * 2 * NT + which * NT + cc) * 4u));
          atomicAdd(a.stats + (size_t)which * a.N + cc, (double)v);
        }
      }
    }
  }
#undef JPB_TILE_DECODE
#undef JPB_KROT
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ======================================================================================== weight gradient
//   dW^T[k, co] = sum_p  im2col[p, k] * dZ[p, co]      (p = output pixel, k = position in the K-chunk order)
// GEMM with the reduction over pixels: both operands are "MN-major" (the gathered channels / the dZ channels are the
// contiguous dimension).  MN-major TF32 operands must use the SWIZZLE_128B_BASE32B layout (cute Swizzle<2,5,2>): shared
// memory holds 512-byte atoms of 4 pixels x 32 channels, 32-byte chunks XOR-ed with the pixel index mod 4.
//   M tile = 128 consecutive K positions (32 chunks of the table), N tile = NT output channels,
//   each pipeline stage = 32 pixels = 4 tcgen05.mma (K = 8 pixels each).
// The pixel range is split over gridDim.z CTAs; partial tiles are accumulated with red.global.add.
// ROWS (the default for stride-1 "same" convolutions whose row length is a multiple of 32): the im2col^T operand arrives by TMA.
// A pipeline step is 32 consecutive pixels of ONE image row, so for a (source, tap, 32-channel block) — one MN group of the A
// tile — the 32 x 32 operand block is a 4-D box {32 channels, 32 pixels, 1 row, 1 image} of the NHWC source at the tap's offset:
// zero padding is the box's out-of-bounds fill; reflection padding mirrors the row coordinate and, at the two row ends, splits the
// box into a 31-pixel box plus the mirrored single pixel (disjoint shared-memory destinations, same barrier).  One thread issues
// at most 8 bulk tensor copies per step in place of 1024 16-byte cp.async gathers issued by 128 threads: measured with the
// tensor pipe in the loop (tools/microbench/mma_tma_gather.cu, profiles/r2_mma_tma_gather.txt) the operand path then sustains
// 650-680 cycles per K block and SM against ~1650 with the gather.  Groups that are not a whole 32-channel block of one source
// (the 1-channel disparity of the iconv layers) keep the gather, in the CTAs that own them.

template <int NT, int STAGES, int MINB, bool ROWS = false>
__global__ void __launch_bounds__(192, MINB) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ WgradRowMaps xmaps,
                                                                  JpbConvWgradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int B_STAGE = NT * BK * 4;
  constexpr int STAGE = A_STAGE + B_STAGE;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = a.B * a.Ho * a.Wo;
  const float acc_scale = a.acc_scale != 0.f ? a.acc_scale : 1.f;
  const int mt = blockIdx.x, n0 = blockIdx.y * NT;
  // pixel range of this split, in 32-pixel steps
  const int steps_total = (P + 31) / 32;
  const int steps_per = (steps_total + (int)gridDim.z - 1) / (int)gridDim.z;
  const int step0 = blockIdx.z * steps_per;
  int nsteps = steps_total - step0;
  if (nsteps > steps_per) nsteps = steps_per;
  if (nsteps < 0) nsteps = 0;

  // ROWS: MN group g of this K tile (chunks 8g .. 8g+7 of the table) is `regular` (one TMA box per step), gathered, or dead
  // (beyond the table: stays zero).  The classification is the same for every step of the CTA.
  int grp_regular = 0, grp_live = 0;
  if (ROWS) {
    for (int g = 0; g < 4; ++g) {
      const int gq = mt * 32 + 8 * g;
      if (gq < a.nchunks && __ldg(a.table + (size_t)gq * 4) >= 0) {
        grp_live |= 1 << g;
        if (a.gflags[gq >> 3]) grp_regular |= 1 << g;
      }
    }
  }
  const bool gather_on = !ROWS || (grp_live & ~grp_regular) != 0;
  if (tid == 160) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], gather_on ? NPROD + 1 : 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (ROWS && !gather_on && grp_live != 0xF) {     // dead groups are never written: zero them once in every stage
    for (int i = tid; i < STAGES * (A_STAGE / 16); i += 192) {
      const int st_ = i / (A_STAGE / 16), o = i % (A_STAGE / 16);
      if (!((grp_live >> (o / 256)) & 1)) sts128(smem_u32(smem + st_ * STAGE) + (uint32_t)o * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    fence_async_proxy();
  }
  if (warp == 4) {
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&dymap)) : "memory");
    constexpr uint32_t cols = NT < 32 ? 32 : NT;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nsteps > 0) {
  if (warp < 4) {
    // ------------------------------------------------ im2col^T gather: thread owns chunk q of the K tile, 8 pixel slots
    const int q = tid & 31, prow = tid >> 5;
    const bool my_copy = !ROWS || !((grp_regular >> (q >> 3)) & 1);     // ROWS: regular groups arrive by TMA
    if (gather_on) {
    const int gq = mt * 32 + q;
    int4 e = make_int4(-1, 0, 0, 0);
    if (gq < a.nchunks) e = __ldg(reinterpret_cast<const int4*>(a.table) + gq);
    const int esrc = e.x < 0 ? 0 : (e.x & 0xff);
    const float* sptr = a.src[esrc];
    const int sC = a.src_C[esrc], sH = a.src_H[esrc], sW = a.src_W[esrc];
    const int sup = a.src_up[esrc];
    const int dy = e.y >> 16, dx = (int)(short)(e.y & 0xffff);
    const int nval = e.w >> 2;
    const int HoWo = a.Ho * a.Wo;
    // pixel coordinates of this thread's 8 slots, advanced incrementally (32 pixels per step): no divisions in the loop
    int pb[8], poy[8], pox[8];
    for (int i = 0; i < 8; ++i) {
      const long long p = (long long)step0 * 32 + prow + 4 * i;
      pb[i] = (int)(p / HoWo);
      const int rem = (int)(p - (long long)pb[i] * HoWo);
      poy[i] = rem / a.Wo;
      pox[i] = rem - poy[i] * a.Wo;
    }
    constexpr int MAXFLY = STAGES < 4 ? STAGES : 4;
    int issued = 0, published = 0;      // same producer state machine as the forward kernel
    while (published < nsteps) {
      bool can = false;
      if (issued < nsteps && issued - published < MAXFLY) {
        const int s = issued % STAGES;
        const uint32_t ph = ((uint32_t)(issued / STAGES) & 1u) ^ 1u;
        if (issued == published) { mbar_wait(&empty_bar[s], ph); can = true; }
        else can = mbar_test(&empty_bar[s], ph);
      }
      if (can) {
        const int s = issued % STAGES;
        const uint32_t abase = smem_u32(smem + s * STAGE);
        for (int i = 0; i < 8; ++i) {
          if (!my_copy) break;
          const int slot = prow + 4 * i;            // pixel slot 0..31 of this step
          // MN group (q>>3) of 4096 B = 8 K-groups of 4 pixels (512 B); row = slot&3; 32-byte chunk ((q&7)>>1) ^ row, 16-byte half q&1
          const uint32_t dst = abase + (uint32_t)((q >> 3) * 4096 + (slot >> 2) * 512 + (slot & 3) * 128 +
                                                   (((((q & 7) >> 1) ^ (slot & 3)) << 5) | ((q & 1) << 4)));
          bool ok = pb[i] < a.B && e.x >= 0;
          int iy = poy[i] * a.stride - a.pad + dy, ix = pox[i] * a.stride - a.pad + dx, b = pb[i];
          if (a.reflect) { iy = jpb_reflect(iy, a.Hin); ix = jpb_reflect(ix, a.Win); }
          else ok = ok && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win;
          if (sup) { iy >>= 1; ix >>= 1; }
          if (!ok) { iy = 0; ix = 0; b = 0; }
          const float* g = sptr + ((size_t)(b * sH + iy) * sW + ix) * sC + e.z;
          if (e.w == 16 || !ok) cp_async16(dst, g, ok ? 16u : 0u);
          else {
            float4 v = make_float4(g[0], nval > 1 ? g[1] : 0.f, nval > 2 ? g[2] : 0.f, 0.f);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
          // advance this slot by 32 pixels
          pox[i] += 32;
          while (pox[i] >= a.Wo) {
            pox[i] -= a.Wo;
            if (++poy[i] >= a.Ho) { poy[i] = 0; ++pb[i]; }
          }
        }
        cp_async_commit();
        ++issued;
      } else {
        cp_async_wait_oldest(issued - published);
        fence_async_proxy();
        mbar_arrive(&full_bar[published % STAGES]);
        ++published;
      }
    }
    }   // gather_on
    // ------------------------------------------------ epilogue: dW[n0 + j][k] (+)= D[k, j]
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    int k = mt * 128 + warp * 32 + lane;
    if (ROWS) {      // column of K position k in dw: the chunk's column (chunk_col, -1 = padding) + the channel inside the chunk
      const int cc = (k >> 2) < a.nchunks ? __ldg(a.chunk_col + (k >> 2)) : -1;
      k = cc < 0 ? a.w_cols : cc + (k & 3);
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool use_atomic = gridDim.z > 1 || a.accumulate;
    for (int j = 0; j < NT; j += 16) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)j, v, acc_scale);
      if (k < a.w_cols) {
        for (int c = 0; c < 16; ++c) {
          const int n = n0 + j + c;
          if (n < a.N) {
            float* d = a.dw + (size_t)n * a.w_row + k;
            if (use_atomic) atomicAdd(d, v[c]); else *d = v[c];
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    if (!ROWS) {
      if (lane == 0) {
        for (int st = 0; st < nsteps; ++st) {
          const int s = st % STAGES;
          const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], (uint32_t)B_STAGE);
          const int p0 = (step0 + st) * 32;
          for (int h = 0; h < NT / 32; ++h)     // one {32 channels x 32 pixels} box per MN group: lands as 8 atoms of 4 pixels
            tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE + h * 4096), &dymap, &full_bar[s], n0 + 32 * h, p0);
        }
      }
    } else {
      // ROWS: lane g < 4 issues the box(es) of MN group g, lane 4 the dZ tile — as ONE 3-D box {32 channels, 32 pixels, NT/32
      // channel blocks} when N % 32 == 0 (a.dz3), which lands block-major exactly like NT/32 separate boxes.  A single thread
      // issuing 12 copies per step (with their address arithmetic) was itself the bottleneck of the first version.
      const int g = lane;
      const bool mine = g < 4 && ((grp_regular >> g) & 1);
      int si = 0, dy = 0, dx = 0, coff = 0;
      if (mine) {
        const int4 e = __ldg(reinterpret_cast<const int4*>(a.table) + (mt * 32 + 8 * g));
        si = e.x & 0xff; dy = e.y >> 16; dx = (int)(short)(e.y & 0xffff); coff = e.z;
      }
      const CUtensorMap* m32 = &xmaps.m[si][0];
      const CUtensorMap* m31 = &xmaps.m[si][1];
      const CUtensorMap* m1 = &xmaps.m[si][2];
      const int nreg = __popc(grp_regular);
      // this step's pixels: row `ry` of image `rb`, pixels rx .. rx + 31 (Wo % 32 == 0: a step never leaves its row)
      const int p00 = step0 * 32;
      int rb = p00 / (a.Ho * a.Wo);
      int ry = (p00 - rb * (a.Ho * a.Wo)) / a.Wo;
      int rx = p00 - (rb * a.Ho + ry) * a.Wo;
      if (lane <= 4) {
        for (int st = 0; st < nsteps; ++st) {
          const int s = st % STAGES;
          const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (lane == 4) {
            mbar_expect_tx(&full_bar[s], (uint32_t)(B_STAGE + nreg * 4096));
            const int p0 = (step0 + st) * 32;
            if (a.dz3) tma_load_3d(smem_u32(smem + s * STAGE + A_STAGE), &dymap, &full_bar[s], 0, p0, n0 >> 5);
            else
              for (int h = 0; h < NT / 32; ++h) tma_load_2d(smem_u32(smem + s * STAGE + A_STAGE + h * 4096), &dymap, &full_bar[s], n0 + 32 * h, p0);
          } else if (mine) {
            int ys = ry + dy - a.pad;
            const int xs = rx + dx - a.pad;
            if (a.reflect) ys = jpb_reflect(ys, a.Hin);
            const uint32_t dst = smem_u32(smem + s * STAGE) + (uint32_t)g * 4096u;
            if (a.reflect && xs < 0) {                       // pixel -1 mirrors to pixel 1
              tma_load_4d(dst + 128u, m31, &full_bar[s], coff, 0, ys, rb);
              tma_load_4d(dst, m1, &full_bar[s], coff, 1, ys, rb);
            } else if (a.reflect && xs + 32 > a.Win) {       // pixel W mirrors to pixel W - 2
              tma_load_4d(dst, m31, &full_bar[s], coff, xs, ys, rb);
              tma_load_4d(dst + 31u * 128u, m1, &full_bar[s], coff, a.Win - 2, ys, rb);
            } else {
              tma_load_4d(dst, m32, &full_bar[s], coff, xs, ys, rb);   // out-of-range rows / pixels: zero fill = zero padding
            }
          }
          rx += 32;
          if (rx >= a.Wo) { rx = 0; if (++ry >= a.Ho) { ry = 0; ++rb; } }
        }
      }
    }
  } else {
    // MN-major x MN-major: a_major = b_major = 1 (bits 15, 16)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int st = 0; st < nsteps; ++st) {
      const int s = st % STAGES;
      const uint32_t ph = (uint32_t)(st / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (a.dbg && st == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {   // debug: dump stage 0 (A then B)
        const float* sa = reinterpret_cast<const float*>(smem);
        for (int i = lane; i < (A_STAGE + B_STAGE) / 4; i += 32) a.dbg[i] = sa[i];
        __syncwarp();
      }
      {
        // descriptor: leading byte offset = distance between 32-channel MN groups (4096 B), stride byte offset = distance
        // between 4-pixel K groups (512 B), layout type 1 = SWIZZLE_128B_BASE32B; one MMA (K = 8 pixels) spans two K groups
        constexpr uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
        constexpr uint32_t lbo = (4096u >> 4) << 16;
        const uint32_t sa = smem_u32(smem + s * STAGE);
        umma_tf32_k4s(tmem_base, ((sa >> 4) & 0x3FFFu) | lbo, (((sa + A_STAGE) >> 4) & 0x3FFFu) | lbo, hi, hi, idesc, st ? 1u : 0u, 64u);
        umma_commit_elect(smem_u32(&empty_bar[s]));
        if (st == nsteps - 1) umma_commit_elect(smem_u32(accum_bar));
      }
    }
  }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    constexpr uint32_t cols = NT < 32 ? 32 : NT;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}


// ======================================================================================== 3x3 / stride 1 / zero pad: A from TMA patches
// The gather kernels above read every input pixel 9 times from L2 (once per filter tap) through 16-byte cp.async — measured to be
// what bounds them (profiles/README.md).  For the convolutions that are plain 3x3 windows over ONE dense NHWC source (every
// non-strided ResNet block convolution of the three encoders, forward and data gradient) no gather is needed at all:
//   * an output tile is 16 image rows x 8 pixels (= 128 GEMM rows); for one block of 32 input channels its whole receptive
//     field is an 18 x 10 pixel patch.  ONE 4-D TMA box {32 ch, 16 px, 18 rows, 1 image} (zero fill outside the image = the
//     convolution's zero padding) lands it in shared memory as 18 x 16 rows of 128 bytes, 128-byte swizzled;
//   * GEMM row i = (tile row i / 8, pixel i % 8) of filter tap (ky, kx) is patch row (i / 8 + ky), pixel (i % 8 + kx): exactly
//     the canonical K-major SWIZZLE_128B operand with an 8-row group stride of 2048 bytes whose start address is moved by
//     ky * 2048 + kx * 128 bytes.  All nine taps of a channel block are tcgen05.mma instructions on the SAME patch.
// L2 -> SM traffic of the A operand drops from 9 x 16 KB to one 36 KB box per channel block, and no thread touches A.
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..), warp 4 TMA (patches + weight tiles), warp 5 MMA issue.
// Persistent over tiles with the accumulator double-buffered in TMEM.
constexpr int PATCH_PX = 16;
// TR: vertically adjacent 16x8 tiles handled together by one CTA (a "super-tile" of 16*TR rows x 8 pixels): one patch of
// 16*TR + 2 rows and ONE weight tile per (channel block, tap) feed TR accumulators — the weight stream, which dominated the L2 -> SM
// traffic of the one-tile version (ncu: 302 of 378 MB for 128 -> 128 channels), is halved per output for TR = 2.


template <int NT, int TR, int PSTAGES, int BSTAGES>
__global__ void __launch_bounds__(192 + 32 * (TR - 1), 1) conv_tc_patch_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                                                                     JpbConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int B_STAGE = NT * BK * 4;
  constexpr int ACC_STRIDE = NT < 32 ? 32 : NT;
  constexpr uint32_t TMEM_COLS = 2 * TR * ACC_STRIDE;
  static_assert(TMEM_COLS <= 512, "TMEM columns");
  constexpr int PATCH_BYTES = (16 * TR + 3) * PATCH_PX * 128;     // buffer size: up to three halo rows (the stem's 4 x 2 taps)
  const int ntaps = a.patch_ntaps;
  const uint32_t patch_tx = (uint32_t)((16 * TR + a.patch_halo) * PATCH_PX * 128);   // bytes of one TMA box
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* s_patch = smem;                                   // [PSTAGES][PATCH_BYTES]
  unsigned char* s_b = smem + PSTAGES * PATCH_BYTES;               // [BSTAGES][B_STAGE]
  float* epi = reinterpret_cast<float*>(s_b + BSTAGES * B_STAGE);  // [4][EPI_WARP_FLOATS]
  uint64_t* pfull = reinterpret_cast<uint64_t*>(epi + 4 * EPI_WARP_FLOATS);
  uint64_t* pempty = pfull + PSTAGES;
  uint64_t* bfull = pempty + PSTAGES;
  uint64_t* bempty = bfull + BSTAGES;
  uint64_t* accf_bar = bempty + BSTAGES;
  uint64_t* acce_bar = accf_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acce_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float acc_scale = a.acc_scale != 0.f ? a.acc_scale : 1.f;
  const int H = a.Ho, W = a.Wo;                   // stride 1, pad 1: output extent == input extent
  const int TX = W >> 3, TY = (H + 16 * TR - 1) / (16 * TR);
  const int ntiles = (a.N + NT - 1) / NT;
  const int total = ntiles * a.B * TY * TX;
  const int ncb = a.src_C[0] >> 5;                // 32-channel blocks
  const int C = a.src_C[0];

  if (tid == 32) {
    // with two tile rows each has its own MMA-issuing warp: barriers released by MMA completion count TR arrivals
    for (int i = 0; i < 2; ++i) { mbar_init(&accf_bar[i], TR); mbar_init(&acce_bar[i], 4); }
    for (int i = 0; i < PSTAGES; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], TR); }
    for (int s = 0; s < BSTAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], TR); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> (n tile, image, tile row, tile column): neighbouring CTAs work on neighbouring patches of the same N tile
#define JPB_PTILE(t)                                   \
  const int tx_ = (t) % TX, r1_ = (t) / TX;            \
  const int ty_ = r1_ % TY, r2_ = r1_ / TY;            \
  const int b_ = r2_ % a.B, n0 = (r2_ / a.B) * NT;     \
  const int ox0 = tx_ * 8, oy0 = ty_ * 16 * TR;

  if (warp == 4) {
    // ===================================================== TMA producer: one patch per channel block, one weight tile per (block, tap)
    if (lane == 0) {
      int pring = 0, bring = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        JPB_PTILE(t)
        for (int cb = 0; cb < ncb; ++cb, ++pring) {
          const int ps = pring % PSTAGES;
          mbar_wait(&pempty[ps], (((uint32_t)(pring / PSTAGES)) & 1u) ^ 1u);
          if (a.dbg_skip & 1) mbar_arrive(&pfull[ps]);            // timing experiment (wrong results): no patch traffic
          else {
            mbar_expect_tx(&pfull[ps], patch_tx);
            tma_load_4d(smem_u32(s_patch + ps * PATCH_BYTES), &xmap, &pfull[ps], cb * 32, ox0 - a.patch_org_x, oy0 - a.patch_org_y, b_);
          }
          for (int tap = 0; tap < ntaps; ++tap, ++bring) {
            if (a.dbg_skip & 8) continue;                          // timing experiment: no per-tap weight hand-shake at all
            const int s = bring % BSTAGES;
            mbar_wait(&bempty[s], (((uint32_t)(bring / BSTAGES)) & 1u) ^ 1u);
            if (a.dbg_skip & 2) { mbar_arrive(&bfull[s]); continue; }   // timing experiment: no weight traffic
            mbar_expect_tx(&bfull[s], (uint32_t)B_STAGE);
            tma_load_2d(smem_u32(s_b + s * B_STAGE), &wmap, &bfull[s], tap * C + cb * 32, n0);
          }
        }
      }
    }
  } else if (warp >= 5) {
    // ===================================================== MMA issuer(s): warp 5 owns tile row 0, warp 6 (TR == 2) tile row 1 — two
    // independent instruction streams into the tensor pipe; each whole warp walks the loop converged, one elected lane issues
    const int mr = warp - 5;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr uint32_t a_hi = DESC_HI_SW128(2048), b_hi = DESC_HI_SW128(1024);
    const uint32_t patch0 = smem_u32(s_patch), b0addr = smem_u32(s_b);
    const uint32_t bempty0 = smem_u32(bempty), pempty0 = smem_u32(pempty), accf0 = smem_u32(accf_bar);
    int ps = 0, bs = 0, it = 0;
    uint32_t pph = 0, bph = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acce_bar[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * TR * ACC_STRIDE);
      const int oy0m = ((t / TX) % TY) * 16 * TR;
      const uint32_t lower_ok = (TR > 1 && oy0m + 16 < H) ? 1u : 0u;   // the lower tile of the last super-tile row may lie outside the image
      for (int cb = 0; cb < ncb; ++cb) {
        mbar_wait(&pfull[ps], pph);
        tc_fence_after();
        const uint32_t a_lo0 = (((patch0 + (uint32_t)(ps * PATCH_BYTES)) >> 4) & 0x3FFFu) | DESC_LO_LBO1;
        for (int tap = 0; tap < ntaps; ++tap) {
          const bool nosync = (a.dbg_skip & 8) != 0;
          if (!nosync) {
            mbar_wait(&bfull[bs], bph);
            tc_fence_after();
          }
          // tap (dy, dx): the same patch, start address moved by dy rows (2048 B) and dx pixels (128 B) — patch_tapoff, in 16-byte
          // units.  The swizzle follows the ABSOLUTE shared-memory address bits (measured: tests/test_conv.py patch cases), so the
          // descriptor's base-offset field stays 0.
          const uint32_t a_lo = a_lo0 + (uint32_t)a.patch_tapoff[tap];
          const uint32_t b_lo = (((b0addr + (uint32_t)(bs * B_STAGE)) >> 4) & 0x3FFFu) | DESC_LO_LBO1;
          const uint32_t first = (cb | tap) ? 1u : 0u;
          umma_tf32_k4(tacc + (uint32_t)(mr * ACC_STRIDE), a_lo + (uint32_t)((mr * 16 * 2048) >> 4), b_lo, a_hi, b_hi, idesc, first,
                       mr == 0 ? 1u : lower_ok);
          if (!nosync) umma_commit_elect(bempty0 + (uint32_t)bs * 8u);
          if (tap == ntaps - 1) {
            umma_commit_elect(pempty0 + (uint32_t)ps * 8u);
            if (cb == ncb - 1) umma_commit_elect(accf0 + (uint32_t)buf * 8u);
          }
          if (++bs == BSTAGES) { bs = 0; bph ^= 1u; }
        }
        if (++ps == PSTAGES) { ps = 0; pph ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 0-3)
    const uint32_t sbuf = smem_u32(epi) + (uint32_t)(warp * EPI_WARP_FLOATS) * 4u;
    const uint32_t rowptr = sbuf + 32 * 36 * 4;
    const int c4 = (lane & 7) * 4, r0 = lane >> 3;
    // fused BatchNorm statistics: when the layer has ONE N tile every tile of this CTA covers the same channels, so the column
    // sums stay in registers across tiles and are folded once per CTA (a flush per tile put thousands of same-address double
    // atomics on 2*N accumulators and tripled the kernel time)
    constexpr int RS_PASSES = NT <= 128 ? (NT + 31) / 32 : 1;
    const bool run_ok = a.stats != nullptr && ntiles == 1 && NT <= 128;
    float run_stats[RS_PASSES][8];
#pragma unroll
    for (int pz = 0; pz < RS_PASSES; ++pz)
#pragma unroll
      for (int qz = 0; qz < 8; ++qz) run_stats[pz][qz] = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      JPB_PTILE(t)
      const int buf = it & 1;
      const int row = warp * 32 + lane;
      int nvalid = a.N - n0;
      if (nvalid > NT) nvalid = NT;
      mbar_wait(&accf_bar[buf], ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < TR; ++r) {
      if (r > 0 && oy0 + 16 * r >= H) break;
      if (a.dbg_skip & 4) break;                                   // timing experiment: no epilogue
      const int oy = oy0 + 16 * r + (row >> 3), ox = ox0 + (row & 7);
      const bool rowok = oy < H;
      const long long pix = ((long long)b_ * H + oy) * W + ox;
      __syncwarp();
      // output / residual row pointers of every tile row (0 = row outside the image)
      sts_row(rowptr, lane, rowok ? (unsigned long long)(uintptr_t)(a.out + (size_t)pix * a.N + n0) : 0ull,
              (unsigned long long)(uintptr_t)(a.residual ? a.residual + (size_t)(rowok ? pix : 0) * a.N + n0 : nullptr));
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((buf * TR + r) * ACC_STRIDE);
      for (int j = 0; j < NT && j < nvalid; j += 32) {
        const int col = j + c4;
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias && col < nvalid) bq = *reinterpret_cast<const float4*>(a.bias + n0 + col);
        float v[32];
        tmem_ld32(taddr + (uint32_t)j, v, acc_scale);
        for (int q = 0; q < 32; q += 4)
          sts128(sbuf + (uint32_t)(lane * 36 + q) * 4u, make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]));
        __syncwarp();
        float4 st1 = make_float4(0.f, 0.f, 0.f, 0.f), st2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < nvalid) epi_rows_dispatch(a.act, a.residual != nullptr, sbuf, rowptr, r0, c4, col, bq, st1, st2);
        if (a.stats) {
          if (run_ok) {
#pragma unroll
            for (int pz = 0; pz < RS_PASSES; ++pz)
              if (pz == j / 32) {
                run_stats[pz][0] += st1.x; run_stats[pz][1] += st1.y; run_stats[pz][2] += st1.z; run_stats[pz][3] += st1.w;
                run_stats[pz][4] += st2.x; run_stats[pz][5] += st2.y; run_stats[pz][6] += st2.z; run_stats[pz][7] += st2.w;
              }
          } else {
            bn_stats_flush(a.stats, a.N, n0 + col, col < nvalid, lane, st1, st2);
          }
        }
        __syncwarp();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acce_bar[buf]);
    }
    if (run_ok) {
      // warp partials -> this warp's (idle) staging tile as [2][NT] -> fold the four warps -> 2*N double-precision adds per CTA
#pragma unroll
      for (int pz = 0; pz < RS_PASSES; ++pz) {
        const float4 a1 = make_float4(run_stats[pz][0], run_stats[pz][1], run_stats[pz][2], run_stats[pz][3]);
        const float4 a2 = make_float4(run_stats[pz][4], run_stats[pz][5], run_stats[pz][6], run_stats[pz][7]);
        bn_stats_to_smem(sbuf + (uint32_t)(pz * 32) * 4u, NT, lane, a1, a2);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      for (int tt = tid; tt < 2 * NT; tt += 128) {
        const int cc = tt % NT, which = tt / NT;
        if (cc < a.N) {
          float v = 0.f;
          for (int ww = 0; ww < 4; ++ww)
            v += __int_as_float(lds32(smem_u32(epi) + (uint32_t)(ww * EPI_WARP_FLOATS + which * NT + cc) * 4u));
          atomicAdd(a.stats + (size_t)which * a.N + cc, (double)v);
        }
      }
    }
  }
#undef JPB_PTILE
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int NT, int STAGES, int MINB>
int launch_fwd2(const JpbConvArgs* a, const CUtensorMap& map, cudaStream_t st) {
  const int smem = STAGES * (A_STAGE + NT * BK * 4) + 4 * EPI_WARP_FLOATS * 4 + (NT <= 64 ? 4 * 2 * NT * 4 : 0) + 1024 + 256 + a->ntaps * a->nsrc * BM * 4;
  static int configured = 0;
  if (smem > 227 * 1024) return JPB_ERR_UNSUPPORTED;
  if (smem > configured) {
    if (cudaFuncSetAttribute(conv_tc_fwd2_kernel<NT, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = smem;
  }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int M = a->B * a->Ho * a->Wo;
  const long long total = (long long)((M + BM - 1) / BM) * ((a->N + NT - 1) / NT) * (a->ksplit > 1 ? a->ksplit : 1);
  const int slots = sms * MINB;
  // balanced persistent grid: every CTA walks ceil(total / grid) or one fewer tiles
  const int waves = (int)((total + slots - 1) / slots);
  const int grid = (int)((total + waves - 1) / waves);
  conv_tc_fwd2_kernel<NT, STAGES, MINB><<<grid, 320, smem, st>>>(map, *a);
  return jpb_status();
}

// Experiment switch (tools/bench_conv.py): JPB_CONV_VARIANT selects the pipeline depth / CTAs-per-SM table below.
int conv_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JPB_CONV_VARIANT"); v = e ? atoi(e) : 5; }
  return v;
}

const WgradRowMaps& no_row_maps() {
  static const WgradRowMaps m = {};
  return m;
}

template <int NT, int STAGES, int MINB, bool ROWS = false>
int launch_fwd(const JpbConvArgs* a, const CUtensorMap& map, cudaStream_t st, const WgradRowMaps* xmaps = nullptr) {
  const int smem = STAGES * (A_STAGE + NT * BK * 4) + 1024 + 256 + (ROWS ? 0 : a->ntaps * a->nsrc * BM * 4);
  static int configured = 0;
  if (smem > 227 * 1024) return JPB_ERR_UNSUPPORTED;
  if (smem > configured) {
    if (cudaFuncSetAttribute(conv_tc_fwd_kernel<NT, STAGES, MINB, 1, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = smem;
  }
  const int M = a->B * a->Ho * (ROWS ? a->rows_wv : a->Wo);
  dim3 grid((M + BM - 1) / BM, (a->N + NT - 1) / NT, a->ksplit > 1 ? a->ksplit : 1);
  conv_tc_fwd_kernel<NT, STAGES, MINB, 1, ROWS><<<grid, 192, smem, st>>>(map, xmaps ? *xmaps : no_row_maps(), *a);
  return jpb_status();
}


// CTA-pair launch: 2-clusters along the M tiles (an odd tile count gets one idle partner: rows >= M gather zeros, store nothing)
template <int NT, int STAGES, int MINB, bool ROWS = false>
int launch_fwd_pair(const JpbConvArgs* a, const CUtensorMap& half_map, cudaStream_t st, const WgradRowMaps* xmaps = nullptr) {
  const int smem = STAGES * (A_STAGE + (NT / 2) * BK * 4) + 1024 + 256 + (ROWS ? 0 : a->ntaps * a->nsrc * BM * 4);
  static int configured = 0;
  if (smem > 227 * 1024) return JPB_ERR_UNSUPPORTED;
  if (smem > configured) {
    if (cudaFuncSetAttribute(conv_tc_fwd_kernel<NT, STAGES, MINB, 2, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = smem;
  }
  const int M = a->B * a->Ho * (ROWS ? a->rows_wv : a->Wo);
  const int mt = (M + BM - 1) / BM;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((mt + 1) & ~1, (a->N + NT - 1) / NT, a->ksplit > 1 ? a->ksplit : 1);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, conv_tc_fwd_kernel<NT, STAGES, MINB, 2, ROWS>, half_map, xmaps ? *xmaps : no_row_maps(), *a) != cudaSuccess)
    return jpb_status() ? jpb_status() : JPB_ERR_UNSUPPORTED;
  return jpb_status();
}

// JPB_CONV_PAIR=1 (two pairs per TPC, 3 stages) / 2 (one pair, 4 stages) switches the CTA-pair schedule of the 256-wide tiles on.
// OFF by default — measured on B200 (tools/gpu_r2r.sh, profiles/r2_conv_pair_ab.txt): parity-exact, but not faster (iconv1 forward
// 556 -> 590 us, merge1 266 -> 278 us; one pair per TPC: 789 / 400 us), i.e. the weight stream it halves is not what bounds these
// layers (the same conclusion as removing the weight TMA altogether, profiles/README.md), and the full multi-stream step did not
// finish with it (bench.py ran into its time limit) — kept as a tested schedule for single-stream use and as the starting point of
// a persistent 2-CTA kernel.
int g_conv_pair = -1;
int conv_pair() {
  if (g_conv_pair < 0) { const char* e = getenv("JPB_CONV_PAIR"); g_conv_pair = e ? atoi(e) : 0; }
  return g_conv_pair;
}

template <int NT, int TR, int PSTAGES, int BSTAGES>
int launch_patch(const JpbConvArgs* a, const CUtensorMap& xmap, const CUtensorMap& wmap, cudaStream_t st) {
  constexpr int smem = PSTAGES * (16 * TR + 3) * PATCH_PX * 128 + BSTAGES * NT * BK * 4 + 4 * EPI_WARP_FLOATS * 4 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "patch kernel shared memory");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_tc_patch_kernel<NT, TR, PSTAGES, BSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = true;
  }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const long long total = (long long)((a->N + NT - 1) / NT) * a->B * ((a->Ho + 16 * TR - 1) / (16 * TR)) * (a->Wo / 8);
  const int waves = (int)((total + sms - 1) / sms);
  const int grid = (int)((total + waves - 1) / waves);
  conv_tc_patch_kernel<NT, TR, PSTAGES, BSTAGES><<<grid, 192 + 32 * (TR - 1), smem, st>>>(xmap, wmap, *a);
  return jpb_status();
}

// 3x3 / stride 1 / zero pad 1 / one dense source with C % 32 == 0 / W % 8 == 0: the TMA-patch kernel
int conv2d_patch(const JpbConvArgs* a, cudaStream_t st) {
  if (a->nsrc != 1 || a->src_up[0] || a->reflect || a->in_div || a->scatter || a->ksplit > 1 || (a->src_C[0] & 31) || (a->Wo & 7) ||
      a->Ho != a->src_H[0] || a->Wo > a->src_W[0] || (a->N & 15) || a->patch_ntaps < 1 || a->patch_ntaps > 16 ||
      a->w_cols != a->patch_ntaps * a->src_C[0] || a->patch_halo < 0 || a->patch_halo > 3)
    return JPB_ERR_ARG;
  for (int t = 0; t < a->patch_ntaps; ++t) {        // every tap must stay inside the 16-pixel x (16 + halo)-row patch
    const int dy = (a->patch_tapoff[t] * 16) / 2048, dx = ((a->patch_tapoff[t] * 16) % 2048) / 128;
    if (a->patch_tapoff[t] < 0 || ((a->patch_tapoff[t] * 16) % 128) || dy > a->patch_halo || dx + 8 > PATCH_PX) return JPB_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(a->src[0]) & 15) || (reinterpret_cast<uintptr_t>(a->weight) & 15) || (a->w_row & 3)) return JPB_ERR_ARG;
  EncodeTiledFn enc = get_encode();
  if (!enc) return JPB_ERR_UNSUPPORTED;
  int nt = 16;
  while (nt < a->N && nt < 256) nt <<= 1;
  if (a->nt) nt = a->nt;
  const int C = a->src_C[0], H = a->src_H[0], W = a->src_W[0];
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  // two tiles per CTA (shared weight tiles) when that still leaves at least ~2 super-tiles per SM; a.patch == 2 / 3 force 1 / 2
  const long long tiles2 = (long long)((a->N + nt - 1) / nt) * a->B * ((H + 31) / 32) * (W / 8);
  // measured (tools/bench_conv.py, JPB_CONV_PATCH_TR): two tiles per CTA halve the weight stream but do not change the time — the
  // kernel sits on the cta_group::1 kind::tf32 issue ceiling (~355 TFLOP/s with every load and the epilogue removed), so one tile
  // per CTA (more CTAs, better balance) stays the default; a.patch == 3 forces two
  (void)tiles2;
  int tr = 1;
  if (a->patch == 3 && nt <= 128) tr = 2;
  CUtensorMap xmap, wmap;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)a->B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)PATCH_PX, (cuuint32_t)(16 * tr + a->patch_halo), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->src[0]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return JPB_ERR_ARG;
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)a->w_cols, (cuuint64_t)a->N};
    const cuuint64_t gstr[1] = {(cuuint64_t)a->w_row * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)nt};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->weight), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return JPB_ERR_ARG;
  }
  // patch buffers x weight stages: a tap of a narrow tile is a ~110 ns MMA burst, so the weight ring must cover the TMA latency
  if (tr == 2) {
    switch (nt) {
      case 16: return launch_patch<16, 2, 2, 8>(a, xmap, wmap, st);
      case 32: return launch_patch<32, 2, 2, 8>(a, xmap, wmap, st);
      case 64: return launch_patch<64, 2, 2, 8>(a, xmap, wmap, st);
      default: return launch_patch<128, 2, 2, 4>(a, xmap, wmap, st);
    }
  }
  switch (nt) {
    case 16: return launch_patch<16, 1, 3, 8>(a, xmap, wmap, st);
    case 32: return launch_patch<32, 1, 3, 8>(a, xmap, wmap, st);
    case 64: return launch_patch<64, 1, 3, 8>(a, xmap, wmap, st);
    case 128: return launch_patch<128, 1, 2, 6>(a, xmap, wmap, st);
    default: return launch_patch<256, 1, 2, 4>(a, xmap, wmap, st);
  }
}

}  // namespace

namespace {
// tensor maps of the TMA-row forward / data-gradient kernel: every source as [C, W, H, B], boxes of 32 channels x {32, 31, 1} pixels,
// 128-byte swizzle (K-major operand rows)
int fwd_row_maps(const JpbConvArgs* a, EncodeTiledFn enc, WgradRowMaps* out) {
  static const cuuint32_t px[3] = {32, 31, 1};
  for (int si = 0; si < a->nsrc; ++si) {
    const int C = a->src_C[si], H = a->src_H[si], W = a->src_W[si];
    if (a->src_up[si] || H != a->Hin || W != a->Win || (C & 31) || (reinterpret_cast<uintptr_t>(a->src[si]) & 15)) return JPB_ERR_ARG;
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)a->B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int v = 0; v < 3; ++v) {
      const cuuint32_t box[4] = {32, px[v], 1, 1};
      if (enc(&out->m[si][v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->src[si]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return JPB_ERR_ARG;
    }
  }
  return JPB_OK;
}
}  // namespace

extern "C" int jpb_conv2d_fwd(const JpbConvArgs* a, void* stream) {
  if (!a || !a->weight || !a->table || (!a->out && !a->scatter) || a->nsrc < 1 || a->nsrc > JPB_CONV_MAX_SRC || a->nkb < 1) return JPB_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(a->weight) & 15) || (a->w_row & 3)) return JPB_ERR_ARG;   // TMA: 16-byte aligned base and row pitch
  if (a->patch) return conv2d_patch(a, (cudaStream_t)stream);
  EncodeTiledFn enc = get_encode();
  if (!enc) return JPB_ERR_UNSUPPORTED;
  int nt = 16;
  while (nt < a->N && nt < 256) nt <<= 1;
  if (a->nt) nt = a->nt;
  if (nt != 16 && nt != 32 && nt != 64 && nt != 128 && nt != 256) return JPB_ERR_ARG;
  if (a->scatter && (a->ndst < 1 || a->ndst > JPB_CONV_MAX_SRC)) return JPB_ERR_ARG;
  if (a->ksplit > 1 && (a->bias || a->residual || a->act || a->stats)) return JPB_ERR_ARG;   // partial tiles cannot run the epilogue
  if (a->stats && ((a->N & 3) || a->scatter)) return JPB_ERR_ARG;
  if (a->ntaps < 1 || a->kw < 1 || a->ntaps * a->nsrc > 64) return JPB_ERR_ARG;
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)a->w_cols, (cuuint64_t)a->N};
  const cuuint64_t gstr[1] = {(cuuint64_t)a->w_row * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)nt};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->weight), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return JPB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // pipeline depth x resident CTAs per SM: several shallow CTAs per SM overlap one tile's prologue (offset table, TMEM
  // allocation) and epilogue (TMEM -> registers -> global) with the other tiles' main loops
  const int nts = a->ntaps * a->nsrc;
  const int var = conv_variant();
  if (a->rows) {
    // TMA-row operand (see conv_tc_fwd_kernel): stride 1, dense sources of whole 32-channel blocks, tile raster of 32-pixel segments
    if (a->stride != 1 || a->in_div || a->patch || (a->rows_wv & 31) || a->rows_wv < a->Wo || (a->reflect && (a->Wo != a->Win || a->pad != 1 || a->Win < 33)) ||
        nt < 64)
      return JPB_ERR_ARG;
    WgradRowMaps xm;
    const int rc = fwd_row_maps(a, enc, &xm);
    if (rc != JPB_OK) return rc;
    const long long tiles = (long long)((a->B * a->Ho * a->rows_wv + BM - 1) / BM) * ((a->N + nt - 1) / nt) * (a->ksplit > 1 ? a->ksplit : 1);
    if (conv_pair() && nt == 256 && tiles > 148) {
      // CTA pairs on top of the TMA-row operand (opt-in, JPB_CONV_PAIR): each CTA streams half of the weight tile, three stages
      CUtensorMap hmap;
      const cuuint32_t hbox[2] = {(cuuint32_t)BK, 128u};
      if (enc(&hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->weight), gdim, gstr, hbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return JPB_ERR_ARG;
      return launch_fwd_pair<256, 3, 2, true>(a, hmap, st, &xm);
    }
    switch (nt) {
      case 64: return launch_fwd<64, 4, 2, true>(a, map, st, &xm);
      case 128: return tiles > 148 ? launch_fwd<128, 3, 2, true>(a, map, st, &xm) : launch_fwd<128, 5, 1, true>(a, map, st, &xm);
      default: return (tiles > 148 && a->rows != 2) ? launch_fwd<256, 2, 2, true>(a, map, st, &xm) : launch_fwd<256, 4, 1, true>(a, map, st, &xm);
    }
  }
  if (var == 5) {
    // measured best per N tile (tools/bench_conv.py, B200): narrow tiles -> persistent kernel, two CTAs per SM; wide tiles ->
    // one tile per CTA, two shallow CTAs per SM when there are enough tiles to fill them, else one deep CTA
    const long long tiles = (long long)((a->B * a->Ho * a->Wo + BM - 1) / BM) * ((a->N + nt - 1) / nt) * (a->ksplit > 1 ? a->ksplit : 1);
    switch (nt) {
      case 16: return launch_fwd2<16, 4, 2>(a, map, st);
      case 32: return launch_fwd2<32, 4, 2>(a, map, st);
      case 64: return nts <= 27 ? launch_fwd2<64, 3, 2>(a, map, st) : launch_fwd2<64, 4, 1>(a, map, st);
      case 128: return (nts <= 27 && tiles > 148) ? launch_fwd<128, 3, 2>(a, map, st) : launch_fwd<128, 5, 1>(a, map, st);
      default:
        if (conv_pair() && nts <= 27 && tiles > 148) {
          // CTA pairs: each CTA streams half of the weight tile (box of 128 rows), three stages, two pairs per TPC
          CUtensorMap hmap;
          const cuuint32_t hbox[2] = {(cuuint32_t)BK, 128u};
          if (enc(&hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->weight), gdim, gstr, hbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return JPB_ERR_ARG;
          return conv_pair() == 2 ? launch_fwd_pair<256, 4, 1>(a, hmap, st) : launch_fwd_pair<256, 3, 2>(a, hmap, st);
        }
        return (nts <= 27 && tiles > 148) ? launch_fwd<256, 2, 2>(a, map, st) : launch_fwd<256, 4, 1>(a, map, st);
    }
  }
  if (var >= 3) {   // persistent kernel
    if (var == 4) {
      switch (nt) {
        case 16: return launch_fwd2<16, 4, 2>(a, map, st);
        case 32: return launch_fwd2<32, 4, 2>(a, map, st);
        case 64: return nts <= 27 ? launch_fwd2<64, 3, 2>(a, map, st) : launch_fwd2<64, 4, 1>(a, map, st);
        case 128: return launch_fwd2<128, 4, 1>(a, map, st);
        default: return launch_fwd2<256, 4, 1>(a, map, st);
      }
    }
    switch (nt) {
      case 16: return launch_fwd2<16, 6, 1>(a, map, st);
      case 32: return launch_fwd2<32, 6, 1>(a, map, st);
      case 64: return launch_fwd2<64, 6, 1>(a, map, st);
      case 128: return launch_fwd2<128, 5, 1>(a, map, st);
      default: return launch_fwd2<256, 4, 1>(a, map, st);
    }
  }
  if (var == 0) {
    switch (nt) {
      case 16: return launch_fwd<16, 6, 1>(a, map, st);
      case 32: return launch_fwd<32, 6, 1>(a, map, st);
      case 64: return launch_fwd<64, 6, 1>(a, map, st);
      case 128: return launch_fwd<128, 5, 1>(a, map, st);
      default: return launch_fwd<256, 4, 1>(a, map, st);
    }
  }
  switch (nt) {
    case 16: return nts <= 27 ? launch_fwd<16, 3, 3>(a, map, st) : launch_fwd<16, 6, 1>(a, map, st);
    case 32: return nts <= 27 ? launch_fwd<32, 3, 3>(a, map, st) : launch_fwd<32, 6, 1>(a, map, st);
    case 64: return nts <= 9 ? launch_fwd<64, 3, 3>(a, map, st) : (nts <= 27 ? launch_fwd<64, 4, 2>(a, map, st) : launch_fwd<64, 6, 1>(a, map, st));
    case 128: return nts <= 27 ? launch_fwd<128, 3, 2>(a, map, st) : launch_fwd<128, 5, 1>(a, map, st);
    default:
      if (var == 2 && nts <= 27) return launch_fwd<256, 2, 2>(a, map, st);
      return launch_fwd<256, 4, 1>(a, map, st);
  }
}

namespace {
template <int NT, int STAGES, int MINB, bool ROWS = false>
int launch_wgrad(const JpbConvWgradArgs* a, const CUtensorMap& map, cudaStream_t st, const WgradRowMaps* xmaps = nullptr) {
  constexpr int smem = STAGES * (A_STAGE + NT * BK * 4) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kernel<NT, STAGES, MINB, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return JPB_ERR_UNSUPPORTED;
    configured = true;
  }
  static const WgradRowMaps no_maps = {};
  dim3 grid((a->nchunks + 31) / 32, (a->N + NT - 1) / NT, a->splits);
  conv_tc_wgrad_kernel<NT, STAGES, MINB, ROWS><<<grid, 192, smem, st>>>(map, xmaps ? *xmaps : no_maps, *a);
  return jpb_status();
}

// tensor maps of the TMA-row weight gradient: every source as [C, W, H, B] with boxes of 32 channels x {32, 31, 1} pixels
int wgrad_row_maps(const JpbConvWgradArgs* a, EncodeTiledFn enc, WgradRowMaps* out) {
  static const cuuint32_t px[3] = {32, 31, 1};
  for (int si = 0; si < a->nsrc; ++si) {
    const int C = a->src_C[si], H = a->src_H[si], W = a->src_W[si];
    if (a->src_up[si] || H != a->Hin || W != a->Win) return JPB_ERR_ARG;
    if (C & 31) {                        // gathered source: its maps are never used (placeholder: the first TMA source's)
      if (si == 0) return JPB_ERR_ARG;   // the host lists a TMA source first
      for (int v = 0; v < 3; ++v) out->m[si][v] = out->m[0][v];
      continue;
    }
    if (reinterpret_cast<uintptr_t>(a->src[si]) & 15) return JPB_ERR_ARG;
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)a->B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int v = 0; v < 3; ++v) {
      const cuuint32_t box[4] = {32, px[v], 1, 1};
      if (enc(&out->m[si][v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->src[si]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return JPB_ERR_ARG;
    }
  }
  return JPB_OK;
}
}  // namespace

extern "C" int jpb_conv_set_pair(int mode) {
  if (mode < 0 || mode > 2) return JPB_ERR_ARG;
  g_conv_pair = mode;
  return JPB_OK;
}

extern "C" int jpb_conv2d_wgrad(const JpbConvWgradArgs* a, void* stream) {
  if (!a || !a->dy || !a->dw || !a->table || a->nsrc < 1 || a->nsrc > JPB_CONV_MAX_SRC || a->nchunks < 1 || a->splits < 1) return JPB_ERR_ARG;
  if ((a->N & 3) || (reinterpret_cast<uintptr_t>(a->dy) & 15)) return JPB_ERR_UNSUPPORTED;   // TMA row pitch of dZ must be 16-byte aligned
  EncodeTiledFn enc = get_encode();
  if (!enc) return JPB_ERR_UNSUPPORTED;
  int nt = 32;
  while (nt < a->N && nt < 256) nt <<= 1;
  const long long P = (long long)a->B * a->Ho * a->Wo;
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)a->N, (cuuint64_t)P};
  const int pitch = a->dy_pitch > 0 ? a->dy_pitch : a->N;
  if (pitch < a->N || (pitch & 3)) return JPB_ERR_ARG;
  const cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  JpbConvWgradArgs local = *a;
  local.dz3 = 0;
  if (a->rows && !(a->N & 31)) {
    // dZ as [32 channels, P pixels, N / 32 channel blocks]: one box per step lands all MN groups of the B tile
    const cuuint64_t gdim3[3] = {32, (cuuint64_t)P, (cuuint64_t)(a->N / 32)};
    const cuuint64_t gstr3[2] = {(cuuint64_t)pitch * 4, 128};
    const cuuint32_t box3[3] = {32, 32, (cuuint32_t)(nt / 32)};
    const cuuint32_t estr3[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a->dy), gdim3, gstr3, box3, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return JPB_ERR_ARG;
    local.dz3 = 1;
  } else if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->dy), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return JPB_ERR_ARG;
  a = &local;
  cudaStream_t st = (cudaStream_t)stream;
  const int var = conv_variant();
  if (a->rows) {
    // TMA-row operand: stride-1 "same" convolution over dense full-resolution sources, rows of a multiple of 32 pixels
    if (a->stride != 1 || a->Ho != a->Hin || a->Wo != a->Win || (a->Wo & 31) || !a->gflags || !a->chunk_col || (a->reflect && a->Win < 33)) return JPB_ERR_ARG;
    WgradRowMaps xm;
    const int rc = wgrad_row_maps(a, enc, &xm);
    if (rc != JPB_OK) return rc;
    switch (nt) {
      case 32: return launch_wgrad<32, 4, 2, true>(a, map, st, &xm);
      case 64: return launch_wgrad<64, 4, 2, true>(a, map, st, &xm);
      case 128: return launch_wgrad<128, 3, 2, true>(a, map, st, &xm);
      default: return a->rows == 2 ? launch_wgrad<256, 4, 1, true>(a, map, st, &xm) : launch_wgrad<256, 2, 2, true>(a, map, st, &xm);
    }
  }
  if (var == 0) {
    switch (nt) {
      case 32: return launch_wgrad<32, 6, 1>(a, map, st);
      case 64: return launch_wgrad<64, 6, 1>(a, map, st);
      case 128: return launch_wgrad<128, 5, 1>(a, map, st);
      default: return launch_wgrad<256, 4, 1>(a, map, st);
    }
  }
  switch (nt) {
    case 32: return launch_wgrad<32, 3, 3>(a, map, st);
    case 64: return launch_wgrad<64, 3, 3>(a, map, st);
    case 128: return launch_wgrad<128, 3, 2>(a, map, st);
    default: return var != 1 ? launch_wgrad<256, 2, 2>(a, map, st) : launch_wgrad<256, 4, 1>(a, map, st);
  }
}

#endif  // JPB_HOST_EMU
