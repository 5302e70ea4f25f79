// Shared device helpers for the jperceiver_b200 CUDA library (sm_100a).
//
// The CUDA-core kernels (loss chain, normalisation, pooling, optimizer) are written so that every
// thread-parallel loop is a blockDim-strided loop and every cross-thread reduction goes through the
// helpers below.  That lets the *same source* be compiled as plain C++ with -DJPB_HOST_EMU (one
// "thread" per block, blocks run sequentially) by tests/emu/ — an authoring aid to check kernel
// logic in the GPU-less container.  The emulation build is test infrastructure only: the product
// library (libjpb200.so) is never built with JPB_HOST_EMU and the Python package never loads the
// emulation library.  tcgen05/TMA kernels are excluded from the emulation build.
// With -DJPB_HOST_EMU_MT on top (tests/emu, C++20) a block runs with its REAL thread count: one OS thread per CUDA
// thread, __syncthreads / __syncwarp are barriers, warp shuffles exchange through a per-warp buffer, atomics are
// serialised — so block-level synchronisation, shuffle reductions and shared-memory indexing are exercised too
// (blocks still run one after the other).  Slow: small cases only.
#pragma once

#include <stdint.h>
#include <math.h>

#ifdef JPB_HOST_EMU
// ------------------------------------------------------------------ host emulation shim
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
struct jpb_dim3 { unsigned x, y, z; jpb_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef jpb_dim3 dim3;
#ifdef JPB_HOST_EMU_MT
#include <barrier>
#include <memory>
#include <mutex>
#include <thread>
// inline variables: ONE instance for all translation units of the emulation library (the inline helpers below are merged
// across translation units by the linker and must all see the same state)
inline thread_local jpb_dim3 threadIdx(0, 0, 0);
inline jpb_dim3 blockIdx(0, 0, 0), blockDim(1, 1, 1), gridDim(1, 1, 1);
struct JpbEmuBlock {   // synchronisation state of the block that is running
  std::unique_ptr<std::barrier<>> block;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<unsigned long long> xchg;   // [warp][32] shuffle exchange slots
  std::mutex atomics;
};
inline JpbEmuBlock jpb_emu_blk;
#else
inline jpb_dim3 threadIdx(0, 0, 0), blockIdx(0, 0, 0), blockDim(1, 1, 1), gridDim(1, 1, 1);
#endif
typedef void* cudaStream_t;
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 v = {x, y}; return v; }
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#ifdef JPB_HOST_EMU_MT
#define __syncthreads() jpb_emu_blk.block->arrive_and_wait()
#define __syncwarp() jpb_emu_blk.warp[threadIdx.x >> 5]->arrive_and_wait()
template <typename T>
static inline T jpb_emu_shfl(T v, int src_lane) {   // every live lane of the warp calls it (full-mask shuffles only)
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* slot = &jpb_emu_blk.xchg[(size_t)w * 32];
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  slot[lane] = bits;
  jpb_emu_blk.warp[w]->arrive_and_wait();
  T r = v;
  if (src_lane >= 0 && src_lane < 32 && w * 32 + src_lane < (int)blockDim.x) memcpy(&r, &slot[src_lane], sizeof(T));
  jpb_emu_blk.warp[w]->arrive_and_wait();
  return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int o) { const int l = (threadIdx.x & 31) + o; return jpb_emu_shfl(v, l < 32 ? l : -1); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return jpb_emu_shfl(v, (int)((threadIdx.x & 31) ^ m)); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int l) { return jpb_emu_shfl(v, l & 31); }
#else
#define __syncthreads() ((void)0)
#define __syncwarp() ((void)0)
#endif
#define __threadfence() ((void)0)
#define __ldg(p) (*(p))
inline std::vector<unsigned char> jpb_emu_dynsmem;
#define JPB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(jpb_emu_dynsmem.data())
#ifdef JPB_HOST_EMU_MT
#define JPB_EMU_ATOMIC_GUARD std::lock_guard<std::mutex> jpb_emu_guard_(jpb_emu_blk.atomics)
#else
#define JPB_EMU_ATOMIC_GUARD ((void)0)
#endif
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { JPB_EMU_ATOMIC_GUARD; T o = *p; *p = o + (T)v; return o; }
static inline int atomicMax(int* p, int v) { JPB_EMU_ATOMIC_GUARD; int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { JPB_EMU_ATOMIC_GUARD; int o = *p; if (v < o) *p = v; return o; }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf expf
#define __logf logf
#define __cosf cosf
#define __sinf sinf
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max;
using std::min;
#ifdef JPB_HOST_EMU_MT
// one OS thread per CUDA thread of the block; a thread that returns from the kernel leaves the barriers (as an exited
// CUDA thread no longer takes part in bar.sync)
template <typename Body>
static inline void jpb_emu_run_block(unsigned nthreads, Body body) {
  const unsigned nwarps = (nthreads + 31) / 32;
  jpb_emu_blk.block.reset(new std::barrier<>(nthreads));
  jpb_emu_blk.warp.clear();
  for (unsigned w = 0; w < nwarps; ++w) jpb_emu_blk.warp.emplace_back(new std::barrier<>(std::min(32u, nthreads - 32 * w)));
  jpb_emu_blk.xchg.assign((size_t)nwarps * 32, 0ull);
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t)
    pool.emplace_back([t, &body]() {
      threadIdx = jpb_dim3(t, 0, 0);
      body();
      jpb_emu_blk.warp[t >> 5]->arrive_and_drop();
      jpb_emu_blk.block->arrive_and_drop();
    });
  for (auto& th : pool) th.join();
}
#define JPB_LAUNCH(kernel, grid, block, smem, stream, ...)                                         \
  do {                                                                                             \
    jpb_dim3 g_ = (grid), b_ = (block);                                                            \
    if (jpb_emu_dynsmem.size() < (size_t)(smem) + 16) jpb_emu_dynsmem.resize((size_t)(smem) + 16); \
    gridDim = g_; blockDim = jpb_dim3(b_.x * b_.y * b_.z, 1, 1);                                   \
    for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                                      \
      for (unsigned by_ = 0; by_ < g_.y; ++by_)                                                    \
        for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                                \
          blockIdx = jpb_dim3(bx_, by_, bz_);                                                      \
          jpb_emu_run_block(blockDim.x, [&]() { kernel(__VA_ARGS__); });                           \
        }                                                                                          \
  } while (0)
#else
#define JPB_LAUNCH(kernel, grid, block, smem, stream, ...)                                         \
  do {                                                                                             \
    jpb_dim3 g_ = (grid);                                                                          \
    if (jpb_emu_dynsmem.size() < (size_t)(smem) + 16) jpb_emu_dynsmem.resize((size_t)(smem) + 16); \
    gridDim = g_; blockDim = jpb_dim3(1, 1, 1); threadIdx = jpb_dim3(0, 0, 0);                     \
    for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                                      \
      for (unsigned by_ = 0; by_ < g_.y; ++by_)                                                    \
        for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                                \
          blockIdx = jpb_dim3(bx_, by_, bz_);                                                      \
          kernel(__VA_ARGS__);                                                                     \
        }                                                                                          \
  } while (0)
#endif
#define JPB_LAST_ERROR() 0
#else
// ------------------------------------------------------------------ CUDA build
#include <cuda_runtime.h>
#define JPB_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char jpb_dynsmem_raw_[]; \
  type* name = reinterpret_cast<type*>(jpb_dynsmem_raw_)
#define JPB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define JPB_LAST_ERROR() ((int)cudaGetLastError())
#endif

// status codes returned by every C-ABI entry point (0 = ok; CUDA launch errors are passed through
// as 1000 + cudaError_t)
#define JPB_OK 0
#define JPB_ERR_ARG 1
#define JPB_ERR_UNSUPPORTED 2
static inline int jpb_status() { int e = JPB_LAST_ERROR(); return e == 0 ? JPB_OK : 1000 + e; }

#define JPB_TID (threadIdx.x)
#define JPB_NT (blockDim.x)

// ------------------------------------------------------------------ block-wide sum
// All threads of the block must call it; `scratch` is >= 32 floats/doubles of shared memory.  Thread 0
// gets the total (other threads get a partial they must not use).
template <typename T>
__device__ __forceinline__ T jpb_block_sum(T v, T* scratch) {
#if defined(JPB_HOST_EMU) && !defined(JPB_HOST_EMU_MT)
  (void)scratch;
  return v;
#else
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();  // protect scratch reuse across consecutive calls
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (wid == 0) {
    v = lane < nw ? scratch[lane] : T(0);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
#endif
}

__device__ __forceinline__ int jpb_reflect(int i, int n) {  // ReflectionPad index, |overhang| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
__device__ __forceinline__ int jpb_clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

// ------------------------------------------------------------------ packed fp32 pairs
// sm_100 issues two fp32 operations per lane with one FADD2 / FMUL2 / FFMA2 instruction (operands in aligned 64-bit
// register pairs).  Kernels that are bound by the FMA pipe process two independent values (the two source frames of a
// snippet) per instruction through these wrappers; the host emulation evaluates the two lanes one after the other.
#ifdef JPB_HOST_EMU
static inline float2 jpb_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 jpb_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 jpb_fma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#else
__device__ __forceinline__ float2 jpb_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 jpb_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 jpb_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif
__device__ __forceinline__ float2 jpb_dup2(float v) { return make_float2(v, v); }

// ------------------------------------------------------------------ Philox4x32-10 counter RNG
struct JpbPhilox {
  uint32_t c[4], k[2];
};
__device__ __forceinline__ void jpb_philox_round(uint32_t* c, const uint32_t* k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// four 32-bit words for (seed, stream, counter)
__device__ __forceinline__ void jpb_philox4(uint64_t seed, uint64_t stream, uint64_t ctr, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int r = 0; r < 10; ++r) {
    jpb_philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ float jpb_u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// one standard normal per (seed, stream, ctr)
__device__ __forceinline__ float jpb_randn(uint64_t seed, uint64_t stream, uint64_t ctr) {
  uint32_t r[4];
  jpb_philox4(seed, stream, ctr, r);
  const float u1 = jpb_u01(r[0]), u2 = jpb_u01(r[1]);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}
