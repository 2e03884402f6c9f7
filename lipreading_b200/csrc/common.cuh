// common.cuh — shared device/host helpers for liblr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <atomic>

#include "../../include/lr_b200.h"

// ----------------------------------------------------------------------------------------
// host side: error text + launch accounting
// ----------------------------------------------------------------------------------------
void lr_set_error(const char* fmt, ...);
void lr_count_launch(int n = 1);

#define LR_CHECK_ARG(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      lr_set_error(__VA_ARGS__);                 \
      return LR_EINVAL;                          \
    }                                            \
  } while (0)

#define LR_CHECK_CUDA(expr)                                                                 \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      lr_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return LR_ECUDA;                                                                      \
    }                                                                                       \
  } while (0)

#define LR_CHECK_LAUNCH()                                                                   \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      lr_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return LR_ECUDA;                                                                      \
    }                                                                                       \
    lr_count_launch();                                                                      \
  } while (0)

static inline cudaStream_t lr_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int lr_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// B200: 148 SMs; grids for streaming kernels are sized in multiples of this.
constexpr int kNumSMs = 148;

// ----------------------------------------------------------------------------------------
// device side
// ----------------------------------------------------------------------------------------
#define LR_NEG_INF (-INFINITY)

__device__ __forceinline__ uint32_t lr_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float lr_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float lr_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log(exp(a)+exp(b)) with -inf handling.  Accurate expf/logf on purpose: the CTC recursions chain these
// T times and the fast intrinsics' error reaches the 1e-4 parity bar at T ~ 300 (measured).
__device__ __forceinline__ float lr_lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == LR_NEG_INF) return LR_NEG_INF;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lr_lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == LR_NEG_INF) return LR_NEG_INF;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// fast variants (MUFU ex2/lg2 based): 4-5x fewer instructions; accurate enough for chains of <= ~128 steps
__device__ __forceinline__ float lr_lse2_fast(float a, float b) {
  float m = fmaxf(a, b);
  if (m == LR_NEG_INF) return LR_NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m));
}
__device__ __forceinline__ float lr_lse3_fast(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == LR_NEG_INF) return LR_NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// ---- mbarrier / bulk-async (TMA) primitives ---------------------------------------------
__device__ __forceinline__ void lr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lr_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lr_fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lr_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void lr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lr_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void lr_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lr_smem_u32(bar)) : "memory");
}
// Bounded wait: a lost transaction must become a trap (error return), never a hung GPU.  The fast path (phase
// already complete) is a single try_wait: no clock read, nothing else between two MMA issue bursts.
__device__ __forceinline__ void lr_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = lr_smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == 0) t0 = clock64();
    else if ((spins & 255u) == 0 && clock64() - t0 > 4000000000LL) break;     // ~2 s at 2 GHz
  }
  printf("lr_b200: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
         blockIdx.y, threadIdx.x, addr, parity);
  __trap();
}
// Same for waiters that are not on the critical path (producer waiting for a free slot, epilogue waiting for an
// accumulator): back off between polls so the spinning warp does not take issue slots from the MMA-issuing warp
// that shares its scheduler.
__device__ __forceinline__ void lr_mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t addr = lr_smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (int spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(100);
    if (spins == 0) t0 = clock64();
    else if ((spins & 1023) == 0 && clock64() - t0 > 4000000000LL) break;
  }
  printf("lr_b200: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
         blockIdx.y, threadIdx.x, addr, parity);
  __trap();
}
// 1-D bulk copy global -> shared (SASS: UBLKCP). 16-byte aligned src/dst/size.
__device__ __forceinline__ void lr_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(lr_smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(lr_smem_u32(bar))
      : "memory");
}

// streaming 128-bit global access (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ uint4 lr_ldg_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void lr_stg_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
