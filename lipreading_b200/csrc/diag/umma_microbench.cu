// umma_microbench.cu — measures the issue-to-retire cost of back-to-back tcgen05.mma of a given shape
// on this GPU (operands resident in shared memory, one accumulator or a ring of accumulators).  Used to
// choose tile orientation / N per layer from measurement instead of the nominal rate (tools/umma_table.py).
#include "../tcgen05.cuh"
#include <string.h>

using namespace lr_tc;

namespace {

__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(uint32_t idesc, uint32_t desc_hi_a, uint32_t desc_hi_b, int a_step16, int b_step16, int n_acc,
                  int acc_cols, int iters, int a_tiles, int a_tile16, int a_shift16, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // deterministic small values (bf16 1.0 = 0x3f80): avoids NaN-path surprises
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { lr_mbar_init(&bar, 1); lr_fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  lr_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const uint64_t a0 = make_desc(lr_smem_u32(base), desc_hi_a) + (uint64_t)a_shift16;   // row-shifted start
    const uint64_t b0 = make_desc(lr_smem_u32(base + 64 * 1024), desc_hi_b);
    long long t0 = 0, t1 = 0;
    // descriptors / accumulators precomputed: the timed loop is nothing but 8 tcgen05.mma per trip
    uint64_t av[8], bv[8];
    uint32_t dv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dv[k] = tmem_base + (uint32_t)((k % n_acc) * acc_cols);
      av[k] = a0 + (uint64_t)((k % a_tiles) * a_tile16) + (uint64_t)((k & 1) * a_step16);
      bv[k] = b0 + (uint64_t)((k & 1) * b_step16);
    }
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(dv[k], av[k], bv[k], idesc, 0u);   // initialise accumulators
      t0 = clock64();
      for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(dv[k], av[k], bv[k], idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    lr_mbar_wait(&bar, 0);
    if (elect_one()) { t1 = clock64(); out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// Pattern variant: the 8 MMAs of a trip have individual N (through their instruction descriptors) and
// accumulator column offsets — measures what differently shaped MMAs on overlapping / disjoint accumulator
// ranges cost when issued back to back (the kt-stacked conv orientation).
struct PatternArgs { uint32_t idesc[8]; uint32_t dcol[8]; uint32_t boff16[8]; };

__global__ void __launch_bounds__(128, 1)
umma_pattern_kernel(PatternArgs pa, uint32_t desc_hi, int iters, int commit_every, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { lr_mbar_init(&bar, 1); lr_mbar_init(&bar2, 1); lr_fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  lr_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const uint64_t a0 = make_desc(lr_smem_u32(base), desc_hi);
    const uint64_t b0 = make_desc(lr_smem_u32(base + 48 * 1024), desc_hi);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      // initialise all 512 columns with two N = 256 MMAs
      const uint32_t id256 = (pa.idesc[0] & ~(0x3Fu << 17)) | ((256u >> 3) << 17);
      umma_bf16(tmem_base, a0, b0, id256, 0u);
      umma_bf16(tmem_base + 256, a0, b0, id256, 0u);
      t0 = clock64();
      for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + pa.dcol[k], a0 + (uint64_t)((k & 3) * 512), b0 + pa.boff16[k], pa.idesc[k], 1u);
        // a tcgen05.commit every `commit_every` MMAs onto a barrier nobody waits for (what a stage release costs)
        if (commit_every > 0 && ((it + 8) % commit_every) == 0) umma_commit(&bar2);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    lr_mbar_wait(&bar, 0);
    if (elect_one()) { t1 = clock64(); out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// Issue-loop variant: the conv kernels' steady-state loop shape — per "tile" n_ops chunks, each KS MMAs of one
// shape, descriptors advanced by loop-carried uniform adds (a += a_step, d += d_step) — with `threads` threads in
// the block (the others wait at the final barrier).  Measures what a single issuing thread sustains when the
// descriptors are computed in the loop instead of sitting in registers.
template <int KS>
__global__ void __launch_bounds__(288, 1)
umma_issue_kernel(uint32_t idesc, uint32_t desc_hi, int tiles, int n_ops, uint32_t a_step16, uint32_t d_step,
                  int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3f803f80u + (uint32_t)i * 2654435761u % 7u;
  if (threadIdx.x == 0) { lr_mbar_init(&bar, 1); lr_fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  lr_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    const uint64_t a0 = make_desc(lr_smem_u32(base), desc_hi);
    const uint64_t b0 = make_desc(lr_smem_u32(base + 64 * 1024), desc_hi);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint32_t id256 = (idesc & ~(0x3Fu << 17)) | ((256u >> 3) << 17);
      umma_bf16(tmem_base, a0, b0, id256, 0u);
      umma_bf16(tmem_base + 256, a0, b0, id256, 0u);
      t0 = clock64();
      if (variant == 0) {
        for (int t = 0; t < tiles; ++t) {
          uint64_t a = a0 + (uint64_t)((t & 3) * 4);
          uint32_t d = tmem_base;
          for (int c = 0; c < n_ops; ++c) {
#pragma unroll
            for (int k = 0; k < KS; ++k) umma_bf16(d, a + 2 * k, b0 + 2 * k, idesc, 1u);
            a += a_step16; d += d_step;
          }
        }
      } else {
        // same MMAs, descriptors constant (the stand-alone rate)
        for (int t = 0; t < tiles; ++t)
          for (int c = 0; c < n_ops; ++c) {
#pragma unroll
            for (int k = 0; k < KS; ++k) umma_bf16(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
          }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    lr_mbar_wait(&bar, 0);
    if (elect_one()) { t1 = clock64(); out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

// Cycles for tiles * n_ops * KS MMAs (M = 128, N, K-major 64-byte rows) issued by one thread of a `threads`-thread
// block from a loop with loop-carried descriptor adds (variant 0) or constant descriptors (variant 1).
extern "C" long long lr_umma_issue_bench(int N, int KS, int tiles, int n_ops, int a_step_bytes, int d_step, int threads,
                                         int variant, void* stream) {
  if (N % 16 != 0 || N < 16 || N > 256 || (KS != 1 && KS != 2) || threads < 64 || threads > 288 || n_ops * d_step + N > 512) {
    lr_set_error("lr_umma_issue_bench: bad args");
    return -1;
  }
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long* d_out = nullptr;
  if (cudaMalloc(&d_out, sizeof(long long)) != cudaSuccess) return -2;
  const size_t smem = 100 * 1024;
  cudaStream_t st = lr_stream(stream);
  if (KS == 1) {
    cudaFuncSetAttribute(umma_issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_issue_kernel<1><<<1, threads, smem, st>>>(idesc, desc_hi_for(64, 8 * 64), tiles, n_ops, (uint32_t)a_step_bytes >> 4,
                                                   (uint32_t)d_step, variant, d_out);
  } else {
    cudaFuncSetAttribute(umma_issue_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_issue_kernel<2><<<1, threads, smem, st>>>(idesc, desc_hi_for(64, 8 * 64), tiles, n_ops, (uint32_t)a_step_bytes >> 4,
                                                   (uint32_t)d_step, variant, d_out);
  }
  long long h = -3;
  if (cudaStreamSynchronize(st) == cudaSuccess) cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  return h;
}

// n[8] = N of each of the 8 MMAs of a trip (M = 128, K-major, 64-byte rows), dcol[8] = accumulator column of each,
// bblk[8] = first 8-row group of the B operand.  Returns SM cycles for `iters` MMAs.
extern "C" long long lr_umma_pattern_bench(const int* n, const int* dcol, const int* bblk, int iters,
                                           int commit_every, void* stream) {
  PatternArgs pa;
  for (int k = 0; k < 8; ++k) {
    if (n[k] % 16 != 0 || n[k] < 16 || n[k] > 256 || dcol[k] < 0 || dcol[k] + n[k] > 512) {
      lr_set_error("lr_umma_pattern_bench: bad pattern");
      return -1;
    }
    pa.idesc[k] = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n[k] >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    pa.dcol[k] = (uint32_t)dcol[k];
    pa.boff16[k] = (uint32_t)(bblk[k] * 8 * 64) >> 4;
  }
  long long* d_out = nullptr;
  if (cudaMalloc(&d_out, sizeof(long long)) != cudaSuccess) return -2;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(umma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st = lr_stream(stream);
  umma_pattern_kernel<<<1, 128, smem, st>>>(pa, desc_hi_for(64, 8 * 64), iters, commit_every, d_out);
  long long h = -3;
  if (cudaStreamSynchronize(st) == cudaSuccess) cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  return h;
}

// a_major / b_major: 0 = K-major, 1 = MN-major.  row_bytes_* select the swizzle (32/64/128).
// a_tiles > 1 cycles the A descriptor over that many different tiles (like the conv kernels' per-frame
// chunks); n_acc accumulators of acc_cols columns are used round-robin.  Returns SM cycles for `iters` MMAs.
extern "C" long long lr_umma_microbench(int M, int N, int row_bytes_a, int row_bytes_b, int a_major, int b_major,
                                        int n_acc, int a_tiles, int iters, int a_shift_rows, void* stream) {
  if (!(M == 64 || M == 128) || N % 16 != 0 || N < 16 || N > 256) { lr_set_error("lr_umma_microbench: bad shape"); return -1; }
  int acc_cols = N < 32 ? 32 : N;
  if (n_acc < 1) n_acc = 1;
  if (n_acc * acc_cols > 512) n_acc = 512 / acc_cols;
  uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_major & 1) << 15) | ((uint32_t)(b_major & 1) << 16) |
                   ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t hi_a = desc_hi_for(row_bytes_a, (uint32_t)(8 * row_bytes_a));
  uint32_t hi_b = desc_hi_for(row_bytes_b, (uint32_t)(8 * row_bytes_b));
  // K-major: the second k-step is 32 B further along the row; MN-major: 16 rows further
  int a_step16 = a_major ? (16 * row_bytes_a) >> 4 : 2;
  int b_step16 = b_major ? (16 * row_bytes_b) >> 4 : 2;
  if (row_bytes_a == 32 && !a_major) a_step16 = 0;
  if (row_bytes_b == 32 && !b_major) b_step16 = 0;
  int a_tile16 = (M * row_bytes_a) >> 4;
  if (a_tiles < 1) a_tiles = 1;
  while (a_tiles > 1 && a_tiles * M * row_bytes_a + 4096 > 60 * 1024) --a_tiles;
  long long* d_out = nullptr;
  if (cudaMalloc(&d_out, sizeof(long long)) != cudaSuccess) return -2;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st = lr_stream(stream);
  umma_bench_kernel<<<1, 128, smem, st>>>(idesc, hi_a, hi_b, a_step16, b_step16, n_acc, acc_cols, iters, a_tiles,
                                          a_tile16, (a_shift_rows * row_bytes_a) >> 4, d_out);
  long long h = -3;
  if (cudaStreamSynchronize(st) == cudaSuccess) cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  return h;
}
