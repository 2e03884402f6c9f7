// proj_tc5.cu — output_proj + masked log-softmax on the 5th-generation tensor cores (throughput path of row a14;
// reference src/models/lipreader/better_model.py:91-94).
//
//   out[m, :] = log_softmax(hidden[m, :] @ W^T + bias + log_mask)        hidden (M,K) f32, W (C,K) f32, C <= 80
//
// The fp32 SIMT kernel (proj_logsoftmax.cu) is the parity path: 93 us at M = 19 200, K = 512, bound by its FMA chains
// (ncu: 28 % SM throughput, 6 % warps active, 5 % of DRAM bandwidth).  Here the contraction runs as tcgen05.mma
// kind::tf32 straight on the fp32 operands — no conversion pass, TF32 (10-bit mantissa) products with fp32
// accumulation, i.e. the same numerical class as the bf16 GEMMs around it on the throughput path:
//   * a CTA owns ONE tile of 128 rows (two CTAs share an SM, so M = 19 200 -> 150 tiles is a single wave on 148 SMs
//     and the loads of one tile overlap the epilogue of the other); TMA streams [128 rows x 32 floats] tiles of `hidden` and [80 x 32] tiles of W (rows past
//     C are zero-filled by the TMA unit) through a 4-stage ring, 128-byte hardware swizzle, K-major;
//   * one thread issues 4 MMAs (M = 128, N = 80, K = 8) per stage into an 80-column TMEM accumulator;
//   * four epilogue warps pull the 80 logits of "their" row with tcgen05.ld (thread = row, so the row-wise
//     log-softmax is register-local: no shuffles), add bias + log-mask, and stage the 65 results per row in shared
//     memory so that the tile leaves as one contiguous, 16-byte coalesced block (rows are 260 bytes apart).
// HBM traffic = M*(K + C)*4 bytes, read once / written once; the op is HBM-bound from M ~ 10^4 rows on.
#include "tcgen05.cuh"
#include <string.h>

using namespace lr_tc;

namespace {

constexpr int kStages = 4;
constexpr int kBM = 128;                 // rows per tile
constexpr int kBK = 32;                  // floats per K tile (128-byte rows)
constexpr int kBN = 80;                  // padded classes (N of the MMA: multiple of 16 for M = 128)
constexpr int kThreads = 32 * 6;         // warp 0 TMA producer, warp 1 MMA issuer (+ TMEM), warps 2-5 epilogue
constexpr int kATile = kBM * kBK * 4;    // 16 KB
constexpr int kBTile = kBN * kBK * 4;    // 10 KB
constexpr int kStageBytes = kATile + kBTile;
constexpr int kTmemCols = 128;

struct Tc5Params {
  int M, K, C;
  int n_tiles, n_ktiles;
  const float* bias;
  const float* log_mask;
  float* out;
  uint32_t idesc, desc_hi;
};

enum { BAR_FULL = 0, BAR_EMPTY = kStages, BAR_ACC_FULL = 2 * kStages, BAR_COUNT = 2 * kStages + 2 };

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
proj_logsoftmax_tc5_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w,
                           const Tc5Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  float* stage_out = reinterpret_cast<float*>(base);       // [128][C] staging: ALIASES the ring (used after the last MMA)
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);                             // [kBN] bias + log-mask
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tile = blockIdx.x;

  if (threadIdx.x < kBN)
    bias_s[threadIdx.x] = (int)threadIdx.x < p.C ? p.bias[threadIdx.x] + p.log_mask[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      lr_mbar_init(&bars[BAR_FULL + s], 1);
      lr_mbar_init(&bars[BAR_EMPTY + s], 1);
    }
    lr_mbar_init(&bars[BAR_ACC_FULL], 1);
    lr_fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    for (int kt = 0; kt < p.n_ktiles; ++kt) {
      const int s = kt % kStages;
      lr_mbar_wait_relaxed(&bars[BAR_EMPTY + s], ((kt / kStages) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = base + (size_t)s * kStageBytes;
        lr_mbar_expect_tx(&bars[BAR_FULL + s], kStageBytes);        // OOB rows/columns are zero-filled and counted
        tma_load_2d(st, &map_h, kt * kBK, tile * kBM, &bars[BAR_FULL + s]);
        tma_load_2d(st + kATile, &map_w, kt * kBK, 0, &bars[BAR_FULL + s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    for (int kt = 0; kt < p.n_ktiles; ++kt) {
      const int s = kt % kStages;
      lr_mbar_wait(&bars[BAR_FULL + s], (kt / kStages) & 1);
      if (elect_one()) {
        const uint32_t a_addr = lr_smem_u32(base + (size_t)s * kStageBytes);
        const uint64_t ad = make_desc(a_addr, p.desc_hi), bd = make_desc(a_addr + kATile, p.desc_hi);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k)                            // K = 8 tf32 = 32 bytes = 2 descriptor units
          umma_tf32(tmem_base, ad + 2 * k, bd + 2 * k, p.idesc, (kt > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars[BAR_EMPTY + s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bars[BAR_ACC_FULL]);
    __syncwarp();
  } else {
    // ===== epilogue: thread = row of the tile; two passes over the TMEM row (online max / sum, then the values) =====
    const int q = warp & 3;                         // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const int etid = (warp - 2) * 32 + lane;        // 0..127 for the coalesced copy-out
    lr_mbar_wait_relaxed(&bars[BAR_ACC_FULL], 0);   // every MMA has retired: the ring is free to hold the staging tile
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + (uint32_t)c0, r);
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.C) cm = fmaxf(cm, __uint_as_float(r[i]) + bias_s[c0 + i]);
      const float nm = fmaxf(mx, cm);
      if (nm == -INFINITY) continue;                // (chunk entirely past C)
      float cs = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.C) cs += expf(__uint_as_float(r[i]) + bias_s[c0 + i] - nm);
      sum = sum * expf(mx - nm) + cs;
      mx = nm;
    }
    const float lse = mx + logf(sum);
    float* dst = stage_out + (size_t)row * p.C;                     // odd C: conflict-free column writes
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + (uint32_t)c0, r);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.C) dst[c0 + i] = __uint_as_float(r[i]) + bias_s[c0 + i] - lse;
    }
    named_bar_sync(1, 128);
    // contiguous copy-out of the tile's rows that exist: rows_here * C floats starting at out + tile*128*C
    const int rows_here = min(kBM, p.M - tile * kBM);
    const int n_f = rows_here * p.C;
    float* og = p.out + (size_t)tile * kBM * p.C;
    if ((((size_t)tile * kBM * p.C) & 3) == 0) {
      for (int i = etid * 4; i + 3 < n_f; i += 128 * 4)
        *reinterpret_cast<float4*>(og + i) = *reinterpret_cast<const float4*>(stage_out + i);
      for (int i = (n_f & ~3) + etid; i < n_f; i += 128) og[i] = stage_out[i];
    } else {
      for (int i = etid; i < n_f; i += 128) og[i] = stage_out[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int make_map_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                    uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { lr_set_error("cuTensorMapEncodeTiled entry point not available"); return LR_ECUDA; }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {inner * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { lr_set_error("cuTensorMapEncodeTiled (f32) failed (%d)", (int)r); return LR_ECUDA; }
  return LR_OK;
}

}  // namespace

// 1 when lr_proj_logsoftmax_fwd(variant 2) can take this shape (else the caller stays on variant 0)
extern "C" int lr_proj_tc5_supported(int M, int K, int C) {
  return (M > 0 && K >= 4 && K % 4 == 0 && C > 0 && C <= kBN) ? 1 : 0;
}

int lr_proj_logsoftmax_fwd_tc5(const float* hidden, const float* weight, const float* bias, const float* log_mask,
                               float* log_probs, int M, int K, int C, void* stream) {
  LR_CHECK_ARG(lr_proj_tc5_supported(M, K, C), "lr_proj_logsoftmax_fwd: the tf32 tensor-core variant needs K %% 4 == 0 and C <= %d",
               kBN);
  LR_CHECK_ARG((reinterpret_cast<uintptr_t>(hidden) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0,
               "lr_proj_logsoftmax_fwd: operands must be 16-byte aligned");
  Tc5Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.K = K; p.C = C;
  p.n_tiles = lr_div_up(M, kBM);
  p.n_ktiles = lr_div_up(K, kBK);
  p.bias = bias; p.log_mask = log_mask; p.out = log_probs;
  // D = f32, A = B = tf32 (format 2), both K-major, N = 80, M = 128
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  p.desc_hi = desc_hi_for(128, 8 * 128);
  CUtensorMap map_h, map_w;
  int rc = make_map_2d_f32(&map_h, hidden, (uint64_t)K, (uint64_t)M, kBK, kBM);
  if (rc != LR_OK) return rc;
  rc = make_map_2d_f32(&map_w, weight, (uint64_t)K, (uint64_t)C, kBK, kBN);
  if (rc != LR_OK) return rc;
  const size_t smem_bytes = (size_t)kStages * kStageBytes + BAR_COUNT * 8 + 16 + kBN * 4 + 1024;     // ~106 KB: two CTAs per SM
  LR_CHECK_CUDA(cudaFuncSetAttribute(proj_logsoftmax_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
  proj_logsoftmax_tc5_kernel<<<p.n_tiles, kThreads, smem_bytes, lr_stream(stream)>>>(map_h, map_w, p);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
