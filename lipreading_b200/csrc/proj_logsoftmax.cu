// proj_logsoftmax.cu — output_proj (Linear) fused with allennlp-style masked log-softmax
// (reference src/models/lipreader/better_model.py:92-93).
//
//   logits = hidden @ W^T + b ;  logits += log(mask + 1e-45) ;  out = log_softmax(logits)
//
// C = vocab+1 = 65 is far too narrow for a tensor-core tile to pay, and the op is bound by reading
// `hidden` once (M*K*4 bytes): fp32 SIMT with register blocking, the softmax done in the GEMM
// epilogue so logits never round-trip through HBM.  fp32 end to end => parity 1e-4 on log-probs.
#include "common.cuh"

namespace {

constexpr int kMaxC = 68;      // 4 column groups x 17 classes
constexpr int kCPerThread = 17;
constexpr int kBM = 128;       // rows per CTA (2 per thread-row-slot)
constexpr int kKT = 32;        // K tile
constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
proj_logsoftmax_fwd_kernel(const float* __restrict__ hidden, const float* __restrict__ weight,
                           const float* __restrict__ bias, const float* __restrict__ log_mask,
                           float* __restrict__ out, int M, int K, int C) {
  __shared__ float Hs[kBM][kKT + 1];
  __shared__ float Ws[kKT][kMaxC];
  const int tid = threadIdx.x;
  const int cg = tid & 3;        // class group: classes cg, cg+4, ...
  const int rs = tid >> 2;       // 0..63 : rows rs and rs+64 of the tile
  const int m0 = blockIdx.x * kBM;

  float acc0[kCPerThread], acc1[kCPerThread];
#pragma unroll
  for (int i = 0; i < kCPerThread; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += kKT) {
    // hidden tile: kBM x kKT, coalesced along k
    for (int i = tid; i < kBM * kKT; i += kThreads) {
      int r = i / kKT, kk = i % kKT;
      int m = m0 + r, k = k0 + kk;
      Hs[r][kk] = (m < M && k < K) ? hidden[(size_t)m * K + k] : 0.f;
    }
    // weight tile transposed: Ws[kk][c]
    for (int i = tid; i < kMaxC * kKT; i += kThreads) {
      int c = i / kKT, kk = i % kKT;
      int k = k0 + kk;
      Ws[kk][c] = (c < C && k < K) ? weight[(size_t)c * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kKT; ++kk) {
      float h0 = Hs[rs][kk], h1 = Hs[rs + 64][kk];
#pragma unroll
      for (int i = 0; i < kCPerThread; ++i) {
        float w = Ws[kk][cg + 4 * i];
        acc0[i] = fmaf(h0, w, acc0[i]);
        acc1[i] = fmaf(h1, w, acc1[i]);
      }
    }
    __syncthreads();
  }

  // epilogue: bias + log-mask, row-wise log-softmax across the 4 lanes that share a row
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float* acc = half ? acc1 : acc0;
    int m = m0 + rs + 64 * half;
    float mx = LR_NEG_INF;
#pragma unroll
    for (int i = 0; i < kCPerThread; ++i) {
      int c = cg + 4 * i;
      if (c < C) {
        acc[i] += bias[c] + log_mask[c];
        mx = fmaxf(mx, acc[i]);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kCPerThread; ++i) {
      int c = cg + 4 * i;
      if (c < C) sum += expf(acc[i] - mx);
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    float lse = mx + logf(sum);
    if (m < M) {
#pragma unroll
      for (int i = 0; i < kCPerThread; ++i) {
        int c = cg + 4 * i;
        if (c < C) out[(size_t)m * C + c] = acc[i] - lse;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward 1: d_logits = g - softmax * sum_c(g); d_bias partial sums (one warp per row)
__global__ void __launch_bounds__(256)
logsoftmax_bwd_kernel(const float* __restrict__ grad_lp, const float* __restrict__ log_probs,
                      float* __restrict__ d_logits, float* __restrict__ d_bias, int M, int C) {
  __shared__ float bias_acc[kMaxC * 2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kMaxC * 2; i += 256) bias_acc[i] = 0.f;
  __syncthreads();
  float local[3] = {0.f, 0.f, 0.f};  // classes lane, lane+32, lane+64
  for (int m = blockIdx.x * 8 + warp; m < M; m += gridDim.x * 8) {
    const float* g = grad_lp + (size_t)m * C;
    const float* lp = log_probs + (size_t)m * C;
    float gv[3], s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int c = lane + 32 * i;
      gv[i] = c < C ? g[c] : 0.f;
      s += gv[i];
    }
    s = lr_warp_sum(s);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float d = gv[i] - expf(lp[c]) * s;
        d_logits[(size_t)m * C + c] = d;
        local[i] += d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int c = lane + 32 * i;
    if (c < C) atomicAdd(&bias_acc[c], local[i]);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) atomicAdd(&d_bias[c], bias_acc[c]);
}

// backward 2: d_hidden (M,K) = d_logits (M,C) @ W (C,K).  CTA = 32 rows x 256 columns.
__global__ void __launch_bounds__(256)
proj_bwd_dhidden_kernel(const float* __restrict__ d_logits, const float* __restrict__ weight,
                        float* __restrict__ d_hidden, int M, int K, int C) {
  extern __shared__ float sm[];
  float* dl = sm;                 // [32][kMaxC]
  float* Ws = sm + 32 * kMaxC;    // [C][256]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 256;
  for (int i = tid; i < 32 * kMaxC; i += 256) {
    int r = i / kMaxC, c = i % kMaxC;
    dl[i] = (m0 + r < M && c < C) ? d_logits[(size_t)(m0 + r) * C + c] : 0.f;
  }
  for (int i = tid; i < C * 256; i += 256) {
    int c = i >> 8, kk = i & 255;
    Ws[i] = (k0 + kk < K) ? weight[(size_t)c * K + k0 + kk] : 0.f;
  }
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  for (int c = 0; c < C; ++c) {
    float w = Ws[c * 256 + tid];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc[r] = fmaf(dl[r * kMaxC + c], w, acc[r]);
  }
  int k = k0 + tid;
  if (k < K) {
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (m0 + r < M) d_hidden[(size_t)(m0 + r) * K + k] = acc[r];
  }
}

// backward 3: d_weight (C,K) += d_logits^T @ hidden over a slab of rows (atomic accumulate across
// slabs; the slab count is fixed by the shape so the summation tree is reproducible up to the
// order of <=148 float adds per element).
__global__ void __launch_bounds__(256)
proj_bwd_dweight_kernel(const float* __restrict__ d_logits, const float* __restrict__ hidden,
                        float* __restrict__ d_weight, int M, int K, int C, int rows_per_slab) {
  __shared__ float dl[16][kMaxC];
  const int tid = threadIdx.x;
  const int k = blockIdx.y * 256 + tid;
  const int m_begin = blockIdx.x * rows_per_slab;
  const int m_end = min(M, m_begin + rows_per_slab);
  float acc[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) acc[c] = 0.f;
  for (int mb = m_begin; mb < m_end; mb += 16) {
    __syncthreads();
    for (int i = tid; i < 16 * kMaxC; i += 256) {
      int r = i / kMaxC, c = i % kMaxC;
      dl[r][c] = (mb + r < m_end && c < C) ? d_logits[(size_t)(mb + r) * C + c] : 0.f;
    }
    __syncthreads();
    int rmax = min(16, m_end - mb);
    for (int r = 0; r < rmax; ++r) {
      float h = (k < K) ? hidden[(size_t)(mb + r) * K + k] : 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) acc[c] = fmaf(dl[r][c], h, acc[c]);
    }
  }
  if (k < K) {
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) atomicAdd(&d_weight[(size_t)c * K + k], acc[c]);
  }
}

}  // namespace

extern "C" int lr_proj_logsoftmax_fwd(const float* hidden, const float* weight, const float* bias,
                                      const float* log_mask, float* log_probs, int M, int K, int C,
                                      void* stream) {
  LR_CHECK_ARG(hidden && weight && bias && log_mask && log_probs, "lr_proj_logsoftmax_fwd: null");
  LR_CHECK_ARG(M > 0 && K > 0 && C > 0 && C <= kMaxC, "lr_proj_logsoftmax_fwd: need 0<C<=%d (C=%d)",
               kMaxC, C);
  proj_logsoftmax_fwd_kernel<<<lr_div_up(M, kBM), kThreads, 0, lr_stream(stream)>>>(
      hidden, weight, bias, log_mask, log_probs, M, K, C);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_proj_logsoftmax_bwd(const float* grad_lp, const float* log_probs,
                                      const float* hidden, const float* weight, float* d_logits,
                                      float* d_hidden, float* d_weight, float* d_bias, int M, int K,
                                      int C, void* stream) {
  LR_CHECK_ARG(grad_lp && log_probs && hidden && weight && d_logits && d_hidden && d_weight && d_bias,
               "lr_proj_logsoftmax_bwd: null");
  LR_CHECK_ARG(M > 0 && K > 0 && C > 0 && C <= kMaxC, "lr_proj_logsoftmax_bwd: need 0<C<=%d", kMaxC);
  cudaStream_t st = lr_stream(stream);
  LR_CHECK_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * C, st));
  LR_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, sizeof(float) * (size_t)C * K, st));
  int g1 = lr_div_up(M, 8);
  if (g1 > kNumSMs * 4) g1 = kNumSMs * 4;
  logsoftmax_bwd_kernel<<<g1, 256, 0, st>>>(grad_lp, log_probs, d_logits, d_bias, M, C);
  LR_CHECK_LAUNCH();
  size_t smem2 = (size_t)(32 * kMaxC + C * 256) * sizeof(float);
  LR_CHECK_CUDA(cudaFuncSetAttribute(proj_bwd_dhidden_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  dim3 grid2(lr_div_up(M, 32), lr_div_up(K, 256));
  proj_bwd_dhidden_kernel<<<grid2, 256, smem2, st>>>(d_logits, weight, d_hidden, M, K, C);
  LR_CHECK_LAUNCH();
  int slabs = kNumSMs;
  int rows_per_slab = lr_div_up(M, slabs);
  rows_per_slab = (rows_per_slab + 15) / 16 * 16;
  slabs = lr_div_up(M, rows_per_slab);
  dim3 grid3(slabs, lr_div_up(K, 256));
  proj_bwd_dweight_kernel<<<grid3, 256, 0, st>>>(d_logits, hidden, d_weight, M, K, C, rows_per_slab);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
