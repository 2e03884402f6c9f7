// attention.cu — the attention core of the character decoder for ALL label positions of a clip at once
// (reference src/models/lipreader/better_model.py:195-223, one position per call there):
//
//   scores[b,l,t] = q[b,l,:] . enc[b,t,:]                      ('dot'; 'general' passes q = attn_proj_general(h))
//   w[b,l,:]      = allennlp masked_softmax(scores, t < len_b)  = softmax(scores*m)*m / (sum(softmax(scores*m)*m) + 1e-13)
//   ctx[b,l,:]    = sum_t w[b,l,t] * enc[b,t,:]
//
// (The masked positions take part in the inner softmax with logit 0 — allennlp's formulation, kept.)
// With teacher forcing the decoder's recurrent state never depends on the attention output, so the L decode steps
// of a clip only share `enc`: one CTA per clip parks the clip's encoder states (T x H fp32, 150 KB for T=75, H=512)
// in shared memory ONCE and serves every label position from there — the reference's step loop re-reads them from
// HBM twice per position.  Positions are processed four at a time so a staged encoder row is used 4x per read.
#include "common.cuh"

namespace {

constexpr int kAttnThreads = 256;
constexpr int kLB = 4;                      // label positions per pass

struct AttnParams {
  const float* q;       // (B,L,H)
  const float* enc;     // (B,T,H)
  const int32_t* lens;  // (B)
  int B, L, T, H;
  int enc_in_smem;
  const float* ext_scores;   // (B,L,T) or null: scores computed by the caller ('1_layer_nn', 'concat') instead of q . enc
  float* d_scores;           // backward of the above: (B,L,T) gradient w.r.t. the scores
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? lr_warp_max(v) : lr_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < kAttnThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// stage the clip's encoder states (or just point at them when they do not fit shared memory)
__device__ __forceinline__ const float* stage_enc(const AttnParams& p, int b, float* smem_enc) {
  const float* g = p.enc + (size_t)b * p.T * p.H;
  if (!p.enc_in_smem) return g;
  const int n = p.T * p.H;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(smem_enc)) & 15) == 0) {
    for (int i = threadIdx.x; i < n / 4; i += kAttnThreads)
      reinterpret_cast<float4*>(smem_enc)[i] = reinterpret_cast<const float4*>(g)[i];
  } else {
    for (int i = threadIdx.x; i < n; i += kAttnThreads) smem_enc[i] = g[i];
  }
  return smem_enc;
}

// rows[j][t] = v[j,:] . e[t,:] for the kLB vectors v (in shared memory); one warp per t
__device__ __forceinline__ void dots(const float* e, const float* v, float* rows, int T, int H, int nl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < T; t += kAttnThreads / 32) {
    float acc[kLB];
#pragma unroll
    for (int j = 0; j < kLB; ++j) acc[j] = 0.f;
    const float* er = e + (size_t)t * H;
    for (int h = lane; h < H; h += 32) {
      const float x = er[h];
#pragma unroll
      for (int j = 0; j < kLB; ++j) acc[j] = fmaf(x, v[j * H + h], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kLB; ++j) {
      const float s = lr_warp_sum(acc[j]);
      if (lane == 0 && j < nl) rows[j * T + t] = s;
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(AttnParams p, float* __restrict__ weights, float* __restrict__ zsum, float* __restrict__ ctx) {
  extern __shared__ __align__(16) float asm_[];
  const int T = p.T, H = p.H, b = blockIdx.x, tid = threadIdx.x;
  float* qs = asm_;                      // [kLB][H]
  float* sc = qs + kLB * H;              // [kLB][T]   scores, then weights
  float* red = sc + kLB * T;             // [8]
  float* es = red + 8;                   // [T][H] when staged
  const float* e = stage_enc(p, b, es);
  const int len = min(max(p.lens[b], 0), T);
  __syncthreads();
  for (int l0 = 0; l0 < p.L; l0 += kLB) {
    const int nl = min(kLB, p.L - l0);
    if (p.ext_scores) {
      for (int i = tid; i < nl * T; i += kAttnThreads) sc[i] = p.ext_scores[((size_t)b * p.L + l0) * T + i];
    } else {
      for (int i = tid; i < kLB * H; i += kAttnThreads) {
        const int j = i / H;
        qs[i] = j < nl ? p.q[((size_t)b * p.L + l0 + j) * H + (i - j * H)] : 0.f;
      }
      __syncthreads();
      dots(e, qs, sc, T, H, nl);
    }
    __syncthreads();
    for (int j = 0; j < nl; ++j) {
      float* s = sc + j * T;
      float mx = -INFINITY;
      for (int t = tid; t < T; t += kAttnThreads) mx = fmaxf(mx, t < len ? s[t] : 0.f);
      mx = block_reduce(mx, red, true);
      float sum = 0.f;
      for (int t = tid; t < T; t += kAttnThreads) {
        const float ex = expf((t < len ? s[t] : 0.f) - mx);
        s[t] = ex;
        sum += ex;
      }
      sum = block_reduce(sum, red, false);
      float kept = 0.f;
      for (int t = tid; t < T; t += kAttnThreads) {
        const float u = t < len ? s[t] / sum : 0.f;
        s[t] = u;
        kept += u;
      }
      kept = block_reduce(kept, red, false);
      const float z = kept + 1e-13f;
      for (int t = tid; t < T; t += kAttnThreads) {
        const float w = s[t] / z;
        s[t] = w;
        weights[((size_t)b * p.L + l0 + j) * T + t] = w;
      }
      if (tid == 0) zsum[(size_t)b * p.L + l0 + j] = z;
    }
    __syncthreads();
    // context rows: thread = hidden unit, kLB label positions per pass over the staged rows
    for (int h = tid; h < H; h += kAttnThreads) {
      float acc[kLB];
#pragma unroll
      for (int j = 0; j < kLB; ++j) acc[j] = 0.f;
      for (int t = 0; t < len; ++t) {
        const float x = e[(size_t)t * H + h];
#pragma unroll
        for (int j = 0; j < kLB; ++j) acc[j] = fmaf(sc[j * T + t], x, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < kLB; ++j)
        if (j < nl) ctx[((size_t)b * p.L + l0 + j) * H + h] = acc[j];
    }
    __syncthreads();
  }
}

// backward: d_q (B,L,H), d_enc (B,T,H) from d_ctx (B,L,H) [+ the saved weights and normalisers]
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_kernel(AttnParams p, const float* __restrict__ weights, const float* __restrict__ zsum,
                const float* __restrict__ d_ctx, float* __restrict__ d_q, float* __restrict__ d_enc) {
  extern __shared__ __align__(16) float asm_[];
  const int T = p.T, H = p.H, L = p.L, b = blockIdx.x, tid = threadIdx.x;
  float* vs = asm_;                      // [kLB][H]   d_ctx rows of the pass
  float* dw = vs + kLB * H;              // [kLB][T]   d_w, then d_s, of the pass
  float* red = dw + kLB * T;             // [8]
  const size_t lt4 = ((size_t)L * T + 3) & ~(size_t)3;   // tables padded so `es` stays 16-byte aligned
  float* W = red + 8;                    // [L][T]  saved weights
  float* DS = W + lt4;                   // [L][T]  d_scores
  float* es = DS + lt4;                  // [T][H] when staged
  const float* e = stage_enc(p, b, es);
  const int len = min(max(p.lens[b], 0), T);
  for (int i = tid; i < L * T; i += kAttnThreads) W[i] = weights[(size_t)b * L * T + i];
  __syncthreads();
  for (int l0 = 0; l0 < L; l0 += kLB) {
    const int nl = min(kLB, L - l0);
    for (int i = tid; i < kLB * H; i += kAttnThreads) {
      const int j = i / H;
      vs[i] = j < nl ? d_ctx[((size_t)b * L + l0 + j) * H + (i - j * H)] : 0.f;
    }
    __syncthreads();
    dots(e, vs, dw, T, H, nl);           // d_w[j][t] = d_ctx_j . enc_t
    __syncthreads();
    for (int j = 0; j < nl; ++j) {
      const float* w = W + (size_t)(l0 + j) * T;
      float* g = dw + j * T;
      const float z = zsum[(size_t)b * L + l0 + j];
      // w = u / z, u = p*m:  d_u = (d_w - sum_t d_w w) / z ;  d_p = d_u * m
      float r = 0.f;
      for (int t = tid; t < len; t += kAttnThreads) r += g[t] * w[t];
      r = block_reduce(r, red, false);
      // softmax backward on the unmasked positions: p = w*z there, d_x = p * (d_p - sum_t p d_p), d_s = d_x * m
      float s2 = 0.f;
      for (int t = tid; t < len; t += kAttnThreads) {
        const float dp = (g[t] - r) / z;
        g[t] = dp;
        s2 += w[t] * z * dp;
      }
      s2 = block_reduce(s2, red, false);
      for (int t = tid; t < T; t += kAttnThreads) {
        const float ds = t < len ? w[t] * z * (g[t] - s2) : 0.f;
        g[t] = ds;
        DS[(size_t)(l0 + j) * T + t] = ds;
        if (p.d_scores) p.d_scores[((size_t)b * L + l0 + j) * T + t] = ds;
      }
    }
    __syncthreads();
    // d_q rows: thread = hidden unit
    for (int h = tid; p.q && h < H; h += kAttnThreads) {
      float acc[kLB];
#pragma unroll
      for (int j = 0; j < kLB; ++j) acc[j] = 0.f;
      for (int t = 0; t < len; ++t) {
        const float x = e[(size_t)t * H + h];
#pragma unroll
        for (int j = 0; j < kLB; ++j) acc[j] = fmaf(dw[j * T + t], x, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < kLB; ++j)
        if (j < nl) d_q[((size_t)b * L + l0 + j) * H + h] = acc[j];
    }
    __syncthreads();
  }
  // d_enc[t,h] = sum_l ( w[l,t] d_ctx[l,h] + d_s[l,t] q[l,h] ): thread = hidden unit, label positions in register
  // chunks of 16 (the two [L][T] coefficient tables are in shared memory, broadcast reads)
  constexpr int LC = 16;
  for (int h = tid; h < H; h += kAttnThreads) {
    for (int lc = 0; lc < L; lc += LC) {
      float dc[LC], qv[LC];
#pragma unroll
      for (int j = 0; j < LC; ++j) {
        const bool on = lc + j < L;
        dc[j] = on ? d_ctx[((size_t)b * L + lc + j) * H + h] : 0.f;
        qv[j] = (on && p.q) ? p.q[((size_t)b * L + lc + j) * H + h] : 0.f;
      }
      for (int t = 0; t < T; ++t) {
        float acc = 0.f;
        if (t < len) {
#pragma unroll
          for (int j = 0; j < LC; ++j)
            if (lc + j < L) acc = fmaf(W[(size_t)(lc + j) * T + t], dc[j], fmaf(DS[(size_t)(lc + j) * T + t], qv[j], acc));
        }
        float* o = d_enc + ((size_t)b * T + t) * H + h;
        *o = lc == 0 ? acc : *o + acc;
      }
    }
  }
}

size_t attn_smem(int L, int T, int H, bool bwd, bool stage) {
  const size_t lt4 = ((size_t)L * T + 3) & ~(size_t)3;
  size_t f = (size_t)kLB * H + (size_t)kLB * T + 8 + (bwd ? 2 * lt4 : 0) + (stage ? (size_t)T * H : 0);
  return f * sizeof(float);
}

}  // namespace

static int attn_fwd_launch(const float* q, const float* ext_scores, const float* enc, const int32_t* lens, int B, int L,
                           int T, int H, float* weights, float* zsum, float* ctx, void* stream) {
  LR_CHECK_ARG((q || ext_scores) && enc && lens && weights && zsum && ctx, "lr_attn_fwd: null pointer");
  LR_CHECK_ARG(B > 0 && L > 0 && T > 0 && H > 0, "lr_attn_fwd: bad shape");
  AttnParams p{q, enc, lens, B, L, T, H, 0, ext_scores, nullptr};
  p.enc_in_smem = attn_smem(L, T, H, false, true) <= 220 * 1024;
  const size_t smem = attn_smem(L, T, H, false, p.enc_in_smem);
  LR_CHECK_ARG(smem <= 220 * 1024, "lr_attn_fwd: T/H too large for the staging buffers");
  LR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<<<B, kAttnThreads, smem, lr_stream(stream)>>>(p, weights, zsum, ctx);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

static int attn_bwd_launch(const float* q, const float* enc, const int32_t* lens, const float* weights,
                           const float* zsum, const float* d_ctx, int B, int L, int T, int H, float* d_q,
                           float* d_scores, float* d_enc, void* stream) {
  LR_CHECK_ARG(enc && lens && weights && zsum && d_ctx && d_enc && ((q && d_q) || d_scores), "lr_attn_bwd: null pointer");
  LR_CHECK_ARG(B > 0 && L > 0 && T > 0 && H > 0, "lr_attn_bwd: bad shape");
  AttnParams p{q, enc, lens, B, L, T, H, 0, nullptr, d_scores};
  p.enc_in_smem = attn_smem(L, T, H, true, true) <= 220 * 1024;
  const size_t smem = attn_smem(L, T, H, true, p.enc_in_smem);
  LR_CHECK_ARG(smem <= 220 * 1024, "lr_attn_bwd: L*T too large for the coefficient tables");
  LR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_kernel<<<B, kAttnThreads, smem, lr_stream(stream)>>>(p, weights, zsum, d_ctx, d_q, d_enc);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_attn_fwd(const float* q, const float* enc, const int32_t* lens, int B, int L, int T, int H,
                           float* weights, float* zsum, float* ctx, void* stream) {
  LR_CHECK_ARG(q, "lr_attn_fwd: null pointer");
  return attn_fwd_launch(q, nullptr, enc, lens, B, L, T, H, weights, zsum, ctx, stream);
}

extern "C" int lr_attn_bwd(const float* q, const float* enc, const int32_t* lens, const float* weights,
                           const float* zsum, const float* d_ctx, int B, int L, int T, int H, float* d_q,
                           float* d_enc, void* stream) {
  LR_CHECK_ARG(q && d_q, "lr_attn_bwd: null pointer");
  return attn_bwd_launch(q, enc, lens, weights, zsum, d_ctx, B, L, T, H, d_q, nullptr, d_enc, stream);
}

// Same attention core with the scores supplied by the caller — the '1_layer_nn' and 'concat' score functions of
// better_model.py:204-221 are small Linear layers over [enc ; q] and are evaluated outside; masked softmax and the
// context sums (forward), d_scores and the context part of d_enc (backward) run here.
extern "C" int lr_attn_scores_fwd(const float* scores, const float* enc, const int32_t* lens, int B, int L, int T,
                                  int H, float* weights, float* zsum, float* ctx, void* stream) {
  LR_CHECK_ARG(scores, "lr_attn_scores_fwd: null pointer");
  return attn_fwd_launch(nullptr, scores, enc, lens, B, L, T, H, weights, zsum, ctx, stream);
}

extern "C" int lr_attn_scores_bwd(const float* enc, const int32_t* lens, const float* weights, const float* zsum,
                                  const float* d_ctx, int B, int L, int T, int H, float* d_scores, float* d_enc,
                                  void* stream) {
  LR_CHECK_ARG(d_scores, "lr_attn_scores_bwd: null pointer");
  return attn_bwd_launch(nullptr, enc, lens, weights, zsum, d_ctx, B, L, T, H, nullptr, d_scores, d_enc, stream);
}
