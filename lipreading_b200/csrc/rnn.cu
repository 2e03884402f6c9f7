// rnn.cu — (bi)directional GRU / LSTM / tanh-RNN layer with packed-sequence semantics
// (reference src/models/lipreader/better_model.py:64-89, which sorts,
// packs, calls nn.{LSTM,GRU,RNN}, unpacks and un-sorts).
//
// Packed semantics are reproduced by masking instead of sorting: a clip of length n only updates
// its state while t < n, emits zeros beyond n, the reverse direction starts from a zero state at
// its own frame n-1, and h_n/c_n are the states after each clip's own last step.
//
// v1 structure (fp32 SIMT, parity path): the input projection for all T and both directions is
// one GEMM done by the caller; this file owns the sequential part.  One launch per time step,
// both directions in the same grid.  A CTA owns NU hidden units (all gates) x BC clips; the W_hh
// rows and the previous hidden state are staged through shared memory in K-chunks and every
// thread keeps a (gates x 4 units x 2 clips) register tile, with 128-bit shared loads
// (W broadcast across the warp, h conflict-free by a +4 pitch).  Gate non-linearities, masking,
// the (B,T,D*H) output write and the activations saved for backward are fused in the epilogue.
#include "common.cuh"

namespace {

constexpr int kNU = 16;        // hidden units per CTA
constexpr int kBC = 64;        // clips per CTA (2 per thread)
constexpr int kKC = 256;       // K chunk staged in shared memory
constexpr int kPitch = kKC + 4;
constexpr int kThreads = 128;

template <int MODE> struct Gates;
template <> struct Gates<LR_RNN_TANH> { static constexpr int G = 1, S = 0; };
template <> struct Gates<LR_RNN_GRU>  { static constexpr int G = 3, S = 4; };
template <> struct Gates<LR_RNN_LSTM> { static constexpr int G = 4, S = 5; };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// acc[g*4+j][s] += sum_k W[(g, ug*4+j)][k] * X[bl + 32*s][k] over K, staged in chunks.
//   W row (g, ul) lives at wbase + (g*w_gate_stride + (u0+ul)) * K   (valid iff u0+ul < n_units)
//   X row bb       lives at xbase + (b0+bb) * x_stride               (valid iff b0+bb < B)
template <int NG>
__device__ __forceinline__ void tile_gemm(float (&acc)[NG * 4][2], float* Ws, float* Xs,
                                          const float* __restrict__ wbase, int w_gate_stride,
                                          int u0, int n_units, const float* __restrict__ xbase,
                                          size_t x_stride, int b0, int B, int K) {
  const int tid = threadIdx.x, bl = tid & 31, ug = tid >> 5;
  for (int k0 = 0; k0 < K; k0 += kKC) {
    const int kc = min(kKC, K - k0);
    const int kc4 = kc >> 2;
    for (int i = tid; i < NG * kNU * kc4; i += kThreads) {
      int r = i / kc4, c4 = i - r * kc4;
      int g = r / kNU, ul = r - g * kNU;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u0 + ul < n_units)
        v = *reinterpret_cast<const float4*>(wbase + ((size_t)g * w_gate_stride + u0 + ul) * K + k0 +
                                             c4 * 4);
      *reinterpret_cast<float4*>(Ws + r * kPitch + c4 * 4) = v;
    }
    for (int i = tid; i < kBC * kc4; i += kThreads) {
      int r = i / kc4, c4 = i - r * kc4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < B)
        v = *reinterpret_cast<const float4*>(xbase + (size_t)(b0 + r) * x_stride + k0 + c4 * 4);
      *reinterpret_cast<float4*>(Xs + r * kPitch + c4 * 4) = v;
    }
    __syncthreads();
    const float* x0 = Xs + bl * kPitch;
    const float* x1 = Xs + (bl + 32) * kPitch;
#pragma unroll 2
    for (int c4 = 0; c4 < kc4; ++c4) {
      float4 h0 = *reinterpret_cast<const float4*>(x0 + c4 * 4);
      float4 h1 = *reinterpret_cast<const float4*>(x1 + c4 * 4);
#pragma unroll
      for (int i = 0; i < NG * 4; ++i) {
        int row = (i >> 2) * kNU + ug * 4 + (i & 3);
        float4 w = *reinterpret_cast<const float4*>(Ws + row * kPitch + c4 * 4);
        acc[i][0] = dot4(w, h0, acc[i][0]);
        acc[i][1] = dot4(w, h1, acc[i][1]);
      }
    }
    __syncthreads();
  }
}

struct FwdParams {
  const float* gi; const float* w_hh; const float* b_hh; const int32_t* lens;
  float* hidden; float* saved; float* h_prev; float* h_next; float* c_state;
  int B, T, H, D, step;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) rnn_fwd_step_kernel(FwdParams p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* Xs = smem + G * kNU * kPitch;
  const int tid = threadIdx.x, bl = tid & 31, ug = tid >> 5;
  const int u0 = blockIdx.x * kNU, d = blockIdx.y, b0 = blockIdx.z * kBC;
  const int B = p.B, T = p.T, H = p.H, D = p.D;
  const int tt = d == 0 ? p.step : T - 1 - p.step;

  float acc[G * 4][2];
#pragma unroll
  for (int i = 0; i < G * 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  if (p.step > 0)
    tile_gemm<G>(acc, Ws, Xs, p.w_hh + (size_t)d * G * H * H, H, u0, H,
                 p.h_prev + (size_t)d * B * H, H, b0, B, H);

  const int ub = u0 + ug * 4;  // this thread's 4 consecutive units
  if (ub >= H) return;
  const int nu = min(4, H - ub);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int b = b0 + bl + 32 * s;
    if (b >= B) continue;
    const bool active = tt < p.lens[b];
    const size_t row = (size_t)b * T + tt;
    const float* gi = p.gi + (row * D + d) * (size_t)G * H;
    const float* bh = p.b_hh + (size_t)d * G * H;
    float* hp_ptr = p.h_prev + ((size_t)d * B + b) * H + ub;
    float* hn_ptr = p.h_next + ((size_t)d * B + b) * H + ub;
    float* out = p.hidden + row * (size_t)D * H + (size_t)d * H + ub;
    float* sv = (S > 0) ? p.saved + (row * D + d) * (size_t)S * H + ub : nullptr;
    for (int j = 0; j < nu; ++j) {
      const int u = ub + j;
      const float hp = p.step > 0 ? hp_ptr[j] : 0.f;
      float hnew;
      if (MODE == LR_RNN_GRU) {
        float ghr = acc[0 * 4 + j][s] + bh[u];
        float ghz = acc[1 * 4 + j][s] + bh[H + u];
        float ghn = acc[2 * 4 + j][s] + bh[2 * H + u];
        float r = sigmoidf_(gi[u] + ghr);
        float z = sigmoidf_(gi[H + u] + ghz);
        float n = tanhf(gi[2 * H + u] + r * ghn);
        hnew = (1.f - z) * n + z * hp;
        if (sv) {
          sv[j] = active ? r : 0.f;
          sv[H + j] = active ? z : 0.f;
          sv[2 * H + j] = active ? n : 0.f;
          sv[3 * H + j] = active ? ghn : 0.f;
        }
      } else if (MODE == LR_RNN_LSTM) {
        float* cptr = p.c_state + ((size_t)d * B + b) * H + u;
        float cp = p.step > 0 ? *cptr : 0.f;
        float ig = sigmoidf_(gi[u] + acc[0 * 4 + j][s] + bh[u]);
        float fg = sigmoidf_(gi[H + u] + acc[1 * 4 + j][s] + bh[H + u]);
        float gg = tanhf(gi[2 * H + u] + acc[2 * 4 + j][s] + bh[2 * H + u]);
        float og = sigmoidf_(gi[3 * H + u] + acc[3 * 4 + j][s] + bh[3 * H + u]);
        float cn = fg * cp + ig * gg;
        hnew = og * tanhf(cn);
        *cptr = active ? cn : cp;
        if (sv) {
          sv[j] = active ? ig : 0.f;
          sv[H + j] = active ? fg : 0.f;
          sv[2 * H + j] = active ? gg : 0.f;
          sv[3 * H + j] = active ? og : 0.f;
          sv[4 * H + j] = active ? cn : 0.f;
        }
      } else {
        hnew = tanhf(gi[u] + acc[j][s] + bh[u]);
      }
      hn_ptr[j] = active ? hnew : hp;
      out[j] = active ? hnew : 0.f;
    }
  }
}

struct BwdParams {
  const float* d_hidden; const float* d_h_n; const float* d_c_n; const float* saved;
  const float* hidden; const float* w_hh_t; const int32_t* lens;
  float* d_gi; float* d_gh; float* h_prev_all; float* dh_direct; float* dc_carry;
  int B, T, H, D, step;
};

// One backward time step.  carry(b,u) = dh_direct(b,u) + sum_r d_gh[prev step](b,r) W_hh[r,u]
// (the matmul runs on W_hh^T rows so it is the same tile routine as forward), then the
// element-wise gate backward of this step.
template <int MODE>
__global__ void __launch_bounds__(kThreads) rnn_bwd_step_kernel(BwdParams p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* Xs = smem + kNU * kPitch;
  const int tid = threadIdx.x, bl = tid & 31, ug = tid >> 5;
  const int u0 = blockIdx.x * kNU, d = blockIdx.y, b0 = blockIdx.z * kBC;
  const int B = p.B, T = p.T, H = p.H, D = p.D;
  const int GH = G * H;
  // backward visits time in the opposite order of forward
  const int tt = d == 0 ? T - 1 - p.step : p.step;
  const int tt_later = d == 0 ? tt + 1 : tt - 1;   // the step processed just before this one

  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  if (p.step > 0)
    tile_gemm<1>(acc, Ws, Xs, p.w_hh_t + (size_t)d * H * GH, 0, u0, H,
                 p.d_gh + ((size_t)tt_later * D + d) * GH, (size_t)T * D * GH, b0, B, GH);

  const int ub = u0 + ug * 4;
  if (ub >= H) return;
  const int nu = min(4, H - ub);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int b = b0 + bl + 32 * s;
    if (b >= B) continue;
    const int len = p.lens[b];
    const bool active = tt < len;
    const size_t row = (size_t)b * T + tt;
    const size_t sidx = ((size_t)d * B + b) * H + ub;
    const float* dout = p.d_hidden + row * (size_t)D * H + (size_t)d * H + ub;
    const float* sv = (S > 0) ? p.saved + (row * D + d) * (size_t)S * H + ub : nullptr;
    float* dgi = p.d_gi + (row * D + d) * (size_t)GH + ub;
    float* dgh = p.d_gh + (row * D + d) * (size_t)GH + ub;
    float* hpa = p.h_prev_all + (row * D + d) * (size_t)H + ub;
    // the hidden state that entered this step in forward order
    const int tt_in = d == 0 ? tt - 1 : tt + 1;
    const bool has_prev = (d == 0) ? (tt_in >= 0) : (tt_in < len);
    const float* hin = has_prev ? p.hidden + ((size_t)b * T + tt_in) * (size_t)D * H + (size_t)d * H + ub
                                : nullptr;
    for (int j = 0; j < nu; ++j) {
      float carry;
      if (p.step == 0) carry = p.d_h_n ? p.d_h_n[sidx + j] : 0.f;
      else carry = p.dh_direct[sidx + j] + acc[j][s];
      if (!active) {
        // state passes through untouched; this (b,t) contributes nothing
        p.dh_direct[sidx + j] = carry;
        if (MODE == LR_RNN_LSTM && p.step == 0) p.dc_carry[sidx + j] = p.d_c_n ? p.d_c_n[sidx + j] : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) { dgi[g * H + j] = 0.f; dgh[g * H + j] = 0.f; }
        hpa[j] = 0.f;
        continue;
      }
      const float dh = dout[j] + carry;
      const float hp = hin ? hin[j] : 0.f;
      hpa[j] = hp;
      if (MODE == LR_RNN_GRU) {
        float r = sv[j], z = sv[H + j], n = sv[2 * H + j], ghn = sv[3 * H + j];
        float dn = dh * (1.f - z);
        float dz = dh * (hp - n);
        float dn_pre = dn * (1.f - n * n);
        float dr_pre = dn_pre * ghn * r * (1.f - r);
        float dz_pre = dz * z * (1.f - z);
        dgi[j] = dr_pre;          dgh[j] = dr_pre;
        dgi[H + j] = dz_pre;      dgh[H + j] = dz_pre;
        dgi[2 * H + j] = dn_pre;  dgh[2 * H + j] = dn_pre * r;
        p.dh_direct[sidx + j] = dh * z;
      } else if (MODE == LR_RNN_LSTM) {
        float ig = sv[j], fg = sv[H + j], gg = sv[2 * H + j], og = sv[3 * H + j], cn = sv[4 * H + j];
        // c that entered this step = saved c of the step before it in forward order
        float cp = 0.f;
        if (has_prev) cp = p.saved[((((size_t)b * T + tt_in) * D + d) * (size_t)S + 4) * H + ub + j];
        float dc_in = (p.step == 0) ? (p.d_c_n ? p.d_c_n[sidx + j] : 0.f) : p.dc_carry[sidx + j];
        float tc = tanhf(cn);
        float dog = dh * tc;
        float dc = dc_in + dh * og * (1.f - tc * tc);
        float di_pre = dc * gg * ig * (1.f - ig);
        float df_pre = dc * cp * fg * (1.f - fg);
        float dg_pre = dc * ig * (1.f - gg * gg);
        float do_pre = dog * og * (1.f - og);
        dgi[j] = di_pre;          dgh[j] = di_pre;
        dgi[H + j] = df_pre;      dgh[H + j] = df_pre;
        dgi[2 * H + j] = dg_pre;  dgh[2 * H + j] = dg_pre;
        dgi[3 * H + j] = do_pre;  dgh[3 * H + j] = do_pre;
        p.dc_carry[sidx + j] = dc * fg;
        p.dh_direct[sidx + j] = 0.f;
      } else {
        float h = p.hidden[row * (size_t)D * H + (size_t)d * H + ub + j];
        float dpre = dh * (1.f - h * h);
        dgi[j] = dpre;
        dgh[j] = dpre;
        p.dh_direct[sidx + j] = 0.f;
      }
    }
  }
}

template <int MODE>
int launch_fwd(FwdParams p, int T, cudaStream_t st) {
  constexpr int G = Gates<MODE>::G;
  size_t smem = (size_t)(G * kNU + kBC) * kPitch * sizeof(float);
  LR_CHECK_CUDA(cudaFuncSetAttribute(rnn_fwd_step_kernel<MODE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(lr_div_up(p.H, kNU), p.D, lr_div_up(p.B, kBC));
  float* ha = p.h_prev;
  float* hb = p.h_next;
  for (int step = 0; step < T; ++step) {
    p.step = step;
    p.h_prev = (step & 1) ? hb : ha;
    p.h_next = (step & 1) ? ha : hb;
    rnn_fwd_step_kernel<MODE><<<grid, kThreads, smem, st>>>(p);
    LR_CHECK_LAUNCH();
  }
  return LR_OK;
}

template <int MODE>
int launch_bwd(BwdParams p, int T, cudaStream_t st) {
  size_t smem = (size_t)(kNU + kBC) * kPitch * sizeof(float);
  LR_CHECK_CUDA(cudaFuncSetAttribute(rnn_bwd_step_kernel<MODE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(lr_div_up(p.H, kNU), p.D, lr_div_up(p.B, kBC));
  for (int step = 0; step < T; ++step) {
    p.step = step;
    rnn_bwd_step_kernel<MODE><<<grid, kThreads, smem, st>>>(p);
    LR_CHECK_LAUNCH();
  }
  return LR_OK;
}

}  // namespace

extern "C" int lr_rnn_saved_per_unit(int mode) {
  switch (mode) {
    case LR_RNN_TANH: return 0;
    case LR_RNN_GRU: return 4;
    case LR_RNN_LSTM: return 5;
  }
  return -1;
}

extern "C" size_t lr_rnn_workspace(int mode, int B, int T, int H, int D) {
  (void)mode; (void)T;
  if (B <= 0 || H <= 0 || D <= 0) return 0;
  return (size_t)3 * D * B * H * sizeof(float);   // fwd: h ping/pong + c ; bwd: dh_direct + dc_carry
}

extern "C" int lr_rnn_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh,
                          const int32_t* lens, int B, int T, int H, int D, float* hidden,
                          float* h_n, float* c_n, float* saved, void* workspace, size_t ws_bytes,
                          void* stream) {
  LR_CHECK_ARG(gi && w_hh && b_hh && lens && hidden && h_n && workspace, "lr_rnn_fwd: null pointer");
  LR_CHECK_ARG(mode >= 0 && mode <= 2, "lr_rnn_fwd: bad mode %d", mode);
  LR_CHECK_ARG(B > 0 && T > 0 && H > 0 && (D == 1 || D == 2), "lr_rnn_fwd: bad shape");
  LR_CHECK_ARG(H % 4 == 0, "lr_rnn_fwd: hidden_size must be a multiple of 4 (got %d)", H);
  LR_CHECK_ARG(mode != LR_RNN_LSTM || c_n, "lr_rnn_fwd: LSTM needs c_n");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_fwd: `saved` required for GRU/LSTM");
  size_t need = lr_rnn_workspace(mode, B, T, H, D);
  if (ws_bytes < need) { lr_set_error("lr_rnn_fwd: workspace %zu < %zu", ws_bytes, need); return LR_EWORKSPACE; }
  cudaStream_t st = lr_stream(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  size_t n = (size_t)D * B * H;
  FwdParams p;
  p.gi = gi; p.w_hh = w_hh; p.b_hh = b_hh; p.lens = lens; p.hidden = hidden; p.saved = saved;
  p.h_prev = ws; p.h_next = ws + n; p.c_state = ws + 2 * n;
  p.B = B; p.T = T; p.H = H; p.D = D; p.step = 0;
  int rc;
  if (mode == LR_RNN_GRU) rc = launch_fwd<LR_RNN_GRU>(p, T, st);
  else if (mode == LR_RNN_LSTM) rc = launch_fwd<LR_RNN_LSTM>(p, T, st);
  else rc = launch_fwd<LR_RNN_TANH>(p, T, st);
  if (rc != LR_OK) return rc;
  // after T steps the final state sits in the buffer written by the last step
  const float* h_final = ((T - 1) & 1) ? ws : ws + n;
  LR_CHECK_CUDA(cudaMemcpyAsync(h_n, h_final, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (mode == LR_RNN_LSTM)
    LR_CHECK_CUDA(cudaMemcpyAsync(c_n, ws + 2 * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return LR_OK;
}

extern "C" int lr_rnn_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
                          const float* saved, const float* hidden, const float* w_hh_t,
                          const int32_t* lens, int B, int T, int H, int D, float* d_gi, float* d_gh,
                          float* h_prev_all, void* workspace, size_t ws_bytes, void* stream) {
  LR_CHECK_ARG(d_hidden && hidden && w_hh_t && lens && d_gi && d_gh && h_prev_all && workspace,
               "lr_rnn_bwd: null pointer");
  LR_CHECK_ARG(mode >= 0 && mode <= 2, "lr_rnn_bwd: bad mode %d", mode);
  LR_CHECK_ARG(B > 0 && T > 0 && H > 0 && (D == 1 || D == 2) && H % 4 == 0, "lr_rnn_bwd: bad shape");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_bwd: `saved` required for GRU/LSTM");
  size_t need = lr_rnn_workspace(mode, B, T, H, D);
  if (ws_bytes < need) { lr_set_error("lr_rnn_bwd: workspace %zu < %zu", ws_bytes, need); return LR_EWORKSPACE; }
  float* ws = reinterpret_cast<float*>(workspace);
  size_t n = (size_t)D * B * H;
  BwdParams p;
  p.d_hidden = d_hidden; p.d_h_n = d_h_n; p.d_c_n = d_c_n; p.saved = saved; p.hidden = hidden;
  p.w_hh_t = w_hh_t; p.lens = lens; p.d_gi = d_gi; p.d_gh = d_gh; p.h_prev_all = h_prev_all;
  p.dh_direct = ws; p.dc_carry = ws + n;
  p.B = B; p.T = T; p.H = H; p.D = D; p.step = 0;
  cudaStream_t st = lr_stream(stream);
  if (mode == LR_RNN_GRU) return launch_bwd<LR_RNN_GRU>(p, T, st);
  if (mode == LR_RNN_LSTM) return launch_bwd<LR_RNN_LSTM>(p, T, st);
  return launch_bwd<LR_RNN_TANH>(p, T, st);
}
