// rnn_grid.cu — grid-persistent recurrent kernels for hidden sizes whose W_hh does not fit a thread-block cluster
// (the BiLSTM-768 of config/archive/experiments/ecd/*, reference
// src/models/lipreader/better_model.py:47-49,74).
//
// rnn_cluster.cu keeps a direction's W_hh (bf16) in the shared memory of ONE 8-CTA cluster and exchanges the state over
// DSMEM.  LSTM-768 has 4*768*768 bf16 = 4.7 MB per direction: a 16-CTA cluster (the hardware maximum) would need 295 KB
// per CTA.  Here a direction's units are spread over H/16 CTAs of a COOPERATIVE launch (48 per direction for H = 768:
// 96 CTAs, all co-resident on the 148 SMs):
//   * CTA r owns the hidden units [16r, 16r+16) of every gate; its W_hh rows (forward: G*16 x H, backward: the 16 x G*H
//     slice of W_hh^T) stay in shared memory as bf16 for the whole sequence (99 KB for LSTM-768);
//   * the per-step operand every CTA needs — h_{t-1} of ALL units (forward) or the gate gradients of ALL units
//     (backward) — lives in a double-buffered bf16 exchange buffer in global memory (it stays in L2): each CTA writes
//     its 16-unit slice, one barrier per step among the CTAs of a direction (an atomic counter; the two directions
//     never wait for each other), then every CTA streams the whole operand through shared memory in 256-column
//     chunks (cp.async.cg, double-buffered) under the mma.sync m16n8k16 of the previous chunk;
//   * 64 clips per pass (8 warps x one n8 tile); larger batches run as consecutive passes inside the same launch;
//   * gate math, length masking, outputs and saved activations are the register-local code of rnn_cluster.cu
//     (same fragment layout: thread = 4 (unit, clip) pairs with all their gates).
// One launch per pass direction pair instead of 2 x 75 per-step launches.
#include "common.cuh"
#include <cooperative_groups.h>
#include <string.h>

namespace {

constexpr int kUH = 16;        // hidden units per CTA (one m16 tile per gate)
constexpr int kBS = 64;        // clips per pass
constexpr int kKC = 256;       // operand columns per streamed chunk
constexpr int kPad = 8;
constexpr int kThreads = 32 * (kBS / 8);

template <int MODE> struct Gates;
template <> struct Gates<LR_RNN_TANH> { static constexpr int G = 1, S = 0; };
template <> struct Gates<LR_RNN_GRU>  { static constexpr int G = 3, S = 4; };
template <> struct Gates<LR_RNN_LSTM> { static constexpr int G = 4, S = 5; };

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 1.f - 2.f / (1.f + __expf(2.f * x)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_cg16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Barrier among the `n_ctas` CTAs of one direction: counter `ctr` only grows, `target` = arrivals expected so far.
// Bounded: a CTA that never arrives becomes a trap (an error return), not a hung GPU.
__device__ __forceinline__ void dir_barrier(unsigned int* ctr, unsigned int target) {
  __threadfence();                 // this thread's exchange-buffer stores are visible device-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ctr, 1u);
    long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (spins == 0) t0 = clock64();
      else if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
        printf("lr_b200: rnn_grid barrier timed out (block %d, counter %u, target %u)\n", blockIdx.x, v, target);
        __trap();
      }
    }
  }
  __syncthreads();
}

// Streams the [kBS x K] bf16 operand `src` (row pitch K) through two shared-memory chunk buffers and accumulates
//   acc[a] += A_a[16 x K] . src^T   for the NA row blocks of `a_addr` (row pitch a_pitch, NA blocks a_block bytes apart).
// Every thread of the CTA calls it (cp.async + __syncthreads inside).
template <int NA>
__device__ __forceinline__ void streamed_mma(float (&acc)[NA][4], const __nv_bfloat16* src, int K, uint32_t a_addr,
                                             uint32_t a_block, int a_pitch, uint32_t x_addr, int lane, int warp) {
  constexpr int xp = kKC + kPad;                                   // chunk row pitch (elements)
  constexpr uint32_t xbuf = (uint32_t)(kBS * xp * 2);
  const int n_ch = K / kKC;
  const int tid = threadIdx.x;
  auto load = [&](int c) {
    const uint32_t dst = x_addr + (uint32_t)(c & 1) * xbuf;
    // kBS rows x kKC*2 bytes = kBS * 32 chunks of 16 bytes
    for (int i = tid; i < kBS * (kKC / 8); i += kThreads) {
      const int row = i / (kKC / 8), ch = i - row * (kKC / 8);
      cp_async_cg16(dst + (uint32_t)((row * xp + ch * 8) * 2), src + (size_t)row * K + (size_t)c * kKC + ch * 8);
    }
    cp_async_commit();
  };
  float acc2[NA][4];
#pragma unroll
  for (int a = 0; a < NA; ++a) acc2[a][0] = acc2[a][1] = acc2[a][2] = acc2[a][3] = 0.f;
  const uint32_t a_lane = (uint32_t)((((lane & 15)) * a_pitch + (lane >> 4) * 8) * 2);
  const uint32_t b_lane = (uint32_t)(((warp * 8 + (lane & 7)) * xp + ((lane >> 3) & 1) * 8) * 2);
  load(0);
  for (int c = 0; c < n_ch; ++c) {
    if (c + 1 < n_ch) {
      load(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t xb = x_addr + (uint32_t)(c & 1) * xbuf + b_lane;
    const uint32_t ab = a_addr + a_lane + (uint32_t)(c * kKC * 2);
#pragma unroll 2
    for (int ks = 0; ks < kKC / 16; ks += 2) {
      uint32_t b0, b1, c0, c1;
      ldsm_x2(xb + ks * 32, b0, b1);
      ldsm_x2(xb + ks * 32 + 32, c0, c1);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        uint32_t a0, a1, a2, a3, e0, e1, e2, e3;
        ldsm_x4(ab + (uint32_t)a * a_block + ks * 32, a0, a1, a2, a3);
        ldsm_x4(ab + (uint32_t)a * a_block + ks * 32 + 32, e0, e1, e2, e3);
        mma_bf16(acc[a], a0, a1, a2, a3, b0, b1);
        mma_bf16(acc2[a], e0, e1, e2, e3, c0, c1);
      }
    }
    __syncthreads();                                               // the buffer is refilled two chunks later
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    acc[a][0] += acc2[a][0]; acc[a][1] += acc2[a][1]; acc[a][2] += acc2[a][2]; acc[a][3] += acc2[a][3];
  }
}

struct GFwd {
  const float* gi; const float* w_hh; const float* b_hh; const int32_t* lens;
  float* hidden; float* saved; float* h_n; float* c_n;
  __nv_bfloat16* xbuf;           // [2][D][kBS][H] exchange buffer (state)
  unsigned int* ctr;             // [D] barrier counters (zeroed by the host call)
  int B, T, H, D, NC;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) rnn_grid_fwd_kernel(GFwd p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, D = p.D, NC = p.NC;
  const int pitch = H + kPad;
  __nv_bfloat16* W_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);                    // [G*16][pitch]
  __nv_bfloat16* x_s = W_s + (size_t)G * kUH * pitch;                                 // [2][kBS][kKC + kPad]
  __nv_bfloat16* stage = x_s + (size_t)2 * kBS * (kKC + kPad);                        // [kBS][16]
  const int d = blockIdx.x / NC, r = blockIdx.x - d * NC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u_base = r * kUH;

  const float* Wd = p.w_hh + (size_t)d * G * H * H;
  for (int i = tid; i < G * kUH * (H / 4); i += kThreads) {
    int row = i / (H / 4), c4 = i - row * (H / 4);
    int g = row / kUH, ul = row - g * kUH;
    float4 v = *reinterpret_cast<const float4*>(Wd + ((size_t)g * H + u_base + ul) * H + c4 * 4);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(W_s + (size_t)row * pitch + c4 * 4) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  __syncthreads();

  const uint32_t W_addr = lr_smem_u32(W_s), x_addr = lr_smem_u32(x_s);
  unsigned int* ctr = p.ctr + d;
  unsigned int arrivals = 0;                                     // barriers passed so far (same on every CTA of d)
  const size_t xdir = (size_t)kBS * H;                           // one direction's buffer (elements)

  for (int b0 = 0; b0 < B; b0 += kBS) {
    int ug[4], bl[4], bg[4], len[4];
    float bh[G][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ug[i] = u_base + (lane >> 2) + 8 * (i >> 1);
      bl[i] = warp * 8 + 2 * (lane & 3) + (i & 1);
      bg[i] = b0 + bl[i];
      len[i] = bg[i] < B ? p.lens[bg[i]] : 0;
#pragma unroll
      for (int g = 0; g < G; ++g) bh[g][i] = p.b_hh[(size_t)d * G * H + g * H + ug[i]];
    }
    float hst[4] = {0.f, 0.f, 0.f, 0.f}, cst[4] = {0.f, 0.f, 0.f, 0.f};
    const int t_first = d == 0 ? 0 : T - 1;
    const long long tdir = d == 0 ? 1 : -1;
    const float* gi_p[4];
    float* hid_p[4];
    float* sv_p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t row0 = (size_t)(bg[i] < B ? bg[i] : 0) * T + t_first;
      gi_p[i] = p.gi + (row0 * D + d) * (size_t)G * H + ug[i];
      hid_p[i] = p.hidden + row0 * (size_t)D * H + (size_t)d * H + ug[i];
      sv_p[i] = S > 0 ? p.saved + (row0 * D + d) * (size_t)S * H + ug[i] : nullptr;
    }
    const long long gi_step = tdir * (long long)D * G * H, hid_step = tdir * (long long)D * H,
                    sv_step = tdir * (long long)D * S * H;

    for (int step = 0; step < T; ++step) {
      const int tt = d == 0 ? step : T - 1 - step;
      const int cur = step & 1;
      float giv[G][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int g = 0; g < G; ++g) giv[g][i] = (bg[i] < B) ? gi_p[i][(size_t)g * H] : 0.f;
        gi_p[i] += gi_step;
      }
      float acc[G][4];
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
      if (step > 0)
        streamed_mma<G>(acc, p.xbuf + ((size_t)cur * D + d) * xdir, H, W_addr, (uint32_t)(kUH * pitch * 2), pitch, x_addr,
                        lane, warp);
      float o_h[4], o_sv[5][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = bg[i] < B;
        const bool active = valid && tt < len[i];
        float hnew, sv0 = 0.f, sv1 = 0.f, sv2 = 0.f, sv3 = 0.f, sv4 = 0.f;
        if (MODE == LR_RNN_GRU) {
          float ghn = acc[2 % G][i] + bh[2 % G][i];
          float rg = sigm(giv[0][i] + acc[0][i] + bh[0][i]);
          float z = sigm(giv[1 % G][i] + acc[1 % G][i] + bh[1 % G][i]);
          float n = tanh_(giv[2 % G][i] + rg * ghn);
          hnew = (1.f - z) * n + z * hst[i];
          sv0 = rg; sv1 = z; sv2 = n; sv3 = ghn;
        } else if (MODE == LR_RNN_LSTM) {
          float ig = sigm(giv[0][i] + acc[0][i] + bh[0][i]);
          float fg = sigm(giv[1 % G][i] + acc[1 % G][i] + bh[1 % G][i]);
          float gg = tanh_(giv[2 % G][i] + acc[2 % G][i] + bh[2 % G][i]);
          float og = sigm(giv[3 % G][i] + acc[3 % G][i] + bh[3 % G][i]);
          float cn = fg * cst[i] + ig * gg;
          hnew = og * tanh_(cn);
          if (active) cst[i] = cn;
          sv0 = ig; sv1 = fg; sv2 = gg; sv3 = og; sv4 = cn;
        } else {
          hnew = tanh_(giv[0][i] + acc[0][i] + bh[0][i]);
        }
        if (active) hst[i] = hnew;
        o_h[i] = active ? hnew : 0.f;
        o_sv[0][i] = active ? sv0 : 0.f; o_sv[1][i] = active ? sv1 : 0.f; o_sv[2][i] = active ? sv2 : 0.f;
        o_sv[3][i] = active ? sv3 : 0.f; o_sv[4][i] = active ? sv4 : 0.f;
        stage[bl[i] * kUH + (ug[i] - u_base)] = __float2bfloat16(hst[i]);
      }
      __syncthreads();
      if (step + 1 < T) {
        // this CTA's [kBS x 16] bf16 slice -> exchange buffer of the next step: 32 contiguous bytes per clip
        __nv_bfloat16* dst = p.xbuf + ((size_t)(cur ^ 1) * D + d) * xdir;
        for (int i = tid; i < kBS * 2; i += kThreads) {
          const int row = i >> 1, half = i & 1;
          *reinterpret_cast<uint4*>(dst + (size_t)row * H + u_base + half * 8) =
              *reinterpret_cast<const uint4*>(stage + row * kUH + half * 8);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (bg[i] < B) {
          *hid_p[i] = o_h[i];
          if (S > 0) {
            float* sv = sv_p[i];
            sv[0] = o_sv[0][i];
            sv[H] = o_sv[1][i];
            sv[2 * H] = o_sv[2][i];
            sv[3 * H] = o_sv[3][i];
            if (S > 4) sv[4 * H] = o_sv[4][i];
          }
        }
        hid_p[i] += hid_step;
        if (S > 0) sv_p[i] += sv_step;
      }
      ++arrivals;
      dir_barrier(ctr, arrivals * (unsigned int)NC);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (bg[i] < B) {
        p.h_n[((size_t)d * B + bg[i]) * H + ug[i]] = hst[i];
        if (MODE == LR_RNN_LSTM) p.c_n[((size_t)d * B + bg[i]) * H + ug[i]] = cst[i];
      }
  }
}

struct GBwd {
  const float* d_hidden; const float* d_h_n; const float* d_c_n; const float* saved; const float* hidden;
  const float* w_hh; const int32_t* lens;
  float* d_gi; float* d_gh; float* h_prev_all;
  __nv_bfloat16* xbuf;           // [2][D][kBS][G*H] exchange buffer (gate gradients)
  unsigned int* ctr;
  int B, T, H, D, NC;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) rnn_grid_bwd_kernel(GBwd p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, D = p.D, NC = p.NC;
  const int GH = G * H;
  const int pitch = GH + kPad;
  __nv_bfloat16* WT_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);                   // [16][pitch]
  __nv_bfloat16* x_s = WT_s + (size_t)kUH * pitch;                                    // [2][kBS][kKC + kPad]
  __nv_bfloat16* stage = x_s + (size_t)2 * kBS * (kKC + kPad);                        // [kBS][G*16]
  const int d = blockIdx.x / NC, r = blockIdx.x - d * NC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u_base = r * kUH;

  // W_hh^T slice: WT_s[ul][row] = W_hh[d][row][u_base+ul]
  const float* Wd = p.w_hh + (size_t)d * GH * H;
  for (int i = tid; i < GH * kUH; i += kThreads) {
    int row = i / kUH, ul = i - row * kUH;
    WT_s[(size_t)ul * pitch + row] = __float2bfloat16(Wd[(size_t)row * H + u_base + ul]);
  }
  __syncthreads();

  const uint32_t WT_addr = lr_smem_u32(WT_s), x_addr = lr_smem_u32(x_s);
  unsigned int* ctr = p.ctr + d;
  unsigned int arrivals = 0;
  const size_t xdir = (size_t)kBS * GH;

  for (int b0 = 0; b0 < B; b0 += kBS) {
    int ug[4], bl[4], bg[4], len[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ug[i] = u_base + (lane >> 2) + 8 * (i >> 1);
      bl[i] = warp * 8 + 2 * (lane & 3) + (i & 1);
      bg[i] = b0 + bl[i];
      len[i] = bg[i] < B ? p.lens[bg[i]] : 0;
    }
    float dh_dir[4], dc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t sidx = ((size_t)d * B + (bg[i] < B ? bg[i] : 0)) * H + ug[i];
      dh_dir[i] = (p.d_h_n && bg[i] < B) ? p.d_h_n[sidx] : 0.f;
      dc[i] = (MODE == LR_RNN_LSTM && p.d_c_n && bg[i] < B) ? p.d_c_n[sidx] : 0.f;
    }
    const int t_first = d == 0 ? T - 1 : 0;                 // reverse of the forward order
    const long long tdir = d == 0 ? -1 : 1;
    const float* dh_p[4];
    const float* sv_p[4];
    const float* hid_p[4];
    float* dgi_p[4];
    float* dgh_p[4];
    float* hp_p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t row0 = (size_t)(bg[i] < B ? bg[i] : 0) * T + t_first;
      dh_p[i] = p.d_hidden + row0 * (size_t)D * H + (size_t)d * H + ug[i];
      hid_p[i] = p.hidden + row0 * (size_t)D * H + (size_t)d * H + ug[i];
      sv_p[i] = S > 0 ? p.saved + (row0 * D + d) * (size_t)S * H + ug[i] : nullptr;
      dgi_p[i] = p.d_gi + (row0 * D + d) * (size_t)GH + ug[i];
      dgh_p[i] = p.d_gh + (row0 * D + d) * (size_t)GH + ug[i];
      hp_p[i] = p.h_prev_all + (row0 * D + d) * (size_t)H + ug[i];
    }
    const long long dh_step = tdir * (long long)D * H, sv_step = tdir * (long long)D * S * H,
                    dg_step = tdir * (long long)D * GH;
    const long long prev_hid = (d == 0 ? -1 : 1) * (long long)D * H, prev_sv = (d == 0 ? -1 : 1) * (long long)D * S * H;

    for (int step = 0; step < T; ++step) {
      const int tt = d == 0 ? T - 1 - step : step;
      const int cur = step & 1;
      float dout[4], svv[5][4], hprev[4], cprev[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = bg[i] < B;
        dout[i] = valid ? *dh_p[i] : 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) svv[k][i] = (k < S && valid) ? sv_p[i][(size_t)k * H] : 0.f;
        const int tt_in = d == 0 ? tt - 1 : tt + 1;
        const bool has_prev = valid && ((d == 0) ? (tt_in >= 0) : (tt_in < len[i]));
        hprev[i] = has_prev ? hid_p[i][prev_hid] : 0.f;
        cprev[i] = (MODE == LR_RNN_LSTM && has_prev) ? sv_p[i][prev_sv + (long long)4 * H] : 0.f;
      }
      float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      if (step > 0)
        streamed_mma<1>(acc, p.xbuf + ((size_t)cur * D + d) * xdir, GH, WT_addr, 0u, pitch, x_addr, lane, warp);
      float o_gi[G][4], o_gh[G][4], o_hp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = bg[i] < B;
        const bool active = valid && tt < len[i];
        const float carry = dh_dir[i] + acc[0][i];
        float dgi[G], dgh[G];
#pragma unroll
        for (int g = 0; g < G; ++g) dgi[g] = dgh[g] = 0.f;
        float hp_out = 0.f;
        if (!active) {
          dh_dir[i] = carry;
        } else {
          const float dh = dout[i] + carry;
          hp_out = hprev[i];
          if (MODE == LR_RNN_GRU) {
            const float rg = svv[0][i], z = svv[1][i], n = svv[2][i], ghn = svv[3][i];
            const float dn_pre = dh * (1.f - z) * (1.f - n * n);
            const float dr_pre = dn_pre * ghn * rg * (1.f - rg);
            const float dz_pre = dh * (hprev[i] - n) * z * (1.f - z);
            dgi[0] = dr_pre; dgh[0] = dr_pre;
            dgi[1 % G] = dz_pre; dgh[1 % G] = dz_pre;
            dgi[2 % G] = dn_pre; dgh[2 % G] = dn_pre * rg;
            dh_dir[i] = dh * z;
          } else if (MODE == LR_RNN_LSTM) {
            const float ig = svv[0][i], fg = svv[1][i], gg = svv[2][i], og = svv[3][i], cn = svv[4][i];
            const float tc = tanh_(cn);
            const float dcc = dc[i] + dh * og * (1.f - tc * tc);
            dgi[0] = dgh[0] = dcc * gg * ig * (1.f - ig);
            dgi[1 % G] = dgh[1 % G] = dcc * cprev[i] * fg * (1.f - fg);
            dgi[2 % G] = dgh[2 % G] = dcc * ig * (1.f - gg * gg);
            dgi[3 % G] = dgh[3 % G] = dh * tc * og * (1.f - og);
            dc[i] = dcc * fg;
            dh_dir[i] = 0.f;
          } else {
            const float h = *hid_p[i];
            dgi[0] = dgh[0] = dh * (1.f - h * h);
            dh_dir[i] = 0.f;
          }
        }
        o_hp[i] = hp_out;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          o_gi[g][i] = dgi[g];
          o_gh[g][i] = dgh[g];
          stage[bl[i] * (G * kUH) + g * kUH + (ug[i] - u_base)] = __float2bfloat16(dgh[g]);
        }
      }
      __syncthreads();
      if (step + 1 < T) {
        // [kBS x G x 16] bf16 -> exchange buffer row [clip][g*H + u_base ..+16): 32 contiguous bytes per (clip, gate)
        __nv_bfloat16* dst = p.xbuf + ((size_t)(cur ^ 1) * D + d) * xdir;
        for (int i = tid; i < kBS * G * 2; i += kThreads) {
          const int row = i / (G * 2), rem = i - row * (G * 2);
          const int g = rem >> 1, half = rem & 1;
          *reinterpret_cast<uint4*>(dst + (size_t)row * GH + (size_t)g * H + u_base + half * 8) =
              *reinterpret_cast<const uint4*>(stage + row * (G * kUH) + g * kUH + half * 8);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (bg[i] < B) {
#pragma unroll
          for (int g = 0; g < G; ++g) { dgi_p[i][(size_t)g * H] = o_gi[g][i]; dgh_p[i][(size_t)g * H] = o_gh[g][i]; }
          *hp_p[i] = o_hp[i];
        }
        dh_p[i] += dh_step; hid_p[i] += dh_step; hp_p[i] += dh_step;
        if (S > 0) sv_p[i] += sv_step;
        dgi_p[i] += dg_step; dgh_p[i] += dg_step;
      }
      ++arrivals;
      dir_barrier(ctr, arrivals * (unsigned int)NC);
    }
  }
}

int gates_of(int mode) { return mode == LR_RNN_GRU ? 3 : (mode == LR_RNN_LSTM ? 4 : 1); }
size_t fwd_smem(int G, int H) {
  return ((size_t)G * kUH * (H + kPad) + (size_t)2 * kBS * (kKC + kPad) + (size_t)kBS * kUH) * 2;
}
size_t bwd_smem(int G, int H) {
  return ((size_t)kUH * (G * H + kPad) + (size_t)2 * kBS * (kKC + kPad) + (size_t)kBS * G * kUH) * 2;
}
size_t xbuf_bytes(int G, int H, int D) { return (size_t)2 * D * kBS * G * H * 2; }     // sized for the backward operand

template <typename K, typename P>
int launch_coop(K kernel, P params, int grid, size_t smem, cudaStream_t st) {
  LR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[1] = {&params};
  LR_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(grid), dim3(kThreads), args, smem, st));
  lr_count_launch();
  return LR_OK;
}

}  // namespace

// 1 when the grid-persistent kernels can run this shape on this device: H a multiple of 256 (the chunk width),
// D * H/16 CTAs co-resident, slices fit shared memory
extern "C" int lr_rnn_grid_supported(int mode, int H, int D) {
  if (mode < 0 || mode > 2 || H <= 0 || H % kKC != 0 || (D != 1 && D != 2)) return 0;
  const int G = gates_of(mode);
  if (D * (H / kUH) > kNumSMs) return 0;
  return fwd_smem(G, H) <= 220 * 1024 && bwd_smem(G, H) <= 220 * 1024;
}
extern "C" size_t lr_rnn_grid_workspace(int mode, int H, int D) {
  if (!lr_rnn_grid_supported(mode, H, D)) return 0;
  return xbuf_bytes(gates_of(mode), H, D) + 256;
}

extern "C" int lr_rnn_grid_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh, const int32_t* lens,
                               int B, int T, int H, int D, float* hidden, float* h_n, float* c_n, float* saved,
                               void* workspace, size_t ws_bytes, void* stream) {
  LR_CHECK_ARG(gi && w_hh && b_hh && lens && hidden && h_n && workspace, "lr_rnn_grid_fwd: null pointer");
  LR_CHECK_ARG(lr_rnn_grid_supported(mode, H, D), "lr_rnn_grid_fwd: unsupported mode/hidden size (%d,%d,%d)", mode, H, D);
  LR_CHECK_ARG(B > 0 && T > 0, "lr_rnn_grid_fwd: bad shape");
  LR_CHECK_ARG(mode != LR_RNN_LSTM || c_n, "lr_rnn_grid_fwd: LSTM needs c_n");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_grid_fwd: `saved` required");
  if (ws_bytes < lr_rnn_grid_workspace(mode, H, D)) { lr_set_error("lr_rnn_grid_fwd: workspace too small"); return LR_EWORKSPACE; }
  cudaStream_t st = lr_stream(stream);
  GFwd p;
  p.gi = gi; p.w_hh = w_hh; p.b_hh = b_hh; p.lens = lens; p.hidden = hidden; p.saved = saved; p.h_n = h_n; p.c_n = c_n;
  p.ctr = reinterpret_cast<unsigned int*>(workspace);
  p.xbuf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  p.B = B; p.T = T; p.H = H; p.D = D; p.NC = H / kUH;
  LR_CHECK_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  const int grid = D * p.NC;
  const size_t smem = fwd_smem(gates_of(mode), H);
  if (mode == LR_RNN_GRU) return launch_coop(rnn_grid_fwd_kernel<LR_RNN_GRU>, p, grid, smem, st);
  if (mode == LR_RNN_LSTM) return launch_coop(rnn_grid_fwd_kernel<LR_RNN_LSTM>, p, grid, smem, st);
  return launch_coop(rnn_grid_fwd_kernel<LR_RNN_TANH>, p, grid, smem, st);
}

extern "C" int lr_rnn_grid_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
                               const float* saved, const float* hidden, const float* w_hh, const int32_t* lens, int B,
                               int T, int H, int D, float* d_gi, float* d_gh, float* h_prev_all, void* workspace,
                               size_t ws_bytes, void* stream) {
  LR_CHECK_ARG(d_hidden && hidden && w_hh && lens && d_gi && d_gh && h_prev_all && workspace, "lr_rnn_grid_bwd: null pointer");
  LR_CHECK_ARG(lr_rnn_grid_supported(mode, H, D), "lr_rnn_grid_bwd: unsupported mode/hidden size (%d,%d,%d)", mode, H, D);
  LR_CHECK_ARG(B > 0 && T > 0, "lr_rnn_grid_bwd: bad shape");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_grid_bwd: `saved` required");
  if (ws_bytes < lr_rnn_grid_workspace(mode, H, D)) { lr_set_error("lr_rnn_grid_bwd: workspace too small"); return LR_EWORKSPACE; }
  cudaStream_t st = lr_stream(stream);
  GBwd p;
  p.d_hidden = d_hidden; p.d_h_n = d_h_n; p.d_c_n = d_c_n; p.saved = saved; p.hidden = hidden; p.w_hh = w_hh;
  p.lens = lens; p.d_gi = d_gi; p.d_gh = d_gh; p.h_prev_all = h_prev_all;
  p.ctr = reinterpret_cast<unsigned int*>(workspace);
  p.xbuf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  p.B = B; p.T = T; p.H = H; p.D = D; p.NC = H / kUH;
  LR_CHECK_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  const int grid = D * p.NC;
  const size_t smem = bwd_smem(gates_of(mode), H);
  if (mode == LR_RNN_GRU) return launch_coop(rnn_grid_bwd_kernel<LR_RNN_GRU>, p, grid, smem, st);
  if (mode == LR_RNN_LSTM) return launch_coop(rnn_grid_bwd_kernel<LR_RNN_LSTM>, p, grid, smem, st);
  return launch_coop(rnn_grid_bwd_kernel<LR_RNN_TANH>, p, grid, smem, st);
}
