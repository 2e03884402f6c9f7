// ctc.cu — CTC alpha/beta dynamic programme with fused gradient.
//
// Replaces the reference's `F.ctc_loss(log_probs.transpose(0,1).cpu(), ...)` + autograd backward
// (src/train/ctc_loss.py:85,100): no transpose, no device->host copy, one kernel for loss+gradient.
//
// One CTA per clip.  The clip's (T,C) log-prob block is contiguous in the batch-first layout the
// encoder emits, so it is pulled into shared memory with ONE bulk-async copy (TMA, UBLKCP) while
// the other threads expand the label into the 2L+1 lattice classes.  Warp 0 runs the alpha
// recursion and warp 1 the beta recursion concurrently out of shared memory (the (T,S) lattices
// never touch HBM when they fit; otherwise they spill to the caller's workspace through the same
// generic pointers).  All warps then build the gradient rows
//     g[t,c] = exp(lp[t,c]) - exp(logsum_{s: l'_s = c}(alpha_t(s)+beta_t(s)) + nll - lp[t,c])
// (torch's native-CTC convention, aten/src/ATen/native/LossCTC.cpp) deterministically: repeated
// label classes are chained through a per-clip "next occurrence" list, no atomics.
//
// HBM traffic is exactly the algorithmic 2*T*C*4 bytes per clip (39 000 B at T=75,C=65).
#include "common.cuh"
#include <stdlib.h>

// kernel choice is a per-call argument (`kernel`): 0 auto (by batch size), 1 always CTA-per-clip, 2 always log-space
// warp-per-clip, 3 always warp-per-clip with the linear-space kernel first (labels <= 31 symbols)

namespace {

constexpr int kCtcThreads = 128;
constexpr int kCtcWarps = kCtcThreads / 32;
constexpr size_t kCtcSmemCap = 200 * 1024;

struct CtcPlan {
  int Smax;
  int Cpad;
  size_t off_cls, off_skip, off_nxt, off_head, off_rowbuf, off_misc, off_lp, off_alpha, off_beta;
  size_t lp_tile_bytes;
  size_t smem_bytes;
  int lp_in_smem;
  int lat_in_smem;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

CtcPlan make_plan(int T, int C, int Lmax) {
  CtcPlan p;
  p.Smax = 2 * Lmax + 1;
  p.Cpad = (C + 31) / 32 * 32;
  size_t off = 16;  // mbarrier
  p.off_cls = off;   off += (size_t)p.Smax * 4;
  p.off_skip = off;  off += (size_t)p.Smax * 4;
  p.off_nxt = off;   off += (size_t)(Lmax + 1) * 4;
  p.off_head = off;  off += (size_t)(Lmax + 1) * 4;
  p.off_rowbuf = off; off += (size_t)kCtcWarps * p.Cpad * 4;
  p.off_misc = off;  off += 16;
  off = align_up(off, 16);
  size_t fixed = off;
  p.lp_tile_bytes = align_up((size_t)T * C * 4, 16) + 32;
  size_t lat = (size_t)T * p.Smax * 4;
  p.lp_in_smem = (fixed + p.lp_tile_bytes) <= kCtcSmemCap;
  p.lat_in_smem = p.lp_in_smem && (fixed + p.lp_tile_bytes + 2 * lat) <= kCtcSmemCap;
  p.off_lp = fixed;
  size_t end = fixed + (p.lp_in_smem ? p.lp_tile_bytes : 0);
  p.off_alpha = end;
  p.off_beta = end + lat;
  if (p.lat_in_smem) end += 2 * lat;
  p.smem_bytes = end;
  return p;
}

__global__ void __launch_bounds__(kCtcThreads)
ctc_alpha_beta_grad_kernel(const float* __restrict__ lp_all, const int32_t* __restrict__ targets,
                           const int32_t* __restrict__ in_lens,
                           const int32_t* __restrict__ tgt_lens, int B, int T, int C, int Lmax,
                           float* __restrict__ nll_out, float* __restrict__ grad,
                           float* __restrict__ ws, CtcPlan plan) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int Tb = in_lens[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int L = tgt_lens[b];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const int S = 2 * L + 1;
  const int Smax = plan.Smax;

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int* cls = reinterpret_cast<int*>(smem_raw + plan.off_cls);
  int* skip = reinterpret_cast<int*>(smem_raw + plan.off_skip);
  int* nxt = reinterpret_cast<int*>(smem_raw + plan.off_nxt);
  int* head = reinterpret_cast<int*>(smem_raw + plan.off_head);
  float* rowbuf = reinterpret_cast<float*>(smem_raw + plan.off_rowbuf) + warp * plan.Cpad;
  float* misc = reinterpret_cast<float*>(smem_raw + plan.off_misc);

  const float* lp_g = lp_all + (size_t)b * T * C;
  const int32_t* tgt = targets + (size_t)b * Lmax;

  // ---- phase 0: bulk-load the clip's log-probs, expand labels --------------------------------
  const float* lp = lp_g;
  uint32_t tail_begin = 0, n_valid = (uint32_t)Tb * C;
  float* lp_s = nullptr;
  if (plan.lp_in_smem && Tb > 0) {
    unsigned char* tile = smem_raw + plan.off_lp;
    uintptr_t g0 = reinterpret_cast<uintptr_t>(lp_g);
    uintptr_t a0 = g0 & ~(uintptr_t)15;
    uint32_t shift = (uint32_t)(g0 - a0);
    // never read past the end of the (B,T,C) tensor: bulk part stops at the last 16-byte boundary
    uintptr_t tensor_end = reinterpret_cast<uintptr_t>(lp_all + (size_t)B * T * C);
    uintptr_t want_end = g0 + (uintptr_t)n_valid * 4;
    uintptr_t a1 = (want_end + 15) & ~(uintptr_t)15;
    if (a1 > tensor_end) a1 = tensor_end & ~(uintptr_t)15;
    uint32_t bulk_bytes = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
    lp_s = reinterpret_cast<float*>(tile + shift);
    lp = lp_s;
    // floats covered by the bulk copy, counted from lp_g
    tail_begin = bulk_bytes > shift ? (bulk_bytes - shift) / 4 : 0;
    if (tail_begin > n_valid) tail_begin = n_valid;
    if (tid == 0) {
      lr_mbar_init(bar, 1);
      lr_fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      if (bulk_bytes > 0) {
        lr_mbar_expect_tx(bar, bulk_bytes);
        lr_bulk_g2s(tile, reinterpret_cast<const void*>(a0), bulk_bytes, bar);
      } else {
        lr_mbar_arrive(bar);
      }
    }
  }

  for (int s = tid; s < S; s += kCtcThreads) {
    int c = (s & 1) ? tgt[s >> 1] : 0;
    cls[s] = c;
  }
  // next-occurrence chains over label positions (O(L^2) but L <= 256 and off the critical path)
  for (int j = tid; j < L; j += kCtcThreads) {
    int cj = tgt[j];
    int h = 1, n = -1;
    for (int k = 0; k < j; ++k)
      if (tgt[k] == cj) { h = 0; break; }
    for (int k = j + 1; k < L; ++k)
      if (tgt[k] == cj) { n = k; break; }
    head[j] = h;
    nxt[j] = n;
  }
  __syncthreads();
  for (int s = tid; s < S; s += kCtcThreads) {
    int f = 0;
    if ((s & 1) && s >= 3 && cls[s] != cls[s - 2]) f |= 1;      // alpha may arrive from s-2
    if ((s & 1) && s + 2 < S && cls[s + 2] != cls[s]) f |= 2;   // beta may arrive from s+2
    skip[s] = f;
  }
  if (lp_s != nullptr) {
    lr_mbar_wait(bar, 0);
    // the <16-byte tail (and anything the clamp cut off) comes in with plain loads
    for (uint32_t i = tail_begin + tid; i < n_valid; i += kCtcThreads) lp_s[i] = lp_g[i];
  }
  __syncthreads();

  float* alpha;
  float* beta;
  if (plan.lat_in_smem) {
    alpha = reinterpret_cast<float*>(smem_raw + plan.off_alpha);
    beta = reinterpret_cast<float*>(smem_raw + plan.off_beta);
  } else {
    alpha = ws + (size_t)b * 2 * T * Smax;
    beta = alpha + (size_t)T * Smax;
  }

  // ---- phase 1: alpha on warp 0, beta on warp 1, concurrently ---------------------------------
  if (Tb > 0) {
    if (warp == 0) {
      for (int s = lane; s < S; s += 32)
        alpha[s] = (s == 0) ? lp[0] : (s == 1 ? lp[cls[1]] : LR_NEG_INF);
      for (int t = 1; t < Tb; ++t) {
        __syncwarp();
        const float* prev = alpha + (size_t)(t - 1) * Smax;
        float* cur = alpha + (size_t)t * Smax;
        const float* row = lp + (size_t)t * C;
        for (int s = lane; s < S; s += 32) {
          float a0 = prev[s];
          float a1 = s >= 1 ? prev[s - 1] : LR_NEG_INF;
          float a2 = (skip[s] & 1) ? prev[s - 2] : LR_NEG_INF;
          cur[s] = lr_lse3(a0, a1, a2) + row[cls[s]];
        }
      }
    } else if (warp == 1) {
      float* last = beta + (size_t)(Tb - 1) * Smax;
      const float* lrow = lp + (size_t)(Tb - 1) * C;
      for (int s = lane; s < S; s += 32)
        last[s] = (s == S - 1) ? lrow[0] : (s == S - 2 ? lrow[cls[s]] : LR_NEG_INF);
      for (int t = Tb - 2; t >= 0; --t) {
        __syncwarp();
        const float* nx = beta + (size_t)(t + 1) * Smax;
        float* cur = beta + (size_t)t * Smax;
        const float* row = lp + (size_t)t * C;
        for (int s = lane; s < S; s += 32) {
          float b0 = nx[s];
          float b1 = s + 1 < S ? nx[s + 1] : LR_NEG_INF;
          float b2 = (skip[s] & 2) ? nx[s + 2] : LR_NEG_INF;
          cur[s] = lr_lse3(b0, b1, b2) + row[cls[s]];
        }
      }
    }
  }
  __syncthreads();

  if (tid == 0) {
    float nll;
    if (Tb == 0) {
      nll = (L == 0) ? 0.f : INFINITY;
    } else {
      const float* last = alpha + (size_t)(Tb - 1) * Smax;
      float l1 = last[S - 1];
      float l2 = S > 1 ? last[S - 2] : LR_NEG_INF;
      nll = -lr_lse2(l1, l2);
    }
    misc[0] = nll;
    nll_out[b] = nll;
  }
  __syncthreads();
  if (grad == nullptr) return;
  const float nll = misc[0];
  const bool dead = !(nll < INFINITY);  // infeasible alignment (or NaN): gradient defined as 0

  // ---- phase 2: gradient rows, one warp per time step ------------------------------------------
  float* g_b = grad + (size_t)b * T * C;
  for (int t = warp; t < T; t += kCtcWarps) {
    float* g = g_b + (size_t)t * C;
    if (t >= Tb || dead) {
      for (int c = lane; c < C; c += 32) g[c] = 0.f;
      continue;
    }
    const float* a = alpha + (size_t)t * Smax;
    const float* be = beta + (size_t)t * Smax;
    const float* row = lp + (size_t)t * C;
    // blank: all even lattice states
    float m = LR_NEG_INF;
    for (int s = 2 * lane; s < S; s += 64) m = fmaxf(m, a[s] + be[s]);
    m = lr_warp_max(m);
    float sum = 0.f;
    if (m != LR_NEG_INF)
      for (int s = 2 * lane; s < S; s += 64) sum += expf(a[s] + be[s] - m);
    sum = lr_warp_sum(sum);
    const float lcab0 = (m == LR_NEG_INF) ? LR_NEG_INF : m + logf(sum);

    for (int c = lane; c < C; c += 32) rowbuf[c] = expf(row[c]);
    __syncwarp();
    if (lane == 0) {
      float lp0 = row[0];
      float occ = (lcab0 == LR_NEG_INF) ? 0.f : expf(lcab0 + nll - lp0);
      rowbuf[0] = expf(lp0) - occ;
    }
    for (int j = lane; j < L; j += 32) {
      if (!head[j]) continue;
      float acc = a[2 * j + 1] + be[2 * j + 1];
      for (int k = nxt[j]; k >= 0; k = nxt[k]) acc = lr_lse2(acc, a[2 * k + 1] + be[2 * k + 1]);
      int c = cls[2 * j + 1];
      if (c > 0 && c < C) {
        float lpc = row[c];
        float occ = (acc == LR_NEG_INF) ? 0.f : expf(acc + nll - lpc);
        rowbuf[c] = expf(lpc) - occ;
      }
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) g[c] = rowbuf[c];
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------------
// Throughput variant: ONE WARP PER CLIP, states in registers, neighbour exchange by warp shuffle.
//
// Lane l owns the K = 2*P consecutive lattice states [l*K, (l+1)*K): P (blank, label) pairs.  The alpha
// step needs exactly one shuffle (the previous lane's last label state), the beta step two (the next
// lane's first blank/label).  Only the alpha lattice is kept (shared memory, T x 32K floats); beta
// lives in registers and the gradient row of frame t is produced in the same backward sweep, so the
// clip's log-probs are read twice from L2/HBM and the gradient written once — nothing else moves.
// A CTA is a single warp: ~20 KB of shared memory per clip lets ~11 clips share an SM, which is what
// hides the 75-step dependent chain (the kernel is latency-bound per clip, throughput-bound per SM).
// GA = 1 keeps the alpha lattice in a global workspace instead (written coalesced as it is produced, read back a few
// rows ahead in the backward sweep; it is consumed ~100 us after it was written, i.e. out of L2): shared memory per
// clip drops from ~23 KB to ~4 KB, so the SM holds 32 clips instead of 9 — the kernel is latency-bound per clip, and
// more clips in flight is what raises its throughput.
template <int P, bool FAST, int GA>
__global__ void __launch_bounds__(32)
ctc_warp_kernel(const float* __restrict__ lp_all, const int32_t* __restrict__ targets,
                const int32_t* __restrict__ in_lens, const int32_t* __restrict__ tgt_lens, int B, int T, int C,
                int Lmax, float* __restrict__ nll_out, float* __restrict__ grad, float* __restrict__ alpha_ws,
                const int32_t* __restrict__ redo) {
  constexpr int K = 2 * P, SP = 32 * K;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, lane = threadIdx.x;
  if (redo && !redo[b]) return;          // second pass behind the linear-space kernel: only the clips it gave up on
  const int Cpad = (C + 31) / 32 * 32;
  constexpr int RING = 8;                                            // log-prob rows in flight (cp.async)
  float* alpha = GA ? alpha_ws + (size_t)b * T * SP : reinterpret_cast<float*>(smem_raw);   // [T][SP]
  float* rows = GA ? reinterpret_cast<float*>(smem_raw) : alpha + (size_t)T * SP;           // [RING][Cpad]
  float* occ = rows + RING * Cpad;                                   // [Cpad]
  float* erow = occ + Cpad;                                          // [32*P] label occupancies of a frame
  int* nxt = reinterpret_cast<int*>(erow + 32 * P);                  // [32*P]
  int* head = nxt + 32 * P;                                          // [32*P]

  int Tb = in_lens[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int L = tgt_lens[b];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const int S = 2 * L + 1;
  const float* lp_g = lp_all + (size_t)b * T * C;
  const int32_t* tgt = targets + (size_t)b * Lmax;
  float* g_b = grad ? grad + (size_t)b * T * C : nullptr;

  // labels of this lane's pairs: pair i <-> label position j = lane*P + i
  int lab[P];
  bool skipa[P], skipb[P], has_lab[P], has_blank[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int j = lane * P + i;
    has_lab[i] = j < L;
    has_blank[i] = j <= L;                                           // blank state 2j exists iff 2j < S
    lab[i] = has_lab[i] ? tgt[j] : 0;
    skipa[i] = has_lab[i] && j >= 1 && lab[i] != tgt[j - 1];          // alpha: s-2 -> s
    skipb[i] = has_lab[i] && j + 1 < L && tgt[j + 1] != lab[i];        // beta:  s+2 -> s
    int h = 1, n = -1;
    if (has_lab[i]) {
      for (int k = 0; k < j; ++k) if (tgt[k] == lab[i]) { h = 0; break; }
      for (int k = j + 1; k < L; ++k) if (tgt[k] == lab[i]) { n = k; break; }
    }
    head[j] = has_lab[i] ? h : 0;
    nxt[j] = n;
  }

  if (Tb == 0) {
    if (lane == 0) nll_out[b] = (L == 0) ? 0.f : INFINITY;
    if (g_b) for (int i = lane; i < T * C; i += 32) g_b[i] = 0.f;
    return;
  }
  // asynchronous row fetch (LDGSTS): row t lands in ring slot t % RING; one commit group per row
  const uint32_t rows_addr = lr_smem_u32(rows);
  auto issue_row = [&](int t) {
    if (t >= 0 && t < Tb) {
      const uint32_t dst = rows_addr + (uint32_t)(((t % RING) * Cpad) * 4);
      for (int c = lane; c < C; c += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + c * 4), "l"(lp_g + (size_t)t * C + c)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto wait_row = [&]() {      // all but the RING-2 most recent groups have landed
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");
    __syncwarp();
  };
  // ---- alpha sweep -------------------------------------------------------------------------------
  float ab[P], al[P];                                                // blank / label state values
  for (int t = 0; t < RING - 1; ++t) issue_row(t);                   // prologue: RING-1 rows in flight
  wait_row();
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int j = lane * P + i;
    ab[i] = (j == 0) ? rows[0] : LR_NEG_INF;
    al[i] = (j == 0 && has_lab[i]) ? rows[lab[i]] : LR_NEG_INF;
    *reinterpret_cast<float2*>(alpha + lane * K + 2 * i) = make_float2(ab[i], al[i]);
  }
  for (int t = 1; t < Tb; ++t) {
    __syncwarp();                                                    // everyone is done with slot (t-2)%RING
    issue_row(t + RING - 2);
    wait_row();
    float* row = rows + (t % RING) * Cpad;
    const float lpb = row[0];
    float prev_lab = __shfl_up_sync(0xffffffffu, al[P - 1], 1);      // label state just left of this lane
    if (lane == 0) prev_lab = LR_NEG_INF;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float nb = has_blank[i] ? (FAST ? lr_lse2_fast : lr_lse2)(ab[i], prev_lab) + lpb : LR_NEG_INF;
      const float nl = has_lab[i] ? (FAST ? lr_lse3_fast : lr_lse3)(al[i], ab[i], skipa[i] ? prev_lab : LR_NEG_INF) + row[lab[i]]
                                  : LR_NEG_INF;
      prev_lab = al[i];
      ab[i] = nb;
      al[i] = nl;
      *reinterpret_cast<float2*>(alpha + (size_t)t * SP + lane * K + 2 * i) = make_float2(nb, nl);
    }
  }
  __syncwarp();
  float nll;
  if (GA) {
    // states S-1 = blank of label position L, S-2 = label of position L-1: pick them out of the registers
    float l1 = LR_NEG_INF, l2 = LR_NEG_INF;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float vb = __shfl_sync(0xffffffffu, ab[i], L / P);
      const float vl = __shfl_sync(0xffffffffu, al[i], L > 0 ? (L - 1) / P : 0);
      if (i == L % P) l1 = vb;
      if (L > 0 && i == (L - 1) % P) l2 = vl;
    }
    nll = -(FAST ? lr_lse2_fast : lr_lse2)(l1, l2);
  } else {
    const float* last = alpha + (size_t)(Tb - 1) * SP;
    const float l1 = last[S - 1];
    const float l2 = S > 1 ? last[S - 2] : LR_NEG_INF;
    nll = -(FAST ? lr_lse2_fast : lr_lse2)(l1, l2);
  }
  if (lane == 0) nll_out[b] = nll;
  if (g_b == nullptr) return;
  const bool dead = !(nll < INFINITY);
  for (int t = T - 1; t >= Tb; --t)
    for (int c = lane; c < C; c += 32) g_b[(size_t)t * C + c] = 0.f;
  if (dead) {
    for (int i = lane; i < Tb * C; i += 32) g_b[i] = 0.f;
    return;
  }
  // ---- beta sweep fused with the gradient rows ------------------------------------------------------
  float bb[P], bl[P];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  for (int k = 0; k < RING - 1; ++k) issue_row(Tb - 1 - k);          // prologue of the backward sweep
  // global-alpha variant: this lane's alpha values of rows t, t-1, t-2 travel in registers (loads issued three
  // iterations before their use cover the L2/HBM latency)
  constexpr int AD = 3;
  float2 apre[AD][P];
  if (GA) {
    __threadfence_block();
#pragma unroll
    for (int d = 0; d < AD; ++d)
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const int tt = Tb - 1 - d;
        apre[d][i] = tt >= 0 ? *reinterpret_cast<const float2*>(alpha + (size_t)tt * SP + lane * K + 2 * i)
                             : make_float2(0.f, 0.f);
      }
  }
  for (int t = Tb - 1; t >= 0; --t) {
    if (t < Tb - 1) { __syncwarp(); issue_row(t - (RING - 2)); }
    wait_row();
    float* row = rows + (t % RING) * Cpad;
    for (int c = lane; c < Cpad; c += 32) occ[c] = 0.f;
    __syncwarp();
    const float lpb = row[0];
    if (t == Tb - 1) {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const int j = lane * P + i;
        bb[i] = (j == L) ? lpb : LR_NEG_INF;                         // state S-1 (final blank)
        bl[i] = (has_lab[i] && j == L - 1) ? row[lab[i]] : LR_NEG_INF;   // state S-2
      }
    } else {
      float nb_blank = __shfl_down_sync(0xffffffffu, bb[0], 1);      // next lane's first blank / label
      float nb_lab = __shfl_down_sync(0xffffffffu, bl[0], 1);
      if (lane == 31) { nb_blank = LR_NEG_INF; nb_lab = LR_NEG_INF; }
      float nbv[P], nlv[P];
#pragma unroll
      for (int i = P - 1; i >= 0; --i) {
        const float right_blank = (i + 1 < P) ? bb[i + 1] : nb_blank;
        const float right_lab = (i + 1 < P) ? bl[i + 1] : nb_lab;
        nlv[i] = has_lab[i] ? (FAST ? lr_lse3_fast : lr_lse3)(bl[i], right_blank, skipb[i] ? right_lab : LR_NEG_INF) + row[lab[i]]
                            : LR_NEG_INF;
        nbv[i] = has_blank[i] ? (FAST ? lr_lse2_fast : lr_lse2)(bb[i], bl[i]) + lpb : LR_NEG_INF;
      }
#pragma unroll
      for (int i = 0; i < P; ++i) { bb[i] = nbv[i]; bl[i] = nlv[i]; }
    }
    // occupancies: exp(alpha + beta + nll - lp) per state (each in [0,1])
    float eb = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float2 a;
      if (GA) {
        a = apre[0][i];
#pragma unroll
        for (int d = 0; d + 1 < AD; ++d) apre[d][i] = apre[d + 1][i];
        const int tt = t - AD;
        if (tt >= 0) apre[AD - 1][i] = *reinterpret_cast<const float2*>(alpha + (size_t)tt * SP + lane * K + 2 * i);
      } else {
        a = *reinterpret_cast<const float2*>(alpha + (size_t)t * SP + lane * K + 2 * i);
      }
      const float sb = a.x + bb[i], sl = a.y + bl[i];
      if (has_blank[i] && sb != LR_NEG_INF) eb += __expf(sb + nll - lpb);
      erow[lane * P + i] = (has_lab[i] && sl != LR_NEG_INF) ? __expf(sl + nll - row[lab[i]]) : 0.f;
    }
    eb = lr_warp_sum(eb);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const int j = lane * P + i;
      if (has_lab[i] && head[j]) {
        float acc = erow[j];
        for (int k = nxt[j]; k >= 0; k = nxt[k]) acc += erow[k];
        if (lab[i] > 0 && lab[i] < C) occ[lab[i]] = acc;
      }
    }
    if (lane == 0) occ[0] = eb;
    __syncwarp();
    for (int c = lane; c < C; c += 32) g_b[(size_t)t * C + c] = __expf(row[c]) - occ[c];
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------
// Linear-space CTC for labels of up to 31 symbols and up to 96 classes (one warp per clip, one (blank, label) state
// pair per lane).  ncu on the log-space warp kernel (profiles/r1_ncu_ctc_*.csv): 34 k warp instructions per clip at
// 75 % issue utilisation, half of them address arithmetic, predicates and branches around the shared-memory row ring —
// the kernel is instruction-bound, not latency- or HBM-bound.  This kernel is written for instruction count:
//   * probabilities instead of log-probabilities, with Rabiner-style rescaling (Graves' original CTC formulation):
//       alpha^_t = (M alpha^_{t-1}) . p_t . k_t        -log p(l|x) = -(log rho - sum_t log k_t)
//       beta~_t  =  M' beta^_{t+1} ,  beta^_t = beta~_t . p_t . k_t   (the SAME k_t: then
//       sum_s alpha^_t(s) beta~_t(s) = rho = alpha^_T(S-1) + alpha^_T(S-2) for every t)
//       grad[t,c] = p_t(c) - sum_{s: l'_s = c} alpha^_t(s) beta~_t(s) / rho
//     (torch's native-CTC gradient exp(lp) - exp(logsum(alpha+beta) + nll - lp): the division by p_t(c) cancels
//     against the emission factor of beta, so no exp/log touches a lattice cell);
//   * k_t = 1 on odd frames and 1 / (lattice mass two frames earlier) on even ones: any positive factor keeps the
//     recursion exact as long as it is recorded, and the warp reduction behind it has two frames of slack instead
//     of sitting on the frame-to-frame dependency chain; sum_t log k_t is accumulated exactly (exponent + mantissa);
//   * a frame's C log-probs live in registers of the lanes that load them (class c in lane c % 32): one coalesced
//     128-byte load per 32 classes, prefetched four frames ahead, exponentiated once; the blank / label emissions
//     reach their states by shuffle, the gradient row is formed where the probabilities already are — no
//     shared-memory row ring, no cp.async bookkeeping;
//   * frames are walked in unrolled groups of four so all of that indexing is static.
// CPU restatement of exactly this recursion: oracle/sequence.py::ctc_linear_rescaled (held to the log-space recursion and
// to torch's native CTC in tests/test_oracle_golden.py, in float64 and in float32 arithmetic).
// A clip whose scale factors leave the fp32 range (infeasible labels, emission probabilities below e^-80, ...) is
// flagged in `redo` and recomputed by the log-space kernel in a second launch.
constexpr int kLinWarps = 2;      // clips per block: 64 resident warps per SM (a one-warp block caps at 32)

template <int CI, int MB>         // 32-class slabs per frame: C <= 32 * CI; MB = resident blocks per SM the registers allow
__global__ void __launch_bounds__(32 * kLinWarps, MB)
ctc_linear_warp_kernel(const float* __restrict__ lp_all, const int32_t* __restrict__ targets,
                       const int32_t* __restrict__ in_lens, const int32_t* __restrict__ tgt_lens, int B, int T, int C,
                       int Lmax, float* __restrict__ nll_out, float* __restrict__ grad,
                       float2* __restrict__ alpha_ws, int32_t* __restrict__ redo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kLinWarps + wib;
  if (b >= B) return;                                                // (no block-wide barriers below)
  const int n_kf = (T + 1) / 2 + 1;
  const size_t per_clip = ((size_t)(CI * 32 + 1 + 33 + n_kf) * sizeof(float) + 15) / 16 * 16;
  float* occ = reinterpret_cast<float*>(smem_raw + (size_t)wib * per_clip);    // [CI*32] class occupancies + 1 dump slot
  float* erow = occ + CI * 32 + 1;                                   // [32] label occupancies of a frame + 1 zero slot
  float* kfac = erow + 33;                                           // [n_kf] scale factor of even frame 2i
  float2* alpha = alpha_ws + (size_t)b * T * 32;

  int Tb = in_lens[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int L = tgt_lens[b];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const float* lp_g = lp_all + (size_t)b * T * C;
  const int32_t* tgt = targets + (size_t)b * Lmax;
  float* g_b = grad + (size_t)b * T * C;
  if (Tb == 0 || L > 31) {                                           // (L > 31 cannot happen: the host checks Lmax)
    if (lane == 0) redo[b] = 1;
    return;
  }
  const int j = lane;
  const bool has_lab = j < L, has_blank = j <= L;
  const int my = has_lab ? tgt[j] : 0;
  const int up = __shfl_up_sync(0xffffffffu, my, 1), dn = __shfl_down_sync(0xffffffffu, my, 1);
  const bool lab_ok = has_lab && my > 0 && my < C;                   // (an out-of-range class can never be emitted)
  const int lab = lab_ok ? my : 0;
  const float hl = lab_ok ? 1.f : 0.f, hb = has_blank ? 1.f : 0.f;
  const float skipa = (has_lab && j >= 1 && my != up) ? 1.f : 0.f;   // alpha: s-2 -> s
  const float skipb = (has_lab && j + 1 < L && dn != my) ? 1.f : 0.f;   // beta:  s+2 -> s
  const int src_lane = lab & 31, src_slab = lab >> 5;
  // occurrences of this lane's class further down the label: first two links in registers (slot 32 of erow is a
  // constant zero), longer chains (a class four or more times in one label) walk the list
  const unsigned same = __match_any_sync(0xffffffffu, has_lab ? my : -1 - lane);
  const unsigned later = same & ~((2u << lane) - 1u);               // lanes > this one with the same class
  const bool is_head = lab_ok && (same & ((1u << lane) - 1u)) == 0;
  const int n1 = later ? __ffs(later) - 1 : 32;
  const unsigned later2 = later & (later - 1);
  const int n2 = later2 ? __ffs(later2) - 1 : 32;
  const unsigned later3 = later2 & (later2 - 1);
  const bool deep = __any_sync(0xffffffffu, is_head && later3 != 0);
  const int occ_dst = is_head ? lab : CI * 32;                       // non-heads write the dump slot
  for (int c = lane; c < CI * 32 + 1; c += 32) occ[c] = 0.f;
  erow[lane] = 0.f;
  if (lane == 0) erow[32] = 0.f;
  __syncwarp();

  // this lane's classes: lane + 32k; beyond C they read as log 0
  const float* lp_lane = lp_g + lane;
  bool cok[CI];
#pragma unroll
  for (int k = 0; k < CI; ++k) cok[k] = lane + 32 * k < C;
  auto load_row = [&](float (&v)[CI], int t) {
#pragma unroll
    for (int k = 0; k < CI; ++k) v[k] = (cok[k] && t >= 0 && t < Tb) ? __ldg(lp_lane + (size_t)t * C + 32 * k) : LR_NEG_INF;
  };
  auto emissions = [&](const float (&pr)[CI], float& pb, float& pl) {
    pb = __shfl_sync(0xffffffffu, pr[0], 0);
    float x = __shfl_sync(0xffffffffu, pr[0], src_lane);
#pragma unroll
    for (int k = 1; k < CI; ++k) {
      const float y = __shfl_sync(0xffffffffu, pr[k], src_lane);
      x = src_slab == k ? y : x;
    }
    pl = x * hl;
    pb *= hb;
  };

  // ---- alpha sweep ---------------------------------------------------------------------------------
  float ab = lane == 0 ? 1.f : 0.f, al = 0.f;                        // "alpha_-1": makes frame 0 the generic step
  float kmant = 1.f;          // product of the mantissas of the applied factors since the last flush (each in [0.5, 1))
  int kexpo = 0;              // sum of their binary exponents
  float klog = 0.f;           // log of the flushed mantissa products
  float m_pending = 1.f, k_pending = 1.f;   // lattice mass at the last even frame and its reciprocal
  bool bad = false;
  float cur[4][CI], nxtv[4][CI];
#pragma unroll
  for (int u = 0; u < 4; ++u) load_row(cur[u], u);
  for (int t0 = 0; t0 < Tb; t0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) load_row(nxtv[u], t0 + 4 + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u;
      if (t < Tb) {                                                  // warp-uniform
        float pr[CI], pb, pl;
#pragma unroll
        for (int k = 0; k < CI; ++k) pr[k] = __expf(cur[u][k]);
        emissions(pr, pb, pl);
        float prev = __shfl_up_sync(0xffffffffu, al, 1);
        prev = lane == 0 ? 0.f : prev;
        float nb = (ab + prev) * pb;
        float nl = fmaf(prev, skipa, al + ab) * pl;
        if (!(u & 1)) {                                              // even frame: rescale (k = 1 at frame 0)
          bad = bad || !(m_pending >= 1e-24f) || !(m_pending < 1e30f);
          const float kf = k_pending;
          nb *= kf;
          nl *= kf;
          if (lane == 0) kfac[t >> 1] = kf;
          const uint32_t kb = __float_as_uint(kf);
          kexpo += (int)((kb >> 23) & 0xffu) - 126;
          kmant *= __uint_as_float((kb & 0x007fffffu) | 0x3f000000u);
          if ((t & 63) == 62) { klog += logf(kmant); kmant = 1.f; }  // 32 mantissas >= 2^-32: no underflow
          m_pending = lr_warp_sum(nb + nl);                          // consumed two frames on
          k_pending = __fdividef(1.f, m_pending);
        }
        ab = nb;
        al = nl;
        alpha[(size_t)t * 32 + lane] = make_float2(ab, al);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < CI; ++k) cur[u][k] = nxtv[u][k];
  }
  const float l1 = __shfl_sync(0xffffffffu, ab, L);
  const float l2 = L > 0 ? __shfl_sync(0xffffffffu, al, L - 1) : 0.f;
  const float rho = l1 + l2;
  bad = bad || !(rho >= 1e-30f) || !(rho < 1e30f);
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) redo[b] = 1;
    return;
  }
  // log p(l|x) = log rho - sum_t log k_t
  const float nll = (klog + logf(kmant) + (float)kexpo * 0.6931471805599453f) - logf(rho);
  const float rinv = 1.f / rho;
  for (int t = T - 1; t >= Tb; --t)
    for (int c = lane; c < C; c += 32) g_b[(size_t)t * C + c] = 0.f;
  // ---- beta sweep fused with the gradient rows -------------------------------------------------------
  __syncwarp();
  float bb = lane == L ? 1.f : 0.f, bl = 0.f;                         // "beta_T": makes frame Tb-1 the generic step
  float2 acur[4], anxt[4];
  const int t_top = (Tb - 1) & ~3;
  auto load_alpha = [&](float2& a, int t) {
    a = (t >= 0 && t < Tb) ? alpha[(size_t)t * 32 + lane] : make_float2(0.f, 0.f);
  };
#pragma unroll
  for (int u = 0; u < 4; ++u) { load_row(cur[u], t_top + u); load_alpha(acur[u], t_top + u); }
  float chk = 0.f;
  for (int t0 = t_top; t0 >= 0; t0 -= 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { load_row(nxtv[u], t0 - 4 + u); load_alpha(anxt[u], t0 - 4 + u); }
#pragma unroll
    for (int u = 3; u >= 0; --u) {
      const int t = t0 + u;
      if (t < Tb) {                                                  // warp-uniform
        float pr[CI], pb, pl;
#pragma unroll
        for (int k = 0; k < CI; ++k) pr[k] = __expf(cur[u][k]);
        emissions(pr, pb, pl);
        const float rb = __shfl_down_sync(0xffffffffu, bb, 1), rl = __shfl_down_sync(0xffffffffu, bl, 1);
        // (lane 31 has no label state when L <= 31, so its wrapped-around neighbours are never used)
        const float tl = hl * (fmaf(rl, skipb, bl + rb));
        const float tb = hb * (bb + bl);
        const float eb = acur[u].x * tb, el = acur[u].y * tl;
        erow[lane] = el;
        const float zb = lr_warp_sum(eb);
        __syncwarp();
        float acc = el + erow[n1] + erow[n2];
        if (deep) {                                                  // warp-uniform, rare
          if (is_head) {
            unsigned rest = later3;
            while (rest) { acc += erow[__ffs(rest) - 1]; rest &= rest - 1; }
          }
        }
        occ[occ_dst] = acc;
        if (lane == 0) occ[0] = zb;
        __syncwarp();
        float* g_row = g_b + (size_t)t * C + lane;
#pragma unroll
        for (int k = 0; k < CI; ++k)
          if (cok[k]) g_row[32 * k] = fmaf(-occ[lane + 32 * k], rinv, pr[k]);
        bb = tb * pb;
        bl = tl * pl;
        if (!(u & 1)) {
          const float kf = kfac[t >> 1];
          bb *= kf;
          bl *= kf;
        }
        chk = fmaxf(chk, bb + bl);                                   // (NaN-propagating check below)
        __syncwarp();
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acur[u] = anxt[u];
#pragma unroll
      for (int k = 0; k < CI; ++k) cur[u][k] = nxtv[u][k];
    }
  }
  // a factor that overflowed poisons bb/bl (inf or NaN) from that frame on: look at the last frame's values
  const float fin = bb + bl;
  if (__any_sync(0xffffffffu, !(fin < 1e30f) || !(chk < 1e30f))) {
    if (lane == 0) redo[b] = 1;
    return;
  }
  if (lane == 0) { nll_out[b] = nll; redo[b] = 0; }
}

size_t linear_kernel_smem(int T, int C) {
  const int CI = (C + 31) / 32;
  const int n_kf = (T + 1) / 2 + 1;
  const size_t per_clip = ((size_t)(CI * 32 + 1 + 33 + n_kf) * sizeof(float) + 15) / 16 * 16;
  return per_clip * kLinWarps;
}

size_t warp_kernel_smem(int T, int C, int P, int global_alpha = 0) {
  const int Cpad = (C + 31) / 32 * 32;
  return ((global_alpha ? 0 : (size_t)T * 64 * P) + 9 * Cpad + 32 * P) * sizeof(float) + (size_t)2 * 32 * P * sizeof(int);
}
// warp-per-clip kernel: label pairs per lane for a target of up to Lmax labels (0: does not fit)
int warp_kernel_pairs(int Lmax) {
  const int S = 2 * Lmax + 1;
  return S <= 64 ? 1 : (S <= 128 ? 2 : (S <= 256 ? 4 : 0));
}
bool warp_kernel_chosen(int B, int T, int C, int Lmax, int lr_ctc_force_block_kernel) {
  const int P = warp_kernel_pairs(Lmax);
  // the linear-space kernel (labels <= 31 symbols, <= 96 classes) beats the CTA-per-clip kernel from a few dozen clips
  // on (50 vs 74 us at B = 256, measured); the log-space warp kernel only pays off once the SMs are full
  const bool linear = P == 1 && Lmax <= 31 && C <= 96 && lr_ctc_force_block_kernel != 2;
  return P && warp_kernel_smem(T, C, P) <= 100 * 1024 && lr_ctc_force_block_kernel != 1 &&
         (B >= 1024 || (linear && B >= 64) || lr_ctc_force_block_kernel >= 2);
}

// Greedy CTC decode (semantics of the reference's GreedyDecoder,
// src/models/lipreader/decoder.py:165-197): per frame arg-max, collapse repeats, drop the blank.
// One warp per clip: lanes split the classes for the arg-max (lowest index wins ties, like
// torch.argmax on CPU), then a ballot + prefix count compacts 32 frames at a time.
__global__ void __launch_bounds__(128)
ctc_greedy_decode_kernel(const float* __restrict__ lp, const int32_t* __restrict__ lens, int B, int T, int C,
                         int32_t* __restrict__ tokens, int32_t* __restrict__ out_lens) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp;
  int Tb = lens[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  const float* base = lp + (size_t)b * T * C;
  int32_t* out = tokens + (size_t)b * T;
  int n_out = 0, prev = -1;
  for (int t0 = 0; t0 < Tb; t0 += 32) {
    // arg-max of frames t0..t0+31: every lane computes all of them cooperatively, lane k keeps frame t0+k
    int mine = 0;
    for (int k = 0; k < 32 && t0 + k < Tb; ++k) {
      const float* row = base + (size_t)(t0 + k) * C;
      float best = -INFINITY;
      int arg = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        float v = row[c];
        if (v > best) { best = v; arg = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      }
      if (lane == k) mine = arg;
    }
    const bool in_range = t0 + lane < Tb;
    int left = __shfl_up_sync(0xffffffffu, mine, 1);
    if (lane == 0) left = prev;
    const bool keep = in_range && mine != 0 && mine != left;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out[n_out + __popc(mask & ((1u << lane) - 1))] = mine;
    n_out += __popc(mask);
    const int last = min(31, Tb - t0 - 1);
    prev = __shfl_sync(0xffffffffu, mine, last);
  }
  for (int i = n_out + lane; i < T; i += 32) out[i] = 0;
  if (lane == 0) out_lens[b] = n_out;
}

}  // namespace

extern "C" int lr_ctc_greedy_decode(const float* log_probs, const int32_t* lens, int B, int T, int C,
                                    int32_t* tokens, int32_t* out_lens, void* stream) {
  LR_CHECK_ARG(log_probs && lens && tokens && out_lens && B > 0 && T > 0 && C > 0, "lr_ctc_greedy_decode: bad args");
  ctc_greedy_decode_kernel<<<lr_div_up(B, 4), 128, 0, lr_stream(stream)>>>(log_probs, lens, B, T, C, tokens, out_lens);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" size_t lr_ctc_workspace(int B, int T, int C, int Lmax, int kernel) {
  if (B <= 0 || T <= 0 || C <= 0 || Lmax < 0) return 0;
  if (Lmax == 0) Lmax = 1;
  if (warp_kernel_chosen(B, T, C, Lmax, kernel))       // alpha lattice of the warp-per-clip kernel: [B][T][64*P] f32
    return (size_t)B * T * 64 * warp_kernel_pairs(Lmax) * sizeof(float) + (size_t)B * sizeof(int32_t);   // + redo flags
  CtcPlan p = make_plan(T, C, Lmax);
  if (p.lat_in_smem) return 16;
  return (size_t)B * 2 * T * p.Smax * sizeof(float);
}

extern "C" int lr_ctc_fwd_bwd(const float* log_probs, const int32_t* targets,
                              const int32_t* input_lens, const int32_t* target_lens, int B, int T,
                              int C, int Lmax, float* nll, float* grad, void* workspace,
                              size_t ws_bytes, int kernel, void* stream) {
  const int lr_ctc_force_block_kernel = kernel;      // per-call choice (0 auto), see lr_b200.h
  LR_CHECK_ARG(log_probs && targets && input_lens && target_lens && nll,
               "lr_ctc_fwd_bwd: null pointer");
  LR_CHECK_ARG(B > 0 && T > 0 && C > 1 && Lmax >= 0, "lr_ctc_fwd_bwd: bad shape B=%d T=%d C=%d L=%d",
               B, T, C, Lmax);
  if (Lmax == 0) Lmax = 1;  // keep array extents non-zero; target_lens still clamp to 0
  // warp-per-clip kernel whenever the alpha lattice of a clip fits a modest slice of shared memory: one warp per
  // clip maximises clips in flight per SM (throughput); with few clips the CTA-per-clip kernel (alpha and beta on two
  // warps, 4-warp gradient) has the shorter critical path.  With a workspace for the lattice it runs the global-alpha
  // variant (32 clips per SM instead of 9).
  if (warp_kernel_chosen(B, T, C, Lmax, kernel)) {
    const int P = warp_kernel_pairs(Lmax);
    const size_t ws_need = (size_t)B * T * 64 * P * sizeof(float);
    const int ga = (grad && workspace && ws_bytes >= ws_need) ? 1 : 0;
    const size_t sm = warp_kernel_smem(T, C, P, ga);
    float* aw = reinterpret_cast<float*>(workspace);
    cudaStream_t st = lr_stream(stream);
    const int32_t* redo_ptr = nullptr;
    // labels of <= 31 symbols: the linear-space kernel first, then the log-space one on the clips it flagged
    if (P == 1 && Lmax <= 31 && C <= 96 && ga && grad && lr_ctc_force_block_kernel != 2 && ws_bytes >= ws_need + (size_t)B * sizeof(int32_t)) {
      int32_t* redo = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + ws_need);
      const dim3 lgrid(lr_div_up(B, kLinWarps));
      const size_t lsm = linear_kernel_smem(T, C);
      float2* a2 = reinterpret_cast<float2*>(aw);
      static const int mb = getenv("LR_CTC_LIN_MB") ? atoi(getenv("LR_CTC_LIN_MB")) : 11;     // tuning hook
#define LR_LAUNCH_LIN(CI_, MB_)                                                                                \
  ctc_linear_warp_kernel<CI_, MB_><<<lgrid, 32 * kLinWarps, lsm, st>>>(log_probs, targets, input_lens, target_lens, B, T, \
                                                                      C, Lmax, nll, grad, a2, redo)
#define LR_LAUNCH_LIN_MB(CI_)                                                                                  \
  do {                                                                                                         \
    if (mb >= 16) LR_LAUNCH_LIN(CI_, 16); else if (mb >= 12) LR_LAUNCH_LIN(CI_, 12); else LR_LAUNCH_LIN(CI_, 11);   \
  } while (0)
      if (C <= 32) LR_LAUNCH_LIN_MB(1);
      else if (C <= 64) LR_LAUNCH_LIN_MB(2);
      else LR_LAUNCH_LIN_MB(3);
#undef LR_LAUNCH_LIN_MB
#undef LR_LAUNCH_LIN
      LR_CHECK_LAUNCH();
      redo_ptr = redo;
    }
#define LR_LAUNCH_WARP2(PP, FF, GG)                                                                            \
  do {                                                                                                         \
    LR_CHECK_CUDA(cudaFuncSetAttribute(ctc_warp_kernel<PP, FF, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)sm));                                                              \
    ctc_warp_kernel<PP, FF, GG><<<B, 32, sm, st>>>(log_probs, targets, input_lens, target_lens, B, T, C, Lmax, \
                                                   nll, grad, aw, redo_ptr);                                   \
  } while (0)
#define LR_LAUNCH_WARP(PP)                                                                                     \
  do {                                                                                                         \
    if (T <= 128) {          /* fast exp/log: chains this short stay inside the 1e-4 parity bar */            \
      if (ga) LR_LAUNCH_WARP2(PP, true, 1); else LR_LAUNCH_WARP2(PP, true, 0);                                 \
    } else {                                                                                                   \
      if (ga) LR_LAUNCH_WARP2(PP, false, 1); else LR_LAUNCH_WARP2(PP, false, 0);                               \
    }                                                                                                          \
  } while (0)
    if (P == 1) LR_LAUNCH_WARP(1);
    else if (P == 2) LR_LAUNCH_WARP(2);
    else LR_LAUNCH_WARP(4);
#undef LR_LAUNCH_WARP
#undef LR_LAUNCH_WARP2
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
  CtcPlan p = make_plan(T, C, Lmax);
  if (!p.lat_in_smem) {
    size_t need = (size_t)B * 2 * T * p.Smax * sizeof(float);
    if (!workspace || ws_bytes < need) {
      lr_set_error("lr_ctc_fwd_bwd: workspace %zu < %zu", ws_bytes, need);
      return LR_EWORKSPACE;
    }
  }
  static thread_local size_t configured = 0;
  if (p.smem_bytes > 48 * 1024 && p.smem_bytes > configured) {
    LR_CHECK_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_grad_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kCtcSmemCap));
    configured = kCtcSmemCap;
  }
  ctc_alpha_beta_grad_kernel<<<B, kCtcThreads, p.smem_bytes, lr_stream(stream)>>>(
      log_probs, targets, input_lens, target_lens, B, T, C, Lmax, nll, grad,
      reinterpret_cast<float*>(workspace), p);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
