// rnn_cluster.cu — persistent thread-block-cluster recurrent kernels,
// the throughput path of the (bi)directional GRU / LSTM / tanh-RNN layer.
//
// The recurrence is sequential in time but independent across clips, and one step is a tiny GEMM
// (B x H) . (H x G*H): launch latency and re-reading W_hh every step dominate a per-step design
// (rnn.cu, the fp32 parity path).  Here ONE launch runs all T steps:
//   * a cluster of 8 CTAs owns one (direction, 32-clip slice); CTA r owns hidden units
//     [r*H/8, (r+1)*H/8) for every gate, with its W_hh rows resident in shared memory as bf16 for
//     the whole sequence (GRU-256: 96 x 256 x 2 B = 48 KB per CTA, 384 KB per cluster);
//   * per step each warp computes a (gates x 16 units x 8 clips) tile with mma.sync m16n8k16
//     (bf16 operands, fp32 accumulate) straight from shared memory via ldmatrix — low-latency
//     warp-level MMA is the right tool for a dependent 75-step chain, the accumulator fragment
//     layout puts all gates of a (unit, clip) pair in the same thread so the gate non-linearities,
//     length masking, output / saved-activation stores are register-local;
//   * the fp32 state lives in registers; its bf16 image is staged in shared memory and pushed to
//     all 8 CTAs of the cluster with 16-byte distributed-shared-memory stores, then ONE cluster
//     barrier per step (no grid-wide sync, no global-memory round trip for h).
// Backward mirrors it: CTA r owns W_hh^T rows of its units, the per-step exchange is the
// gate-gradient slice (G*H/8 values per clip), dh is carried in registers.
//
// Supported when H % 128 == 0 and the slices fit shared memory (GRU/LSTM-256, GRU-512, ...);
// other shapes use the per-step kernels.
#include "common.cuh"
#include <cooperative_groups.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kCS = 8;         // CTAs per cluster
constexpr int kBS = 32;        // clips per cluster
constexpr int kPad = 8;        // bf16 elements of row padding (16 B): conflict-free ldmatrix

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
// address of `local_smem_addr` inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct CFwd {
  const float* gi; const float* w_hh; const float* b_hh; const int32_t* lens;
  float* hidden; float* saved; float* h_n; float* c_n;
  int B, T, H, D;
};

// ------------------------------------------------------------------------------------------------
// forward: grid = kCS * (D * ceil(B/kBS)), cluster (kCS,1,1), threads = (UH/16)*(kBS/8)*32
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void rnn_cluster_fwd_kernel(CFwd p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, D = p.D;
  const int UH = H / kCS;
  const int pitch = H + kPad;                       // bf16 elements per smem row
  __nv_bfloat16* W_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);                    // [G*UH][pitch]
  __nv_bfloat16* h_s = W_s + (size_t)G * UH * pitch;                                  // [2][kBS][pitch]
  __nv_bfloat16* stage = h_s + (size_t)2 * kBS * pitch;                               // [kBS][UH]

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int cid = blockIdx.x / kCS;
  const int d = cid % D, b0 = (cid / D) * kBS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_nt = kBS / 8;                         // n-tiles (8 clips each)
  const int ub = warp / n_nt, nt = warp % n_nt;     // this warp: units [ub*16,+16) x clips [nt*8,+8)
  const int u_base = rank * UH;                     // first global unit of this CTA

  // ---- one-time: W_hh slice -> bf16 smem, zero both h buffers -----------------------------------
  const float* Wd = p.w_hh + (size_t)d * G * H * H;
  for (int i = tid; i < G * UH * (H / 4); i += blockDim.x) {
    int row = i / (H / 4), c4 = i - row * (H / 4);
    int g = row / UH, ul = row - g * UH;
    float4 v = *reinterpret_cast<const float4*>(Wd + ((size_t)g * H + u_base + ul) * H + c4 * 4);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    *reinterpret_cast<uint2*>(W_s + (size_t)row * pitch + c4 * 4) = pk;
  }
  for (int i = tid; i < 2 * kBS * pitch / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(h_s)[i] = 0u;
  __syncthreads();
  cluster_sync_();

  // ---- per-thread constants: the 4 accumulator elements this thread owns --------------------------
  // element i: unit row = (lane/4) + 8*(i/2), clip col = 2*(lane%4) + (i%2)
  int ug[4], bl[4], bg[4], len[4];
  float bh[G][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ug[i] = u_base + ub * 16 + (lane >> 2) + 8 * (i >> 1);
    bl[i] = nt * 8 + 2 * (lane & 3) + (i & 1);
    bg[i] = b0 + bl[i];
    len[i] = bg[i] < B ? p.lens[bg[i]] : 0;
#pragma unroll
    for (int g = 0; g < G; ++g) bh[g][i] = p.b_hh[(size_t)d * G * H + g * H + ug[i]];
  }
  float hst[4] = {0.f, 0.f, 0.f, 0.f}, cst[4] = {0.f, 0.f, 0.f, 0.f};

  const uint32_t W_addr = lr_smem_u32(W_s), h_addr = lr_smem_u32(h_s), stage_addr = lr_smem_u32(stage);
  // ldmatrix lane addresses (bytes)
  const uint32_t a_lane = (uint32_t)(((ub * 16 + (lane & 15)) * pitch + (lane >> 4) * 8) * 2);
  const uint32_t b_lane = (uint32_t)(((nt * 8 + (lane & 7)) * pitch + ((lane >> 3) & 1) * 8) * 2);
  const int ksteps = H / 16;

  // ---- step-invariant addressing, hoisted out of the 75-step chain --------------------------------
  // global rows of this thread's 4 elements: pointers advance by one time step per iteration
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
  // the per-step push of this CTA's [kBS x UH] state slice: 4 16-byte chunks per thread, source offset and
  // remote (DSMEM) destination are the same every step (the integer divisions used to sit inside the time loop)
  constexpr int kPush = 4;
  uint32_t push_src[kPush], push_dst[kPush];
  {
    const int chunks_per_row = UH / 8;                 // 16-byte chunks
    const int n_chunks = kBS * chunks_per_row;
#pragma unroll
    for (int k = 0; k < kPush; ++k) {
      const int i = tid + k * (int)blockDim.x;
      const int dst = i / n_chunks, c = i - dst * n_chunks;
      const int rowb = c / chunks_per_row, ch = c - rowb * chunks_per_row;
      push_src[k] = stage_addr + (uint32_t)((rowb * UH + ch * 8) * 2);
      push_dst[k] = mapa(h_addr + (uint32_t)((rowb * pitch + u_base + ch * 8) * 2), (uint32_t)dst);
    }
  }
  const uint32_t buf_bytes = (uint32_t)(kBS * pitch * 2);

  for (int step = 0; step < T; ++step) {
    const int tt = d == 0 ? step : T - 1 - step;
    const int cur = step & 1;
    // (A) issue this step's gi loads early
    float giv[G][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) giv[g][i] = (bg[i] < B) ? gi_p[i][(size_t)g * H] : 0.f;
      gi_p[i] += gi_step;
    }
    // (B) gates_pre = W_slice . h_prev^T; two accumulator chains per gate (even / odd k-steps) halve the
    // dependent HMMA chain of the step
    float acc[G][4], acc2[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
      acc2[g][0] = acc2[g][1] = acc2[g][2] = acc2[g][3] = 0.f;
    }
    if (step > 0) {
      const uint32_t hb = h_addr + (uint32_t)cur * buf_bytes + b_lane;
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ks += 2) {
        uint32_t b0r, b1r, c0r, c1r;
        ldsm_x2(hb + ks * 32, b0r, b1r);
        ldsm_x2(hb + ks * 32 + 32, c0r, c1r);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          uint32_t a0, a1, a2, a3, e0, e1, e2, e3;
          ldsm_x4(W_addr + (uint32_t)(g * UH * pitch * 2) + a_lane + ks * 32, a0, a1, a2, a3);
          ldsm_x4(W_addr + (uint32_t)(g * UH * pitch * 2) + a_lane + ks * 32 + 32, e0, e1, e2, e3);
          mma_bf16(acc[g], a0, a1, a2, a3, b0r, b1r);
          mma_bf16(acc2[g], e0, e1, e2, e3, c0r, c1r);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        acc[g][0] += acc2[g][0]; acc[g][1] += acc2[g][1]; acc[g][2] += acc2[g][2]; acc[g][3] += acc2[g][3];
      }
    }
    // (C) gate math, masking; the outputs stay in registers until after the cluster arrive
    float o_h[4], o_sv[5][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = bg[i] < B;
      const bool active = valid && tt < len[i];
      float hnew, sv0 = 0.f, sv1 = 0.f, sv2 = 0.f, sv3 = 0.f, sv4 = 0.f;
      if (MODE == LR_RNN_GRU) {
        float ghn = acc[2][i] + bh[2][i];
        float r = sigm(giv[0][i] + acc[0][i] + bh[0][i]);
        float z = sigm(giv[1][i] + acc[1][i] + bh[1][i]);
        float n = tanh_(giv[2][i] + r * ghn);
        hnew = (1.f - z) * n + z * hst[i];
        sv0 = r; sv1 = z; sv2 = n; sv3 = ghn;
      } else if (MODE == LR_RNN_LSTM) {
        float ig = sigm(giv[0][i] + acc[0][i] + bh[0][i]);
        float fg = sigm(giv[1][i] + acc[1][i] + bh[1][i]);
        float gg = tanh_(giv[2][i] + acc[2][i] + bh[2][i]);
        float og = sigm(giv[3][i] + acc[3][i] + bh[3][i]);
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
      stage[bl[i] * UH + (ug[i] - u_base)] = __float2bfloat16(hst[i]);
    }
    __syncthreads();
    // (D) push this CTA's [kBS x UH] bf16 slice into h_s[next] of every CTA of the cluster
    if (step + 1 < T) {
      const uint32_t nxt_off = (uint32_t)(cur ^ 1) * buf_bytes;
#pragma unroll
      for (int k = 0; k < kPush; ++k) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(push_src[k]));
        st_cluster_v4(push_dst[k] + nxt_off, v);
      }
    }
    // (E) one cluster barrier per step (also orders the stage buffer reuse).  The release only has to cover the
    // DSMEM pushes: this step's global stores are issued between arrive and wait, so the barrier never waits for
    // HBM write acknowledgements (they were 14 % of the kernel's stall samples in front of the arrive)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
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
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // final states
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (bg[i] < B) {
      p.h_n[((size_t)d * B + bg[i]) * H + ug[i]] = hst[i];
      if (MODE == LR_RNN_LSTM) p.c_n[((size_t)d * B + bg[i]) * H + ug[i]] = cst[i];
    }
}

struct CBwd {
  const float* d_hidden; const float* d_h_n; const float* d_c_n; const float* saved; const float* hidden;
  const float* w_hh; const int32_t* lens;
  float* d_gi; float* d_gh; float* h_prev_all;
  int B, T, H, D;
};

// ------------------------------------------------------------------------------------------------
// backward: same clustering; CTA r owns W_hh^T rows of its units (K = G*H) and exchanges the
// gate-gradient slice every step.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void rnn_cluster_bwd_kernel(CBwd p) {
  constexpr int G = Gates<MODE>::G, S = Gates<MODE>::S;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, D = p.D;
  const int UH = H / kCS, GH = G * H;
  const int pitch = GH + kPad;
  __nv_bfloat16* WT_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);                   // [UH][pitch]
  __nv_bfloat16* g_s = WT_s + (size_t)UH * pitch;                                     // [2][kBS][pitch]
  __nv_bfloat16* stage = g_s + (size_t)2 * kBS * pitch;                               // [kBS][G*UH]

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int cid = blockIdx.x / kCS;
  const int d = cid % D, b0 = (cid / D) * kBS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_nt = kBS / 8;
  const int ub = warp / n_nt, nt = warp % n_nt;
  const int u_base = rank * UH;

  // W_hh^T slice: WT_s[ul][r] = W_hh[d][r][u_base+ul]  (coalesced over ul)
  const float* Wd = p.w_hh + (size_t)d * GH * H;
  for (int i = tid; i < GH * UH; i += blockDim.x) {
    int r = i / UH, ul = i - r * UH;
    WT_s[(size_t)ul * pitch + r] = __float2bfloat16(Wd[(size_t)r * H + u_base + ul]);
  }
  for (int i = tid; i < 2 * kBS * pitch / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(g_s)[i] = 0u;
  __syncthreads();
  cluster_sync_();

  int ug[4], bl[4], bg[4], len[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ug[i] = u_base + ub * 16 + (lane >> 2) + 8 * (i >> 1);
    bl[i] = nt * 8 + 2 * (lane & 3) + (i & 1);
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
  const uint32_t WT_addr = lr_smem_u32(WT_s), g_addr = lr_smem_u32(g_s), stage_addr = lr_smem_u32(stage);
  const uint32_t a_lane = (uint32_t)(((ub * 16 + (lane & 15)) * pitch + (lane >> 4) * 8) * 2);
  const uint32_t b_lane = (uint32_t)(((nt * 8 + (lane & 7)) * pitch + ((lane >> 3) & 1) * 8) * 2);
  const int ksteps = GH / 16;

  // ---- step-invariant addressing (see the forward kernel) ------------------------------------------
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
  // offset (in elements) from time tt to the forward-previous time step tt_in = tt -/+ 1
  const long long prev_hid = (d == 0 ? -1 : 1) * (long long)D * H, prev_sv = (d == 0 ? -1 : 1) * (long long)D * S * H;
  // per-step push of this CTA's [kBS x G*UH] gate-gradient slice: 4*G 16-byte chunks per thread, fixed addresses
  constexpr int kPush = 4 * G;
  uint32_t push_src[kPush], push_dst[kPush];
  {
    const int cpr = UH / 8;                              // 16-byte chunks per (clip, gate)
    const int n_chunks = kBS * G * cpr;
#pragma unroll
    for (int k = 0; k < kPush; ++k) {
      const int i = tid + k * (int)blockDim.x;
      const int dst = i / n_chunks, c = i - dst * n_chunks;
      const int rowb = c / (G * cpr), rem = c - rowb * (G * cpr);
      const int g = rem / cpr, ch = rem - g * cpr;
      push_src[k] = stage_addr + (uint32_t)((rowb * (G * UH) + g * UH + ch * 8) * 2);
      push_dst[k] = mapa(g_addr + (uint32_t)((rowb * pitch + g * H + u_base + ch * 8) * 2), (uint32_t)dst);
    }
  }
  const uint32_t buf_bytes = (uint32_t)(kBS * pitch * 2);

  for (int step = 0; step < T; ++step) {
    const int tt = d == 0 ? T - 1 - step : step;          // reverse of the forward order
    const int cur = step & 1;
    // prefetch what the element-wise part needs
    float dout[4], svv[5][4], hprev[4], cprev[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = bg[i] < B;
      dout[i] = valid ? *dh_p[i] : 0.f;
      if (S > 0) {
#pragma unroll
        for (int k = 0; k < S; ++k) svv[k][i] = valid ? sv_p[i][(size_t)k * H] : 0.f;
      }
      const int tt_in = d == 0 ? tt - 1 : tt + 1;
      const bool has_prev = valid && ((d == 0) ? (tt_in >= 0) : (tt_in < len[i]));
      hprev[i] = has_prev ? hid_p[i][prev_hid] : 0.f;
      cprev[i] = (MODE == LR_RNN_LSTM && has_prev) ? sv_p[i][prev_sv + (long long)4 * H] : 0.f;
    }
    // dh contribution through W_hh: acc = WT_slice . dgh_prev^T — four independent accumulator chains
    // (k-step mod 4): a single chain of G*H/16 dependent HMMAs was the longest latency of the step
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (step > 0) {
      float ac1[4] = {0.f, 0.f, 0.f, 0.f}, ac2[4] = {0.f, 0.f, 0.f, 0.f}, ac3[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t gb = g_addr + (uint32_t)cur * buf_bytes + b_lane;
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ks += 4) {
        uint32_t bq[4][2], aq[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ldsm_x2(gb + (ks + q) * 32, bq[q][0], bq[q][1]);
          ldsm_x4(WT_addr + a_lane + (ks + q) * 32, aq[q][0], aq[q][1], aq[q][2], aq[q][3]);
        }
        mma_bf16(acc, aq[0][0], aq[0][1], aq[0][2], aq[0][3], bq[0][0], bq[0][1]);
        mma_bf16(ac1, aq[1][0], aq[1][1], aq[1][2], aq[1][3], bq[1][0], bq[1][1]);
        mma_bf16(ac2, aq[2][0], aq[2][1], aq[2][2], aq[2][3], bq[2][0], bq[2][1]);
        mma_bf16(ac3, aq[3][0], aq[3][1], aq[3][2], aq[3][3], bq[3][0], bq[3][1]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = (acc[i] + ac1[i]) + (ac2[i] + ac3[i]);
    }
    float o_gi[G][4], o_gh[G][4], o_hp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = bg[i] < B;
      const bool active = valid && tt < len[i];
      const float carry = dh_dir[i] + acc[i];
      float dgi[G], dgh[G];
#pragma unroll
      for (int g = 0; g < G; ++g) dgi[g] = dgh[g] = 0.f;
      float hp_out = 0.f;
      if (!active) {
        dh_dir[i] = carry;            // state passed through untouched
      } else {
        const float dh = dout[i] + carry;
        hp_out = hprev[i];
        if (MODE == LR_RNN_GRU) {
          const float r = svv[0][i], z = svv[1][i], n = svv[2][i], ghn = svv[3][i];
          const float dn_pre = dh * (1.f - z) * (1.f - n * n);
          const float dr_pre = dn_pre * ghn * r * (1.f - r);
          const float dz_pre = dh * (hprev[i] - n) * z * (1.f - z);
          dgi[0] = dr_pre; dgh[0] = dr_pre;
          dgi[1] = dz_pre; dgh[1] = dz_pre;
          dgi[2] = dn_pre; dgh[2] = dn_pre * r;
          dh_dir[i] = dh * z;
        } else if (MODE == LR_RNN_LSTM) {
          const float ig = svv[0][i], fg = svv[1][i], gg = svv[2][i], og = svv[3][i], cn = svv[4][i];
          const float tc = tanh_(cn);
          const float dcc = dc[i] + dh * og * (1.f - tc * tc);
          dgi[0] = dgh[0] = dcc * gg * ig * (1.f - ig);
          dgi[1] = dgh[1] = dcc * cprev[i] * fg * (1.f - fg);
          dgi[2] = dgh[2] = dcc * ig * (1.f - gg * gg);
          dgi[3] = dgh[3] = dh * tc * og * (1.f - og);
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
        stage[bl[i] * (G * UH) + g * UH + (ug[i] - u_base)] = __float2bfloat16(dgh[g]);
      }
    }
    __syncthreads();
    if (step + 1 < T) {
      const uint32_t nxt_off = (uint32_t)(cur ^ 1) * buf_bytes;
#pragma unroll
      for (int k = 0; k < kPush; ++k) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(push_src[k]));
        st_cluster_v4(push_dst[k] + nxt_off, v);
      }
    }
    // the global stores of the step sit between arrive and wait (see the forward kernel)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
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
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

size_t fwd_smem(int G, int H) {
  const int UH = H / kCS, pitch = H + kPad;
  return ((size_t)G * UH * pitch + (size_t)2 * kBS * pitch + (size_t)kBS * UH) * 2;
}
size_t bwd_smem(int G, int H) {
  const int UH = H / kCS, pitch = G * H + kPad;
  return ((size_t)UH * pitch + (size_t)2 * kBS * pitch + (size_t)kBS * G * UH) * 2;
}
int gates_of(int mode) { return mode == LR_RNN_GRU ? 3 : (mode == LR_RNN_LSTM ? 4 : 1); }

template <typename K, typename P>
int launch_cluster(K kernel, P params, int n_clusters, int threads, size_t smem, cudaStream_t st) {
  LR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_clusters * kCS);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, params));
  lr_count_launch();
  return LR_OK;
}

}  // namespace

// 1 when the persistent cluster kernels can run this shape (else use lr_rnn_fwd / lr_rnn_bwd)
extern "C" int lr_rnn_cluster_supported(int mode, int H) {
  if (mode < 0 || mode > 2 || H <= 0 || H % (16 * kCS) != 0 || H / kCS > 64) return 0;
  const int G = gates_of(mode);
  if (!(fwd_smem(G, H) <= 220 * 1024 && bwd_smem(G, H) <= 220 * 1024)) return 0;
  // the block is (H/kCS/16) * 4 warps: both kernels must fit the register file at that size
  const int threads = (H / kCS / 16) * (kBS / 8) * 32;
  const void* fns[2] = {
      mode == LR_RNN_GRU ? (const void*)rnn_cluster_fwd_kernel<LR_RNN_GRU>
                         : mode == LR_RNN_LSTM ? (const void*)rnn_cluster_fwd_kernel<LR_RNN_LSTM>
                                               : (const void*)rnn_cluster_fwd_kernel<LR_RNN_TANH>,
      mode == LR_RNN_GRU ? (const void*)rnn_cluster_bwd_kernel<LR_RNN_GRU>
                         : mode == LR_RNN_LSTM ? (const void*)rnn_cluster_bwd_kernel<LR_RNN_LSTM>
                                               : (const void*)rnn_cluster_bwd_kernel<LR_RNN_TANH>};
  for (int i = 0; i < 2; ++i) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, fns[i]) != cudaSuccess) { cudaGetLastError(); return threads <= 256; }
    if ((long long)fa.numRegs * threads > 65536) return 0;
  }
  return 1;
}

extern "C" int lr_rnn_cluster_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh,
                                  const int32_t* lens, int B, int T, int H, int D, float* hidden, float* h_n,
                                  float* c_n, float* saved, void* stream) {
  LR_CHECK_ARG(gi && w_hh && b_hh && lens && hidden && h_n, "lr_rnn_cluster_fwd: null pointer");
  LR_CHECK_ARG(lr_rnn_cluster_supported(mode, H), "lr_rnn_cluster_fwd: unsupported mode/hidden size (%d,%d)", mode, H);
  LR_CHECK_ARG(B > 0 && T > 0 && (D == 1 || D == 2), "lr_rnn_cluster_fwd: bad shape");
  LR_CHECK_ARG(mode != LR_RNN_LSTM || c_n, "lr_rnn_cluster_fwd: LSTM needs c_n");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_cluster_fwd: `saved` required");
  CFwd p;
  p.gi = gi; p.w_hh = w_hh; p.b_hh = b_hh; p.lens = lens; p.hidden = hidden; p.saved = saved; p.h_n = h_n;
  p.c_n = c_n; p.B = B; p.T = T; p.H = H; p.D = D;
  const int n_clusters = D * lr_div_up(B, kBS);
  const int threads = (H / kCS / 16) * (kBS / 8) * 32;
  const size_t smem = fwd_smem(gates_of(mode), H);
  cudaStream_t st = lr_stream(stream);
  if (mode == LR_RNN_GRU) return launch_cluster(rnn_cluster_fwd_kernel<LR_RNN_GRU>, p, n_clusters, threads, smem, st);
  if (mode == LR_RNN_LSTM) return launch_cluster(rnn_cluster_fwd_kernel<LR_RNN_LSTM>, p, n_clusters, threads, smem, st);
  return launch_cluster(rnn_cluster_fwd_kernel<LR_RNN_TANH>, p, n_clusters, threads, smem, st);
}

extern "C" int lr_rnn_cluster_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
                                  const float* saved, const float* hidden, const float* w_hh,
                                  const int32_t* lens, int B, int T, int H, int D, float* d_gi, float* d_gh,
                                  float* h_prev_all, void* stream) {
  LR_CHECK_ARG(d_hidden && hidden && w_hh && lens && d_gi && d_gh && h_prev_all, "lr_rnn_cluster_bwd: null pointer");
  LR_CHECK_ARG(lr_rnn_cluster_supported(mode, H), "lr_rnn_cluster_bwd: unsupported mode/hidden size (%d,%d)", mode, H);
  LR_CHECK_ARG(B > 0 && T > 0 && (D == 1 || D == 2), "lr_rnn_cluster_bwd: bad shape");
  LR_CHECK_ARG(mode == LR_RNN_TANH || saved, "lr_rnn_cluster_bwd: `saved` required");
  CBwd p;
  p.d_hidden = d_hidden; p.d_h_n = d_h_n; p.d_c_n = d_c_n; p.saved = saved; p.hidden = hidden; p.w_hh = w_hh;
  p.lens = lens; p.d_gi = d_gi; p.d_gh = d_gh; p.h_prev_all = h_prev_all; p.B = B; p.T = T; p.H = H; p.D = D;
  const int n_clusters = D * lr_div_up(B, kBS);
  const int threads = (H / kCS / 16) * (kBS / 8) * 32;
  const size_t smem = bwd_smem(gates_of(mode), H);
  cudaStream_t st = lr_stream(stream);
  if (mode == LR_RNN_GRU) return launch_cluster(rnn_cluster_bwd_kernel<LR_RNN_GRU>, p, n_clusters, threads, smem, st);
  if (mode == LR_RNN_LSTM) return launch_cluster(rnn_cluster_bwd_kernel<LR_RNN_LSTM>, p, n_clusters, threads, smem, st);
  return launch_cluster(rnn_cluster_bwd_kernel<LR_RNN_TANH>, p, n_clusters, threads, smem, st);
}
