// conv3d_sm100.cu — spatio-temporal conv front-end on tcgen05 / TMEM / TMA.
//
// North-star extension (the reference has no Conv3d; oracle = torch.nn.functional.conv3d fp32).
//
// Formulation: "shifted-window implicit GEMM over the zero-padded channels-last volume".
//   The activation lives in HBM as a zero-padded NDHWC bf16 volume X[B][Tp][Hp][Wp][Cin] with
//   Wp a power of two, i.e. a 2-D matrix X2D[rows = B*Tp*Hp*Wp][Cin].  For a stride-1 conv the
//   A-operand row of output pixel (t,y,x) and tap (kt,ky,kx) is X2D[p(t,y,x) + kt*Hp*Wp + ky*Wp + kx]:
//   every tap is a pure ROW SHIFT.  So a CTA
//     1. TMA-loads, once per work item, the KT+J-1 plane "chunks" (128 + halo rows x Cin, hardware
//        swizzled, K-major) that cover J consecutive output frames of one 128-position tile;
//     2. streams the per-tap weight tiles [Cout x Cin] through a small TMA ring;
//     3. issues, per tap, J x Cin/16 tcgen05.mma (M=128, N=Cout, K=16) whose A descriptors simply
//        start (ky*Wp+kx) rows into the resident chunk — the input is read from L2 once, not once
//        per tap, and the weights once per J tiles;
//     4. keeps the J accumulators (128 lanes x Cout fp32 columns each) in TMEM;
//     5. drains them with tcgen05.ld in a 4-warp epilogue that fuses bias + ReLU + MaxPool(1,2,2)
//        (+ pool-argmax bytes for backward) and writes the pooled bf16 tile straight into the
//        interior of the NEXT layer's zero-padded volume.
//   The tile is 128/Wp full padded rows; positions x >= W and y >= H are junk lanes whose results
//   are never stored (their windows wrap across rows, which is harmless).
// The same kernel run on the un-pooled output gradient with flipped/transposed weights is dgrad.
//
// Warp roles (288 threads): warp 0 = TMA producer, warps 1-4 = MMA issuers (one elected lane each,
// accumulator j belongs to issuer j % 4; warp 1 also owns the TMEM allocation), warps 5-8 = epilogue
// (TMEM lane quarter = warp_id % 4).
#include "tcgen05.cuh"
#include <string.h>
#include <stdlib.h>

using namespace lr_tc;

namespace {

constexpr int kMmaWarps = 4;                      // MMA-issuing warps (accumulator j is issued by warp j % 4)
constexpr int kEpiGroups = 2;                     // epilogue warp quartets (each covers the four TMEM lane quarters)
constexpr int kExtraWarps = 2;                    // warps 13, 14: with the issuing warps 3, 4 that orientation 3 leaves idle
                                                  // they form a THIRD epilogue quartet (lane quarters 3, 0, 1, 2)
constexpr int kThreads = 32 * (1 + kMmaWarps + 4 * kEpiGroups + kExtraWarps);   // producer + issuers + epilogue warps
// (15 warps: no scheduler hosts more than four of them, so the 128-register cap per thread is unchanged)
constexpr int kMaxChunks = 16;   // (J + KT - 1) * channel groups
constexpr int kWStages = 4;
constexpr int kMaxTaps = 80;
constexpr int kUnpoolIters = 8; // epi_mode 2: (row, channel-group) items per epilogue thread = Cout/8 <= 8
constexpr int kEpiIters = 4;    // pooled (row,x,channel-group) items per epilogue thread: <= 32*16/128

struct ConvParams {
  int B, T, H, W;              // valid (un-padded) extents, same for input and conv output
  int Tp, Hp, Wp;              // padded input extents
  int KT, KH, KW;
  int Cin;                     // channels per group row: 16, 32 or 64 (32/64/128-byte rows)
  int CG;                      // channel groups (input stored group-major: CG volumes)
  int Cout;                    // N, multiple of 32, <= 128
  int R;                       // tile rows = 128 / Wp
  int J;                       // accumulators (consecutive frames) per work item
  int n_sets;                  // TMEM accumulator sets (2 = epilogue of item i overlaps MMAs of item i+1)
  int a_set_bytes;             // bytes of one chunk set
  int a_sets;                  // chunk-set buffers (2 = next item's planes load while this item computes)
  int kxs;                     // kx taps stacked on N (1, or KW in mode 2: one MMA per (kt,ky), N = KW*Cout; the
                               // epilogue adds the KW column blocks with a row shift of kx each)
  int n_eff_taps;              // weight tiles per group = KT*KH*KW / kxs
  int smem_off_halo;           // mode 2: rows a warp's first lanes hand to the previous warp
  int swap;                    // 1: D^T = W . X^T  (M = Cout lanes, N = 128 positions): weights are the A operand
  int Mt;                      // UMMA M in swap mode (64 or 128)
  int acc_cols;                // TMEM columns per accumulator (Cout, or 128 in swap mode)
  int tps;                     // filter taps per weight stage (amortises the per-stage barrier round trip)
  int w_stages;                // weight ring depth (<= kWStages)
  int n_ytiles, n_tgroups, n_items;
  int CH;                      // chunk rows
  int chunk_bytes;             // 1024-aligned
  int wtile_bytes;             // 1024-aligned
  int tmem_cols;
  int epi_mode;                // 0: bias+ReLU+pool(1,2,2) ; 1: plain store of valid positions ; 2: fused un-pooling
                               // (dgrad output routed to the arg-max slot of its 2x2 window + bias gradient)
  const uint8_t* am_in;        // epi_mode 2: arg-max bytes (B,T,H,W,Cout) of the pooling layer being undone
  float* d_bias;               // epi_mode 2: per-channel sum of the routed gradient (atomically accumulated)
  int has_bias;
  int oTp, oHp, oWp, o_t, o_y, o_x;   // output volume geometry / interior offset
  long long rows_per_group;    // B*Tp*Hp*Wp
  const float* bias;
  long long* dbg;              // optional [grid][8] cycle counters (lr_conv3d_set_debug): where the roles wait
  const uint8_t* w_packed;     // [CG][taps] tile images of wtile_bytes each (swizzled like a TMA box load)
  __nv_bfloat16* y;
  uint8_t* argmax;
  uint32_t idesc;
  uint32_t desc_hi;            // SBO | version | layout type (upper 32 bits of the smem descriptor)
  int row_bytes;
  int smem_off_w, smem_off_stage, smem_off_bar;
  int stage_pitch;             // bytes per staging row
  int stage_bufs;              // 1 or 2 staging tiles per epilogue group (2: one named barrier per accumulator instead of two)
  int epi_groups;              // epilogue warp quartets in use: group e drains the accumulators j = e, e + groups, ...
  int cg_shift, wp_shift;      // log2(Cout/8) (or -1) and log2(Wp) for the plain-store epilogue
  uint32_t tap_off[kMaxTaps];  // per tap: descriptor offset (16-byte units) of its window inside a chunk set
                               // = kt * chunk + (ky*Wp + kx) rows — read with uniform constant loads
  int n_issuers;               // MMA-issuing warps that take part in the barrier protocol
  int Jg;                      // mode 3: frames per issuing warp (J split into n_issuers runs)
  int seam;                    // mode 3: 1 = the item's chunks form ONE list shared out between the issuing warps (every
                               // input plane is multiplied once, with the widest window it has); the accumulators of the
                               // KT-1 frames at a hand-over point then receive MMAs from two threads (sum order not fixed)
  int stage_fence;             // diagnostics: tcgen05.fence::after_thread_sync after every weight-stage wait (old behaviour)
  int skip;                    // diagnostics (lr_conv3d_set_debug_skip): 1 no epilogue work, 2 weights loaded once,
                               // 4 input chunks loaded once — wrong results, used to find the limiting role
};

__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
  if (!on) { lr_mbar_wait(bar, parity); return; }
  const long long t = clock64();
  lr_mbar_wait(bar, parity);
  acc += clock64() - t;
}
// off the critical path (producer / epilogue): polls with a back-off
__device__ __forceinline__ void timed_wait_relaxed(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
  if (!on) { lr_mbar_wait_relaxed(bar, parity); return; }
  const long long t = clock64();
  lr_mbar_wait_relaxed(bar, parity);
  acc += clock64() - t;
}

// barrier block layout (uint64 each)
enum { BAR_A_FULL = 0, BAR_A_EMPTY = 2, BAR_ACC_FULL = 4, BAR_ACC_EMPTY = 6, BAR_W_FULL = 8,
       BAR_W_EMPTY = BAR_W_FULL + kWStages, BAR_COUNT = BAR_W_EMPTY + kWStages };

// Epilogue of the swapped orientation: the accumulator is D^T [channel lanes x 128 position columns].
// Thread = one output channel (M=64: rows live in lanes 0-15 of each 32-lane quarter, channel = 16q+lane;
// M=128: channel = 32q+lane).  64 columns (= 64/WP tile rows, always whole pooling row-pairs) are pulled per
// round; bias + ReLU + MaxPool(1,2,2) + arg-max are register-local; a warp stores one pixel's channels
// contiguously.
template <int WP>
__device__ __forceinline__ void epilogue_swapped(const ConvParams& p, uint32_t tcol, int q, int lane, int b, int t,
                                                 int y0) {
  const int ch = p.Mt == 64 ? 16 * q + lane : 32 * q + lane;
  const bool ch_ok = ch < p.Cout && (p.Mt == 128 || lane < 16);
  const float bias = (p.has_bias && ch_ok) ? p.bias[ch] : 0.f;
  const int PW = p.W >> 1, PH = p.H >> 1;
  constexpr int ROWS = 64 / WP;                      // tile rows per 64-column round
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    uint32_t v[64];
    tmem_ld32(tcol + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
    tmem_ld32(tcol + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
    if (!ch_ok) continue;
    if (p.epi_mode == 0) {
#pragma unroll
      for (int pr = 0; pr < ROWS / 2; ++pr) {
        const int gy = ((y0 + half * ROWS) >> 1) + pr;                    // pooled row
        if (gy >= PH) continue;
#pragma unroll
        for (int px = 0; px < WP / 2; ++px) {
          if (px >= PW) continue;
          const int c00 = (2 * pr) * WP + 2 * px;
          float m = fmaxf(__uint_as_float(v[c00]) + bias, 0.f);
          int a = 0;
          // the stored activation is bf16: compare the rounded values so ties resolve like the oracle
          m = __bfloat162float(__float2bfloat16(m));
          float c1 = __bfloat162float(__float2bfloat16(fmaxf(__uint_as_float(v[c00 + 1]) + bias, 0.f)));
          float c2 = __bfloat162float(__float2bfloat16(fmaxf(__uint_as_float(v[c00 + WP]) + bias, 0.f)));
          float c3 = __bfloat162float(__float2bfloat16(fmaxf(__uint_as_float(v[c00 + WP + 1]) + bias, 0.f)));
          if (c1 > m) { m = c1; a = 1; }
          if (c2 > m) { m = c2; a = 2; }
          if (c3 > m) { m = c3; a = 3; }
          const size_t opix = (((size_t)b * p.oTp + (t + p.o_t)) * p.oHp + (gy + p.o_y)) * p.oWp + (px + p.o_x);
          p.y[opix * p.Cout + ch] = __float2bfloat16(m);
          if (p.argmax) {
            const size_t apix = (((size_t)b * p.T + t) * PH + gy) * PW + px;
            p.argmax[apix * p.Cout + ch] = (uint8_t)(m > 0.f ? a : 4);
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int gy = y0 + half * ROWS + r;
        if (gy >= p.H) continue;
#pragma unroll
        for (int x = 0; x < WP; ++x) {
          if (x >= p.W) continue;
          const size_t opix = (((size_t)b * p.oTp + (t + p.o_t)) * p.oHp + (gy + p.o_y)) * p.oWp + (x + p.o_x);
          p.y[opix * p.Cout + ch] = __float2bfloat16(__uint_as_float(v[r * WP + x]) + bias);
        }
      }
    }
  }
}

__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// KS = Cin/16 K-steps per tap (1, 2, 4); MODE 0 = positions on M, 1 = swapped orientation (weights on M),
// 2 = positions on M with the KW kx-taps of a filter row stacked on N (plain-store epilogue only),
// 3 = positions on M with the KT kt-taps stacked on N: input plane c times [W(kt=KT-1); ...; W(kt=0)] lands on the
//     accumulators of the consecutive output frames c-KT+1 .. c, which sit side by side in TMEM — N = KT*Cout
//     per MMA with no epilogue change.  All MMAs accumulate; the epilogue re-zeroes each accumulator after
//     draining it (tcgen05.st), so no "first MMA" bookkeeping exists and one thread issues everything in order.
template <int KS, int MODE>
__device__ __forceinline__ void mma_tap(uint32_t d, uint64_t a, uint64_t w, uint32_t idesc, uint32_t acc) {
  constexpr bool SWAP = MODE == 1;
  const uint64_t ma = SWAP ? w : a, mb = SWAP ? a : w;
  umma_bf16(d, ma, mb, idesc, acc);
#pragma unroll
  for (int k = 1; k < KS; ++k) umma_bf16(d, ma + 2 * k, mb + 2 * k, idesc, 1u);
}

// (13 warps: the sub-partition that hosts four of them has 16 K registers -> 128 per thread is the hardware cap)
// UNPOOL: the fused un-pooling epilogue (epi_mode 2) is a compile-time variant so that its arg-max prefetch registers and
// bias partial sums do not weigh on the pooling epilogue of the forward layers
template <int KS, int MODE, bool UNPOOL>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ ConvParams p) {
  constexpr bool SWAP = MODE == 1;
  constexpr int kOpsUnroll = KS >= 4 ? 2 : 8 / KS;      // mode 3: chunk ops per unrolled group (~8 MMAs)
  extern __shared__ __align__(1024) uint8_t smem[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = base;
  uint8_t* w_smem = base + p.smem_off_w;
  uint8_t* stage = base + p.smem_off_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + p.smem_off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* bias_s = reinterpret_cast<float*>(base + p.smem_off_bar + 256);      // [128], zeros without a bias
  if (threadIdx.x < 128) bias_s[threadIdx.x] = (p.has_bias && (int)threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;

  // warp index through a shuffle: tells the compiler it is warp-uniform (role dispatch, descriptor math on
  // the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_taps = p.n_eff_taps;          // weight tiles per channel group

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      lr_mbar_init(&bars[BAR_A_FULL + i], 1);
      lr_mbar_init(&bars[BAR_A_EMPTY + i], p.n_issuers);
      lr_mbar_init(&bars[BAR_ACC_FULL + i], p.n_issuers);
      lr_mbar_init(&bars[BAR_ACC_EMPTY + i], 4 * p.epi_groups);
    }
    for (int s = 0; s < kWStages; ++s) {
      lr_mbar_init(&bars[BAR_W_FULL + s], 1);
      lr_mbar_init(&bars[BAR_W_EMPTY + s], p.n_issuers);
    }
    lr_fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (MODE == 3) {
    // every MMA accumulates: the accumulators start (and are left by the epilogue) at zero
    if (warp > kMmaWarps && warp <= kMmaWarps + 4) {
      for (int c = 0; c < p.tmem_cols; c += 32) tmem_zero32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const bool dbg_on = p.dbg != nullptr;
  long long dbg0 = 0, dbg1 = 0, dbg2 = 0;
  const long long t_start = clock64();

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    uint32_t wn = 0;   // global weight-stage counter
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int b = item / (p.n_tgroups * p.n_ytiles);
      const int rem = item - b * (p.n_tgroups * p.n_ytiles);
      const int tg = rem / p.n_ytiles, yt = rem - tg * p.n_ytiles;
      const int t0 = tg * p.J, jn = min(p.J, p.T - t0), y0 = yt * p.R;
      const int n_chunks = jn + p.KT - 1;
      const int aset = it & (p.a_sets - 1);
      uint8_t* a_set = a_smem + (size_t)aset * p.a_set_bytes;
      timed_wait_relaxed(&bars[BAR_A_EMPTY + aset], ((it / p.a_sets) & 1) ^ 1, dbg0, dbg_on);
      if ((p.skip & 4) && it >= p.a_sets) {
        if (elect_one()) lr_mbar_arrive(&bars[BAR_A_FULL + aset]);
      } else if (elect_one()) {
        lr_mbar_expect_tx(&bars[BAR_A_FULL + aset], (uint32_t)(n_chunks * p.CG * p.CH * p.row_bytes));
        for (int g = 0; g < p.CG; ++g)
          for (int c = 0; c < n_chunks; ++c) {
            long long row0 = (long long)g * p.rows_per_group +
                             (((long long)b * p.Tp + (t0 + c)) * p.Hp + y0) * p.Wp;
            tma_load_2d(a_set + (size_t)(g * (p.J + p.KT - 1) + c) * p.chunk_bytes, &map_x, 0, (int)row0,
                        &bars[BAR_A_FULL + aset]);
          }
      }
      __syncwarp();
      for (int g = 0; g < p.CG; ++g)
        for (int tap0 = 0; tap0 < n_taps; tap0 += p.tps, ++wn) {
          const int s = wn % p.w_stages;
          const int nt = min(p.tps, n_taps - tap0);
          timed_wait_relaxed(&bars[BAR_W_EMPTY + s], ((wn / p.w_stages) & 1) ^ 1, dbg1, dbg_on);
          if ((p.skip & 2) && wn >= (uint32_t)p.w_stages) {
            if (elect_one()) lr_mbar_arrive(&bars[BAR_W_FULL + s]);
          } else if (elect_one()) {
            // weights arrive as pre-swizzled tile images (lr_pack_conv_weights): one contiguous bulk copy
            // per stage instead of Cout narrow strided rows per tap
            const uint32_t bytes = (uint32_t)(nt * p.wtile_bytes);
            lr_mbar_expect_tx(&bars[BAR_W_FULL + s], bytes);
            lr_bulk_g2s(w_smem + (size_t)(s * p.tps) * p.wtile_bytes,
                        p.w_packed + (size_t)(g * n_taps + tap0) * p.wtile_bytes, bytes, &bars[BAR_W_FULL + s]);
          }
          __syncwarp();
        }
    }
  } else if (MODE == 3 && warp <= kMmaWarps && (warp <= p.n_issuers || p.epi_groups < 3)) {
    // ===================== mode 3: ONE issuing warp, kt-stacked MMAs in program order ===================
    // A single thread feeds the tensor pipe, so the scalar work per MMA is a handful of uniform adds: no
    // tables, no divisions, no constant-bank loads in the inner loops.  Per spatial tap the chunks c = 0 ..
    // jn+KT-2 are walked in three phases (hi = newest kt a chunk feeds, lo = oldest):
    //   ramp-up   c <  KT-1      : N grows by Cout, the weight window slides towards kt = 0, D stays at frame 0
    //   steady    KT-1 <= c < jn : N = KT*Cout, whole tile, D advances one frame per chunk
    //   ramp-down c >= jn        : N shrinks by Cout, D advances
    // (groups with fewer than KT-1 frames — a clip's tail — fall back to one MMA per (frame, kt).)
    // n_issuers = G warps split the J frames of an item into G runs of Jg consecutive frames; each warp stacks
    // inside its own run and owns its accumulators, so the result does not depend on how the warps interleave, and
    // one warp's stage-boundary bookkeeping (barrier wait, commit) is covered by the others' queued MMAs.
    if (warp <= p.n_issuers) {
      const int g_first = (warp - 1) * p.Jg;      // first frame (= first chunk) of this warp's run
      int it = 0;
      int ws = 0;                     // weight ring slot and its phase bit, advanced without divisions
      uint32_t wphase = 0;
      const uint64_t a_desc0 = make_desc(lr_smem_u32(a_smem), p.desc_hi);
      const uint64_t w_desc0 = make_desc(lr_smem_u32(w_smem), p.desc_hi);
      const uint32_t wtile16 = (uint32_t)p.wtile_bytes >> 4;
      const uint32_t group16 = (uint32_t)(p.J + p.KT - 1) * ((uint32_t)p.chunk_bytes >> 4);
      // (the skip bits 8..128 freeze one descriptor component each — diagnostics, wrong results)
      const uint32_t chunk16 = (p.skip & 16) ? 0u : (uint32_t)p.chunk_bytes >> 4;
      const uint32_t blk16 = (p.skip & 64) ? 0u : ((uint32_t)p.Cout * (uint32_t)p.row_bytes) >> 4;   // one kt block of a tile
      const uint32_t istep = (p.skip & 128) ? 0u : ((uint32_t)p.Cout >> 3) << 17;   // +Cout on the descriptor's N
      const uint32_t idesc1 = p.idesc;                                              // N = Cout
      const uint32_t cout = (p.skip & 32) ? 0u : (uint32_t)p.Cout;
      const uint32_t tapmask = (p.skip & 8) ? 0u : 0xffffffffu;
      const int KT = p.KT;
      const int tiles_per_group = p.n_tgroups * p.n_ytiles;
      int rem = blockIdx.x % tiles_per_group;                                        // item index inside its clip
      const int rem_step = gridDim.x % tiles_per_group;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int tg = rem / p.n_ytiles;
        rem += rem_step;
        if (rem >= tiles_per_group) rem -= tiles_per_group;
        const int jn_item = min(p.J, p.T - tg * p.J);
        // run mode: this warp's frames [g_first, g_first + jn) are a self-contained sub-item (chunks 0 .. jn+KT-2 of it);
        // seam mode: jn = all frames of the item, this warp walks chunks [c_lo, c_hi) of the item's single chunk list
        const int f0 = p.seam ? 0 : g_first;
        int jn = p.seam ? jn_item : min(p.Jg, jn_item - g_first);         // (<= 0: nothing to do)
        int c_lo = 0, c_hi = jn + KT - 1;
        if (p.seam) {
          if (jn >= KT - 1) {
            const int per = (c_hi + p.n_issuers - 1) / p.n_issuers;
            c_lo = min(c_hi, (warp - 1) * per);
            c_hi = min(c_hi, c_lo + per);
          } else if (warp != 1) {
            jn = 0;                                                       // tail fallback: the first warp does it all
          }
        }
        const int set = it & (p.n_sets - 1);
        const uint32_t d_base = tmem_base + (uint32_t)((set * p.J + f0) * p.acc_cols);
        timed_wait(&bars[BAR_ACC_EMPTY + set], ((it / p.n_sets) & 1) ^ 1, dbg0, dbg_on);
        const int aset = it & (p.a_sets - 1);
        timed_wait(&bars[BAR_A_FULL + aset], (it / p.a_sets) & 1, dbg1, dbg_on);
        tc_fence_after();
        const uint64_t a_item = a_desc0 + (uint64_t)(aset * (p.a_set_bytes >> 4)) +
                                (uint64_t)((uint32_t)f0 * ((uint32_t)p.chunk_bytes >> 4));
        const int n_steady = jn - (KT - 1);                 // chunks that feed all KT taps (< 0: tail fallback)
        // state of the chunk walk at its first chunk c_lo: window = frames [lo_f, hi_f]
        const int lo_f = max(0, c_lo - (KT - 1)), hi_f = min(jn - 1, c_lo);
        const uint32_t a_first = (uint32_t)c_lo * chunk16;
        const uint32_t b_first = (uint32_t)(KT - 1 - (c_lo - lo_f)) * blk16;
        const uint32_t id_first = idesc1 + (uint32_t)(hi_f - lo_f) * istep;
        const uint32_t d_first = d_base + (uint32_t)lo_f * cout;
        for (int g = 0; g < p.CG; ++g) {
          const uint32_t goff = (uint32_t)g * group16;
          for (int tap0 = 0; tap0 < n_taps; tap0 += p.tps) {
            const int nt = min(p.tps, n_taps - tap0);
            // (no tcgen05 fence: the weights were written by the async proxy, the mbarrier's complete_tx orders them)
            timed_wait(&bars[BAR_W_FULL + ws], wphase, dbg2, dbg_on);
            if (elect_one()) {
              uint64_t wd = w_desc0 + (uint64_t)((uint32_t)(ws * p.tps) * wtile16);
              uint32_t off_next = p.tap_off[tap0];
              for (int i = 0; i < nt; ++i, wd += wtile16) {
                uint64_t a = a_item + (uint64_t)((off_next & tapmask) + goff);
                if (i + 1 < nt) off_next = p.tap_off[tap0 + i + 1];      // fetched a tile ahead of its use
                if (n_steady >= 0) {
                  // one flat loop over the run's chunks; the phase only changes which increments apply.  It is
                  // unrolled so that ~8 MMAs are issued between two rewrites of the same descriptor registers:
                  // a tcgen05.mma holds its uniform-register operands until the tensor pipe accepts it (the next
                  // write to them waits on the short scoreboard), so short groups serialise issue behind the pipe
                  uint64_t b = wd + (uint64_t)b_first;
                  uint32_t d = d_first, id = id_first;
                  a += (uint64_t)a_first;
#pragma unroll kOpsUnroll
                  for (int c = c_lo; c < c_hi; ++c) {
#pragma unroll
                    for (int k = 0; k < KS; ++k) umma_bf16(d, a + 2 * k, b + 2 * k, id, 1u);
                    const bool up = c < KT - 1;               // ramp-up: the next chunk also feeds an older tap
                    a += chunk16;
                    b -= up ? blk16 : 0u;
                    id += up ? istep : 0u;
                    d += up ? 0u : cout;
                    if (c + 1 >= jn) id -= istep;             // ramp-down: the newest taps have no frame left
                  }
                } else {
                  for (int j = 0; j < jn; ++j)
                    for (int kt = 0; kt < KT; ++kt) {
                      const uint64_t aa = a + (uint64_t)((uint32_t)(j + kt) * chunk16);
                      const uint64_t bb = wd + (uint64_t)((uint32_t)(KT - 1 - kt) * blk16);
#pragma unroll
                      for (int k = 0; k < KS; ++k) umma_bf16(d_base + (uint32_t)j * cout, aa + 2 * k, bb + 2 * k, idesc1, 1u);
                    }
                }
              }
              umma_commit(&bars[BAR_W_EMPTY + ws]);
            }
            __syncwarp();
            if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
          }
        }
        if (elect_one()) {
          umma_commit(&bars[BAR_A_EMPTY + aset]);
          umma_commit(&bars[BAR_ACC_FULL + set]);
        }
        __syncwarp();
      }
    }
  } else if (MODE != 3 && warp <= kMmaWarps) {
    // ===================== MMA issuers: warp m issues accumulators j = m and m+4 ========================
    // (one thread cannot feed the tensor pipe for N < 128; the per-MMA scalar work is kept to one 64-bit
    // add: per-item descriptor bases + a per-tap offset table in constant memory)
    const int mw = warp - 1;
    uint32_t wn = 0;
    int it = 0;
    const uint64_t a_desc0 = make_desc(lr_smem_u32(a_smem), p.desc_hi);
    const uint64_t w_desc0 = make_desc(lr_smem_u32(w_smem), p.desc_hi);
    const uint32_t chunk16 = (uint32_t)p.chunk_bytes >> 4, wtile16 = (uint32_t)p.wtile_bytes >> 4;
    const uint32_t group16 = (uint32_t)(p.J + p.KT - 1) * chunk16;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int rem = item % (p.n_tgroups * p.n_ytiles);
      const int tg = rem / p.n_ytiles;
      const int t0 = tg * p.J, jn = min(p.J, p.T - t0);
      const int set = it & (p.n_sets - 1);
      const uint32_t d_base = tmem_base + (uint32_t)(set * p.J * p.acc_cols);
      timed_wait(&bars[BAR_ACC_EMPTY + set], ((it / p.n_sets) & 1) ^ 1, dbg0, dbg_on);
      const int aset = it & (p.a_sets - 1);
      timed_wait(&bars[BAR_A_FULL + aset], (it / p.a_sets) & 1, dbg1, dbg_on);
      tc_fence_after();
      // this warp's (at most two) accumulators: frames mw and mw + 4 of the item
      const uint64_t aj0 = a_desc0 + (uint64_t)(aset * (p.a_set_bytes >> 4)) + (uint64_t)(mw * chunk16);
      const uint64_t aj1 = aj0 + (uint64_t)(kMmaWarps * chunk16);
      const uint32_t d0 = d_base + (uint32_t)(mw * p.acc_cols), d1 = d0 + (uint32_t)(kMmaWarps * p.acc_cols);
      const bool has0 = mw < jn, has1 = mw + kMmaWarps < jn;
      uint32_t first = 0;      // accumulate flag: 0 for the very first tap of the item
      for (int g = 0; g < p.CG; ++g) {
        const uint32_t goff = (uint32_t)g * group16;
        for (int tap0 = 0; tap0 < n_taps; tap0 += p.tps, ++wn) {
          const int s = wn % p.w_stages;
          const int nt = min(p.tps, n_taps - tap0);
          timed_wait(&bars[BAR_W_FULL + s], (wn / p.w_stages) & 1, dbg2, dbg_on);
          if (p.stage_fence) tc_fence_after();
          if (elect_one()) {
            uint64_t wd = w_desc0 + (uint64_t)((uint32_t)(s * p.tps) * wtile16);
            for (int i = 0; i < nt; ++i, wd += wtile16) {
              const uint32_t off = p.tap_off[tap0 + i] + goff;
              if (has0) mma_tap<KS, MODE>(d0, aj0 + off, wd, p.idesc, first);
              if (has1) mma_tap<KS, MODE>(d1, aj1 + off, wd, p.idesc, first);
              first = 1u;
            }
            umma_commit(&bars[BAR_W_EMPTY + s]);    // weight stage free once these MMAs retire
          }
          __syncwarp();
        }
      }
      if (elect_one()) {
        umma_commit(&bars[BAR_A_EMPTY + aset]);
        umma_commit(&bars[BAR_ACC_FULL + set]);
      }
      __syncwarp();
    }
  } else if (warp <= kMmaWarps ? (MODE == 3 && p.epi_groups == 3 && warp > 2)
                               : ((warp - 1 - kMmaWarps) >> 2 < (p.epi_groups < 3 ? p.epi_groups : 2) ||
                                  (p.epi_groups == 3 && warp > kMmaWarps + 4 * kEpiGroups))) {
    // ===================== epilogue (quartets of warps; quartet e owns accumulators j = e mod groups) ==========
    // quartets 0, 1 = warps 5-8, 9-12; quartet 2 (orientation 3 with <= 2 issuing warps) = warps 3, 4, 13, 14
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const bool third = warp <= kMmaWarps || warp > kMmaWarps + 4 * kEpiGroups;
    const int eg = third ? 2 : (warp - 1 - kMmaWarps) >> 2;        // epilogue group (quartet)
    const int wq = third ? (warp <= kMmaWarps ? warp - 3 : warp - (kMmaWarps + 4 * kEpiGroups + 1) + 2)
                         : (warp - 1 - kMmaWarps) & 3;             // this warp's slot 0..3 inside its quartet
    const int etid = wq * 32 + lane;              // 0..127 inside the quartet
    const int bar_id = 1 + eg;                    // named barrier of this quartet
    stage += (size_t)eg * p.stage_bufs * (128 * p.stage_pitch);
    const int row = q * 32 + lane;                // accumulator row (tile position) held by this thread
    const int PW = p.W >> 1;
    const int cgroups = p.Cout >> 3;
    // pooling work items of this thread (idx = etid + 128k -> pooled row, x, channel group): fixed per layer
    int e_src[kEpiIters], e_out[kEpiIters], e_am[kEpiIters], e_py[kEpiIters], e_cg[kEpiIters];
    const int epi_n = ((p.R >> 1) * PW * cgroups + 127) / 128;
#pragma unroll
    for (int k = 0; k < (UNPOOL ? 0 : kEpiIters); ++k) {
      const int idx = etid + 128 * k;
      const int cg = idx % cgroups, pp = idx / cgroups;
      const int px = pp % (PW > 0 ? PW : 1), py = pp / (PW > 0 ? PW : 1);
      e_cg[k] = cg;
      e_py[k] = idx < (p.R >> 1) * PW * cgroups ? py : (1 << 30);      // beyond the tile: never valid
      e_src[k] = ((2 * py) * p.Wp + 2 * px) * p.stage_pitch + cg * 16;
      e_out[k] = py * p.oWp + px;
      e_am[k] = py * PW + px;
    }
    int ebuf = 0;
    int it = 0;
    float bsum[UNPOOL ? 8 : 1];            // epi_mode 2: this thread's 8 channels (its channel group is fixed: 128 % cgroups == 0)
#pragma unroll
    for (int e = 0; e < (UNPOOL ? 8 : 1); ++e) bsum[e] = 0.f;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int b = item / (p.n_tgroups * p.n_ytiles);
      const int rem = item - b * (p.n_tgroups * p.n_ytiles);
      const int tg = rem / p.n_ytiles, yt = rem - tg * p.n_ytiles;
      const int t0 = tg * p.J, jn = min(p.J, p.T - t0), y0 = yt * p.R;
      const int set = it & (p.n_sets - 1);
      const uint32_t d_base = tmem_base + (uint32_t)(set * p.J * p.acc_cols);
      timed_wait_relaxed(&bars[BAR_ACC_FULL + set], (it / p.n_sets) & 1, dbg0, dbg_on);
      tc_fence_after();
      if (p.skip & 1) {
        tc_fence_before();
        if (lane == 0) lr_mbar_arrive(&bars[BAR_ACC_EMPTY + set]);
        continue;
      }
      if (SWAP) {
        for (int j = 0; j < jn; ++j) {
          const uint32_t tcol = d_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 128);
          if (p.Wp == 8) epilogue_swapped<8>(p, tcol, q, lane, b, t0 + j, y0);
          else if (p.Wp == 16) epilogue_swapped<16>(p, tcol, q, lane, b, t0 + j, y0);
          else epilogue_swapped<32>(p, tcol, q, lane, b, t0 + j, y0);
        }
        tc_fence_before();
        if (lane == 0) lr_mbar_arrive(&bars[BAR_ACC_EMPTY + set]);
        continue;
      }
      const size_t obase = (((size_t)b * p.oTp + (t0 + p.o_t)) * p.oHp + ((p.epi_mode == 0 ? (y0 >> 1) : y0) + p.o_y)) *
                               p.oWp + p.o_x;                       // output pixel of (t0, tile row 0, x 0)
      const size_t oframe = (size_t)p.oHp * p.oWp;
      const int rows_left = (p.epi_mode == 0 ? (p.H >> 1) - (y0 >> 1) : p.H - y0);   // valid output rows from y0 on
      for (int j = eg; j < jn; j += p.epi_groups) {
        const int t = t0 + j;
        uint8_t* stg = stage + (size_t)((ebuf++) & (p.stage_bufs - 1)) * (128 * p.stage_pitch);
        // epi_mode 2: the arg-max bytes of this thread's (up to kUnpoolIters) items are fetched before the TMEM drain
        // and the staging barrier, so their global-load latency is off the critical path (0x04.. = nothing to route)
        uint2 amv[UNPOOL ? kUnpoolIters : 1];
        if (UNPOOL) {
          const size_t am_base = (((size_t)b * p.T + t) * p.H + y0) * p.W;
#pragma unroll
          for (int k = 0; k < kUnpoolIters; ++k) {
            const int idx = etid + 128 * k;
            const int cg = idx & (cgroups - 1), r = idx >> p.cg_shift;
            const int yl = r >> p.wp_shift, x = r & (p.Wp - 1);
            const bool ok = k < cgroups && x < p.W && yl < rows_left;
            amv[k] = ok ? __ldg(reinterpret_cast<const uint2*>(p.am_in + (am_base + (size_t)yl * p.W + x) * p.Cout + cg * 8))
                        : make_uint2(0x04040404u, 0x04040404u);
          }
        }
        if (MODE == 2) {
          // kx-stacked accumulator: lane r holds P[r][kx][n] (partial sums of filter column kx evaluated at
          // window row r); out[r][n] = sum_kx P[r + kx][kx][n].  Rows r+kx of the same warp come by shuffle,
          // the first kx rows of the next warp through a small shared-memory halo.  (Rows past the tile are
          // only ever needed by junk positions x >= W.)  Cout == 32 here.
          float* halo = reinterpret_cast<float*>(base + p.smem_off_halo);
          const uint32_t tj = d_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.acc_cols);
          uint32_t v[32];
          float acc[32];
          tmem_ld32(tj, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v[i]);
          for (int kx = 1; kx < p.kxs; ++kx) {
            tmem_ld32(tj + (uint32_t)(kx * 32), v);
            if (lane < kx) {
              float4* h4 = reinterpret_cast<float4*>(halo + ((q * 4 + (kx - 1)) * 4 + lane) * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                h4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                    __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            const bool in_warp = lane + kx < 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = __shfl_down_sync(0xffffffffu, __uint_as_float(v[i]), kx);
              if (in_warp) acc[i] += t;
            }
          }
          named_bar_sync(bar_id, 128);
          if (q < 3) {
            for (int kx = 1; kx < p.kxs; ++kx) {
              if (lane + kx >= 32) {
                const float4* h4 =
                    reinterpret_cast<const float4*>(halo + (((q + 1) * 4 + (kx - 1)) * 4 + (lane + kx - 32)) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 h = h4[i];
                  acc[4 * i] += h.x; acc[4 * i + 1] += h.y; acc[4 * i + 2] += h.z; acc[4 * i + 3] += h.w;
                }
              }
            }
          }
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + 2 * i);
            __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i] + bb.x, acc[2 * i + 1] + bb.y);
            packed[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)row * p.stage_pitch);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        } else
        // TMEM -> registers -> (bias, ReLU) -> bf16 staging tile [128][Cout]
        for (int cc = 0; cc < p.Cout; cc += 32) {
          uint32_t v[32];
          tmem_ld32(d_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.acc_cols + cc), v);
          if (MODE == 3) tmem_zero32(d_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.acc_cols + cc));
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + cc + 2 * i);     // smem broadcast
            float f0 = __uint_as_float(v[2 * i]) + bb.x, f1 = __uint_as_float(v[2 * i + 1]) + bb.y;
            if (!UNPOOL && p.epi_mode == 0) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            packed[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)row * p.stage_pitch + cc * 2);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
        named_bar_sync(bar_id, 128);
        const size_t ob = obase + (size_t)j * oframe;
        if (!UNPOOL && p.epi_mode == 0) {
          // MaxPool(1,2,2) over the staged tile; 8 channels (16 B) per thread-item; the (row, x, channel
          // group) decomposition of each item was hoisted out of the item loop (no divisions here)
          const size_t ab = (((size_t)b * p.T + t) * (p.H >> 1) + (y0 >> 1)) * PW;
#pragma unroll
          for (int k = 0; k < kEpiIters; ++k) {
            if (k >= epi_n || e_py[k] >= rows_left) continue;
            const uint8_t* src = stg + e_src[k];
            uint4 q4[4];
            q4[0] = *reinterpret_cast<const uint4*>(src);
            q4[1] = *reinterpret_cast<const uint4*>(src + p.stage_pitch);
            q4[2] = *reinterpret_cast<const uint4*>(src + (size_t)p.Wp * p.stage_pitch);
            q4[3] = *reinterpret_cast<const uint4*>(src + (size_t)(p.Wp + 1) * p.stage_pitch);
            const __nv_bfloat16* e0 = reinterpret_cast<const __nv_bfloat16*>(&q4[0]);
            const __nv_bfloat16* e1 = reinterpret_cast<const __nv_bfloat16*>(&q4[1]);
            const __nv_bfloat16* e2 = reinterpret_cast<const __nv_bfloat16*>(&q4[2]);
            const __nv_bfloat16* e3 = reinterpret_cast<const __nv_bfloat16*>(&q4[3]);
            __align__(16) __nv_bfloat16 outv[8];
            __align__(8) uint8_t am[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float m = __bfloat162float(e0[e]);
              int a = 0;
              float c1 = __bfloat162float(e1[e]), c2 = __bfloat162float(e2[e]), c3 = __bfloat162float(e3[e]);
              if (c1 > m) { m = c1; a = 1; }
              if (c2 > m) { m = c2; a = 2; }
              if (c3 > m) { m = c3; a = 3; }
              outv[e] = __float2bfloat16(m);
              am[e] = (uint8_t)(m > 0.f ? a : 4);
            }
            *reinterpret_cast<uint4*>(p.y + (ob + e_out[k]) * p.Cout + e_cg[k] * 8) =
                *reinterpret_cast<const uint4*>(outv);
            if (p.argmax)
              *reinterpret_cast<uint2*>(p.argmax + (ab + e_am[k]) * p.Cout + e_cg[k] * 8) =
                  *reinterpret_cast<const uint2*>(am);
          }
        } else if (UNPOOL) {
          // fused un-pooling (backward of ReLU + MaxPool(1,2,2)): the gradient of pooled pixel (y,x) goes to the
          // arg-max slot of its 2x2 window in the padded dY volume of the layer below, zeros to the other three
          // (slot 4 = ReLU-dead: all zeros); the conv bias gradient is the per-channel sum of what was routed
          const size_t oplane = ((size_t)b * p.oTp + (t + p.o_t)) * p.oHp;
#pragma unroll
          for (int k = 0; k < kUnpoolIters; ++k) {
            const int idx = etid + 128 * k;
            const int cg = idx & (cgroups - 1), r = idx >> p.cg_shift;
            const int yl = r >> p.wp_shift, x = r & (p.Wp - 1);
            if (k >= cgroups || x >= p.W || yl >= rows_left) continue;
            const uint2 a8 = amv[k];
            const uint4 v8 = *reinterpret_cast<const uint4*>(stg + (size_t)r * p.stage_pitch + cg * 16);
            const uint32_t vw[4] = {v8.x, v8.y, v8.z, v8.w};
            const uint32_t aw[2] = {a8.x, a8.y};
            uint32_t o[4][4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {              // two channels per 32-bit word
              const uint32_t a_lo = (aw[h >> 1] >> (16 * (h & 1))) & 0xffu, a_hi = (aw[h >> 1] >> (16 * (h & 1) + 8)) & 0xffu;
              const uint32_t lo = vw[h] & 0xffffu, hi = vw[h] & 0xffff0000u;
#pragma unroll
              for (int w = 0; w < 4; ++w) o[w][h] = (a_lo == (uint32_t)w ? lo : 0u) | (a_hi == (uint32_t)w ? hi : 0u);
              bsum[2 * h] += a_lo < 4u ? __uint_as_float(lo << 16) : 0.f;
              bsum[2 * h + 1] += a_hi < 4u ? __uint_as_float(hi) : 0.f;
            }
            __nv_bfloat16* orow = p.y + ((oplane + (size_t)(2 * (y0 + yl) + p.o_y)) * p.oWp + (2 * x + p.o_x)) * p.Cout + cg * 8;
            *reinterpret_cast<uint4*>(orow) = make_uint4(o[0][0], o[0][1], o[0][2], o[0][3]);
            *reinterpret_cast<uint4*>(orow + p.Cout) = make_uint4(o[1][0], o[1][1], o[1][2], o[1][3]);
            *reinterpret_cast<uint4*>(orow + (size_t)p.oWp * p.Cout) = make_uint4(o[2][0], o[2][1], o[2][2], o[2][3]);
            *reinterpret_cast<uint4*>(orow + (size_t)(p.oWp + 1) * p.Cout) = make_uint4(o[3][0], o[3][1], o[3][2], o[3][3]);
          }
        } else {
          const int n_out = 128 * cgroups;
          for (int idx = etid; idx < n_out; idx += 128) {
            int cg, r;
            if (p.cg_shift >= 0) { cg = idx & (cgroups - 1); r = idx >> p.cg_shift; }
            else { cg = idx % cgroups; r = idx / cgroups; }
            const int yl = r >> p.wp_shift, x = r & (p.Wp - 1);
            if (x >= p.W || yl >= rows_left) continue;
            *reinterpret_cast<uint4*>(p.y + (ob + (size_t)yl * p.oWp + x) * p.Cout + cg * 8) =
                *reinterpret_cast<const uint4*>(stg + (size_t)r * p.stage_pitch + cg * 16);
          }
        }
        if (p.stage_bufs == 1) named_bar_sync(bar_id, 128);     // single staging tile: reused by the next accumulator
      }
      if (MODE == 3) tmem_wait_st();
      tc_fence_before();
      if (lane == 0) lr_mbar_arrive(&bars[BAR_ACC_EMPTY + set]);
    }
    if (UNPOOL && p.d_bias) {
      // lanes with the same (lane & (cgroups-1)) hold the same 8 channels: butterfly over the other lane bits
      for (int o = cgroups; o < 32; o <<= 1)
#pragma unroll
        for (int e = 0; e < 8; ++e) bsum[e] += __shfl_xor_sync(0xffffffffu, bsum[e], o);
      if (lane < cgroups)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(p.d_bias + lane * 8 + e, bsum[e]);
    }
  }

  if (dbg_on && lane == 0 && (warp <= 1 || warp == 1 + kMmaWarps)) {   // (first quartet's wait is reported)
    long long* o = p.dbg + (size_t)blockIdx.x * 8;
    if (warp == 0) { o[0] = dbg0; o[1] = dbg1; }                       // producer: a_empty, w_empty
    if (warp == 1) { o[2] = dbg0; o[3] = dbg1; o[4] = dbg2; o[7] = clock64() - t_start; }   // mma: acc_empty, a_full, w_full
    if (warp == 1 + kMmaWarps) { o[5] = dbg0; }                        // epilogue: acc_full
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// ---- weight packing: [Cout][CG][taps][Cin] bf16 -> per-(group,tap) tile images [Cout][Cin] with the
// 16-byte chunks XOR-swizzled exactly as a TMA box load with the matching swizzle mode would leave them
// in shared memory (address bits [4:6] ^= bits [7:9] for 128 B rows, [4:5] ^= [7:8] for 64 B, [4] ^= [7]
// for 32 B).  The conv kernel can then fetch a whole stage of taps with one contiguous bulk copy.
// KT > 0 selects the kt-stacked order of conv mode 3: source tap (kt, s) -> tile s*KT + (KT-1-kt).
__global__ void pack_conv_weights_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                         int Cout, int CG, int taps, int Cin, int KT) {
  const int chunks = Cin / 8;
  const long long total = (long long)CG * taps * Cout * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long r = i / chunks;
    const int row = (int)(r % Cout); r /= Cout;
    const int tap = (int)(r % taps);
    const int g = (int)(r / taps);
    const int row_bytes = Cin * 2;
    const int sw = row_bytes == 128 ? (row & 7) : (row_bytes == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
    const uint4 v = *reinterpret_cast<const uint4*>(w + (((size_t)row * CG + g) * taps + tap) * Cin + ch * 8);
    int dtap = tap;
    if (KT > 0) { const int khw = taps / KT, kt = tap / khw, sp = tap - kt * khw; dtap = sp * KT + (KT - 1 - kt); }
    *reinterpret_cast<uint4*>(out + (((size_t)g * taps + dtap) * Cout + row) * Cin + (ch ^ sw) * 8) = v;
  }
}

// ---- clip preparation: u8 NDHWC -> /255 -> 2x2 space-to-depth -> zero-padded bf16 volume -------
// out (B, T+2, Hp, Wp, 16): channel = (dy*2+dx)*3 + c for c<3, 12..15 = 0; interior at (1,1,1).
__global__ void __launch_bounds__(256)
clip_s2d_kernel(const uint8_t* __restrict__ clip, __nv_bfloat16* __restrict__ out, int B, int T, int H,
                int W, int Hp, int Wp) {
  // u8 -> bf16(x/255) through an exact 256-entry table (no per-element division)
  __shared__ __nv_bfloat16 lut[256];
  if (threadIdx.x < 256) lut[threadIdx.x] = __float2bfloat16((float)threadIdx.x / 255.0f);
  __syncthreads();
  const int H2 = H >> 1, W2 = W >> 1;
  const int bt = blockIdx.y;
  const int b = bt / T, t = bt - b * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H2 * W2; i += gridDim.x * blockDim.x) {
    const int Y = i / W2, X = i - Y * W2;
    const uint8_t* src = clip + ((((size_t)b * T + t) * H + 2 * Y) * W + 2 * X) * 3;
    __align__(16) __nv_bfloat16 v[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      // the 2 pixels x 3 channels of this row are 6 contiguous bytes at an even address
      const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src + (size_t)dy * W * 3);
      const uint16_t p0 = s16[0], p1 = s16[1], p2 = s16[2];
      v[dy * 6 + 0] = lut[p0 & 0xff]; v[dy * 6 + 1] = lut[p0 >> 8];
      v[dy * 6 + 2] = lut[p1 & 0xff]; v[dy * 6 + 3] = lut[p1 >> 8];
      v[dy * 6 + 4] = lut[p2 & 0xff]; v[dy * 6 + 5] = lut[p2 >> 8];
    }
#pragma unroll
    for (int c = 12; c < 16; ++c) v[c] = __float2bfloat16(0.f);
    size_t opix = (((size_t)b * (T + 2) + (t + 1)) * Hp + (Y + 1)) * Wp + (X + 1);
    uint4* dst = reinterpret_cast<uint4*>(out + opix * 16);
    dst[0] = reinterpret_cast<const uint4*>(v)[0];
    dst[1] = reinterpret_cast<const uint4*>(v)[1];
  }
}

// ---- backward helper: route pooled gradients through the stored argmax (ReLU'd max-pool) --------
// d_pooled (B,T,H/2,W/2,C) bf16 + argmax u8 -> d_conv_out written into the interior of the
// zero-padded, channel-grouped volume the dgrad/wgrad passes read:
//   out[g][b][t+pt][y+ph][x+pw][c % Cg], g = c / Cg.
// One thread owns one POOLED pixel x 8 channels: it reads the gradient (16 B) and the arg-max bytes (8 B)
// once and writes the four positions of the 2x2 window (the arg-max one gets the gradient, the others 0).
// Interior rows/columns beyond the last whole window (odd H or W) are never written by anyone and keep the
// zeros they were allocated with, like the borders.  blockIdx.y strides over frames; a thread's
// (pooled row, x, channel group) is fixed, so there is no index arithmetic in the loop, and kUnpoolBatch
// frames are loaded before any store is issued (memory-level parallelism).
constexpr int kUnpoolBatch = 4;
__global__ void __launch_bounds__(256)
unpool_kernel(const __nv_bfloat16* __restrict__ d_pooled, const uint8_t* __restrict__ argmax,
              __nv_bfloat16* __restrict__ out, float* __restrict__ d_bias, int B, int T, int H, int W, int C,
              int Cg, int Tp, int Hp, int Wp, int pt, int ph, int pw) {
  __shared__ float bias_acc[128];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) bias_acc[i] = 0.f;
  __syncthreads();
  const int PH = H >> 1, PW = W >> 1;
  const int c8 = C >> 3;
  const int per_frame = PH * PW * c8;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = item < per_frame;
  const int cg = item % c8, pp = item / c8;
  const int px = pp % PW, py = pp / PW;
  const int c0 = cg * 8, g = c0 / Cg, cl = c0 - g * Cg;
  const long long rows_per_group = (long long)B * Tp * Hp * Wp;
  const size_t frame_in = (size_t)per_frame * 8;                       // elements per frame of d_pooled / argmax
  const size_t in_off = (size_t)item * 8;
  // output element offset of window position (0,0) in frame (b=0,t=0)
  const size_t out_off = ((size_t)g * rows_per_group + ((size_t)pt * Hp + (2 * py + ph)) * Wp + (2 * px + pw)) * Cg + cl;
  const size_t out_frame = (size_t)Hp * Wp * Cg, out_clip = (size_t)Tp * out_frame;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int n_frames = B * T;
  if (active) {
    for (int bt0 = blockIdx.y * kUnpoolBatch; bt0 < n_frames; bt0 += gridDim.y * kUnpoolBatch) {
      uint4 gr[kUnpoolBatch];
      uint2 am[kUnpoolBatch];
#pragma unroll
      for (int k = 0; k < kUnpoolBatch; ++k) {
        const int bt = bt0 + k;
        if (bt < n_frames) {
          gr[k] = *reinterpret_cast<const uint4*>(d_pooled + (size_t)bt * frame_in + in_off);
          am[k] = *reinterpret_cast<const uint2*>(argmax + (size_t)bt * frame_in + in_off);
        }
      }
#pragma unroll
      for (int k = 0; k < kUnpoolBatch; ++k) {
        const int bt = bt0 + k;
        if (bt >= n_frames) break;
        const int b = bt / T, t = bt - b * T;
        const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gr[k]);
        const uint8_t* ae = reinterpret_cast<const uint8_t*>(&am[k]);
        __nv_bfloat16* o = out + (size_t)b * out_clip + (size_t)t * out_frame + out_off;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          __align__(16) __nv_bfloat16 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ae[e] == w ? ge[e] : __float2bfloat16(0.f);
          *reinterpret_cast<uint4*>(o + ((size_t)(w >> 1) * Wp + (w & 1)) * Cg) = *reinterpret_cast<const uint4*>(v);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (ae[e] < 4) bsum[e] += __bfloat162float(ge[e]);
      }
    }
  }
  if (d_bias) {
    if (active) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&bias_acc[cg * 8 + e], bsum[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&d_bias[i], bias_acc[i]);
  }
}

}  // namespace

#ifdef LR_DIAG
// liblr_b200_diag.so only (include/lr_b200_diag.h): measurement hooks with process-global state
static long long* g_conv_dbg = nullptr;
static int g_conv_skip = 0;
extern "C" void lr_conv3d_set_debug_skip(int mask) { g_conv_skip = mask; }
// diagnostics: device buffer of 148*8 int64 that the next conv launches fill with per-role wait cycles
// [producer a_empty, producer w_empty, mma acc_empty, mma a_full, mma w_full, epilogue acc_full, -, mma total]
extern "C" void lr_conv3d_set_debug(long long* device_buffer) { g_conv_dbg = device_buffer; }
#else
static long long* const g_conv_dbg = nullptr;
static const int g_conv_skip = 0;
#endif

extern "C" int lr_conv3d_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}

extern "C" int lr_clip_s2d(const uint8_t* clip, void* out_bf16, int B, int T, int H, int W, int Hp,
                           int Wp, void* stream) {
  LR_CHECK_ARG(clip && out_bf16 && B > 0 && T > 0 && H > 0 && W > 0 && (H % 2) == 0 && (W % 2) == 0,
               "lr_clip_s2d: H and W must be even");
  LR_CHECK_ARG(Wp >= W / 2 + 2 && Hp >= H / 2 + 2, "lr_clip_s2d: padded extents too small");
  LR_CHECK_ARG(B * T <= 65535, "lr_clip_s2d: B*T must be <= 65535");
  dim3 grid(lr_div_up((H / 2) * (W / 2), 256), B * T);
  clip_s2d_kernel<<<grid, 256, 0, lr_stream(stream)>>>(clip, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, T,
                                                       H, W, Hp, Wp);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_pack_conv_weights(const void* w, void* out, int Cout, int CG, int taps, int Cin, void* stream) {
  LR_CHECK_ARG(w && out && (Cin == 16 || Cin == 32 || Cin == 64) && Cout > 0 && CG > 0 && taps > 0,
               "lr_pack_conv_weights: bad args");
  const long long total = (long long)CG * taps * Cout * (Cin / 8);
  pack_conv_weights_kernel<<<lr_div_up(total, 256), 256, 0, lr_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(out), Cout, CG, taps, Cin, 0);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_pack_conv_weights_kt(const void* w, void* out, int Cout, int CG, int KT, int KHW, int Cin,
                                       void* stream) {
  LR_CHECK_ARG(w && out && (Cin == 16 || Cin == 32 || Cin == 64) && Cout > 0 && CG > 0 && KT > 0 && KHW > 0,
               "lr_pack_conv_weights_kt: bad args");
  const long long total = (long long)CG * KT * KHW * Cout * (Cin / 8);
  pack_conv_weights_kernel<<<lr_div_up(total, 256), 256, 0, lr_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(out), Cout, CG, KT * KHW, Cin, KT);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_unpool(const void* d_pooled, const uint8_t* argmax, void* out, float* d_bias, int B, int T,
                         int H, int W, int C, int Cg, int Tp, int Hp, int Wp, int pt, int ph, int pw,
                         void* stream) {
  LR_CHECK_ARG(d_pooled && argmax && out && C % 8 == 0 && Cg % 8 == 0 && C % Cg == 0 && C <= 128,
               "lr_unpool: bad args");
  if (d_bias) LR_CHECK_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * C, lr_stream(stream)));
  const int per_frame = (H / 2) * (W / 2) * (C / 8);
  if (per_frame == 0) return LR_OK;
  const int gx = lr_div_up(per_frame, 256);
  int gy = lr_div_up(kNumSMs * 8, gx);                    // ~8 resident blocks per SM
  const int max_gy = lr_div_up(B * T, kUnpoolBatch);
  if (gy > max_gy) gy = max_gy;
  if (gy > 65535) gy = 65535;
  dim3 grid(gx, gy);
  unpool_kernel<<<grid, 256, 0, lr_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(d_pooled), argmax,
                                                     reinterpret_cast<__nv_bfloat16*>(out), d_bias, B, T, H, W,
                                                     C, Cg, Tp, Hp, Wp, pt, ph, pw);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

// x: zero-padded, channel-grouped bf16 volume [CG][B][Tp][Hp][Wp][Cin] with Tp=T+KT-1, Hp >= H+KH-1
// (a multiple of 128/Wp keeps tiles inside their plane), Wp = power of two >= W+KW-1.
// w: tile images from lr_pack_conv_weights: [CG][KT*KH*KW][Cout rows x Cin] bf16, 16-byte chunks swizzled.
static int conv3d_launch(const void* x, const void* w, const float* bias, void* y, uint8_t* argmax,
                         const uint8_t* am_in, float* d_bias,
                         int B, int T, int H, int W, int Hp, int Wp, int Cin, int CG, int Cout, int KT,
                         int KH, int KW, int epi_mode, int oTp, int oHp, int oWp, int o_t, int o_y, int o_x,
                         int J, int swap, void* stream) {
  LR_CHECK_ARG(x && w && y, "lr_conv3d_fwd: null pointer");
  LR_CHECK_ARG(Cin == 16 || Cin == 32 || Cin == 64, "lr_conv3d_fwd: Cin per group must be 16/32/64 (got %d)", Cin);
  LR_CHECK_ARG(Cout % 32 == 0 && Cout >= 32 && Cout <= 128, "lr_conv3d_fwd: Cout must be 32..128, %%32 (got %d)", Cout);
  LR_CHECK_ARG(Wp == 8 || Wp == 16 || Wp == 32 || Wp == 64 || Wp == 128, "lr_conv3d_fwd: Wp must be 8..128 pow2");
  LR_CHECK_ARG(Wp >= W + KW - 1 && Hp >= H + KH - 1, "lr_conv3d_fwd: padded extents smaller than H+KH-1 / W+KW-1");
  LR_CHECK_ARG(KT >= 1 && KH >= 1 && KW >= 1 && CG >= 1 && B > 0 && T > 0 && H > 0 && W > 0, "lr_conv3d_fwd: bad shape");
  LR_CHECK_ARG(epi_mode >= 0 && epi_mode <= 2, "lr_conv3d_fwd: bad epilogue mode");
  LR_CHECK_ARG(epi_mode != 2 || (am_in && (Cout == 32 || Cout == 64) && swap != 1),
               "lr_conv3d_dgrad_unpool: needs the arg-max bytes, Cout in {32,64}, positions on M");
  LR_CHECK_ARG(KT * KH * KW <= kMaxTaps, "lr_conv3d_fwd: more than %d taps", kMaxTaps);
  if (!lr_conv3d_supported()) { lr_set_error("lr_conv3d_fwd needs an sm_100 device"); return LR_EARCH; }

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.Tp = T + KT - 1; p.Hp = Hp; p.Wp = Wp;
  p.KT = KT; p.KH = KH; p.KW = KW; p.Cin = Cin; p.CG = CG; p.Cout = Cout;
  p.R = 128 / Wp;
  LR_CHECK_ARG(epi_mode != 0 || (p.R % 2 == 0), "lr_conv3d_fwd: pooling needs an even number of tile rows");
  p.CH = 128 + (KH - 1) * Wp + (KW - 1);
  LR_CHECK_ARG(p.CH <= 256, "lr_conv3d_fwd: halo too large for one TMA box (CH=%d)", p.CH);
  p.row_bytes = Cin * 2;
  p.chunk_bytes = (p.CH * p.row_bytes + 1023) / 1024 * 1024;
  const int single_writer = (swap >> 8) & 1;   // LR_CONV_SINGLE_WRITER: every accumulator written by one thread
  swap &= 0xff;
  const int mode = swap;                       // 0 positions on M, 1 swapped, 2 kx-stacked
  LR_CHECK_ARG(mode >= 0 && mode <= 3, "lr_conv3d_fwd: orientation must be 0, 1, 2 or 3");
  swap = mode == 1;
  p.kxs = mode == 2 ? KW : 1;
  if (mode == 2) {
    LR_CHECK_ARG(epi_mode >= 1 && Cout == 32 && KW >= 2 && KW <= 5 && KW * Cout <= 256 && (KW * Cout) % 16 == 0,
                 "lr_conv3d_fwd: kx-stacking needs the plain-store epilogue, Cout = 32 and 2 <= KW <= 5");
  }
  p.n_eff_taps = KT * KH * KW / p.kxs;
  p.wtile_bytes = p.kxs * Cout * p.row_bytes;   // one weight tile = the kxs consecutive per-tap images of a filter row
  if (mode == 3) {                              // one weight tile = the KT kt-taps of a spatial tap (kt descending)
    p.n_eff_taps = KH * KW;
    p.wtile_bytes = KT * Cout * p.row_bytes;
  }
  p.n_issuers = kMmaWarps;     // mode 3: set with J below
  LR_CHECK_ARG(p.wtile_bytes % 1024 == 0, "lr_conv3d_fwd: Cout*Cin*2 must be a multiple of 1024");
  p.stage_pitch = Cout * 2 + 16;
  p.swap = swap ? 1 : 0;
  p.Mt = Cout <= 64 ? 64 : 128;
  p.acc_cols = swap ? 128 : p.kxs * Cout;
  // swap mode pools in registers (no staging tile) but reads Mt weight rows per MMA: keep that many
  // bytes of slack behind the weight ring
  const int stage_bytes = swap ? p.Mt * p.row_bytes : 128 * p.stage_pitch;
  const int halo_bytes = mode == 2 ? 4 * 4 * 4 * 32 * (int)sizeof(float) : 0;
  const int smem_cap = 227 * 1024 - 1024 - 512 - halo_bytes;     // minus alignment slack, bias copy, halo
  // a second epilogue quartet when the MMA work per frame is short enough for the epilogue to be the limiter
  // (bias/ReLU/pool of one 128 x Cout accumulator costs ~1.4 k cycles on one quartet, measured on conv1)
  {
    const long long mma_cycles = (long long)KT * KH * KW * (Cin / 16) * CG * (32 + Cout / 4) * 2 / (mode == 3 ? 3 : 2);
    // (the un-pooling epilogue stores four rows per position and reads the arg-max map: ~2.5x the plain one)
    p.epi_groups = (mode == 0 || mode == 3) && mma_cycles < (epi_mode == 2 ? 9000 : 2500) ? 2 : 1;   // (conv2 dgrad: 2.29 -> 2.24 ms)
    // a third quartet (the two issuing warps orientation 3 does not use + two extra warps) for the layers whose
    // epilogue is still the limiter with two: conv1 (9 N = 96 MMAs per frame against a full pooling epilogue)
    if (mode == 3 && p.epi_groups == 2 && mma_cycles < (epi_mode == 2 ? 0 : 1500)) p.epi_groups = 3;
    if (getenv("LR_CONV_EPI_GROUPS")) { const int v = atoi(getenv("LR_CONV_EPI_GROUPS")); if (v == 1 || ((v == 2 || (v == 3 && mode == 3)) && mode != 1 && mode != 2)) p.epi_groups = v; }
  }
  int fixed = 2 * p.wtile_bytes + p.epi_groups * stage_bytes + 256;      // at least a 2-deep ring of single taps
  // accumulators per item: bounded by TMEM (512 columns), chunk slots and shared memory
  int n_sets = (512 / p.acc_cols) >= 4 ? 2 : 1;      // double-buffer TMEM when >= 2 accumulators per set fit
  if (getenv("LR_CONV_SETS")) { int v = atoi(getenv("LR_CONV_SETS")); if (v == 1 || (v == 2 && 512 / p.acc_cols >= 2)) n_sets = v; }   // tuning hook
  if (mode == 2) n_sets = 1;                         // wide accumulators: weight reuse (J) beats overlap
  int Jmax = 512 / p.acc_cols / n_sets;
  if (J <= 0 || J > Jmax) J = Jmax;
  if (J > 2 * kMmaWarps) J = 2 * kMmaWarps;      // each issuing warp owns at most two accumulators
  int G = 1;                    // mode 3: issuing warps, each stacking inside its own run of Jg frames
  if (mode == 3) {
    G = 2;
    if (getenv("LR_CONV_ISSUERS")) { const int v = atoi(getenv("LR_CONV_ISSUERS")); if (v >= 1 && v <= kMmaWarps) G = v; }   // tuning hook
  }
  while (J > 1 && ((J + KT - 1) * CG > kMaxChunks || (J + KT - 1) * CG * p.chunk_bytes + fixed > smem_cap)) --J;
  LR_CHECK_ARG((J + KT - 1) * CG * p.chunk_bytes + fixed <= smem_cap && (J + KT - 1) * CG <= kMaxChunks,
               "lr_conv3d_fwd: tile does not fit shared memory");
  if (J > T) J = T;
  p.J = J;
  if (mode == 3) {
    if (G > J) G = J;
    p.Jg = (J + G - 1) / G;
    // a stacked MMA spans min(KT, Jg) weight blocks: N = blocks * Cout <= 256
    while ((KT < p.Jg ? KT : p.Jg) * Cout > 256 && G < kMmaWarps && G < J) { ++G; p.Jg = (J + G - 1) / G; }
    LR_CHECK_ARG((KT < p.Jg ? KT : p.Jg) * Cout <= 256, "lr_conv3d_fwd: kt-stacked N exceeds 256");
    G = (J + p.Jg - 1) / p.Jg;
    p.n_issuers = G;
    if (p.epi_groups == 3 && G > 2) p.epi_groups = 2;      // the third quartet borrows issuing warps 3 and 4
    // shared chunk list: a stacked MMA spans min(KT, J) weight blocks
    // (not for 32-byte rows: conv1's one-k-step MMAs are bound by shared-memory bandwidth, not by the pipe, and the
    // unequal chunk shares made it 5 % slower — measured)
    const bool seam_ok = G > 1 && (KT < J ? KT : J) * Cout <= 256;
    p.seam = !single_writer && seam_ok && Cin >= 32;
    if (getenv("LR_CONV_SEAM")) p.seam = atoi(getenv("LR_CONV_SEAM")) != 0 && seam_ok && (Cin >= 32 || atoi(getenv("LR_CONV_SEAM")) > 1);
  }
  if (mode != 3 && p.epi_groups == 3) p.epi_groups = 2;
  p.n_sets = n_sets;
  p.n_ytiles = lr_div_up(H, p.R);
  p.n_tgroups = lr_div_up(T, J);
  p.n_items = B * p.n_ytiles * p.n_tgroups;
  int cols = 32;
  while (cols < n_sets * J * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.epi_mode = epi_mode;
  p.am_in = am_in;
  p.d_bias = d_bias;
  if (epi_mode == 2 && d_bias) LR_CHECK_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * Cout, lr_stream(stream)));
  p.has_bias = bias != nullptr;
  p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.argmax = argmax;
  p.oTp = oTp; p.oHp = oHp; p.oWp = oWp; p.o_t = o_t; p.o_y = o_y; p.o_x = o_x;
  p.rows_per_group = (long long)B * p.Tp * p.Hp * p.Wp;
  p.idesc = swap ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(p.Mt >> 4) << 24))
                 : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.acc_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
  const uint32_t layout = p.row_bytes == 32 ? 6u : (p.row_bytes == 64 ? 4u : 2u);
  const uint32_t sbo = (uint32_t)(8 * p.row_bytes) >> 4;
  p.desc_hi = sbo | (1u << 14) | (layout << 29);
  p.a_set_bytes = (J + KT - 1) * CG * p.chunk_bytes;
  // second chunk set when it still leaves room for a 2-deep ring of whole filter rows (or 2 taps)
  {
    const int want_w = 2 * (mode == 3 ? 1 : (KW < 2 || mode == 2 ? 2 : KW)) * p.wtile_bytes;
    p.a_sets = (2 * p.a_set_bytes + p.epi_groups * stage_bytes + 256 + want_w <= smem_cap) ? 2 : 1;
    if (getenv("LR_CONV_ASETS") && atoi(getenv("LR_CONV_ASETS")) == 1) p.a_sets = 1;       // tuning hook
  }
  p.smem_off_w = p.a_sets * p.a_set_bytes;
  {
    // weight ring: as many tiles per stage as fit (whole filter rows when possible), 3 stages deep
    const int n_taps_h = p.n_eff_taps;
    const int avail = smem_cap - p.smem_off_w - p.epi_groups * stage_bytes - 256;
    int stages = 3;
    int tps = avail / (stages * p.wtile_bytes);
    if (tps < 1) { stages = 2; tps = avail / (stages * p.wtile_bytes); }
    if (tps > n_taps_h) tps = n_taps_h;
    if (tps > 32) tps = 32;
    if (mode < 2 && tps >= KW) tps = tps / KW * KW;
    LR_CHECK_ARG(tps >= 1, "lr_conv3d_fwd: weight ring does not fit shared memory");
    if (tps == n_taps_h && stages > 2) stages = 2;
    p.tps = tps;
    p.w_stages = stages;
  }
  p.smem_off_stage = p.smem_off_w + p.w_stages * p.tps * p.wtile_bytes;
  // a second staging tile (drops one of the two named barriers per accumulator) when shared memory is left over
  p.stage_bufs = (!swap && p.smem_off_stage + 2 * p.epi_groups * stage_bytes + 256 <= smem_cap) ? 2 : 1;
  p.smem_off_halo = p.smem_off_stage + p.epi_groups * p.stage_bufs * stage_bytes;
  p.smem_off_bar = p.smem_off_halo + halo_bytes;
  p.cg_shift = -1;
  for (int sft = 0; sft < 8; ++sft) {
    if ((Cout >> 3) == (1 << sft)) p.cg_shift = sft;
    if (Wp == (1 << sft)) p.wp_shift = sft;
  }
  const size_t smem_bytes = (size_t)p.smem_off_bar + 256 + 512 + 1024;

  CUtensorMap map_x;
  int rc = make_map_2d(&map_x, x, (uint64_t)Cin, (uint64_t)p.rows_per_group * CG, (uint32_t)Cin, (uint32_t)p.CH,
                       p.row_bytes);
  if (rc != LR_OK) return rc;
  p.w_packed = reinterpret_cast<const uint8_t*>(w);
  p.dbg = g_conv_dbg;
  p.skip = g_conv_skip;
  p.stage_fence = getenv("LR_CONV_STAGE_FENCE") ? atoi(getenv("LR_CONV_STAGE_FENCE")) : 0;

  for (int kt = 0, tap = 0; kt < KT; ++kt)
    for (int ky = 0; ky < KH; ++ky)
      for (int kx = 0; kx < KW; kx += p.kxs, ++tap)
        p.tap_off[tap] = (uint32_t)kt * ((uint32_t)p.chunk_bytes >> 4) +
                         (((uint32_t)(ky * Wp + kx) * (uint32_t)p.row_bytes) >> 4);
  if (mode == 3) {
    for (int ky = 0, tap = 0; ky < KH; ++ky)
      for (int kx = 0; kx < KW; ++kx, ++tap)
        p.tap_off[tap] = ((uint32_t)(ky * Wp + kx) * (uint32_t)p.row_bytes) >> 4;
  }

  int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
#define LR_LAUNCH_CONV(KS, MD, UP)                                                                          \
  do {                                                                                                      \
    LR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tcgen05_kernel<KS, MD, UP>,                                   \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));      \
    conv3d_tcgen05_kernel<KS, MD, UP><<<grid, kThreads, smem_bytes, lr_stream(stream)>>>(map_x, p);         \
  } while (0)
#define LR_LAUNCH_CONV_KS(MD, UP)                                                                           \
  do {                                                                                                      \
    if (ks == 1) LR_LAUNCH_CONV(1, MD, UP); else if (ks == 2) LR_LAUNCH_CONV(2, MD, UP); else LR_LAUNCH_CONV(4, MD, UP); \
  } while (0)
  const int ks = Cin / 16;
  if (epi_mode == 2) {
    if (mode == 2) LR_LAUNCH_CONV_KS(2, true);
    else if (mode == 3) LR_LAUNCH_CONV_KS(3, true);
    else LR_LAUNCH_CONV_KS(0, true);
  } else if (mode == 1) LR_LAUNCH_CONV_KS(1, false);
  else if (mode == 2) LR_LAUNCH_CONV_KS(2, false);
  else if (mode == 3) LR_LAUNCH_CONV_KS(3, false);
  else LR_LAUNCH_CONV_KS(0, false);
#undef LR_LAUNCH_CONV_KS
#undef LR_LAUNCH_CONV
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_conv3d_fwd(const void* x, const void* w, const float* bias, void* y, uint8_t* argmax,
                             int B, int T, int H, int W, int Hp, int Wp, int Cin, int CG, int Cout, int KT,
                             int KH, int KW, int epi_mode, int oTp, int oHp, int oWp, int o_t, int o_y, int o_x,
                             int J, int swap, void* stream) {
  LR_CHECK_ARG(epi_mode == 0 || epi_mode == 1, "lr_conv3d_fwd: bad epilogue mode");
  return conv3d_launch(x, w, bias, y, argmax, nullptr, nullptr, B, T, H, W, Hp, Wp, Cin, CG, Cout, KT, KH, KW, epi_mode,
                       oTp, oHp, oWp, o_t, o_y, o_x, J, swap, stream);
}

// dgrad of one conv layer fused with the backward of the ReLU + MaxPool(1,2,2) in front of it: instead of the
// pooled-resolution gradient (B,T,H,W,Cout), the epilogue writes the un-pooled gradient straight into the padded
// dY volume (B,oTp,oHp,oWp,Cout) of the layer below (2x2 window of pooled pixel (y,x) at rows 2y+o_y.., columns
// 2x+o_x..) and accumulates that layer's bias gradient — lr_conv3d_fwd(epi_mode 1) + lr_unpool in one pass.
extern "C" int lr_conv3d_dgrad_unpool(const void* dy, const void* w, const uint8_t* argmax, void* dy_below,
                                      float* d_bias, int B, int T, int H, int W, int Hp, int Wp, int Cin, int CG,
                                      int Cout, int KT, int KH, int KW, int oTp, int oHp, int oWp, int o_t,
                                      int o_y, int o_x, int J, int swap, void* stream) {
  LR_CHECK_ARG(argmax && dy_below, "lr_conv3d_dgrad_unpool: null pointer");
  LR_CHECK_ARG(oHp >= 2 * H + o_y && oWp >= 2 * W + o_x, "lr_conv3d_dgrad_unpool: output volume too small");
  return conv3d_launch(dy, w, nullptr, dy_below, nullptr, argmax, d_bias, B, T, H, W, Hp, Wp, Cin, CG, Cout, KT, KH, KW,
                       2, oTp, oHp, oWp, o_t, o_y, o_x, J, swap, stream);
}
