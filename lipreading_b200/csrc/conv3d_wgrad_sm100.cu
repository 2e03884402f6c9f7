// conv3d_wgrad_sm100.cu — weight gradient of the conv front-end on tcgen05.
//
//   dW[co, tap, ci] = sum over padded positions p of  dY[p, co] * X[p + shift(tap), ci]
//
// with dY stored top-left aligned in the SAME zero-padded flattened geometry as X (it is the dgrad
// input volume viewed through a constant row offset), so that once more every tap is a pure ROW
// SHIFT of a shared-memory-resident tile.  The reduction (K) dimension is "positions", i.e. both
// operands are MN-major for the tensor core (channels contiguous, positions strided): the
// instruction descriptor sets a_major = b_major = MN and the shared-memory descriptors use the
// MN-major canonical layouts (row = 32/64/128 B of channels, 8-row groups SBO apart, further
// 64-/32-/16-channel blocks LBO apart — which is how three 32-channel gradient volumes become one
// N = 96 operand).
//
// One MMA atom is M x N x K=16 positions; per 128-position tile a CTA issues units_in_group x 8 of
// them into units_in_group TMEM accumulators (M lanes x N fp32 columns each).  M = 64 is half rate
// on the tensor pipe (profiles/r1_umma_table.txt), so the layers whose M side is dY stack S = 128/Cy
// filter rows (ky taps) on M = 128: M block b is the SAME dY tile shifted up by b image rows
// (descriptor LBO = Wp * row pitch), exactly like the kx taps are stacked on N.
// Grid = tap groups x splits; every CTA walks its share of the B*T*ytiles position tiles through a
// 3-stage TMA ring, then dumps its partial accumulators (fp32) to a workspace; a second kernel
// reduces the splits in a fixed order (deterministic).
#include "tcgen05.cuh"
#include <string.h>

using namespace lr_tc;

namespace {

constexpr int kMmaWarps = 4;                       // issuing warps: accumulator unit j belongs to warp j % 4
constexpr int kThreads = 32 * (1 + kMmaWarps + 4);
constexpr int kStages = 3;
constexpr int kMaxGroups = 32;

struct Side {
  int row_bytes;        // 32 / 64 / 128
  int blocks;           // MN blocks (channel groups stored as separate volumes, or duplicates)
  int rows;             // smem rows per block tile (128, or 128 + halo for the shifted side)
  int shifted;          // 1: this operand is X (taps shift it), 0: it is dY
  int tile_bytes;       // 1024-aligned bytes per block tile (= LBO when blocks > 1 and !dup)
  int lbo_bytes;        // descriptor LBO (0 duplicates block 0)
  long long group_rows; // global rows between channel-group volumes
  long long base_off;   // constant global row offset (dY: interior offset of the dgrad volume)
  uint32_t desc_hi;
};

struct WgradParams {
  int B, T, H, W, Tp, Hp, Wp, KH, KW;
  int R, n_ytiles, n_tiles;
  int pair_x;             // m_is_x with two taps per M = 128 MMA: unit u = flat (ky,kx) taps 2u, 2u+1 of a kt plane;
                          // M block 1 is the X chunk shifted by the second tap (descriptor LBO = shift difference)
  int Hv;                 // image rows of a plane in which dY (incl. its ky-stacking shifts) can be non-zero
  int n_groups, splits;
  int group_kt[kMaxGroups], group_lo[kMaxGroups], group_n[kMaxGroups], group_tap0[kMaxGroups];
  int group_nkt;          // kt planes per group (1, or KT when the whole filter depth is fused into one CTA)
  int per_kt;             // accumulator units per kt plane
  int Mrows;              // 64 or 128 (M of the MMA)
  int m_stack;            // S: ky taps stacked on M (1 = off); unit u covers ky = u*S .. u*S+S-1
  int n_region_bytes;     // smem bytes of one kt plane of the N-side operand
  Side m, n;
  int Nc;                 // N of the MMA = n.blocks * n.row_bytes / 2 (x KW when kx taps are stacked)
  int stack;              // 1: one accumulator per tap; KW: the KW kx-taps of a (kt,ky) row share one MMA
                          //    (N blocks = the same chunk shifted by one row each: LBO = row pitch)
  int n_taps_total;
  int tmem_cols;
  int stage_bytes;
  int smem_off_bar;
  uint32_t idesc;
  float* ws;              // [splits][n_taps_total][64][Nc]
};

enum { BAR_FULL = 0, BAR_EMPTY = kStages, BAR_ACC = 2 * kStages, BAR_COUNT = 2 * kStages + 1 };

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

__global__ void __launch_bounds__(kThreads, 1)
conv3d_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_n,
                            const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + p.smem_off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = blockIdx.x % p.n_groups, split = blockIdx.x / p.n_groups;
  const int kt = p.group_kt[gi], tap_lo = p.group_lo[gi], ntap = p.group_n[gi];
  const int m_bytes = p.m.blocks * p.m.tile_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      lr_mbar_init(&bars[BAR_FULL + s], 1);
      lr_mbar_init(&bars[BAR_EMPTY + s], kMmaWarps);
    }
    lr_mbar_init(&bars[BAR_ACC], kMmaWarps);
    lr_fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues
    uint32_t n = 0;
    const uint32_t tx = (uint32_t)(p.m.blocks * p.m.rows * p.m.row_bytes +
                                   p.group_nkt * p.n.blocks * p.n.rows * p.n.row_bytes);
    for (int tile = split; tile < p.n_tiles; tile += p.splits, ++n) {
      const int s = n % kStages;
      const int yt = tile % p.n_ytiles;
      const int bt = tile / p.n_ytiles;
      const int t = bt % p.T, b = bt / p.T;
      const long long p0 = (((long long)b * p.Tp + t) * p.Hp + (long long)yt * p.R) * p.Wp;
      lr_mbar_wait(&bars[BAR_EMPTY + s], ((n / kStages) & 1) ^ 1);
      if (elect_one()) {
        lr_mbar_expect_tx(&bars[BAR_FULL + s], tx);
        uint8_t* st = base + (size_t)s * p.stage_bytes;
        for (int blk = 0; blk < p.m.blocks; ++blk) {
          long long row = p0 + p.m.base_off + (p.m.shifted ? (long long)kt * p.Hp * p.Wp : 0) + blk * p.m.group_rows;
          tma_load_2d(st + (size_t)blk * p.m.tile_bytes, &map_m, 0, (int)row, &bars[BAR_FULL + s]);
        }
        for (int kk = 0; kk < p.group_nkt; ++kk)
          for (int blk = 0; blk < p.n.blocks; ++blk) {
            long long row = p0 + p.n.base_off + (p.n.shifted ? (long long)(kt + kk) * p.Hp * p.Wp : 0) +
                            blk * p.n.group_rows;
            tma_load_2d(st + m_bytes + (size_t)kk * p.n_region_bytes + (size_t)blk * p.n.tile_bytes, &map_n, 0,
                        (int)row, &bars[BAR_FULL + s]);
          }
      }
      __syncwarp();
    }
  } else if (warp <= kMmaWarps) {
    // MMA issuers: warp-uniform loops, one elected lane each; unit j is issued by warp j % kMmaWarps
    const int mw = warp - 1;
    uint32_t n = 0;
    const uint32_t m_step = (uint32_t)(16 * p.m.row_bytes) >> 4, n_step = (uint32_t)(16 * p.n.row_bytes) >> 4;
    for (int tile = split; tile < p.n_tiles; tile += p.splits, ++n) {
      const int s = n % kStages;
      lr_mbar_wait(&bars[BAR_FULL + s], (n / kStages) & 1);      // TMA data: ordered by the mbarrier, no tcgen05 fence
      const uint32_t m_addr = lr_smem_u32(base + (size_t)s * p.stage_bytes);
      const uint64_t md0 = make_desc_lbo(m_addr, (uint32_t)p.m.lbo_bytes, p.m.desc_hi);
      const uint64_t nd0 = make_desc_lbo(m_addr + (uint32_t)m_bytes, (uint32_t)p.n.lbo_bytes, p.n.desc_hi);
      const uint32_t acc0 = n == 0 ? 0u : 1u;
      // K = positions: the k-steps (16 positions = 16/Wp image rows) past the last row where dY can be non-zero
      // multiply zeros — the last tile of a plane issues only the steps it needs
      const int rows_here = min(p.R, p.Hv - (tile % p.n_ytiles) * p.R);
      const int nks = min(8, (rows_here * p.Wp + 15) >> 4);
      if (elect_one()) {
        for (int j = mw; j < ntap; j += kMmaWarps) {
          const int kk = j / p.per_kt;                 // kt plane inside a fused group (else 0: groups never span)
          const int unit = p.group_nkt > 1 ? j - kk * p.per_kt : tap_lo + j;
          const int u0 = p.pair_x ? 2 * unit : unit;
          const int ky = p.stack > 1 ? unit * p.m_stack : u0 / p.KW, kx = p.stack > 1 ? 0 : u0 - ky * p.KW;
          const uint32_t shift = (uint32_t)(ky * p.Wp + kx);
          uint64_t md = md0 + (uint64_t)(p.m.shifted ? (shift * p.m.row_bytes) >> 4 : 0u);
          if (p.pair_x) {
            const int u1 = u0 + 1, ky1 = u1 / p.KW, kx1 = u1 - ky1 * p.KW;
            const uint32_t d = u1 < p.KH * p.KW ? (uint32_t)(ky1 * p.Wp + kx1) - shift : 0u;   // odd tap out: duplicate
            md = make_desc_lbo(m_addr + shift * (uint32_t)p.m.row_bytes, d * (uint32_t)p.m.row_bytes, p.m.desc_hi);
          }
          const uint64_t nd = nd0 + (uint64_t)((uint32_t)(kk * p.n_region_bytes) >> 4) +
                              (uint64_t)(p.n.shifted ? (shift * p.n.row_bytes) >> 4 : 0u);
          const uint32_t d = tmem_base + (uint32_t)(j * p.Nc);
          umma_bf16(d, md, nd, p.idesc, acc0);
#pragma unroll
          for (int ks = 1; ks < 8; ++ks)
            if (ks < nks) umma_bf16(d, md + ks * m_step, nd + ks * n_step, p.idesc, 1u);
        }
        umma_commit(&bars[BAR_EMPTY + s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bars[BAR_ACC]);
    __syncwarp();
  } else {
    // epilogue: M=64 accumulators live in lanes {0-15, 32-47, 64-79, 96-111}: row m = 16*q + lane;
    // M=128: row m = lane of the TMEM = 32*q + lane
    const int q = warp & 3;
    lr_mbar_wait(&bars[BAR_ACC], 0);
    tc_fence_after();
    const bool any_tile = split < p.n_tiles;
    const bool wide = p.Mrows == 128;
    const int row = wide ? q * 32 + lane : q * 16 + lane;
    for (int j = 0; j < ntap; ++j) {
      const int tap_global = p.group_tap0[gi] + j;
      float* dst = p.ws + (((size_t)split * p.n_taps_total + tap_global) * p.Mrows + row) * p.Nc;
      for (int cc = 0; cc < p.Nc; cc += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.Nc + cc), v);
        if (wide || lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = any_tile ? make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                              __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst + cc + i) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// out[tap][m][n] = sum_split ws[split][tap][m][n]   (fixed order => deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int splits,
                                    long long per_split) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_split;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(size_t)s * per_split + i];
    out[i] = acc;
  }
}

void fill_side(Side& s, int ch_per_block, int blocks, int rows, int shifted, int dup, long long group_rows,
               long long base_off) {
  s.row_bytes = ch_per_block * 2;
  s.blocks = dup ? 1 : blocks;
  s.rows = rows;
  s.shifted = shifted;
  s.tile_bytes = (rows * s.row_bytes + 1023) / 1024 * 1024;
  s.lbo_bytes = dup ? 0 : s.tile_bytes;
  s.group_rows = group_rows;
  s.base_off = base_off;
  s.desc_hi = desc_hi_for(s.row_bytes, (uint32_t)(8 * s.row_bytes));
}

}  // namespace

// Floats per split (= size of `out`): without ky stacking [KT*KH*KW][64][Nc] (Nc = Cx, or Gy*Cy when m_is_x);
// with stack_ky = S > 1 (m_is_x = 0, kx stacked): [KT*ceil(KH/S)][128][KW*Cx].
extern "C" size_t lr_conv3d_wgrad_out_floats(int KT, int KH, int KW, int Nc, int stack_ky) {
  if (stack_ky < 0) return (size_t)KT * ((KH * KW + 1) / 2) * 128 * Nc;        // m_is_x tap pairs
  if (stack_ky > 1) return (size_t)KT * ((KH + stack_ky - 1) / stack_ky) * 128 * KW * Nc;
  return (size_t)KT * KH * KW * 64 * Nc;
}
extern "C" size_t lr_conv3d_wgrad_workspace(int KT, int KH, int KW, int Nc, int splits, int stack_ky) {
  return (size_t)splits * lr_conv3d_wgrad_out_floats(KT, KH, KW, Nc, stack_ky) * sizeof(float);
}

// dW partials for one layer.
//   x  : padded channels-last bf16 volume [B][Tp][Hp][Wp][Cx]  (Cx in {16,32,64}), Tp = T+KT-1
//   dy : [Gy][B][Tp][Hp][Wp][Cy] zero-padded volume(s) holding the conv-output gradient with its
//        interior at row offset dy_off (= pt*Hp*Wp + ph*Wp + pw), zeros everywhere else; Cy in {32,64}
//   m_is_x: 0 -> D[tap] is [co(64 lanes) x ci]  (M side = dy, needs Gy*Cy <= 64; Cy=32,Gy=1 is
//                duplicated to fill M=64), 1 -> D[tap] is [ci(64) x co] (M side = x, needs Cx = 64)
//   out: fp32 [KT*KH*KW][64][Nc], Nc = m_is_x ? Gy*Cy : Cx;  with stack_kx (m_is_x = 0 only) the layout is
//        [KT*KH][64][KW][Cx] (the kx taps of a row are the column blocks of one accumulator);
//        with stack_ky = S = 128/Cy as well: [KT][U = ceil(KH/S)][S][Cy][KW][Cx], where M block b of unit u
//        holds filter row ky = u*S + (S-1-b) (rows with ky >= KH are scratch).
//        stack_ky = -1 with m_is_x = 1: two taps (flat (ky,kx) index 2u, 2u+1 of a kt plane) share one M = 128 MMA —
//        M block 1 is the X chunk shifted by the second tap; out is [KT][ceil(KH*KW/2)][2][64][Nc] (an odd last
//        tap is duplicated into its second block).
//   fuse_kt (stack_ky only): one CTA accumulates all KT planes of a position tile (needs KT*U*KW*Cx <= 512
//        TMEM columns) instead of one tap group per kt -> the dY tile is fetched once, not KT times.
extern "C" int lr_conv3d_wgrad(const void* x, const void* dy, float* out, void* workspace, size_t ws_bytes,
                               int B, int T, int H, int W, int Hp, int Wp, int Cx, int Cy, int Gy,
                               long long dy_off, int KT, int KH, int KW, int m_is_x, int stack_kx, int stack_ky,
                               int fuse_kt, int splits, void* stream) {
  LR_CHECK_ARG(x && dy && out && workspace, "lr_conv3d_wgrad: null pointer");
  LR_CHECK_ARG(Cx == 16 || Cx == 32 || Cx == 64, "lr_conv3d_wgrad: Cx must be 16/32/64");
  LR_CHECK_ARG(Cy == 32 || Cy == 64, "lr_conv3d_wgrad: Cy must be 32/64");
  LR_CHECK_ARG(Wp == 8 || Wp == 16 || Wp == 32 || Wp == 64 || Wp == 128, "lr_conv3d_wgrad: Wp pow2 8..128");
  LR_CHECK_ARG(Wp >= W + KW - 1 && Hp >= H + KH - 1, "lr_conv3d_wgrad: padded extents too small");
  const int pair_x = (m_is_x && stack_ky == -1) ? 1 : 0;
  const int S = (stack_ky > 1 && !pair_x) ? stack_ky : 1;
  if (S > 1) {
    LR_CHECK_ARG(!m_is_x && stack_kx && Gy == 1 && S * Cy == 128,
                 "lr_conv3d_wgrad: stack_ky needs m_is_x = 0, stack_kx and stack_ky * Cy == 128");
  }
  LR_CHECK_ARG(!fuse_kt || S > 1, "lr_conv3d_wgrad: fuse_kt needs stack_ky");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.W = W; p.Tp = T + KT - 1; p.Hp = Hp; p.Wp = Wp; p.KH = KH; p.KW = KW;
  p.R = 128 / Wp;
  LR_CHECK_ARG(Hp % p.R == 0, "lr_conv3d_wgrad: Hp must be a multiple of %d (tiles may not straddle planes)", p.R);
  p.n_ytiles = lr_div_up(H, p.R);
  // ky stacking pairs dY image row (y - sh) with tile row y, sh = 0 .. min(S,KH)-1 for the filter rows that
  // exist: the tiles of a plane must reach row H-1+sh
  if (S > 1) p.n_ytiles = lr_div_up(H + (S < KH ? S : KH) - 1, p.R);
  LR_CHECK_ARG(p.n_ytiles * p.R <= Hp, "lr_conv3d_wgrad: padded plane too short for the stacked tiles");
  p.Hv = S > 1 ? H + (S < KH ? S : KH) - 1 : H;
  p.n_tiles = B * T * p.n_ytiles;
  const int U = (KH + S - 1) / S;                       // units per kt plane when ky is stacked
  const int CH = 128 + (S > 1 ? U * S - 1 : KH - 1) * Wp + (KW - 1);
  LR_CHECK_ARG(CH <= 256, "lr_conv3d_wgrad: halo too large");
  const long long vol_rows = (long long)B * p.Tp * Hp * Wp;
  p.Mrows = (S > 1 || pair_x) ? 128 : 64;
  p.pair_x = pair_x;
  p.m_stack = S;
  if (m_is_x) {
    LR_CHECK_ARG(Cx == 64, "lr_conv3d_wgrad: m_is_x needs Cx = 64");
    fill_side(p.m, Cx, 1, CH, 1, 0, 0, 0);
    fill_side(p.n, Cy, Gy, 128, 0, 0, vol_rows, dy_off);
    p.Nc = Gy * Cy;
  } else {
    LR_CHECK_ARG(Gy == 1 && (Cy == 64 || Cy == 32), "lr_conv3d_wgrad: M side = dy needs one group of 32/64 channels");
    if (S > 1) {
      // one TMA box of 128 + (S-1)*Wp rows starting (S-1) image rows above the tile; M block b = rows shifted
      // down by b*Wp, i.e. dY[p - (S-1-b)*Wp]  <->  filter row ky0 + (S-1-b)
      fill_side(p.m, Cy, 1, 128 + (S - 1) * Wp, 0, 0, vol_rows, dy_off - (long long)(S - 1) * Wp);
      p.m.lbo_bytes = Wp * p.m.row_bytes;
    } else {
      fill_side(p.m, Cy, Cy == 32 ? 2 : 1, 128, 0, Cy == 32, vol_rows, dy_off);
    }
    fill_side(p.n, Cx, 1, CH, 1, 0, 0, 0);
    p.Nc = Cx;
    if (stack_kx && KW > 1) {          // N = KW*Cx: block j of the N operand = the chunk shifted by j rows
      p.stack = KW;
      p.Nc = KW * Cx;
      p.n.lbo_bytes = p.n.row_bytes;
    }
  }
  if (p.stack == 0) p.stack = 1;
  LR_CHECK_ARG(S == 1 || p.stack > 1, "lr_conv3d_wgrad: stack_ky needs KW > 1");
  LR_CHECK_ARG(p.Nc % 16 == 0 && p.Nc <= 256, "lr_conv3d_wgrad: bad N (%d)", p.Nc);
  // tap groups: one kt each (or all kt when fused), at most 512/Nc accumulators
  const int per_kt = pair_x ? (KH * KW + 1) / 2 : (S > 1 ? U : (p.stack > 1 ? KH : KH * KW));   // accumulator units per kt plane
  const int max_taps = 512 / p.Nc;
  LR_CHECK_ARG(max_taps >= 1, "lr_conv3d_wgrad: N too wide for TMEM");
  p.per_kt = per_kt;
  p.group_nkt = 1;
  int g = 0, tap0 = 0;
  if (fuse_kt) {
    LR_CHECK_ARG(KT * per_kt <= max_taps, "lr_conv3d_wgrad: fuse_kt does not fit TMEM");
    p.group_nkt = KT;
    p.group_kt[0] = 0; p.group_lo[0] = 0; p.group_n[0] = KT * per_kt; p.group_tap0[0] = 0;
    g = 1;
  } else {
    for (int kt = 0; kt < KT; ++kt)
      for (int lo = 0; lo < per_kt; lo += max_taps) {
        LR_CHECK_ARG(g < kMaxGroups, "lr_conv3d_wgrad: too many tap groups");
        p.group_kt[g] = kt; p.group_lo[g] = lo; p.group_n[g] = per_kt - lo < max_taps ? per_kt - lo : max_taps;
        p.group_tap0[g] = tap0;
        tap0 += p.group_n[g];
        ++g;
      }
  }
  p.n_groups = g;
  p.n_taps_total = KT * per_kt;          // accumulator units (taps, or (kt,ky) rows when stacked)
  if (splits <= 0) splits = kNumSMs / p.n_groups;
  if (splits < 1) splits = 1;
  if (splits > p.n_tiles) splits = p.n_tiles;
  p.splits = splits;
  size_t need = (size_t)splits * p.n_taps_total * p.Mrows * p.Nc * sizeof(float);
  if (ws_bytes < need) { lr_set_error("lr_conv3d_wgrad: workspace %zu < %zu", ws_bytes, need); return LR_EWORKSPACE; }
  p.ws = reinterpret_cast<float*>(workspace);
  int max_units = 0;
  for (int i = 0; i < g; ++i) max_units = p.group_n[i] > max_units ? p.group_n[i] : max_units;
  int cols = 32;
  while (cols < max_units * p.Nc && cols < 512) cols <<= 1;
  p.tmem_cols = cols;
  p.n_region_bytes = p.n.blocks * p.n.tile_bytes;
  p.stage_bytes = p.m.blocks * p.m.tile_bytes + p.group_nkt * p.n_region_bytes;
  p.smem_off_bar = kStages * p.stage_bytes;
  const size_t smem_bytes = (size_t)p.smem_off_bar + 256 + 1024;
  LR_CHECK_ARG(smem_bytes <= 227 * 1024, "lr_conv3d_wgrad: stages do not fit shared memory");
  // instruction descriptor: f32 accum, bf16 x bf16, both operands MN-major, M = 64 / 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Nc >> 3) << 17) |
            ((uint32_t)(p.Mrows >> 4) << 24);

  const void* m_base = m_is_x ? x : dy;
  const void* n_base = m_is_x ? dy : x;
  const int m_ch = p.m.row_bytes / 2, n_ch = p.n.row_bytes / 2;
  const long long m_rows_total = m_is_x ? vol_rows : vol_rows * Gy;
  const long long n_rows_total = m_is_x ? vol_rows * Gy : vol_rows;
  CUtensorMap map_m, map_n;
  int rc = make_map_2d(&map_m, m_base, (uint64_t)m_ch, (uint64_t)m_rows_total, (uint32_t)m_ch, (uint32_t)p.m.rows,
                       p.m.row_bytes);
  if (rc != LR_OK) return rc;
  rc = make_map_2d(&map_n, n_base, (uint64_t)n_ch, (uint64_t)n_rows_total, (uint32_t)n_ch, (uint32_t)p.n.rows,
                   p.n.row_bytes);
  if (rc != LR_OK) return rc;
  LR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
  cudaStream_t st = lr_stream(stream);
  conv3d_wgrad_tcgen05_kernel<<<p.n_groups * splits, kThreads, smem_bytes, st>>>(map_m, map_n, p);
  LR_CHECK_LAUNCH();
  const long long per_split = (long long)p.n_taps_total * p.Mrows * p.Nc;
  wgrad_reduce_kernel<<<lr_div_up(per_split, 256), 256, 0, st>>>(p.ws, out, splits, per_split);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
