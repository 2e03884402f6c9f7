// tapgemm_sm100.cu — "tap GEMM": 2-D convolutions / transposed convolutions / plain GEMMs over channels-last
// bf16 volumes on tcgen05 + TMEM + TMA (the body of the position-map CNN,
// reference src/models/face/prnet.py:211-280; also the input GEMM of row a13).
//
//   acc[q, n] = sum_g sum_k  A[q + off_g][k] * W_g[n][k]            q = flattened position of the padded grid
//
// The activation is a zero-padded NHWC volume = a 2-D matrix [positions][C]; a filter tap is a constant ROW
// OFFSET, so the A tile of tap g for the 128 positions of a tile is one TMA box at row q0 + off_g (rows outside
// the matrix are zero-filled by the TMA unit).  No im2col buffer, no per-tap index arithmetic on the device.
//   * K tile = 64 bf16 = one 128-byte swizzled row.  Layers with C < 64 channels read a box whose ROW PITCH is
//     the position pitch (C*2 bytes < 128 bytes): the 128-byte row of position q then holds the channels of
//     positions q .. q+64/C-1, i.e. 64/C horizontally adjacent taps are fused into one K = 64 tile (16-channel
//     4x4 conv: 4 MMAs-groups of K = 64 instead of 16 of K = 16).
//   * transposed stride-2 convs run as 4 output phases of 2x2 taps each (phase = a work-item dimension, the
//     epilogue scatters to (2y+py, 2x+px)); stride-2 convs read a space-to-depth volume the PRODUCING layer's
//     epilogue wrote (store mode 2), so they are 2x2-tap stride-1 convs over 4C channels — no FLOP inflation.
//   * epilogue (4 warps, thread = position): tcgen05.ld -> alpha[c]*acc + beta[c] (folded inference batch-norm)
//     [+ gamma[c]*residual] -> none / ReLU / sigmoid -> bf16 into the interior of the next layer's padded volume
//     (or fp32 for the final layer / the plain-GEMM mode), optionally a second, even-subsampled copy for the
//     stride-2 shortcut of the next residual block.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-5 = epilogue; persistent CTAs
// (grid = min(items, 148)), two TMEM accumulators so the epilogue of item i overlaps the MMAs of item i+1.
#include "tcgen05.cuh"
#include <string.h>

using namespace lr_tc;

namespace {

constexpr int kBM = 128;
constexpr int kMaxBK = 64;                    // bf16 per K tile: 64 (128-byte swizzled rows), or 32 / 16 (64- / 32-byte rows)
constexpr int kMaxStages = 24;                // small-K layers: many small stages keep enough bytes in flight
constexpr int kMaxAcc = 8;                    // TMEM accumulator buffers (Ntile <= 64: 8, 128: 4, 256: 2)
constexpr int kEpiGroups = 2;                 // epilogue warp quartets (alternate work items)
constexpr int kThreads = 32 * (2 + 4 * kEpiGroups + 1);   // + a second MMA-issuing warp (warp 10)
constexpr int kMaxN = 256;                    // widest N tile
constexpr int kMaxTapGroups = 64;             // phases * groups
constexpr int kSmemCap = 192 * 1024;

struct TgParams {
  int n_mtiles, n_ntiles, n_phases, n_groups, n_chunks, n_items;
  int Ntile, Cout_pad, Cout;
  int Nrow;                      // weight rows per (phase, group) = pack * Cout_pad
  int a_chunked;                 // 1: A column coordinate = chunk*Kt (C >= Kt); 0: always 0 (fused-tap rows)
  int Kt, a_tile_bytes;          // K tile (elements) and bytes of the A tile
  int stages, stage_bytes;
  int n_acc;                     // accumulator buffers in use (power of two)
  // resident mode (small K, many taps): ALL weight tiles stay in shared memory, and per work item ONE chunk of
  // 128 + (max_off - min_off) rows is loaded; every tap's A operand is a descriptor into that chunk
  int resident, n_wtiles, w_tile_bytes, w_bytes, chunk_boxes, min_off;
  int rows, HpWp, Wp, vy0, vx0, H, W;
  int mode, act;
  int oHp, oWp, oC, opy, opx;
  int aHp, aWp, aC, apad;
  int resC;
  float out_scale;
  const float* alpha;
  const float* beta;
  const float* gamma;
  void* out;
  const __nv_bfloat16* res;
  __nv_bfloat16* aux;
  uint32_t idesc, desc_hi;
  struct { uint32_t m, s; } d_mtiles, d_ntiles, d_hpwp, d_wp;   // magic numbers: exact division of values < 2^31
  int tap_off[kMaxTapGroups];
};

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

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// exact q / d for q < 2^31: (q * m) >> s with m = ceil(2^s / d), s = 31 + ceil(log2 d)
template <typename F>
__device__ __forceinline__ uint32_t fdiv(uint32_t q, F f) {
  return (uint32_t)(((uint64_t)q * f.m) >> f.s);
}

template <int MODE, int PACK>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint8_t* wbase = base + (size_t)p.stages * p.stage_bytes;               // resident weight tiles (resident mode)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + p.w_bytes);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kMaxStages;
  uint64_t* bar_acc_full = bars + 2 * kMaxStages;
  uint64_t* bar_acc_empty = bars + 2 * kMaxStages + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAcc);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tmem_slot + 4);
  float* vecs = reinterpret_cast<float*>(bar_w + 2);       // [kEpiGroups][3][kMaxN]: alpha, beta, gamma of the current N tile
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.n_acc * p.Ntile) tmem_cols <<= 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      lr_mbar_init(&bar_full[s], 1);
      lr_mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < kMaxAcc; ++b) {
      lr_mbar_init(&bar_acc_full[b], 1);
      lr_mbar_init(&bar_acc_empty[b], 128);
    }
    lr_mbar_init(bar_w, 1);
    lr_fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_k = p.n_groups * p.n_chunks;                  // stages (K tiles) per work item

  if (warp == 0) {
    // ===== TMA producer =====
    uint32_t it = 0;
    if (p.resident) {
      if (elect_one()) {
        lr_mbar_expect_tx(bar_w, p.w_bytes);
        for (int t = 0; t < p.n_wtiles; ++t)
          tma_load_2d(wbase + (size_t)t * p.w_tile_bytes, &map_w, 0, t * p.Nrow, bar_w);
      }
      __syncwarp();
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int r = (int)fdiv((uint32_t)item, p.d_mtiles), m = item - r * p.n_mtiles;
        const int s = it % p.stages;
        lr_mbar_wait(&bar_empty[s], ((it / p.stages) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* st = base + (size_t)s * p.stage_bytes;
          lr_mbar_expect_tx(&bar_full[s], p.stage_bytes);
          for (int bx = 0; bx < p.chunk_boxes; ++bx)
            tma_load_2d(st + (size_t)bx * p.a_tile_bytes, &map_a, 0, m * kBM + p.min_off + bx * kBM, &bar_full[s]);
        }
        __syncwarp();
      }
    } else
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int r = (int)fdiv((uint32_t)item, p.d_mtiles), m = item - r * p.n_mtiles;
      const int ph = (int)fdiv((uint32_t)r, p.d_ntiles), nt = r - ph * p.n_ntiles;
      const int q0 = m * kBM;
      for (int g = 0; g < p.n_groups; ++g) {
        const int pg = ph * p.n_groups + g;
        const int arow = q0 + p.tap_off[pg];
        const int wrow = pg * p.Nrow + nt * p.Ntile;
        for (int kc = 0; kc < p.n_chunks; ++kc, ++it) {
          const int s = it % p.stages;
          lr_mbar_wait(&bar_empty[s], ((it / p.stages) & 1) ^ 1);     // (tiles are short: no back-off sleeps)
          if (elect_one()) {
            uint8_t* st = base + (size_t)s * p.stage_bytes;
            lr_mbar_expect_tx(&bar_full[s], p.stage_bytes);          // OOB rows / columns are zero-filled and counted
            tma_load_2d(st, &map_a, p.a_chunked ? kc * p.Kt : 0, arow, &bar_full[s]);
            tma_load_2d(st + p.a_tile_bytes, &map_w, kc * p.Kt, wrow, &bar_full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1 || warp == 2 + 4 * kEpiGroups) {
    // ===== MMA issuers: two warps take alternate work items (disjoint accumulators and ring stages) =====
    // (one thread sustains about one tcgen05.mma per ~100 cycles from this loop; an N <= 64 MMA occupies the pipe for
    //  <= 48 — a second issuing thread keeps the pipe fed on the narrow layers)
    // Only in resident mode (one ring stage per work item, an EVEN number of stages: each issuer then meets "its" stages
    // and accumulators in order, one mbarrier phase at a time).  With several stages per item the second issuer would
    // wait on phases two laps ahead of the barrier — a parity wait cannot tell those apart — so the streamed mode keeps
    // one issuer.
    const uint32_t iw = warp == 1 ? 0u : 1u;
    const bool two = p.resident != 0;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
      if (two ? (n & 1u) != iw : iw != 0u) continue;
      uint32_t it = p.resident ? n : n * (uint32_t)n_k;
      const uint32_t buf = n & (uint32_t)(p.n_acc - 1);
      lr_mbar_wait(&bar_acc_empty[buf], ((n / (uint32_t)p.n_acc) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * (uint32_t)p.Ntile;
      if (p.resident) {
        lr_mbar_wait(bar_w, 0);                       // (completes once; later waits return at the first poll)
        const int ph = (int)fdiv(fdiv((uint32_t)item, p.d_mtiles), p.d_ntiles);
        const int s = it % p.stages;
        lr_mbar_wait(&bar_full[s], (it / p.stages) & 1);
        if (elect_one()) {
          const uint32_t c_addr = lr_smem_u32(base + (size_t)s * p.stage_bytes);
          const uint32_t w_addr = lr_smem_u32(wbase);
          for (int g = 0; g < p.n_groups; ++g) {
            const int pg = ph * p.n_groups + g;
            // tap = row offset into the resident chunk (rows are Kt*2 bytes; the swizzle follows absolute addresses)
            const uint64_t ad = make_desc(c_addr + (uint32_t)(p.tap_off[pg] - p.min_off) * (uint32_t)(p.Kt * 2), p.desc_hi);
            const uint64_t bd = make_desc(w_addr + (uint32_t)pg * (uint32_t)p.w_tile_bytes, p.desc_hi);
            const int n_kk = p.Kt >> 4;
#pragma unroll 4
            for (int kk = 0; kk < n_kk; ++kk) umma_bf16(d, ad + 2 * kk, bd + 2 * kk, p.idesc, (g > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&bar_empty[s]);
          umma_commit(&bar_acc_full[buf]);
        }
        __syncwarp();
        continue;
      }
      for (int k = 0; k < n_k; ++k, ++it) {
        const int s = it % p.stages;
        lr_mbar_wait(&bar_full[s], (it / p.stages) & 1);
        if (elect_one()) {
          const uint32_t a_addr = lr_smem_u32(base + (size_t)s * p.stage_bytes);
          const uint64_t ad = make_desc(a_addr, p.desc_hi), bd = make_desc(a_addr + p.a_tile_bytes, p.desc_hi);
          const int n_kk = p.Kt >> 4;
#pragma unroll 4
          for (int kk = 0; kk < n_kk; ++kk)                           // K = 16 bf16 = 32 bytes = 2 descriptor units
            umma_bf16(d, ad + 2 * kk, bd + 2 * kk, p.idesc, (k > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&bar_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar_acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ===== epilogue: two warp quartets take alternate work items; thread = one position (TMEM lane) =====
    // (a tile's epilogue is a dependent chain of a few hundred instructions on ONE warp per scheduler: with 16- or
    //  32-column tiles it, not the tensor pipe or HBM, bounds the kernel — hence magic-number divisions, the
    //  per-channel vectors in shared memory, a compile-time store mode and two quartets)
    const int quarter = warp & 3;
    const int eg = (warp - 2) >> 2;                             // quartet 0: warps 2-5, quartet 1: warps 6-9
    float* v_alpha = vecs + eg * 3 * kMaxN;
    float* v_beta = v_alpha + kMaxN;
    float* v_gamma = v_beta + kMaxN;
    const int etid = ((warp - 2) & 3) * 32 + lane;
    int cur_nt = -1;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
      if ((int)(n & 1) != eg) continue;
      const int r = (int)fdiv((uint32_t)item, p.d_mtiles), m = item - r * p.n_mtiles;
      const int ph = (int)fdiv((uint32_t)r, p.d_ntiles), nt = r - ph * p.n_ntiles;
      if (nt != cur_nt) {                                       // (rare: items run over the M tiles first)
        named_bar_sync(1 + eg, 128);
        const int nv = PACK > 1 ? p.Cout_pad : p.Ntile, v0 = PACK > 1 ? 0 : nt * p.Ntile;
        for (int c = etid; c < nv; c += 128) {
          v_alpha[c] = p.alpha ? p.alpha[v0 + c] : 1.f;
          v_beta[c] = p.beta ? p.beta[v0 + c] : 0.f;
          v_gamma[c] = p.gamma ? p.gamma[v0 + c] : 1.f;
        }
        named_bar_sync(1 + eg, 128);
        cur_nt = nt;
      }
      const uint32_t buf = n & (uint32_t)(p.n_acc - 1);
      const int q = m * kBM + quarter * 32 + lane;           // matrix row of this thread = PACK consecutive positions
      // ---- where the PACK positions of this row go -------------------------------------------------------------
      bool valid[PACK];
      long long orow[PACK], arow[PACK];
      int coff[PACK];
#pragma unroll
      for (int j = 0; j < PACK; ++j) {
        const int qp = q * PACK + j;
        valid[j] = q < p.rows;
        orow[j] = qp;                            // mode 3: the GEMM's own row
        arow[j] = -1;
        coff[j] = nt * p.Ntile;                  // channel offset inside the output row (PACK > 1: one N tile)
        if (MODE != 3) {
          const int b = (int)fdiv((uint32_t)qp, p.d_hpwp);
          const int rr = qp - b * p.HpWp;
          const int y = (int)fdiv((uint32_t)rr, p.d_wp);
          const int yy = y - p.vy0;
          const int xx = rr - y * p.Wp - p.vx0;
          valid[j] = valid[j] && (unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
          if (MODE == 0) {
            orow[j] = ((long long)b * p.oHp + yy + p.opy) * p.oWp + xx + p.opx;
            if (p.aux != nullptr && valid[j] && !((yy | xx) & 1))
              arow[j] = ((long long)b * p.aHp + (yy >> 1) + p.apad) * p.aWp + (xx >> 1) + p.apad;
          } else if (MODE == 1) {
            orow[j] = ((long long)b * p.oHp + 2 * yy + (ph >> 1) + p.opy) * p.oWp + 2 * xx + (ph & 1) + p.opx;
          } else if (MODE == 2) {
            orow[j] = ((long long)b * p.oHp + ((yy + 1) >> 1) + p.opy) * p.oWp + ((xx + 1) >> 1) + p.opx;
            coff[j] += ((((yy + 1) & 1) << 1) | ((xx + 1) & 1)) * p.Cout_pad;
          } else {                               // mode 4: compact fp32 (B,H,W,Cout)
            orow[j] = ((long long)b * p.H + yy) * p.W + xx;
          }
        }
      }
      // columns of the accumulator per position: PACK > 1 -> Cout_pad each (position j at j*Cout_pad); else the N tile
      const int n_cols = PACK > 1 ? p.Cout_pad : p.Ntile;
      // the residual of the first 16 channels is requested BEFORE the accumulator wait, every later chunk's while the one
      // before it is processed: its L2 / HBM latency is off the tile-to-tile chain
      const bool has_res = MODE == 0 && p.res != nullptr;
      uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
      if (has_res && valid[0]) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + orow[0] * p.resC + coff[0]);
        r0 = __ldg(rp);
        r1 = __ldg(rp + 1);
      }
      lr_mbar_wait(&bar_acc_full[buf], (n / (uint32_t)p.n_acc) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + buf * (uint32_t)p.Ntile + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int j = 0; j < PACK; ++j) {
#pragma unroll 1
        for (int c0 = 0; c0 < n_cols; c0 += 16) {
          uint32_t v[16];
          uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
          if (has_res) {                         // next chunk of this position, or the first chunk of the next one
            if (c0 + 16 < n_cols) {
              if (valid[j]) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + orow[j] * p.resC + coff[j] + c0 + 16);
                n0 = __ldg(rp);
                n1 = __ldg(rp + 1);
              }
            } else if (j + 1 < PACK) {
              if (valid[j + 1 < PACK ? j + 1 : j]) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + orow[j + 1 < PACK ? j + 1 : j] * p.resC +
                                                                 coff[j + 1 < PACK ? j + 1 : j]);
                n0 = __ldg(rp);
                n1 = __ldg(rp + 1);
              }
            }
          }
          tmem_ld16(trow + (uint32_t)(j * n_cols + c0), v);
          if (valid[j]) {
            const int cn = (PACK > 1 ? 0 : nt * p.Ntile) + c0;      // channel index in [0, Cout_pad)
            float f[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = *reinterpret_cast<const float4*>(v_alpha + c0 + 4 * i);
              const float4 b4 = *reinterpret_cast<const float4*>(v_beta + c0 + 4 * i);
              f[4 * i] = fmaf(__uint_as_float(v[4 * i]), a4.x, b4.x);
              f[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), a4.y, b4.y);
              f[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), a4.z, b4.z);
              f[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), a4.w, b4.w);
            }
            if (has_res) {
              const uint32_t ru[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 g4 = *reinterpret_cast<const float4*>(v_gamma + c0 + 4 * i);
                f[4 * i] = fmaf(bf16_lo(ru[2 * i]), g4.x, f[4 * i]);
                f[4 * i + 1] = fmaf(bf16_hi(ru[2 * i]), g4.y, f[4 * i + 1]);
                f[4 * i + 2] = fmaf(bf16_lo(ru[2 * i + 1]), g4.z, f[4 * i + 2]);
                f[4 * i + 3] = fmaf(bf16_hi(ru[2 * i + 1]), g4.w, f[4 * i + 3]);
              }
            }
            if (p.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.act == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (MODE != 4 || cn + i < p.Cout) f[i] = 1.f / (1.f + __expf(-f[i]));
            }
            if (MODE == 3) {
              float* o = reinterpret_cast<float*>(p.out) + orow[j] * p.oC + coff[j] + c0;
              if (cn + 16 <= p.Cout && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (cn + i < p.Cout) o[i] = f[i];
              }
            } else if (MODE == 4) {
              float* o = reinterpret_cast<float*>(p.out) + orow[j] * p.Cout;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (cn + i < p.Cout) o[cn + i] = f[i] * p.out_scale;
            } else {
              uint4 s0, s1;
              s0.x = pack_bf16x2(f[0], f[1]);   s0.y = pack_bf16x2(f[2], f[3]);
              s0.z = pack_bf16x2(f[4], f[5]);   s0.w = pack_bf16x2(f[6], f[7]);
              s1.x = pack_bf16x2(f[8], f[9]);   s1.y = pack_bf16x2(f[10], f[11]);
              s1.z = pack_bf16x2(f[12], f[13]); s1.w = pack_bf16x2(f[14], f[15]);
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow[j] * p.oC + coff[j] + c0);
              o[0] = s0;
              o[1] = s1;
              if (MODE == 0 && arow[j] >= 0) {
                uint4* a2 = reinterpret_cast<uint4*>(p.aux + arow[j] * p.aC + cn);
                a2[0] = s0;
                a2[1] = s1;
              }
            }
          }
          r0 = n0;
          r1 = n1;
          __syncwarp();                          // tcgen05.ld is warp-collective: reconverge before the next one
        }
      }
      tc_fence_before();
      lr_mbar_arrive(&bar_acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int make_map_bf16_strided(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t stride_bytes,
                          uint32_t box_inner, uint32_t box_outer, int row_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { lr_set_error("cuTensorMapEncodeTiled entry point not available"); return LR_ECUDA; }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                              : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lr_set_error("cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu stride %llu box %u x %u", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride_bytes, box_inner, box_outer);
    return LR_ECUDA;
  }
  return LR_OK;
}

}  // namespace

extern "C" int lr_tapgemm(const lr_tapgemm_desc* d, void* stream) {
  LR_CHECK_ARG(d != nullptr && d->a && d->w && d->out, "lr_tapgemm: null descriptor / operand");
  LR_CHECK_ARG(d->rows > 0 && d->rows < (1ll << 31) - 4096, "lr_tapgemm: rows out of range");
  LR_CHECK_ARG(d->C >= 8 && d->C % 8 == 0, "lr_tapgemm: C must be a multiple of 8 (16-byte position pitch)");
  LR_CHECK_ARG(d->n_phases >= 1 && d->n_groups >= 1 && d->n_phases * d->n_groups <= kMaxTapGroups && d->tap_off,
               "lr_tapgemm: 1..%d tap groups over all phases", kMaxTapGroups);
  const int Kt = d->Kt;
  LR_CHECK_ARG(Kt == 16 || Kt == 32 || Kt == kMaxBK, "lr_tapgemm: K tile must be 16, 32 or 64");
  LR_CHECK_ARG(d->Kg >= 1 && (d->C * (d->pack > 1 ? d->pack : 1) >= Kt ? d->Kg <= d->C * (d->pack > 1 ? d->pack : 1) : d->Kg == Kt),
               "lr_tapgemm: Kg = Kt (fused taps, C < Kt) or <= the channels of a matrix row");
  LR_CHECK_ARG(d->Cout_pad >= 16 && d->Cout_pad % 16 == 0 && d->Cout >= 1 && d->Cout <= d->Cout_pad,
               "lr_tapgemm: Cout_pad must be a multiple of 16 >= Cout");
  LR_CHECK_ARG(d->mode >= 0 && d->mode <= 4 && d->act >= 0 && d->act <= 2, "lr_tapgemm: mode 0..4, act 0..2");
  LR_CHECK_ARG(d->mode == 1 ? d->n_phases == 4 : d->n_phases == 1, "lr_tapgemm: 4 phases in mode 1, else 1");
  LR_CHECK_ARG(d->res == nullptr || d->mode == 0, "lr_tapgemm: residual input only in store mode 0");
  LR_CHECK_ARG(d->aux == nullptr || d->mode == 0, "lr_tapgemm: subsampled copy only in store mode 0");
  LR_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "lr_tapgemm: operands must be 16-byte aligned");
  if (d->mode != 3)
    LR_CHECK_ARG(d->Hp > 0 && d->Wp > 0 && d->H > 0 && d->W > 0 && (d->rows * (d->pack > 1 ? d->pack : 1)) % ((long long)d->Hp * d->Wp) == 0,
                 "lr_tapgemm: rows must be whole Hp x Wp grids");
  if (d->mode <= 2) LR_CHECK_ARG(d->oC % 8 == 0 && d->oC >= (d->mode == 2 ? 4 : 1) * d->Cout_pad, "lr_tapgemm: output row too narrow");

  const int PK = d->pack > 0 ? d->pack : 1;
  LR_CHECK_ARG(PK == 1 || ((PK == 2 || PK == 4) && d->mode != 3 && PK * d->C == d->Kt && PK * d->Cout_pad <= kMaxN &&
                           d->Wp % PK == 0),
               "lr_tapgemm: pack 2 / 4 needs pack*C == Kt, pack*Cout_pad <= %d, Wp %% pack == 0 and a grid store mode", kMaxN);
  const int Crow = d->C * PK;        // channels per matrix row
  TgParams p;
  memset(&p, 0, sizeof(p));
  // N tile: the largest divisor of the row block that is a multiple of 16 and <= 256
  p.Nrow = PK * d->Cout_pad;
  int ntile = p.Nrow <= 256 ? p.Nrow : 256;
  while (p.Nrow % ntile) ntile -= 16;
  p.Ntile = ntile;
  p.n_ntiles = p.Nrow / ntile;
  p.n_mtiles = lr_div_up(d->rows, kBM);
  p.n_phases = d->n_phases;
  p.n_groups = d->n_groups;
  p.n_chunks = lr_div_up(d->Kg, Kt);
  p.a_chunked = Crow >= Kt;
  p.Kt = Kt;
  p.a_tile_bytes = kBM * Kt * 2;
  const long long items = (long long)p.n_mtiles * p.n_ntiles * p.n_phases;
  LR_CHECK_ARG(items < (1ll << 31), "lr_tapgemm: too many work items");
  p.n_items = (int)items;
  p.Cout_pad = d->Cout_pad;
  p.Cout = d->Cout;
  p.stage_bytes = p.a_tile_bytes + ntile * Kt * 2;
  p.stages = kSmemCap / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  {
    // resident mode: one K tile per group, several groups, one N tile, and everything fits: weights + >= 2 chunks
    int mn = d->tap_off[0], mx = d->tap_off[0];
    for (int i = 1; i < d->n_phases * d->n_groups; ++i) {
      mn = d->tap_off[i] < mn ? d->tap_off[i] : mn;
      mx = d->tap_off[i] > mx ? d->tap_off[i] : mx;
    }
    const int boxes = lr_div_up(kBM + mx - mn, kBM);
    const int wt = p.Nrow * Kt * 2, nw = d->n_phases * d->n_groups;
    const long long need = (long long)nw * wt + 2ll * boxes * p.a_tile_bytes;
    if (!(d->flags & 1) && p.n_chunks == 1 && p.n_ntiles == 1 && d->n_groups >= 3 && Crow >= Kt && need <= kSmemCap) {
      p.resident = 1;
      p.n_wtiles = nw;
      p.w_tile_bytes = wt;
      p.w_bytes = nw * wt;
      p.chunk_boxes = boxes;
      p.min_off = mn;
      p.stage_bytes = boxes * p.a_tile_bytes;
      p.stages = (kSmemCap - p.w_bytes) / p.stage_bytes;
      if (p.stages > 4) p.stages = 4;
      if (p.stages == 3) p.stages = 2;      // even: the two MMA-issuing warps alternate stages
    }
  }
  p.n_acc = ntile <= 64 ? 8 : (ntile <= 128 ? 4 : 2);
  p.rows = (int)d->rows;
  p.HpWp = d->Hp * d->Wp; p.Wp = d->Wp; p.vy0 = d->vy0; p.vx0 = d->vx0; p.H = d->H; p.W = d->W;
  p.mode = d->mode; p.act = d->act;
  p.oHp = d->oHp; p.oWp = d->oWp; p.oC = d->oC; p.opy = d->opy; p.opx = d->opx;
  p.aHp = d->aHp; p.aWp = d->aWp; p.aC = d->aC; p.apad = d->apad;
  p.resC = d->resC;
  p.out_scale = d->out_scale;
  p.alpha = d->alpha; p.beta = d->beta; p.gamma = d->gamma;
  p.out = d->out;
  p.res = reinterpret_cast<const __nv_bfloat16*>(d->res);
  p.aux = reinterpret_cast<__nv_bfloat16*>(d->aux);
  // D = f32, A = B = bf16, both K-major, N = ntile, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ntile >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  p.desc_hi = desc_hi_for(Kt * 2, 8 * Kt * 2);
  for (int i = 0; i < d->n_phases * d->n_groups; ++i) p.tap_off[i] = d->tap_off[i];
  auto magic = [](uint32_t dv, uint32_t& m, uint32_t& sh) {
    if (dv == 0) dv = 1;
    uint32_t l = 0;
    while ((1ull << l) < dv) ++l;
    sh = 31 + l;
    m = (uint32_t)(((1ull << sh) + dv - 1) / dv);
  };
  magic((uint32_t)p.n_mtiles, p.d_mtiles.m, p.d_mtiles.s);
  magic((uint32_t)p.n_ntiles, p.d_ntiles.m, p.d_ntiles.s);
  magic((uint32_t)p.HpWp, p.d_hpwp.m, p.d_hpwp.s);
  magic((uint32_t)p.Wp, p.d_wp.m, p.d_wp.s);

  CUtensorMap map_a, map_w;
  // A: [rows][C] with the position pitch as row stride; the inner extent is C (>= Kt: chunks of Kt, columns past C
  // are zero-filled) or Kt (C < Kt: the row runs over the next positions — fused horizontal taps)
  int rc = make_map_bf16_strided(&map_a, d->a, (uint64_t)(Crow >= Kt ? Crow : Kt), (uint64_t)d->rows,
                                 (uint64_t)Crow * 2, Kt, kBM, Kt * 2);
  if (rc != LR_OK) return rc;
  const uint64_t w_rows = (uint64_t)d->n_phases * d->n_groups * p.Nrow;
  LR_CHECK_ARG(d->w_pitch >= d->Kg && d->w_pitch % 8 == 0, "lr_tapgemm: weight row pitch must be >= Kg and a multiple of 8");
  rc = make_map_bf16_strided(&map_w, d->w, (uint64_t)d->Kg, w_rows, (uint64_t)d->w_pitch * 2, Kt, (uint32_t)ntile, Kt * 2);
  if (rc != LR_OK) return rc;
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + p.w_bytes + (2 * kMaxStages + 2 * kMaxAcc + 2) * 8 + 16 +
                            kEpiGroups * 3 * kMaxN * sizeof(float) + 1024;
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
#define LR_TG_LAUNCH(M, P)                                                                                             \
  case M * 8 + P:                                                                                                      \
    LR_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)); \
    tapgemm_kernel<M, P><<<grid, kThreads, smem_bytes, lr_stream(stream)>>>(map_a, map_w, p);                          \
    break;
  switch (d->mode * 8 + PK) {
    LR_TG_LAUNCH(0, 1) LR_TG_LAUNCH(1, 1) LR_TG_LAUNCH(2, 1) LR_TG_LAUNCH(3, 1) LR_TG_LAUNCH(4, 1)
    LR_TG_LAUNCH(0, 2) LR_TG_LAUNCH(1, 2) LR_TG_LAUNCH(2, 2) LR_TG_LAUNCH(4, 2)
    LR_TG_LAUNCH(0, 4) LR_TG_LAUNCH(1, 4) LR_TG_LAUNCH(2, 4) LR_TG_LAUNCH(4, 4)
  }
#undef LR_TG_LAUNCH
  LR_CHECK_LAUNCH();
  return LR_OK;
}

// ---- lr_pack_image16: (N,H,W,3) f32 -> interior of the zero-padded (N,H+2p,W+2p,16) bf16 volume ----------------
__global__ void pack_image16_kernel(const float* __restrict__ img, uint4* __restrict__ out, long long n_px, int H, int W,
                                    int pad) {
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / W;
    const int x = (int)(i - row * W);
    const long long b = row / H;
    const int y = (int)(row - b * H);
    const float r = img[3 * i], g = img[3 * i + 1], bl = img[3 * i + 2];
    uint4 lo = make_uint4(pack_bf16x2(r, g), pack_bf16x2(bl, 0.f), 0u, 0u);
    uint4* o = out + 2 * ((b * Hp + y + pad) * Wp + x + pad);
    o[0] = lo;
    o[1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

extern "C" int lr_pack_image16(const float* img, void* out_bf16, int N, int H, int W, int pad, void* stream) {
  LR_CHECK_ARG(img && out_bf16 && N > 0 && H > 0 && W > 0 && pad >= 0, "lr_pack_image16: bad arguments");
  const long long n_px = (long long)N * H * W;
  const int grid = (int)(lr_div_up(n_px, 256) < kNumSMs * 16 ? lr_div_up(n_px, 256) : kNumSMs * 16);
  pack_image16_kernel<<<grid, 256, 0, lr_stream(stream)>>>(img, reinterpret_cast<uint4*>(out_bf16), n_px, H, W, pad);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
