// vision.cu — per-frame geometry kernels of the landmark dataview path (SURVEY §8 rows a2-a4,
// a6-a9, a11, N2).  All of them are HBM/latency-bound byte and index work: batched over frames,
// coalesced 128-bit accesses where the layout allows it, grids sized in multiples of 148 SMs.
//
// Reference arithmetic restated (file:line in the reference checkout):
//   _applyPadding        src/utils/data/face.py:76-90        -> rect_geometry_kernel
//   PRN.process crop box src/models/face/prnet.py:112-119    -> rect_geometry_kernel
//   estimate_transform / image/255. / warp   prnet.py:137-143 -> warp256_kernel
//   restore + get_landmarks/get_vertices     prnet.py:151-156,169,179-180
//   getFace translate    src/utils/data/face.py:171-174      -> posmap_gather_kernel
//   _collate_fn._pad     src/data/data_loader.py:124-137     -> collate_pad_kernel
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// a2 + a3.  Python's int(0.3*bw) == (3*bw)/10 and int(((r-l)+(b-t))/2*1.6) == (4*k)/5 for every
// non-negative extent below 2000 / 6000 (checked exhaustively against the float64 restatement in
// tests/test_vision_oracle.py), so the kernel is integer-exact.
__global__ void rect_geometry_kernel(const int32_t* __restrict__ rects, int N, int img_h, int img_w,
                                     int32_t* __restrict__ rect_pad, int32_t* __restrict__ crop) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int4 rc = reinterpret_cast<const int4*>(rects)[i];
  int l = rc.x, r = rc.y, t = rc.z, b = rc.w;
  int bw = r - l, bh = b - t;
  int pw = (3 * bw) / 10, ph = (3 * bh) / 10;
  int4 o;
  o.x = max(0, l - pw);
  o.y = min(img_w, r + pw);
  o.z = max(0, t - ph);
  o.w = min(img_h, b + ph);
  reinterpret_cast<int4*>(rect_pad)[i] = o;
  int k = bw + bh;
  int4 c;
  c.x = r + l;          // 2*cx
  c.y = b + t;          // 2*cy
  c.z = (4 * k) / 5;    // size
  c.w = 0;
  reinterpret_cast<int4*>(crop)[i] = c;
}

// ---------------------------------------------------------------------------------------------
// a4.  One thread per output pixel (u fastest).
// Source sampling follows skimage 0.14 `_warp_fast` bilinear, mode='constant', cval=0:
//   (x,y) = T^-1 (u,v) ; floor/ceil taps ; out-of-image taps contribute 0.
// The similarity fitted to the three crop corners is an exact axis-aligned scale + shift, so
//   x = u*size/255 + (cx - size/2),  y = v*size/255 + (cy - size/2)
// evaluated in fp64 like the reference (fp32 coordinates at x~1000 would already cost 1e-4); the one
// fp64 division is done once per CTA.  The interpolation itself runs in fp32 on the raw bytes and is
// scaled by 1/255 at the end: the output is float32, and this differs from the reference's float64
// `image/255.` + interpolation by at most an ulp of that float32 (test tolerance 1e-6).
// Only the <= size^2 window of the frame is ever read (the reference divides the whole frame); the two
// taps of a source row are 6 adjacent bytes, fetched as aligned 32-bit words instead of 6 byte loads
// (the kernel is LSU-bound, not HBM-bound, with byte loads).

// bytes off .. off+5 of buf (zero outside [0, total)); buf is 4-byte aligned
__device__ __forceinline__ void load6(const uint8_t* __restrict__ buf, long long off, long long total,
                                      uint32_t (&b)[6]) {
  const long long w = off & ~3LL;
  const int sh = (int)(off & 3);
  uint32_t x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long wi = w + 4 * i;
    if (wi >= 0 && wi + 4 <= total) {
      x[i] = *reinterpret_cast<const uint32_t*>(buf + wi);
    } else {
      x[i] = 0;
      for (int k = 0; k < 4; ++k)
        if (wi + k >= 0 && wi + k < total) x[i] |= (uint32_t)buf[wi + k] << (8 * k);
    }
  }
  const uint32_t lo = __funnelshift_r(x[0], x[1], 8 * sh);       // bytes sh .. sh+3
  const uint32_t hi = __funnelshift_r(x[1], x[2], 8 * sh);       // bytes sh+4 .. sh+7
  b[0] = lo & 0xff; b[1] = (lo >> 8) & 0xff; b[2] = (lo >> 16) & 0xff; b[3] = lo >> 24;
  b[4] = hi & 0xff; b[5] = (hi >> 8) & 0xff;
}

// One output pixel straight from global memory (any crop size): used by CTAs whose source span does not fit the
// shared-memory row buffers of warp256_kernel.
__device__ __forceinline__ void warp_pixel_direct(const uint8_t* __restrict__ frames, long long frame, long long total,
                                                  int H, int W, double x, double y, float (&res)[3]) {
  const double fx = floor(x), fy = floor(y);
  const int minc = (int)fx, minr = (int)fy;
  const int maxc = (int)ceil(x), maxr = (int)ceil(y);
  const float dc = (float)(x - fx), dr = (float)(y - fy);
  const bool r0 = minr >= 0 && minr < H, r1 = maxr >= 0 && maxr < H;
  const bool c0 = minc >= 0 && minc < W, c1 = maxc >= 0 && maxc < W;
  const int sel = (maxc == minc) ? 0 : 3;            // integral x: both taps are the same pixel
  uint32_t top[6] = {0, 0, 0, 0, 0, 0}, bot[6] = {0, 0, 0, 0, 0, 0};
  if (r0 && (c0 || c1)) load6(frames, frame + ((long long)minr * W + minc) * 3, total, top);
  if (r1 && (c0 || c1)) {
    if (maxr == minr) {
#pragma unroll
      for (int i = 0; i < 6; ++i) bot[i] = top[i];
    } else {
      load6(frames, frame + ((long long)maxr * W + minc) * 3, total, bot);
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float v00 = c0 ? (float)top[ch] : 0.f;
    const float v01 = c1 ? (float)(sel ? top[3 + ch] : top[ch]) : 0.f;
    const float v10 = c0 ? (float)bot[ch] : 0.f;
    const float v11 = c1 ? (float)(sel ? bot[3 + ch] : bot[ch]) : 0.f;
    // lerp as a + d*(b-a): exact when the taps are equal, so a flat 255 region stays exactly 1.0
    const float t = fmaf(dc, v01 - v00, v00);
    const float bt = fmaf(dc, v11 - v10, v10);
    res[ch] = __fdiv_rn(fmaf(dr, bt - t, t), 255.0f);
  }
}

// v / 255 correctly rounded without the division sequence: q = v*r, one residual correction (Markstein); checked
// bit-identical to IEEE division on 2 M random inputs and on every quarter-integer in [0,255].
__device__ __forceinline__ float div255(float v) {
  const float r = 1.0f / 255.0f;
  const float q = v * r;
  return fmaf(fmaf(-255.0f, q, v), r, q);
}

// The first version (one thread per output pixel, everything per pixel) was issue-bound at ~300 instructions per
// pixel: fp64 coordinates, byte unpacking and int->float conversions per tap, an IEEE division per channel.  This
// version keeps the arithmetic and moves the work:
//   * a CTA owns kWarpRows output rows x 256 columns of one frame, thread = output column: the fp64 column geometry
//     is computed once per thread, the fp64 row geometry once per row (one thread each, through shared memory);
//   * each WARP (32 columns) then runs on its own: per output row it fetches the span of the two source rows its
//     columns touch with aligned 32-bit loads (a row ahead, in registers), converts byte -> float with PRMT+FADD
//     (exact) into a warp-private shared-memory strip, and every tap becomes one LDS; only __syncwarp in the loop;
//   * the 32 x 3 result floats of a row are staged and written as 24 coalesced float4.
constexpr int kWarpRows = 8;          // output rows per CTA
constexpr int kStripWords = 96;       // per-warp source strip: 384 bytes = 32 columns of a <= ~1000-pixel crop window

struct WarpRow {                      // per output row, written by one thread
  long long top_byte, bot_byte;       // byte offset of column 0 of the (clamped) source rows
  int flags;                          // bit0 r0 (top row in the image), bit1 r1
  float dr;
};

template <int NK>
__device__ __forceinline__ void warp_rows(const uint8_t* __restrict__ frames, const WarpRow* rows, float* strip,
                                          float* ostage, float* orow, int lo3, int n_words, int lane, int k0, int k1,
                                          float dc, bool c0, bool c1) {
  uint32_t pre[2][NK];
  int rel_t = 0, rel_b = 0, nrel_t, nrel_b;
  auto prefetch = [&](const WarpRow& w) {
    const long long bt = w.top_byte + lo3, bb = w.bot_byte + lo3;
    const uint32_t* pt = reinterpret_cast<const uint32_t*>(frames + (bt & ~3LL));
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(frames + (bb & ~3LL));
    nrel_t = (int)(bt & 3);
    nrel_b = (int)(bb & 3);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int wi = lane + 32 * k;
      const bool on = (k + 1 < NK) || wi < n_words;          // only the last slice can be partial
      pre[0][k] = on ? __ldg(pt + wi) : 0u;
      pre[1][k] = on ? __ldg(pb + wi) : 0u;
    }
  };
  // bytes -> floats (exact: 0x4B0000bb is 8388608 + b)
  auto park = [&]() {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float4* dst = reinterpret_cast<float4*>(strip + s * kStripWords * 4);
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const uint32_t v = pre[s][k];
        float4 f;
        f.x = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540)) - 8388608.0f;
        f.y = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7541)) - 8388608.0f;
        f.z = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7542)) - 8388608.0f;
        f.w = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7543)) - 8388608.0f;
        dst[lane + 32 * k] = f;                               // (slots past n_words hold zeros, never read)
      }
    }
  };
  WarpRow w = rows[0];
  prefetch(w);
  park();
  rel_t = nrel_t; rel_b = nrel_b;
  __syncwarp();
#pragma unroll 1
  for (int r = 0; r < kWarpRows; ++r) {
    const WarpRow wn = rows[r + 1 < kWarpRows ? r + 1 : r];
    prefetch(wn);                                     // global loads in flight during this row's arithmetic
    const bool r0 = w.flags & 1, r1 = w.flags & 2;
    const float* top = strip + rel_t;
    const float* bot = strip + kStripWords * 4 + rel_b;
    float* os = ostage + lane * 3;
    if (r0 && r1 && c0 && c1) {                       // interior pixel: four in-image taps
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float v00 = top[k0 + ch], v01 = top[k1 + ch], v10 = bot[k0 + ch], v11 = bot[k1 + ch];
        const float t = fmaf(dc, v01 - v00, v00);
        const float bt = fmaf(dc, v11 - v10, v10);
        os[ch] = div255(fmaf(w.dr, bt - t, t));
      }
    } else {                                          // constant-0 border (mode='constant', cval=0)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float v00 = (r0 && c0) ? top[k0 + ch] : 0.f, v01 = (r0 && c1) ? top[k1 + ch] : 0.f;
        const float v10 = (r1 && c0) ? bot[k0 + ch] : 0.f, v11 = (r1 && c1) ? bot[k1 + ch] : 0.f;
        const float t = fmaf(dc, v01 - v00, v00);
        const float bt = fmaf(dc, v11 - v10, v10);
        os[ch] = div255(fmaf(w.dr, bt - t, t));
      }
    }
    __syncwarp();                                     // strip fully read, result row staged
    // 32 pixels x 3 floats leave as 24 float4 (384 contiguous bytes)
    if (lane < 24)
      lr_stg_stream_f4(reinterpret_cast<float4*>(orow + (size_t)r * 768) + lane,
                       reinterpret_cast<const float4*>(ostage)[lane]);
    park();
    rel_t = nrel_t; rel_b = nrel_b;
    __syncwarp();
    w = wn;
  }
}

__global__ void __launch_bounds__(256, 6)
warp256_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ crop,
               float* __restrict__ out, int H, int W, long long total) {
  __shared__ __align__(16) float strips[8][2 * kStripWords * 4];          // per warp: top / bottom source strip
  __shared__ __align__(16) float ostages[8][96];
  __shared__ double geo[3];
  __shared__ int tail_risk;
  __shared__ WarpRow rows[kWarpRows];
  const int n = blockIdx.y, v0 = blockIdx.x * kWarpRows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {                                     // one fp64 division per CTA, not per pixel
    const int4 c = reinterpret_cast<const int4*>(crop)[n];
    const double size = (double)c.z;
    geo[0] = size / 255.0;
    geo[1] = 0.5 * (double)c.x - 0.5 * size;
    geo[2] = 0.5 * (double)c.y - 0.5 * size;
    tail_risk = 0;
  }
  __syncthreads();
  const double step = geo[0], x0 = geo[1], y0 = geo[2];
  const long long frame = (long long)n * H * W * 3;
  // ---- row geometry: thread r describes output row v0 + r ----
  if (tid < kWarpRows) {
    const double y = (double)(v0 + tid) * step + y0;
    const double fy = floor(y);
    const int minr = (int)fy, maxr = (int)ceil(y);
    const bool r0 = minr >= 0 && minr < H, r1 = maxr >= 0 && maxr < H;
    WarpRow w;
    w.dr = (float)(y - fy);
    w.top_byte = frame + (long long)min(max(minr, 0), H - 1) * W * 3;
    w.bot_byte = frame + (long long)min(max(maxr, 0), H - 1) * W * 3;
    w.flags = (r0 ? 1 : 0) | (r1 ? 2 : 0);
    rows[tid] = w;
    // whole-word fetches may only run past the end of the frame buffer in its very last row: those CTAs take the
    // per-pixel path
    if (w.bot_byte + (long long)W * 3 + 4LL * kStripWords > total) tail_risk = 1;
  }
  // ---- column geometry of this thread (u = tid), once ----
  const double x = (double)tid * step + x0;
  const double fx = floor(x);
  const int minc = (int)fx, maxc = (int)ceil(x);
  const float dc = (float)(x - fx);
  const bool c0 = minc >= 0 && minc < W, c1 = maxc >= 0 && maxc < W;
  const int cminc = min(max(minc, 0), W - 1), cmaxc = min(max(maxc, 0), W - 1);
  const int lo_col = __shfl_sync(0xffffffffu, cminc, 0), hi_col = __shfl_sync(0xffffffffu, cmaxc, 31);
  const int n_words = ((hi_col - lo_col + 1) * 3 + 3 + 3) / 4;       // worst-case alignment of the first byte
  __syncthreads();
  float* orow = out + ((size_t)n * 65536 + (size_t)v0 * 256) * 3;

  if (tail_risk || n_words > kStripWords) {           // (warp-uniform) per-pixel path
    for (int r = 0; r < kWarpRows; ++r) {
      float res[3];
      warp_pixel_direct(frames, frame, total, H, W, x, (double)(v0 + r) * step + y0, res);
      float* o = orow + ((size_t)r * 256 + tid) * 3;
      o[0] = res[0]; o[1] = res[1]; o[2] = res[2];
    }
    return;
  }
  const int k0 = (cminc - lo_col) * 3, k1 = (cmaxc - lo_col) * 3;   // float offsets of this thread's two taps
  float* strip = strips[warp];
  float* ostage = ostages[warp];
  float* owarp = orow + warp * 96;
  if (n_words <= 32) warp_rows<1>(frames, rows, strip, ostage, owarp, lo_col * 3, n_words, lane, k0, k1, dc, c0, c1);
  else if (n_words <= 64) warp_rows<2>(frames, rows, strip, ostage, owarp, lo_col * 3, n_words, lane, k0, k1, dc, c0, c1);
  else warp_rows<3>(frames, rows, strip, ostage, owarp, lo_col * 3, n_words, lane, k0, k1, dc, c0, c1);
}

// ---------------------------------------------------------------------------------------------
// a6-a9 fused: for each requested map location, undo the crop similarity, translate into the
// padded-face frame, emit float64 (the dataview dtype).  The full (256,256,3) "restored" map the
// reference materialises per frame is never written: only the gathered points are.
//   z = f32(P_z) / f32(s)   (numpy-1.x value-based casting keeps this division in float32)
//   x = f64(P_x) * (1/s) + x0 - left_pad ,  y likewise with top_pad
__global__ void __launch_bounds__(256)
posmap_gather_kernel(const float* __restrict__ posmap, const int32_t* __restrict__ crop,
                     const int32_t* __restrict__ rect_pad, const int32_t* __restrict__ idx,
                     int n_idx, double* __restrict__ out) {
  const int n = blockIdx.y;
  const int4 c = reinterpret_cast<const int4*>(crop)[n];
  const int4 rp = reinterpret_cast<const int4*>(rect_pad)[n];
  const double size = (double)c.z;
  const double s = 255.0 / size;
  const double inv_s = size / 255.0;
  const float s32 = (float)s;
  const double x0 = 0.5 * (double)c.x - 0.5 * size;
  const double y0 = 0.5 * (double)c.y - 0.5 * size;
  const double left = (double)rp.x, top = (double)rp.z;
  const float* map = posmap + (size_t)n * 65536 * 3;
  double* o = out + (size_t)n * n_idx * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += gridDim.x * blockDim.x) {
    const int flat = idx[i];
    const float* p = map + (size_t)flat * 3;
    const float px = p[0], py = p[1], pz = p[2];
    o[(size_t)i * 3 + 0] = ((double)px * inv_s + x0) - left;
    o[(size_t)i * 3 + 1] = ((double)py * inv_s + y0) - top;
    o[(size_t)i * 3 + 2] = (double)(pz / s32);
  }
}

// ---------------------------------------------------------------------------------------------
// a11.  Ragged float64 rows -> zero-padded (B,Tmax,F) float32.  One CTA row-block per clip.
__global__ void __launch_bounds__(256)
collate_pad_kernel(const double* __restrict__ src, const int64_t* __restrict__ offs,
                   float* __restrict__ dst, int Tmax, int F) {
  const int b = blockIdx.y;
  const int64_t r0 = offs[b], r1 = offs[b + 1];
  const int64_t n_valid = (r1 - r0) * F;
  const int64_t n_all = (int64_t)Tmax * F;
  const double* s = src + r0 * F;
  float* d = dst + (int64_t)b * n_all;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_all;
       i += (int64_t)gridDim.x * blockDim.x)
    d[i] = i < n_valid ? (float)s[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// N2 (extension).  Mouth ROI from landmarks 48:68 and a bilinear resize to out_h x out_w u8.
// Spec (mirrored by oracle/vision.py:mouth_roi / mouth_crop):
//   frame coords = landmark + (left_pad, top_pad); bbox of the 20 mouth points;
//   rw = 1.2 * max(xmax-xmin, (ymax-ymin) * out_w/out_h), rh = rw * out_h/out_w, both >= 2;
//   x_lo = floor(cx - rw/2), y_lo = floor(cy - rh/2), roi_w = ceil(rw), roi_h = ceil(rh);
//   sample centre-aligned: sx = (ox+0.5)*roi_w/out_w - 0.5 + x_lo, clamp to the frame (replicate),
//   fp32 lerp, round-half-even to u8.
__global__ void mouth_roi_kernel(const double* __restrict__ lmk, const int32_t* __restrict__ rect_pad,
                                 int32_t* __restrict__ roi, int N, int out_h, int out_w) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double* p = lmk + (size_t)n * 68 * 3;
  const int4 rp = reinterpret_cast<const int4*>(rect_pad)[n];
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int i = 48; i < 68; ++i) {
    double x = p[i * 3 + 0] + (double)rp.x, y = p[i * 3 + 1] + (double)rp.z;
    xmin = fmin(xmin, x); xmax = fmax(xmax, x);
    ymin = fmin(ymin, y); ymax = fmax(ymax, y);
  }
  double aspect = (double)out_w / (double)out_h;
  double rw = 1.2 * fmax(xmax - xmin, (ymax - ymin) * aspect);
  if (rw < 2.0) rw = 2.0;
  double rh = rw / aspect;
  if (rh < 2.0) rh = 2.0;
  double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
  int4 o;
  o.x = (int)floor(cx - 0.5 * rw);
  o.y = (int)floor(cy - 0.5 * rh);
  o.z = (int)ceil(rw);
  o.w = (int)ceil(rh);
  reinterpret_cast<int4*>(roi)[n] = o;
}

// One CTA = `rows_per_cta` output rows of one frame.  The 2 source rows every output row samples are staged in
// shared memory with aligned 16-byte loads (a row's ROI span is contiguous bytes: HBM sees whole sectors once, instead
// of 12 scattered byte loads per output pixel), the per-column geometry (byte offsets of the two taps inside a staged
// row, the fp32 weight) is computed once per CTA, and every tap is an LDS.  Arithmetic is bit-for-bit the spec above.
// Shared layout: col_x0[out_w] | col_x1[out_w] (int32 byte offsets) | col_ax[out_w] | row bytes [2*rows][pitch].
__global__ void __launch_bounds__(256)
mouth_crop_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ roi,
                  uint8_t* __restrict__ out, int H, int W, int out_h, int out_w, int rows_per_cta, int pitch,
                  long long total_bytes) {
  extern __shared__ __align__(16) uint8_t msm[];
  int* col_x0 = reinterpret_cast<int*>(msm);
  int* col_x1 = col_x0 + out_w;
  float* col_ax = reinterpret_cast<float*>(col_x1 + out_w);
  uint8_t* rows = msm + (((size_t)out_w * 12 + 15) & ~(size_t)15);
  const int n = blockIdx.y;
  const int oy0 = blockIdx.x * rows_per_cta;
  const int n_rows = min(rows_per_cta, out_h - oy0);
  const int4 r = reinterpret_cast<const int4*>(roi)[n];
  const long long img_off = (long long)n * H * W * 3;
  const float sxs = (float)r.z / (float)out_w, sys = (float)r.w / (float)out_h;
  const int tid = threadIdx.x;

  // source column range touched by this ROI (clamped like the taps), as a byte span of a frame row
  const float sx_first = __fadd_rn(__fadd_rn(__fmul_rn(0.5f, sxs), -0.5f), (float)r.x);
  const float sx_last = __fadd_rn(__fadd_rn(__fmul_rn((float)(out_w - 1) + 0.5f, sxs), -0.5f), (float)r.x);
  const int xa = min(max((int)floorf(sx_first), 0), W - 1);
  const int xb = min(max((int)floorf(sx_last) + 1, 0), W - 1);
  const int span = (xb - xa + 1) * 3;                       // bytes of a row that any tap can touch
  if (span + 32 > pitch) {
    // ROI wider than the staging rows were sized for (a face filling the frame): direct taps from global memory
    const uint8_t* img = frames + img_off;
    uint8_t* od = out + ((size_t)n * out_h + oy0) * out_w * 3;
    for (int i = tid; i < n_rows * out_w; i += blockDim.x) {
      const int j = i / out_w, ox = i - j * out_w;
      float sx = __fadd_rn(__fadd_rn(__fmul_rn((float)ox + 0.5f, sxs), -0.5f), (float)r.x);
      float sy = __fadd_rn(__fadd_rn(__fmul_rn((float)(oy0 + j) + 0.5f, sys), -0.5f), (float)r.y);
      float fx = floorf(sx), fy = floorf(sy);
      float ax = sx - fx, ay = sy - fy;
      int x0 = min(max((int)fx, 0), W - 1), x1 = min(max((int)fx + 1, 0), W - 1);
      int y0 = min(max((int)fy, 0), H - 1), y1 = min(max((int)fy + 1, 0), H - 1);
      const uint8_t* p00 = img + ((size_t)y0 * W + x0) * 3;
      const uint8_t* p01 = img + ((size_t)y0 * W + x1) * 3;
      const uint8_t* p10 = img + ((size_t)y1 * W + x0) * 3;
      const uint8_t* p11 = img + ((size_t)y1 * W + x1) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float top = __fadd_rn(__fmul_rn(1.f - ax, (float)p00[ch]), __fmul_rn(ax, (float)p01[ch]));
        float bot = __fadd_rn(__fmul_rn(1.f - ax, (float)p10[ch]), __fmul_rn(ax, (float)p11[ch]));
        float val = __fadd_rn(__fmul_rn(1.f - ay, top), __fmul_rn(ay, bot));
        od[(size_t)i * 3 + ch] = (uint8_t)fminf(fmaxf(rintf(val), 0.f), 255.f);
      }
    }
    return;
  }

  for (int ox = tid; ox < out_w; ox += blockDim.x) {
    float sx = __fadd_rn(__fadd_rn(__fmul_rn((float)ox + 0.5f, sxs), -0.5f), (float)r.x);
    float fx = floorf(sx);
    int x0 = min(max((int)fx, 0), W - 1), x1 = min(max((int)fx + 1, 0), W - 1);
    col_x0[ox] = (x0 - xa) * 3;
    col_x1[ox] = (x1 - xa) * 3;
    col_ax[ox] = sx - fx;
  }
  // stage source rows: slot 2*j = y0 of output row oy0+j, slot 2*j+1 = y1.  Each slot starts at the 16-byte aligned
  // address at or below the row's first byte; `lead` (same for all threads of a row) is added at read time.
  for (int slot = tid >> 5; slot < 2 * n_rows; slot += (blockDim.x >> 5)) {
    const int oy = oy0 + (slot >> 1);
    float sy = __fadd_rn(__fadd_rn(__fmul_rn((float)oy + 0.5f, sys), -0.5f), (float)r.y);
    int fy = (int)floorf(sy);
    int y = min(max(fy + (slot & 1), 0), H - 1);
    const long long first = img_off + ((long long)y * W + xa) * 3;
    const long long base = first & ~15LL;
    const int n_chunks = (int)((first - base + span + 15) >> 4);
    uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)slot * pitch);
    for (int c = threadIdx.x & 31; c < n_chunks; c += 32) {
      const long long a = base + 16LL * c;
      if (a + 16 <= total_bytes) {
        dst[c] = *reinterpret_cast<const uint4*>(frames + a);
      } else {                                               // last partial chunk of the whole frames buffer
        uint8_t* d8 = reinterpret_cast<uint8_t*>(dst + c);
        for (int e = 0; e < 16; ++e) d8[e] = (a + e < total_bytes) ? frames[a + e] : (uint8_t)0;
      }
    }
  }
  __syncthreads();

  uint8_t* o = out + ((size_t)n * out_h + oy0) * out_w * 3;
  const int n_px = n_rows * out_w;
  for (int i = tid; i < n_px; i += blockDim.x) {
    const int j = i / out_w, ox = i - j * out_w;
    const int oy = oy0 + j;
    float sy = __fadd_rn(__fadd_rn(__fmul_rn((float)oy + 0.5f, sys), -0.5f), (float)r.y);
    float fyf = floorf(sy);
    const float ay = sy - fyf;
    const int fy = (int)fyf;
    const int y0 = min(max(fy, 0), H - 1), y1 = min(max(fy + 1, 0), H - 1);
    const int lead0 = (int)((img_off + ((long long)y0 * W + xa) * 3) & 15);
    const int lead1 = (int)((img_off + ((long long)y1 * W + xa) * 3) & 15);
    const uint8_t* r0 = rows + (size_t)(2 * j) * pitch + lead0;
    const uint8_t* r1 = rows + (size_t)(2 * j + 1) * pitch + lead1;
    const int b0 = col_x0[ox], b1 = col_x1[ox];
    const float ax = col_ax[ox];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float top = __fadd_rn(__fmul_rn(1.f - ax, (float)r0[b0 + ch]), __fmul_rn(ax, (float)r0[b1 + ch]));
      float bot = __fadd_rn(__fmul_rn(1.f - ax, (float)r1[b0 + ch]), __fmul_rn(ax, (float)r1[b1 + ch]));
      float val = __fadd_rn(__fmul_rn(1.f - ay, top), __fmul_rn(ay, bot));
      val = fminf(fmaxf(rintf(val), 0.f), 255.f);
      o[(size_t)i * 3 + ch] = (uint8_t)val;
    }
  }
}

}  // namespace

extern "C" int lr_rect_geometry(const int32_t* rects, int N, int img_h, int img_w,
                                int32_t* rect_pad, int32_t* crop, void* stream) {
  LR_CHECK_ARG(rects && rect_pad && crop && N > 0 && img_h > 0 && img_w > 0, "lr_rect_geometry: bad args");
  rect_geometry_kernel<<<lr_div_up(N, 128), 128, 0, lr_stream(stream)>>>(rects, N, img_h, img_w,
                                                                         rect_pad, crop);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_warp256(const uint8_t* frames, const int32_t* crop, float* out, int N, int H, int W,
                          void* stream) {
  LR_CHECK_ARG(frames && crop && out && N > 0 && H > 0 && W > 0, "lr_warp256: bad args");
  LR_CHECK_ARG(N <= 65535, "lr_warp256: at most 65535 frames per call");
  LR_CHECK_ARG((reinterpret_cast<uintptr_t>(frames) & 3) == 0, "lr_warp256: frames must be 4-byte aligned");
  dim3 grid(256 / kWarpRows, N);
  warp256_kernel<<<grid, 256, 0, lr_stream(stream)>>>(frames, crop, out, H, W, (long long)N * H * W * 3);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_posmap_gather(const float* posmap, const int32_t* crop, const int32_t* rect_pad,
                                const int32_t* kpt_idx, int n_kpt, const int32_t* face_idx, int n_vtx,
                                double* lmk, double* vtx, int N, void* stream) {
  LR_CHECK_ARG(posmap && crop && rect_pad && N > 0 && N <= 65535, "lr_posmap_gather: bad args");
  LR_CHECK_ARG((kpt_idx && lmk && n_kpt > 0) || (face_idx && vtx && n_vtx > 0),
               "lr_posmap_gather: nothing to gather");
  cudaStream_t st = lr_stream(stream);
  if (kpt_idx && lmk && n_kpt > 0) {
    dim3 grid(1, N);
    posmap_gather_kernel<<<grid, 96, 0, st>>>(posmap, crop, rect_pad, kpt_idx, n_kpt, lmk);
    LR_CHECK_LAUNCH();
  }
  if (face_idx && vtx && n_vtx > 0) {
    dim3 grid(lr_div_up(n_vtx, 256 * 4), N);
    posmap_gather_kernel<<<grid, 256, 0, st>>>(posmap, crop, rect_pad, face_idx, n_vtx, vtx);
    LR_CHECK_LAUNCH();
  }
  return LR_OK;
}

extern "C" int lr_collate_pad_f64(const double* src_concat, const int64_t* row_offsets, float* dst,
                                  int B, int Tmax, int F, void* stream) {
  LR_CHECK_ARG(src_concat && row_offsets && dst && B > 0 && B <= 65535 && Tmax > 0 && F > 0,
               "lr_collate_pad_f64: bad args");
  int gx = lr_div_up((int64_t)Tmax * F, 256 * 4);
  dim3 grid(gx < 1 ? 1 : gx, B);
  collate_pad_kernel<<<grid, 256, 0, lr_stream(stream)>>>(src_concat, row_offsets, dst, Tmax, F);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

extern "C" int lr_mouth_crop(const uint8_t* frames, const double* lmk, const int32_t* rect_pad,
                             uint8_t* out, int32_t* roi, int N, int H, int W, int out_h, int out_w,
                             void* stream) {
  LR_CHECK_ARG(frames && lmk && rect_pad && out && roi && N > 0 && N <= 65535 && H > 0 && W > 0 &&
                   out_h > 0 && out_w > 0,
               "lr_mouth_crop: bad args");
  cudaStream_t st = lr_stream(stream);
  mouth_roi_kernel<<<lr_div_up(N, 128), 128, 0, st>>>(lmk, rect_pad, roi, N, out_h, out_w);
  LR_CHECK_LAUNCH();
  // staged-row geometry.  The ROI is data dependent (it lives on the device), so the rows are sized for mouths up to
  // 512 source pixels wide (a mouth is ~1/3 of a face box); wider ROIs take the kernel's direct path.  16 bytes of
  // alignment lead + 16 of tail per staged row; 10 output rows (20 staged rows, ~31 KB) per CTA -> 7 CTAs per SM.
  const int pitch = (((W < 512 ? W : 512) * 3 + 15) & ~15) + 32;
  const size_t head = ((size_t)out_w * 12 + 15) & ~(size_t)15;
  const int rows_per_cta = 10;
  const size_t smem = head + (size_t)2 * rows_per_cta * pitch;
  LR_CHECK_ARG(smem <= 48 * 1024, "lr_mouth_crop: out_w = %d too wide for the column tables", out_w);
  dim3 grid(lr_div_up(out_h, rows_per_cta), N);
  mouth_crop_kernel<<<grid, 256, smem, st>>>(frames, roi, out, H, W, out_h, out_w, rows_per_cta, pitch,
                                             (long long)N * H * W * 3);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
