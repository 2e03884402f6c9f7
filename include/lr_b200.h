/*
 * lr_b200.h — C-ABI of liblr_b200.so: the B200 (sm_100a) hot path of joseph-zhong/LipReading.
 *
 * The reference has no FFI of its own (it is pure Python calling third-party wheels), so the
 * boundary is the set of Python call sites listed per entry point below ("replaces: file:line",
 * paths relative to the reference checkout).  INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add at each site.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless the name ends in _host;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates,
 *     never synchronises, never throws; returns 0 or a negative LR_E* code, text in lr_last_error();
 *   - tensors are dense row-major with the shapes given in the comments;
 *   - re-entrant across streams and devices (one process per GPU under data parallelism).
 */
#ifndef LR_B200_H_
#define LR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LR_OK            0
#define LR_EINVAL       -1   /* bad argument / unsupported shape          */
#define LR_ECUDA        -2   /* CUDA runtime / driver error at launch     */
#define LR_EWORKSPACE   -3   /* workspace too small                       */
#define LR_EARCH        -4   /* device is not sm_100 (tcgen05 paths)      */

#define LR_RNN_TANH 0
#define LR_RNN_GRU  1
#define LR_RNN_LSTM 2

#define LR_F32  0
#define LR_BF16 1

/* -------- library ---------------------------------------------------------------------- */
int         lr_abi_version(void);
const char* lr_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t    lr_launch_count(void);

/* -------- a15: CTC alpha/beta + gradient ----------------------------------------------- */
/* replaces: src/train/ctc_loss.py:85,100 (F.ctc_loss(...cpu()) and its autograd backward).
 * log_probs (B,T,C) f32 batch-first (the layout VideoEncoder emits; no transpose, no D2H),
 * targets (B,Lmax) i32 CTC classes (label+1, blank=0), input_lens/target_lens (B) i32.
 * nll (B) f32 = -log p(target|input), +inf when infeasible.
 * grad (B,T,C) f32 or NULL: d nll_b / d log_probs with torch's native-CTC convention
 *   exp(lp) - exp(log(alpha*beta summed per class) + nll - lp), zero for t >= input_len.     */
/* `kernel` (per call, same value for lr_ctc_workspace and lr_ctc_fwd_bwd):
 * 0 (default): warp-per-clip kernels for large batches (>= 1024 clips) whose lattice fits shared
 * memory, else CTA-per-clip.  Among the warp kernels, labels of <= 31 symbols run the linear-space
 * (per-frame rescaled) recursion first and the log-space kernel only on the clips that one flags
 * (underflow / infeasible).  1: always CTA-per-clip; 2: log-space warp-per-clip whenever it fits;
 * 3: warp-per-clip whenever it fits, linear-space first (2, 3: test hooks).                          */
size_t lr_ctc_workspace(int B, int T, int C, int Lmax, int kernel);
int lr_ctc_fwd_bwd(const float* log_probs, const int32_t* targets, const int32_t* input_lens,
                   const int32_t* target_lens, int B, int T, int C, int Lmax,
                   float* nll, float* grad, void* workspace, size_t ws_bytes, int kernel, void* stream);
/* Greedy CTC decode for the inference stream (semantics of src/models/lipreader/decoder.py:165-197):
 * arg-max per frame, collapse repeats, drop blank.  tokens (B,T) i32 zero padded, out_lens (B).   */
int lr_ctc_greedy_decode(const float* log_probs, const int32_t* lens, int B, int T, int C,
                         int32_t* tokens, int32_t* out_lens, void* stream);
/* out[b,:,:] = in[b,:,:] * scale[b]  (chain rule for the per-sample upstream gradient).      */
int lr_scale_rows(const float* in, const float* scale, float* out, int B, int64_t row_elems,
                  void* stream);

/* -------- a14: output_proj + masked log-softmax ------------------------------------------ */
/* replaces: src/models/lipreader/better_model.py:92-93 (nn.Linear + allennlp masked_log_softmax).
 * hidden (M,K) f32, weight (C,K) f32, bias (C) f32, log_mask (C) f32 additive term
 * (log(mask+1e-45)); out log_probs (M,C) f32.                                                */
/* `variant` (per call): 0 = fp32 SIMT kernel (the parity path); 1 = 3xTF32 tensor-core kernel (mma.sync, W resident
 * in shared memory) whenever C <= 68, K % 16 == 0 and K <= 688 — 1.2x faster, 3e-5 instead of 1e-5 from the exact
 * logits (legacy TF32 mma.sync on sm_100: slow, and its accumulator adds do not round to nearest);
 * 2 = tcgen05 kind::tf32 kernel (TMA ring, TMEM accumulator, register-local row softmax): the throughput path,
 * TF32 products (1e-3 relative on the logits), needs lr_proj_tc5_supported(M, K, C).                             */
int lr_proj_logsoftmax_fwd(const float* hidden, const float* weight, const float* bias,
                           const float* log_mask, float* log_probs, int M, int K, int C,
                           int variant, void* stream);
int lr_proj_tc5_supported(int M, int K, int C);
/* The log-softmax half of the backward alone: d_logits (M,C) = g - softmax*sum(g), d_bias (C) = its column sums. */
int lr_logsoftmax_bwd(const float* grad_lp, const float* log_probs, float* d_logits, float* d_bias, int M, int C,
                      void* stream);
/* grad_lp (M,C) upstream; writes d_logits (M,C) = g - softmax*sum(g), d_bias (C) (zeroed by the
 * call), and d_hidden (M,K) = d_logits @ weight, d_weight (C,K) = d_logits^T @ hidden (zeroed by
 * the call).                                                                                 */
int lr_proj_logsoftmax_bwd(const float* grad_lp, const float* log_probs, const float* hidden,
                           const float* weight, float* d_logits, float* d_hidden,
                           float* d_weight, float* d_bias, int M, int K, int C, void* stream);

/* -------- a12/a13: (bi)directional recurrent layer, packed-sequence semantics ------------ */
/* replaces: src/models/lipreader/better_model.py:64-89 (sort -> pack -> nn.{LSTM,GRU,RNN} ->
 * unpack -> unsort).  gi (B,T,D,G*H) f32 = x @ W_ih^T + b_ih precomputed by the caller (one
 * GEMM for all T and both directions), w_hh (D,G*H,H), b_hh (D,G*H), lens (B) i32.
 * out hidden (B,T,D*H) zero beyond each length; h_n/c_n (D,B,H) state at each sample's own last
 * valid frame; `saved` (B,T,D,S*H) activations kept for backward (S = lr_rnn_saved_per_unit).  */
int    lr_rnn_saved_per_unit(int mode);
size_t lr_rnn_workspace(int mode, int B, int T, int H, int D);
int lr_rnn_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh,
               const int32_t* lens, int B, int T, int H, int D,
               float* hidden, float* h_n, float* c_n, float* saved,
               void* workspace, size_t ws_bytes, void* stream);
/* d_hidden (B,T,D*H), d_h_n/d_c_n (D,B,H) or NULL -> d_gi (B,T,D,G*H), d_gh (B,T,D,G*H)
 * (gradient w.r.t. h@W_hh^T+b_hh pre-activations; the caller forms d_w_hh = d_gh^T @ h_prev).
 * h_prev_all (B,T,D,H) is also written (the h that entered each step) for that GEMM.          */
int lr_rnn_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
               const float* saved, const float* hidden, const float* w_hh, const int32_t* lens,
               int B, int T, int H, int D, float* d_gi, float* d_gh, float* h_prev_all,
               void* workspace, size_t ws_bytes, void* stream);

/* Throughput path of the same layer: ONE persistent launch per pass, 8-CTA clusters keep the W_hh
 * slices resident in shared memory as bf16, per-step state exchange over distributed shared memory
 * (csrc/rnn_cluster.cu).  Same tensors as lr_rnn_fwd / lr_rnn_bwd (w_hh un-transposed for both);
 * bf16 operands, fp32 accumulation and state.  lr_rnn_cluster_supported: H % 128 == 0 and the
 * slices fit shared memory.                                                                    */
int lr_rnn_cluster_supported(int mode, int H);
int lr_rnn_cluster_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh,
                       const int32_t* lens, int B, int T, int H, int D, float* hidden, float* h_n,
                       float* c_n, float* saved, void* stream);
int lr_rnn_cluster_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
                       const float* saved, const float* hidden, const float* w_hh,
                       const int32_t* lens, int B, int T, int H, int D, float* d_gi, float* d_gh,
                       float* h_prev_all, void* stream);

/* Hidden sizes whose W_hh does not fit a cluster's shared memory (the reference's BiLSTM-768,
 * config/archive/experiments/ecd/*): the same layer as ONE cooperative launch per pass — a direction's units are
 * spread over H/16 co-resident CTAs (W_hh slices resident as bf16), the per-step operand (state / gate gradients) is
 * exchanged through a double-buffered bf16 buffer in `workspace` (L2-resident) with one barrier per step among the
 * CTAs of a direction; 64 clips per pass (csrc/rnn_grid.cu).  Same tensors and semantics as lr_rnn_cluster_*;
 * workspace of lr_rnn_grid_workspace bytes (contents need not be initialised).
 * lr_rnn_grid_supported: H % 256 == 0, D*H/16 <= 148 CTAs, slices fit shared memory.                  */
int    lr_rnn_grid_supported(int mode, int H, int D);
size_t lr_rnn_grid_workspace(int mode, int H, int D);
int lr_rnn_grid_fwd(int mode, const float* gi, const float* w_hh, const float* b_hh,
                    const int32_t* lens, int B, int T, int H, int D, float* hidden, float* h_n,
                    float* c_n, float* saved, void* workspace, size_t ws_bytes, void* stream);
int lr_rnn_grid_bwd(int mode, const float* d_hidden, const float* d_h_n, const float* d_c_n,
                    const float* saved, const float* hidden, const float* w_hh,
                    const int32_t* lens, int B, int T, int H, int D, float* d_gi, float* d_gh,
                    float* h_prev_all, void* workspace, size_t ws_bytes, void* stream);

/* -------- a11: collate / pad -------------------------------------------------------------- */
/* replaces: src/data/data_loader.py:124-137 (_pad of ragged (T_i,68,3) f64 rows to (B,Tmax,F) f32).
 * src_concat f64 rows back to back, offsets (B+1) i64 in rows, dst (B,Tmax,F) f32 zero padded.  */
int lr_collate_pad_f64(const double* src_concat, const int64_t* row_offsets, float* dst,
                       int B, int Tmax, int F, void* stream);

/* -------- a2/a3: rectangle geometry (integer exact) --------------------------------------- */
/* replaces: src/utils/data/face.py:76-90 (_applyPadding, padding 0.3) and
 * src/models/face/prnet.py:112-119 (old_size, center, size=int(old_size*1.6)).
 * rects (N,4) i32 (left,right,top,bottom); out rect_pad (N,4) i32, crop (N,4) i32 = (2*cx,2*cy,
 * size, 0) with the centre doubled so it stays integral.                                      */
int lr_rect_geometry(const int32_t* rects, int N, int img_h, int img_w,
                     int32_t* rect_pad, int32_t* crop, void* stream);

/* -------- a4: /255 + bilinear similarity warp to 256x256 ----------------------------------- */
/* replaces: src/models/face/prnet.py:137-143 (estimate_transform + image/255. + skimage warp).
 * frames (N,H,W,3) u8, crop (N,4) i32 from lr_rect_geometry -> out (N,256,256,3) f32 in [0,1].  */
int lr_warp256(const uint8_t* frames, const int32_t* crop, float* out, int N, int H, int W,
               void* stream);

/* -------- a6-a9: position-map restore + landmark/vertex gather + translate ---------------- */
/* replaces: src/models/face/prnet.py:151-156,169,179-180 and src/utils/data/face.py:171-174.
 * posmap (N,256,256,3) f32 (CNN output * 281.6), crop/rect_pad from lr_rect_geometry,
 * kpt_idx (68) i32 flat indices row*256+col, face_idx (V) i32 or NULL.
 * out lmk (N,68,3) f64, vtx (N,V,3) f64 or NULL: x' = (u/s + tx') - left_pad etc.              */
int lr_posmap_gather(const float* posmap, const int32_t* crop, const int32_t* rect_pad,
                     const int32_t* kpt_idx, int n_kpt, const int32_t* face_idx, int n_vtx,
                     double* lmk, double* vtx, int N, void* stream);

/* -------- N2: mouth ROI crop + bilinear resize (extension, no reference code) -------------- */
/* lmk (N,68,3) f64 face-relative landmarks + rect_pad -> roi from landmarks 48:68
 * (face.py:21 `_mouth`), fixed aspect out_w:out_h, bilinear -> clips (N,out_h,out_w,3) u8.     */
int lr_mouth_crop(const uint8_t* frames, const double* lmk, const int32_t* rect_pad,
                  uint8_t* out, int32_t* roi, int N, int H, int W, int out_h, int out_w,
                  void* stream);

/* -------- N1: spatio-temporal conv front-end on tcgen05 ------------------------------------ */
/* extension behind VideoEncoder(frame_processing='conv3d'); oracle = torch.nn.functional.conv3d
 * fp32 (oracle/conv3d.py).  See lipreading_b200/csrc/conv3d_sm100.cu for the tile geometry.     */
int lr_conv3d_supported(void);
/* u8 NDHWC clip (B,T,H,W,3) -> /255 -> 2x2 space-to-depth -> zero-padded bf16 volume
 * (B,T+2,Hp,Wp,16), Hp >= H/2+2, interior at (1,1,1); the caller zero-fills the volume once.             */
int lr_clip_s2d(const uint8_t* clip, void* out_bf16, int B, int T, int H, int W, int Hp, int Wp,
                void* stream);
/* Stride-1 "same" conv k=(KT,KH,KW) as a shifted-window implicit GEMM on tcgen05.
 *   x : zero-padded channel-grouped bf16 volume [CG][B][T+KT-1][Hp][Wp][Cin], Hp >= H+KH-1 (a
 *       multiple of 128/Wp), Wp a power of two >= W+KW-1, Cin in {16,32,64} per group;
 *   w : weight tile images from lr_pack_conv_weights; bias f32 (Cout) or NULL; Cout in {32,64,96,128};
 *   epi_mode 0: bias + ReLU + MaxPool(1,2,2) -> bf16 written at offset (o_t,o_y,o_x) inside the
 *               output volume (B,oTp,oHp,oWp,Cout) [the next layer's padded input], plus one
 *               arg-max byte per pooled element (0..3, 4 = ReLU-dead) into argmax (may be NULL);
 *   epi_mode 1: plain bf16 store of the valid (t,y,x) positions (used for dgrad).
 *   J = accumulators (consecutive frames) per CTA work item, 0 = choose.
 *   swap (orientation) = 0: positions on M, N = Cout.  1: D^T = W . X^T (channels on the M lanes, the
 *   128 tile positions on N).  2 (epi_mode 1, Cout = 32, 2 <= KW <= 5): positions on M and the KW
 *   kx-taps of a filter row stacked on N = KW*Cout — one MMA per (kt,ky) instead of KW narrow ones
 *   (an N = 32 MMA is operand-fetch bound at 40 % of the pipe); the epilogue adds the KW column
 *   blocks with a row shift of kx each (warp shuffles + a 4-row shared-memory halo).
 *   3: positions on M and the KT kt-taps of a spatial tap stacked on N: input plane c times
 *   [W(kt=KT-1); ...; W(kt=0)] lands on the accumulators of the consecutive output frames c-KT+1 .. c
 *   (side by side in TMEM), N = KT*Cout (<= 256) per MMA, same epilogues as 0; `w` must come from
 *   lr_pack_conv_weights_kt.                                                                       */
int lr_conv3d_fwd(const void* x, const void* w, const float* bias, void* y, uint8_t* argmax,
                  int B, int T, int H, int W, int Hp, int Wp, int Cin, int CG, int Cout, int KT,
                  int KH, int KW, int epi_mode, int oTp, int oHp, int oWp, int o_t, int o_y,
                  int o_x, int J, int swap, void* stream);
/* dgrad of a conv layer fused with the backward of the ReLU + MaxPool(1,2,2) in front of it
 * (= lr_conv3d_fwd with epi_mode 1 followed by lr_unpool, in one pass): `dy`/`w` as for a dgrad call of
 * lr_conv3d_fwd (Cout = the layer's INPUT channels, 32 or 64); argmax (B,T,H,W,Cout) are the pooling
 * layer's arg-max bytes; the un-pooled gradient is written into the interior (o_t,o_y,o_x) of the
 * zero-padded volume dy_below (B,oTp,oHp,oWp,Cout) — only the 2x2 windows of pooled pixels, the rest
 * must already be zero — and d_bias (Cout) f32 (or NULL) receives its per-channel sum.           */
int lr_conv3d_dgrad_unpool(const void* dy, const void* w, const uint8_t* argmax, void* dy_below,
                           float* d_bias, int B, int T, int H, int W, int Hp, int Wp, int Cin, int CG,
                           int Cout, int KT, int KH, int KW, int oTp, int oHp, int oWp, int o_t,
                           int o_y, int o_x, int J, int swap, void* stream);
/* Backward of ReLU+MaxPool(1,2,2): d_pooled (B,T,H/2,W/2,C) bf16 + argmax -> gradient w.r.t. the
 * conv output, written into the interior (pt,ph,pw) of a zero-padded channel-grouped volume
 * [C/Cg][B][Tp][Hp][Wp][Cg] ready to be the `x` of a dgrad pass; d_bias (C) f32 or NULL receives
 * the per-channel sum (the conv bias gradient).  Only the 2x2 windows of the pooled pixels are
 * written: `out` must already be zero everywhere else (borders, and the last row/column of an odd
 * H or W) — allocate it zeroed once and reuse it.                                              */
/* bf16 [Cout][CG][taps][Cin] -> [CG][taps][Cout x Cin] tile images, 16-byte chunks pre-swizzled so a
 * stage of taps is ONE contiguous bulk copy into shared memory (what lr_conv3d_fwd expects as `w`). */
int lr_pack_conv_weights(const void* w, void* out, int Cout, int CG, int taps, int Cin, void* stream);
/* Same tile images in the order orientation 3 reads them: [CG][KH*KW][KT, kt descending][Cout x Cin]. */
int lr_pack_conv_weights_kt(const void* w, void* out, int Cout, int CG, int KT, int KHW, int Cin,
                            void* stream);
/* Weight gradient: out[tap][64][Nc] fp32 = sum_p dy[p,:] (x) x[p+shift(tap),:] on tcgen05 (M=64,
 * MN-major operands), split over CTAs and reduced in a fixed order.  x [B][T+KT-1][Hp][Wp][Cx];
 * dy [Gy][B][T+KT-1][Hp][Wp][Cy] zero except its interior, which starts dy_off rows in.
 * m_is_x = 0: rows of out[tap] are output channels (needs Gy*Cy <= 64), columns the Cx inputs;
 * m_is_x = 1: rows are the Cx = 64 inputs, columns the Gy*Cy outputs.
 * stack_kx (m_is_x = 0): the KW kx-taps of a filter row share one MMA (N = KW*Cx, operand blocks
 * are the same tile shifted by one row each); out is then [KT*KH][64][KW][Cx].
 * stack_ky = S = 128/Cy (with stack_kx): S filter rows share one M = 128 MMA (M block b = the dY tile
 * shifted by b image rows); out is [KT][ceil(KH/S)][S][Cy][KW][Cx] with block b of unit u holding
 * ky = u*S + (S-1-b) (ky >= KH: scratch).  fuse_kt: one CTA accumulates all KT planes of a tile.
 * stack_ky = -1 with m_is_x = 1: taps 2u, 2u+1 (flat (ky,kx) order inside a kt plane) share one M = 128
 * MMA (M block 1 = the x chunk shifted by the second tap); out is [KT][ceil(KH*KW/2)][2][64][Nc].
 * The k-steps of a plane's last tile that lie past the last row where dy can be non-zero are skipped.
 * lr_conv3d_wgrad_out_floats = elements of `out` (and of one split of the workspace).          */
size_t lr_conv3d_wgrad_out_floats(int KT, int KH, int KW, int Nc, int stack_ky);
size_t lr_conv3d_wgrad_workspace(int KT, int KH, int KW, int Nc, int splits, int stack_ky);
int lr_conv3d_wgrad(const void* x, const void* dy, float* out, void* workspace, size_t ws_bytes,
                    int B, int T, int H, int W, int Hp, int Wp, int Cx, int Cy, int Gy,
                    long long dy_off, int KT, int KH, int KW, int m_is_x, int stack_kx, int stack_ky,
                    int fuse_kt, int splits, void* stream);
int lr_unpool(const void* d_pooled, const uint8_t* argmax, void* out, float* d_bias, int B, int T,
              int H, int W, int C, int Cg, int Tp, int Hp, int Wp, int pt, int ph, int pw,
              void* stream);

/* -------- f1: attention core of the character decoder, all label positions at once ---------- */
/* reference: src/models/lipreader/better_model.py:195-223 (one position per call, stock ops).
 *   scores[b,l,t] = q[b,l,:].enc[b,t,:] ; w = allennlp masked_softmax(scores, t < lens[b]) ;
 *   ctx[b,l,:] = sum_t w[b,l,t] enc[b,t,:].
 * q (B,L,H) f32 is the decoder state ('dot') or attn_proj_general(state) ('general'); enc (B,T,H) f32.
 * One CTA per clip keeps the clip's encoder states in shared memory for all L positions.
 * weights (B,L,T), zsum (B,L) [the +1e-13 normaliser] and ctx (B,L,H) are outputs; bwd returns d_q, d_enc
 * (fully written) from d_ctx.                                                                      */
int lr_attn_fwd(const float* q, const float* enc, const int32_t* lens, int B, int L, int T, int H,
                float* weights, float* zsum, float* ctx, void* stream);
int lr_attn_bwd(const float* q, const float* enc, const int32_t* lens, const float* weights,
                const float* zsum, const float* d_ctx, int B, int L, int T, int H, float* d_q,
                float* d_enc, void* stream);
/* The same core with caller-supplied scores (B,L,T): the '1_layer_nn' / 'concat' score functions of
 * better_model.py:204-221 (the reference's default is attention_type='1_layer_nn', src/scripts/train.py:151).
 * Forward: masked softmax + context.  Backward: d_scores (B,L,T) and the context part of d_enc (B,T,H).        */
int lr_attn_scores_fwd(const float* scores, const float* enc, const int32_t* lens, int B, int L, int T, int H,
                       float* weights, float* zsum, float* ctx, void* stream);
int lr_attn_scores_bwd(const float* enc, const int32_t* lens, const float* weights, const float* zsum,
                       const float* d_ctx, int B, int L, int T, int H, float* d_scores, float* d_enc, void* stream);

/* -------- a5 / f4: position-map CNN body (and plain GEMMs) as "tap GEMMs" on tcgen05 --------- */
/* replaces: the tcl.conv2d / tcl.conv2d_transpose + batch_norm + relu/sigmoid calls of resfcn256
 * (src/models/face/prnet.py:211-280, one TF session.run per frame at :305-309) and, in store mode 3, the
 * input GEMM x @ W_ih^T + b_ih in front of nn.{GRU,LSTM,RNN} (src/models/lipreader/better_model.py:47-49,74).
 *
 *   acc[q, n] = sum over groups g, k < Kg of  A[q + tap_off[ph*n_groups+g]][k] * W[(ph*n_groups+g)*Cout_pad + n][k]
 *
 * a   : bf16 channels-last volume = matrix [rows][C] (row = one position of a zero-padded (B,Hp,Wp) grid; C*2 bytes
 *       apart).  C >= Kt: a group's K runs over the channels of ONE position (Kg <= C).  C < Kt: the K = Kt
 *       elements of row q are the channels of positions q .. q+Kt/C-1 (horizontally adjacent taps fused into one
 *       K tile; Kg == Kt); the allocation must extend (Kt - C)*2 bytes past the last row.
 * w   : bf16 [n_phases*n_groups*Cout_pad][w_pitch], K-major; tap_off (HOST pointer): row offset per (phase, group).
 * epilogue per position: v = alpha[n]*acc + beta[n] (NULL: 1 / 0) [+ gamma[n]*res[same row as out][n]] ; act 0 none,
 *       1 ReLU, 2 sigmoid; positions outside vy0 <= y < vy0+H, vx0 <= x < vx0+W of their grid are not stored.
 * store modes (yy = y-vy0, xx = x-vx0):
 *   0  bf16 -> out[(b*oHp + yy+opy)*oWp + xx+opx][n]                       (+ aux: even (yy,xx) also to
 *      aux[(b*aHp + yy/2+apad)*aWp + xx/2+apad][n], the input of a stride-2 1x1 shortcut)
 *   1  4 phases (py,px) = (ph>>1, ph&1) of a stride-2 transposed conv -> out[(b*oHp + 2yy+py+opy)*oWp + 2xx+px+opx][n]
 *   2  space-to-depth for a following stride-2 4x4 conv: block ((yy+1)>>1, (xx+1)>>1), channel slot
 *      (((yy+1)&1)*2 + ((xx+1)&1))*Cout_pad + n of rows oC = 4*Cout_pad wide
 *   3  plain GEMM: fp32 out[q*oC + n] for n < Cout, q < rows (no grid)
 *   4  fp32 compact out[((b*H + yy)*W + xx)*Cout + n] * out_scale for n < Cout (the final position map)
 * Asynchronous on `stream`; the descriptor is read before the call returns.                                     */
typedef struct lr_tapgemm_desc {
  const void* a; long long rows; int C; int Hp, Wp, vy0, vx0, H, W;
  const void* w; int w_pitch, Kg, Kt, n_phases, n_groups; const int32_t* tap_off;
  int Cout_pad, Cout; const float* alpha; const float* beta; const float* gamma; int act, mode;
  void* out; int oHp, oWp, oC, opy, opx;
  const void* res; int resC;
  void* aux; int aHp, aWp, aC, apad;
  float out_scale;
  int flags;            /* bit 0: never keep the weights / input chunk resident (test hook: stream every tap's tiles) */
  int pack;             /* 0/1, or 2 / 4: a matrix row holds `pack` horizontally adjacent positions (pack*C == Kt = 64):
                         * `rows` counts matrix rows, tap_off is in matrix rows, w has pack*Cout_pad rows per
                         * (phase, group) — accumulator columns [j*Cout_pad, (j+1)*Cout_pad) are position q*pack + j */
} lr_tapgemm_desc;
int lr_tapgemm(const lr_tapgemm_desc* desc, void* stream);
/* (N,H,W,3) f32 image -> interior (pad,pad) of the zero-padded bf16 volume (N,H+2*pad,W+2*pad,16), channels 3..15 zero */
int lr_pack_image16(const float* img, void* out_bf16, int N, int H, int W, int pad, void* stream);

/* No entry point of this header keeps state between calls: kernel variants are chosen per call (`kernel`, `variant`,
 * the flag bits of `swap`).  The measurement hooks and micro-benchmarks live in include/lr_b200_diag.h and are
 * compiled into a separate liblr_b200_diag.so.                                                                  */

#ifdef __cplusplus
}
#endif
#endif /* LR_B200_H_ */
