"""torch.autograd.Function wrappers over the C-ABI kernels (include/lr_b200.h).

PyTorch here is plumbing: device memory, the current stream and autograd bookkeeping.  The
arithmetic of every op below runs in liblr_b200.so; the only library calls are the plain GEMMs
around the recurrent kernel (x @ W_ih^T and the weight-gradient reductions), which go to cuBLAS.
"""
import torch

from . import native as N


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# a15: CTC
# ------------------------------------------------------------------------------------------------
# per-call kernel choices handed to the C-ABI (see include/lr_b200.h); 0 = the library's own choice / the parity path
CTC_KERNEL = 0        # 1 CTA-per-clip, 2 log-space warp-per-clip, 3 linear-space warp kernel first (tests walk all of them)
PROJ_VARIANT = 0      # 1 = 3xTF32 mma.sync forward; 2 = tcgen05 kind::tf32 forward + low-precision library GEMMs in the
                      # backward (the throughput path; falls back to 0 for shapes the kernel does not take)


class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lens, target_lens):
        N.require_cuda(log_probs, targets, input_lens, target_lens)
        lp = N.cont(log_probs, torch.float32)
        tg = N.cont(targets, torch.int32)
        il = N.cont(input_lens, torch.int32)
        tl = N.cont(target_lens, torch.int32)
        B, T, C = lp.shape
        Lmax = max(int(tg.shape[1]), 1)
        if tg.shape[1] == 0:
            tg = torch.zeros((B, 1), dtype=torch.int32, device=lp.device)
        nll = torch.empty(B, dtype=torch.float32, device=lp.device)
        need_grad = log_probs.requires_grad
        grad = torch.empty_like(lp) if need_grad else None
        L = N.lib()
        ws = _ws(L.lr_ctc_workspace(B, T, C, Lmax, CTC_KERNEL), lp.device)
        N.check(L.lr_ctc_fwd_bwd(N.ptr(lp), N.ptr(tg), N.ptr(il), N.ptr(tl), B, T, C, Lmax,
                                 N.ptr(nll), N.ptr(grad), N.ptr(ws), ws.numel(), CTC_KERNEL, N.stream()),
                "lr_ctc_fwd_bwd")
        ctx.unit_grad = grad
        return nll

    @staticmethod
    def backward(ctx, g):
        grad = ctx.unit_grad
        if grad is None:
            return None, None, None, None
        g = N.cont(g, torch.float32)
        out = torch.empty_like(grad)
        B = grad.shape[0]
        N.check(N.lib().lr_scale_rows(N.ptr(grad), N.ptr(g), N.ptr(out), B, grad[0].numel(), N.stream()),
                "lr_scale_rows")
        return out, None, None, None


def ctc_nll(log_probs, targets, input_lens, target_lens):
    """Per-sample CTC negative log-likelihood.  log_probs (B,T,C) f32 batch-first, targets (B,Lmax)
    CTC classes (label+1, blank=0), lens (B).  Differentiable w.r.t. log_probs (torch's native-CTC
    gradient convention).  Replaces F.ctc_loss(..., reduction='none') at src/train/ctc_loss.py:85."""
    return _CTC.apply(log_probs, targets, input_lens, target_lens)


def ctc_greedy_decode(log_probs, lens):
    """(B,T,C) log-probs -> (tokens (B,T) int32 zero padded, out_lens (B)) : arg-max, collapse repeats,
    drop blank (GreedyDecoder semantics, decoder.py:165-197)."""
    N.require_cuda(log_probs, lens)
    lp = N.cont(log_probs.detach(), torch.float32)
    B, T, C = lp.shape
    tokens = torch.empty((B, T), dtype=torch.int32, device=lp.device)
    out_lens = torch.empty(B, dtype=torch.int32, device=lp.device)
    N.check(N.lib().lr_ctc_greedy_decode(N.ptr(lp), N.ptr(N.cont(lens, torch.int32)), B, T, C, N.ptr(tokens),
                                         N.ptr(out_lens), N.stream()), "lr_ctc_greedy_decode")
    return tokens, out_lens


# ------------------------------------------------------------------------------------------------
# f1: attention core of the character decoder, all label positions of a clip at once
# ------------------------------------------------------------------------------------------------
class _AttnContext(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, enc, lens):
        N.require_cuda(q, enc, lens)
        q, enc = N.cont(q, torch.float32), N.cont(enc, torch.float32)
        lens32 = N.cont(lens, torch.int32)
        B, L, H = q.shape
        T = enc.shape[1]
        assert enc.shape[0] == B and enc.shape[2] == H
        w = torch.empty((B, L, T), dtype=torch.float32, device=q.device)
        z = torch.empty((B, L), dtype=torch.float32, device=q.device)
        c = torch.empty((B, L, H), dtype=torch.float32, device=q.device)
        N.check(N.lib().lr_attn_fwd(N.ptr(q), N.ptr(enc), N.ptr(lens32), B, L, T, H, N.ptr(w), N.ptr(z), N.ptr(c),
                                    N.stream()), "lr_attn_fwd")
        ctx.save_for_backward(q, enc, lens32, w, z)
        ctx.mark_non_differentiable(w)
        return c, w

    @staticmethod
    def backward(ctx, d_c, _d_w):
        q, enc, lens32, w, z = ctx.saved_tensors
        B, L, H = q.shape
        T = enc.shape[1]
        d_c = N.cont(d_c, torch.float32)
        d_q = torch.empty_like(q)
        d_enc = torch.empty_like(enc)
        N.check(N.lib().lr_attn_bwd(N.ptr(q), N.ptr(enc), N.ptr(lens32), N.ptr(w), N.ptr(z), N.ptr(d_c), B, L, T, H,
                                    N.ptr(d_q), N.ptr(d_enc), N.stream()), "lr_attn_bwd")
        return d_q, d_enc, None


class _AttnContextScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, enc, lens):
        N.require_cuda(scores, enc, lens)
        scores, enc = N.cont(scores, torch.float32), N.cont(enc, torch.float32)
        lens32 = N.cont(lens, torch.int32)
        B, L, T = scores.shape
        H = enc.shape[2]
        assert enc.shape[0] == B and enc.shape[1] == T
        w = torch.empty((B, L, T), dtype=torch.float32, device=enc.device)
        z = torch.empty((B, L), dtype=torch.float32, device=enc.device)
        c = torch.empty((B, L, H), dtype=torch.float32, device=enc.device)
        N.check(N.lib().lr_attn_scores_fwd(N.ptr(scores), N.ptr(enc), N.ptr(lens32), B, L, T, H, N.ptr(w), N.ptr(z),
                                           N.ptr(c), N.stream()), "lr_attn_scores_fwd")
        ctx.save_for_backward(enc, lens32, w, z)
        ctx.mark_non_differentiable(w)
        return c, w

    @staticmethod
    def backward(ctx, d_c, _d_w):
        enc, lens32, w, z = ctx.saved_tensors
        B, L, T = w.shape
        H = enc.shape[2]
        d_c = N.cont(d_c, torch.float32)
        d_scores = torch.empty_like(w)
        d_enc = torch.empty_like(enc)
        N.check(N.lib().lr_attn_scores_bwd(N.ptr(enc), N.ptr(lens32), N.ptr(w), N.ptr(z), N.ptr(d_c), B, L, T, H,
                                           N.ptr(d_scores), N.ptr(d_enc), N.stream()), "lr_attn_scores_bwd")
        return d_scores, d_enc, None


def attn_context_scores(scores, enc, lens):
    """The attention core with caller-computed scores (B,L,T) — '1_layer_nn' / 'concat' (better_model.py:204-223):
    allennlp masked softmax over t < len, context = weights @ enc.  -> (context (B,L,H), weights (B,L,T))."""
    return _AttnContextScores.apply(scores, enc, lens)


def attn_context(q, enc, lens):
    """Dot-product attention of L queries per clip over the clip's encoder states with allennlp's masked softmax
    (better_model.py:195-223): q (B,L,H), enc (B,T,H), lens (B) -> (context (B,L,H), weights (B,L,T))."""
    return _AttnContext.apply(q, enc, lens)


# ------------------------------------------------------------------------------------------------
# a14: Linear + masked log-softmax
# ------------------------------------------------------------------------------------------------
class _ProjLogSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, bias, log_mask):
        N.require_cuda(hidden, weight, bias, log_mask)
        shp = hidden.shape
        h2 = N.cont(hidden.reshape(-1, shp[-1]), torch.float32)
        w, b, lm = N.cont(weight, torch.float32), N.cont(bias, torch.float32), N.cont(log_mask, torch.float32)
        M, K = h2.shape
        C = w.shape[0]
        out = torch.empty((M, C), dtype=torch.float32, device=h2.device)
        variant = PROJ_VARIANT
        if variant == 2 and not N.lib().lr_proj_tc5_supported(M, K, C):
            variant = 0
        N.check(N.lib().lr_proj_logsoftmax_fwd(N.ptr(h2), N.ptr(w), N.ptr(b), N.ptr(lm), N.ptr(out),
                                               M, K, C, variant, N.stream()), "lr_proj_logsoftmax_fwd")
        ctx.save_for_backward(h2, w, out)
        ctx.in_shape = shp
        ctx.variant = variant
        return out.reshape(shp[:-1] + (C,))

    @staticmethod
    def backward(ctx, g):
        h2, w, out = ctx.saved_tensors
        M, K = h2.shape
        C = w.shape[0]
        g2 = N.cont(g.reshape(M, C), torch.float32)
        d_logits = torch.empty_like(out)
        if ctx.variant == 2:
            # throughput path: the log-softmax backward is the kernel, the two plain GEMMs go to the library with
            # bf16 operands and fp32 accumulation / output (like the GEMMs around the recurrent kernel)
            d_b = torch.empty(C, dtype=torch.float32, device=w.device)
            N.check(N.lib().lr_logsoftmax_bwd(N.ptr(g2), N.ptr(out), N.ptr(d_logits), N.ptr(d_b), M, C, N.stream()),
                    "lr_logsoftmax_bwd")
            # (the class dimension is padded to a multiple of 16: with C = 65 the library otherwise falls back to a legacy
            #  unaligned kernel — cutlass_75 s1688 align1, 119 + 26 us per step in the launch list against ~15 us aligned)
            Cp = (C + 15) // 16 * 16
            dl = torch.zeros((M, Cp), dtype=torch.bfloat16, device=w.device)
            dl[:, :C] = d_logits
            wp = torch.zeros((Cp, K), dtype=torch.bfloat16, device=w.device)
            wp[:C] = w
            d_hidden = torch.mm(dl, wp, out_dtype=torch.float32)
            d_w = torch.mm(dl.t(), h2.to(torch.bfloat16), out_dtype=torch.float32)[:C]
            return d_hidden.reshape(ctx.in_shape), d_w, d_b, None
        d_hidden = torch.empty_like(h2)
        d_w = torch.empty_like(w)
        d_b = torch.empty(C, dtype=torch.float32, device=w.device)
        N.check(N.lib().lr_proj_logsoftmax_bwd(N.ptr(g2), N.ptr(out), N.ptr(h2), N.ptr(w), N.ptr(d_logits),
                                               N.ptr(d_hidden), N.ptr(d_w), N.ptr(d_b), M, K, C, N.stream()),
                "lr_proj_logsoftmax_bwd")
        return d_hidden.reshape(ctx.in_shape), d_w, d_b, None


def proj_masked_log_softmax(hidden, weight, bias, log_mask):
    """log_softmax(hidden @ weight^T + bias + log_mask) fused (better_model.py:92-93)."""
    return _ProjLogSoftmax.apply(hidden, weight, bias, log_mask)


# ------------------------------------------------------------------------------------------------
# a12/a13: one (bi)directional recurrent layer with packed-sequence semantics
# ------------------------------------------------------------------------------------------------
# dtype of the plain GEMMs around the recurrent kernel (x @ W_ih^T, dW_ih, dW_hh, dX).  float32 is the
# parity path; bfloat16 operands with fp32 accumulation is the throughput path (BASELINE config "bf16").
GEMM_DTYPE = torch.float32
# True: run the recurrence on the persistent cluster kernels (bf16 operands) when the shape allows it
RNN_CLUSTER = False
# True (with GEMM_DTYPE = bfloat16): the four plain GEMMs around the recurrent kernel (x @ W_ih^T + b_ih, dX, dW_ih, dW_hh)
# run on this repo's tcgen05 tap-GEMM kernel (lr_tapgemm, store mode 3) instead of cuBLAS
GEMM_TCGEN05 = False


def _tc_ok(K, Nn):
    return GEMM_TCGEN05 and GEMM_DTYPE == torch.bfloat16 and K >= 64 and K % 8 == 0 and Nn % 16 == 0


def _mm_nt(a, w, bias=None, out_dtype=torch.float32):
    """a (M,K) @ w (N,K)^T [+ bias]: both operands K-major, the orientation the tensor-core kernel reads."""
    if _tc_ok(a.shape[1], w.shape[0]):
        r = tap_linear(N.cont(_lowp(a)), N.cont(_lowp(w)), bias)
        return r if out_dtype == torch.float32 else r.to(out_dtype)
    return _mm(a, w.t(), bias=bias, out_dtype=out_dtype)


def _lowp(t):
    """operand of a throughput-path GEMM: GEMM_DTYPE, converted at most once"""
    return t if t.dtype == GEMM_DTYPE else t.to(GEMM_DTYPE)


def _mm(a, b, bias=None, out_dtype=torch.float32):
    """plain GEMM -> cuBLAS.  fp32 path: a @ b [+ bias].  bf16 path: bf16 operands, fp32 accumulation, the result
    written once in `out_dtype` (no bf16 round trip, no separate bias pass)."""
    if GEMM_DTYPE == torch.float32:
        r = a.float() @ b.float()
        if bias is not None:
            r = r + bias
        return r if out_dtype == torch.float32 else r.to(out_dtype)
    a, b = _lowp(a), _lowp(b)
    if out_dtype == GEMM_DTYPE:
        return torch.mm(a, b) if bias is None else torch.addmm(bias.to(GEMM_DTYPE), a, b)
    if bias is None:
        return torch.mm(a, b, out_dtype=out_dtype)
    return torch.addmm(bias.to(out_dtype), a, b, out_dtype=out_dtype)


def tap_linear(a, w, bias=None):
    """a (M,K) bf16 @ w (N,K) bf16 ^T [+ bias (N) f32] -> (M,N) f32 on the tcgen05 tap-GEMM kernel (lr_tapgemm, store
    mode 3: one tap, fp32 accumulation, the bias added in the epilogue).  K a multiple of 8 and >= 64, N a multiple
    of 16: the shapes of the recurrent layers' input GEMM x @ W_ih^T + b_ih (better_model.py:47-49,74)."""
    import ctypes
    N.require_cuda(a, w, bias)
    M, K = a.shape
    Nn = w.shape[0]
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.shape[1] == K
    assert K >= 64 and K % 8 == 0 and Nn % 16 == 0, (K, Nn)
    a, w = N.cont(a), N.cont(w)
    out = torch.empty((M, Nn), dtype=torch.float32, device=a.device)
    d = N.TapGemmDesc()
    d.a, d.rows, d.C = a.data_ptr(), M, K
    d.w, d.w_pitch, d.Kg, d.Kt, d.n_phases, d.n_groups = w.data_ptr(), K, K, 64, 1, 1
    off = (ctypes.c_int32 * 1)(0)
    d.tap_off = ctypes.cast(off, ctypes.POINTER(ctypes.c_int32))
    d.Cout_pad, d.Cout = Nn, Nn
    d.beta = N.cont(bias, torch.float32).data_ptr() if bias is not None else None
    d.act, d.mode = 0, 3
    d.out, d.oC = out.data_ptr(), Nn
    d.out_scale = 1.0
    N.check(N.lib().lr_tapgemm(ctypes.byref(d), N.stream()), "lr_tapgemm")
    return out


class _RNNLayer(torch.autograd.Function):
    """inputs: x (B,T,I), lens (B) int32, mode str, then per direction (w_ih, w_hh, b_ih, b_hh)."""

    @staticmethod
    def forward(ctx, x, lens, mode, *weights):
        N.require_cuda(x, lens, *weights)
        D = len(weights) // 4
        G = N.RNN_GATES[mode]
        B, T, I = x.shape
        H = weights[1].shape[1]
        # throughput path: a bf16 feature tensor (conv front-end) feeds the input GEMM as it is
        x2 = N.cont(x.reshape(B * T, I), torch.float32 if GEMM_DTYPE == torch.float32 else GEMM_DTYPE)
        w_ih = torch.cat([weights[4 * d + 0] for d in range(D)], 0)            # (D*G*H, I)
        b_ih = torch.cat([weights[4 * d + 2] for d in range(D)], 0)
        w_hh = torch.stack([weights[4 * d + 1] for d in range(D)], 0).contiguous()   # (D,G*H,H)
        b_hh = torch.stack([weights[4 * d + 3] for d in range(D)], 0).contiguous()
        if GEMM_DTYPE != torch.float32:
            w_ih = _lowp(w_ih)            # converted while contiguous (a .to() of the transposed view is a strided copy)
        gi = _mm_nt(x2, w_ih, bias=b_ih)                                        # plain GEMM: lr_tapgemm or cuBLAS
        lens32 = N.cont(lens, torch.int32)
        dev = x.device
        hidden = torch.empty((B, T, D * H), dtype=torch.float32, device=dev)
        h_n = torch.empty((D, B, H), dtype=torch.float32, device=dev)
        c_n = torch.empty((D, B, H), dtype=torch.float32, device=dev) if mode == "LSTM" else None
        L = N.lib()
        S = L.lr_rnn_saved_per_unit(N.RNN_MODES[mode])
        saved = torch.empty((B, T, D, S * H), dtype=torch.float32, device=dev) if S > 0 else None
        use_cluster = bool(RNN_CLUSTER and L.lr_rnn_cluster_supported(N.RNN_MODES[mode], H))
        # hidden sizes whose W_hh does not fit one cluster (BiLSTM-768): grid-persistent kernels (csrc/rnn_grid.cu)
        use_grid = bool(RNN_CLUSTER and not use_cluster and L.lr_rnn_grid_supported(N.RNN_MODES[mode], H, D))
        if use_cluster:
            N.check(L.lr_rnn_cluster_fwd(N.RNN_MODES[mode], N.ptr(gi), N.ptr(w_hh), N.ptr(b_hh), N.ptr(lens32), B, T,
                                         H, D, N.ptr(hidden), N.ptr(h_n), N.ptr(c_n), N.ptr(saved), N.stream()),
                    "lr_rnn_cluster_fwd")
        elif use_grid:
            ws = _ws(L.lr_rnn_grid_workspace(N.RNN_MODES[mode], H, D), dev)
            N.check(L.lr_rnn_grid_fwd(N.RNN_MODES[mode], N.ptr(gi), N.ptr(w_hh), N.ptr(b_hh), N.ptr(lens32), B, T, H, D,
                                      N.ptr(hidden), N.ptr(h_n), N.ptr(c_n), N.ptr(saved), N.ptr(ws), ws.numel(),
                                      N.stream()), "lr_rnn_grid_fwd")
        else:
            ws = _ws(L.lr_rnn_workspace(N.RNN_MODES[mode], B, T, H, D), dev)
            N.check(L.lr_rnn_fwd(N.RNN_MODES[mode], N.ptr(gi), N.ptr(w_hh), N.ptr(b_hh), N.ptr(lens32), B, T, H, D,
                                 N.ptr(hidden), N.ptr(h_n), N.ptr(c_n), N.ptr(saved), N.ptr(ws), ws.numel(),
                                 N.stream()), "lr_rnn_fwd")
        ctx.use_grid = use_grid
        ctx.use_cluster = use_cluster
        ctx.mode, ctx.dims = mode, (B, T, I, H, D, G)
        ctx.save_for_backward(x2, lens32, w_ih, w_hh, hidden, saved if saved is not None else torch.empty(0, device=dev))
        ctx.x_needs_grad = x.requires_grad
        ctx.x_dtype = x.dtype
        if mode == "LSTM":
            return hidden, h_n, c_n
        return hidden, h_n

    @staticmethod
    def backward(ctx, d_hidden, d_h_n, d_c_n=None):
        mode = ctx.mode
        B, T, I, H, D, G = ctx.dims
        x2, lens32, w_ih, w_hh, hidden, saved = ctx.saved_tensors
        dev = x2.device
        d_hidden = N.cont(d_hidden, torch.float32) if d_hidden is not None else torch.zeros_like(hidden)
        d_h_n = N.cont(d_h_n, torch.float32) if d_h_n is not None else None
        d_c_n = N.cont(d_c_n, torch.float32) if d_c_n is not None else None
        d_gi = torch.empty((B, T, D, G * H), dtype=torch.float32, device=dev)
        d_gh = torch.empty((B, T, D, G * H), dtype=torch.float32, device=dev)
        h_prev = torch.empty((B, T, D, H), dtype=torch.float32, device=dev)
        L = N.lib()
        if ctx.use_cluster:
            N.check(L.lr_rnn_cluster_bwd(N.RNN_MODES[mode], N.ptr(d_hidden), N.ptr(d_h_n), N.ptr(d_c_n),
                                         N.ptr(saved) if saved.numel() else None, N.ptr(hidden), N.ptr(w_hh),
                                         N.ptr(lens32), B, T, H, D, N.ptr(d_gi), N.ptr(d_gh), N.ptr(h_prev),
                                         N.stream()), "lr_rnn_cluster_bwd")
        elif ctx.use_grid:
            ws = _ws(L.lr_rnn_grid_workspace(N.RNN_MODES[mode], H, D), dev)
            N.check(L.lr_rnn_grid_bwd(N.RNN_MODES[mode], N.ptr(d_hidden), N.ptr(d_h_n), N.ptr(d_c_n),
                                      N.ptr(saved) if saved.numel() else None, N.ptr(hidden), N.ptr(w_hh),
                                      N.ptr(lens32), B, T, H, D, N.ptr(d_gi), N.ptr(d_gh), N.ptr(h_prev), N.ptr(ws),
                                      ws.numel(), N.stream()), "lr_rnn_grid_bwd")
        else:
            w_hh_t = w_hh.transpose(1, 2).contiguous()                          # (D,H,G*H)
            ws = _ws(L.lr_rnn_workspace(N.RNN_MODES[mode], B, T, H, D), dev)
            N.check(L.lr_rnn_bwd(N.RNN_MODES[mode], N.ptr(d_hidden), N.ptr(d_h_n), N.ptr(d_c_n),
                                 N.ptr(saved) if saved.numel() else None, N.ptr(hidden), N.ptr(w_hh_t),
                                 N.ptr(lens32), B, T, H, D, N.ptr(d_gi), N.ptr(d_gh), N.ptr(h_prev), N.ptr(ws),
                                 ws.numel(), N.stream()), "lr_rnn_bwd")
        # weight-gradient reductions over B*T: plain GEMMs -> cuBLAS
        d_gi2 = d_gi.reshape(B * T, D * G * H)
        d_gh2 = d_gh.reshape(B * T, D * G * H)
        h_prev3 = h_prev.reshape(B * T, D, H)
        d_b_ih = d_gi2.sum(0)
        d_b_hh = d_gh2.sum(0)
        if GEMM_DTYPE != torch.float32:                                         # one conversion per operand
            d_gi2, d_gh2, h_prev3 = _lowp(d_gi2), _lowp(d_gh2), _lowp(h_prev3)
        tc = _tc_ok(B * T, I) and _tc_ok(B * T, H)
        # (the K-major operands the tensor-core kernel wants are the transposes: one bf16 copy each)
        d_w_ih = _mm_nt(d_gi2.t().contiguous(), x2.t().contiguous()) if tc else _mm(d_gi2.t(), x2)       # (D*G*H, I)
        d_gh3 = d_gh2.reshape(B * T, D, G * H)
        grads = []
        for d in range(D):
            d_w_hh = _mm_nt(d_gh3[:, d].t().contiguous(), h_prev3[:, d].t().contiguous()) if tc \
                else _mm(d_gh3[:, d].t(), h_prev3[:, d])                        # (G*H, H)
            grads += [d_w_ih[d * G * H:(d + 1) * G * H], d_w_hh, d_b_ih[d * G * H:(d + 1) * G * H],
                      d_b_hh[d * G * H:(d + 1) * G * H]]
        d_x = None
        if ctx.x_needs_grad:
            d_x = (_mm_nt(d_gi2, w_ih.t().contiguous(), out_dtype=ctx.x_dtype) if _tc_ok(D * G * H, I)
                   else _mm(d_gi2, w_ih, out_dtype=ctx.x_dtype)).reshape(B, T, I)
        return (d_x, None, None) + tuple(grads)


def rnn_layer(x, lens, mode, weights):
    """weights: flat list per direction [w_ih, w_hh, b_ih, b_hh] (+ the reverse direction's four).
    Returns (hidden (B,T,D*H), h_n (D,B,H)[, c_n])."""
    return _RNNLayer.apply(x, lens, mode, *weights)


# ------------------------------------------------------------------------------------------------
# vision ops (no autograd)
# ------------------------------------------------------------------------------------------------
def rect_geometry(rects, img_h, img_w):
    """rects (N,4) int32 (left,right,top,bottom) -> rect_pad (N,4), crop (N,4)=(2cx,2cy,size,0).
    face.py:76-90 + prnet.py:112-119."""
    N.require_cuda(rects)
    r = N.cont(rects, torch.int32)
    n = r.shape[0]
    rect_pad = torch.empty_like(r)
    crop = torch.empty_like(r)
    N.check(N.lib().lr_rect_geometry(N.ptr(r), n, img_h, img_w, N.ptr(rect_pad), N.ptr(crop), N.stream()),
            "lr_rect_geometry")
    return rect_pad, crop


def warp256(frames, crop):
    """frames (N,H,W,3) u8, crop (N,4) -> (N,256,256,3) f32 in [0,1] (prnet.py:137-143)."""
    N.require_cuda(frames, crop)
    f = N.cont(frames, torch.uint8)
    n, H, W, _ = f.shape
    out = torch.empty((n, 256, 256, 3), dtype=torch.float32, device=f.device)
    N.check(N.lib().lr_warp256(N.ptr(f), N.ptr(N.cont(crop, torch.int32)), N.ptr(out), n, H, W, N.stream()),
            "lr_warp256")
    return out


def posmap_gather(posmap, crop, rect_pad, kpt_idx, face_idx=None):
    """posmap (N,256,256,3) f32 -> landmarks (N,68,3) f64 [, vertices (N,V,3) f64], face-relative
    (prnet.py:151-156,169,179-180 + face.py:171-174)."""
    N.require_cuda(posmap, crop, rect_pad, kpt_idx, face_idx)
    pm = N.cont(posmap, torch.float32)
    n = pm.shape[0]
    k = N.cont(kpt_idx, torch.int32)
    lmk = torch.empty((n, k.numel(), 3), dtype=torch.float64, device=pm.device)
    vtx = None
    fi = None
    if face_idx is not None:
        fi = N.cont(face_idx, torch.int32)
        vtx = torch.empty((n, fi.numel(), 3), dtype=torch.float64, device=pm.device)
    N.check(N.lib().lr_posmap_gather(N.ptr(pm), N.ptr(N.cont(crop, torch.int32)), N.ptr(N.cont(rect_pad, torch.int32)),
                                     N.ptr(k), k.numel(), N.ptr(fi), fi.numel() if fi is not None else 0,
                                     N.ptr(lmk), N.ptr(vtx), n, N.stream()), "lr_posmap_gather")
    return (lmk, vtx) if face_idx is not None else lmk


def mouth_crop(frames, lmk, rect_pad, out_h=50, out_w=100):
    """N2 extension: (N,H,W,3) u8 + landmarks -> (N,out_h,out_w,3) u8 mouth clips, roi (N,4)."""
    N.require_cuda(frames, lmk, rect_pad)
    f = N.cont(frames, torch.uint8)
    n, H, W, _ = f.shape
    out = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=f.device)
    roi = torch.empty((n, 4), dtype=torch.int32, device=f.device)
    N.check(N.lib().lr_mouth_crop(N.ptr(f), N.ptr(N.cont(lmk, torch.float64)), N.ptr(N.cont(rect_pad, torch.int32)),
                                  N.ptr(out), N.ptr(roi), n, H, W, out_h, out_w, N.stream()), "lr_mouth_crop")
    return out, roi


def collate_pad(src_concat, row_offsets, B, Tmax, F):
    """Ragged f64 rows (sum T_i, F) + offsets (B+1) i64 -> (B,Tmax,F) f32 zero padded
    (data_loader.py:124-137)."""
    N.require_cuda(src_concat, row_offsets)
    s = N.cont(src_concat, torch.float64)
    o = N.cont(row_offsets, torch.int64)
    dst = torch.empty((B, Tmax, F), dtype=torch.float32, device=s.device)
    N.check(N.lib().lr_collate_pad_f64(N.ptr(s), N.ptr(o), N.ptr(dst), B, Tmax, F, N.stream()),
            "lr_collate_pad_f64")
    return dst
