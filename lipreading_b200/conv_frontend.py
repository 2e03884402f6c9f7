"""Spatio-temporal conv front-end (SURVEY §8 row N1 — north-star extension, no reference code).

Spec (LipNet STCNN, the paper the reference cites at README.md:211-214), input clip (B,T,H,W,3) u8:
    x/255 -> Conv3d(3->32,k(3,5,5),s(1,2,2),p(1,2,2)) + ReLU + MaxPool(1,2,2)
          -> Conv3d(32->64,k(3,5,5),p(1,2,2))         + ReLU + MaxPool(1,2,2)
          -> Conv3d(64->96,k(3,3,3),p(1,1,1))         + ReLU + MaxPool(1,2,2)
          -> per-frame flatten in (h,w,c) order -> (B,T,96*h*w)   (1728 for 100x50 frames)
Oracle: oracle/conv3d.py (torch.nn.functional.conv3d fp32 on bf16-rounded operands).

All three layers run on ONE tcgen05 kernel (csrc/conv3d_sm100.cu): the strided first layer is turned
into a stride-1 3x3x3 conv over a 2x2 space-to-depth'ed clip (12 -> 16 channels), so every layer is a
"shifted-window implicit GEMM" over a zero-padded channels-last bf16 volume.  dgrad is the same
kernel on the un-pooled output gradient with flipped/transposed weights.
Parameters keep nn.Conv3d's names/shapes (conv{1,2,3}.weight (Cout,Cin,KT,KH,KW), .bias).
"""
import torch
import torch.nn as nn

from . import native as N
from .native import _i, _vp

N.register("lr_conv3d_supported", _i, [])
N.register("lr_clip_s2d", _i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp])
N.register("lr_unpool", _i, [_vp, _vp, _vp, _vp] + [_i] * 12 + [_vp])
N.register("lr_conv3d_fwd", _i, [_vp, _vp, _vp, _vp, _vp] + [_i] * 21 + [_vp])
N.register("lr_conv3d_dgrad_unpool", _i, [_vp, _vp, _vp, _vp, _vp] + [_i] * 20 + [_vp])
N.register("lr_pack_conv_weights", _i, [_vp, _vp, _i, _i, _i, _i, _vp])
N.register("lr_pack_conv_weights_kt", _i, [_vp, _vp, _i, _i, _i, _i, _i, _vp])
N.register("lr_conv3d_wgrad_out_floats", N._sz, [_i] * 5)
N.register("lr_conv3d_wgrad_workspace", N._sz, [_i] * 6)
N.register("lr_conv3d_wgrad", _i, [_vp, _vp, _vp, _vp, N._sz] + [_i] * 9 + [N._i64] + [_i] * 8 + [_vp])

LAYERS = (  # name, Cin, Cout, kernel, stride, pad
    ("conv1", 3, 32, (3, 5, 5), (1, 2, 2), (1, 2, 2)),
    ("conv2", 32, 64, (3, 5, 5), (1, 1, 1), (1, 2, 2)),
    ("conv3", 64, 96, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
)


def _pow2_at_least(n):
    p = 8
    while p < n:
        p *= 2
    return p


def _plane_rows(h_valid, k, wp):
    """Allocated padded plane height: >= h_valid + k - 1, rounded up to whole 128-position tiles so
    no tile straddles two planes (required by the weight-gradient pass, harmless for the others)."""
    r = 128 // wp
    return (h_valid + k - 1 + r - 1) // r * r


STACK_KX = True      # wgrad: fold the KW kx-taps of a filter row into one N = KW*Cx MMA
STACK_KY = True      # wgrad: also fold 128/Cy filter rows into one M = 128 MMA (M = 64 is half rate)


def conv3d_wgrad_native(x, dy, B, T, H, W, Hp, Wp, Cx, Cy, Gy, dy_off, K, m_is_x, splits=0, stack_kx=None,
                        stack_ky=None, tag="wgrad", algo_macs=None):
    """Thin call into lr_conv3d_wgrad -> fp32 [taps][rows][Nc]: rows = 64 input channels, Nc = Gy*Cy if
    m_is_x; else rows = output channels (64, or Cy when ky-stacked) and Nc = Cx."""
    L = N.lib()
    Nc = Gy * Cy if m_is_x else Cx
    KT, KH, KW = K
    taps = KT * KH * KW
    use_ky = STACK_KY if stack_ky is None else stack_ky
    stack = (STACK_KX if stack_kx is None else stack_kx) and not m_is_x and KW > 1 and KW * Cx <= 256
    S = 128 // Cy if (stack and Gy == 1 and use_ky) else 1
    U = -(-KH // S)
    if S > 1 and (128 + (U * S - 1) * Wp + KW - 1 > 256 or -(-(H + min(S, KH) - 1) // (128 // Wp)) * (128 // Wp) > Hp):
        S, U = 1, KH                                            # halo / plane height do not allow stacking
    pair = bool(m_is_x and use_ky and KH * KW > 1)              # two taps per M = 128 MMA (M = 64 runs at half rate)
    P = -(-KH * KW // 2)
    fuse_kt = S > 1 and KT * U * KW * Cx <= 512
    if splits <= 0:
        if pair:
            groups = KT * -(-P // (512 // Nc))
        elif S > 1:
            groups = 1 if fuse_kt else KT * -(-U // (512 // (KW * Cx)))
        else:
            units, n_mma = (KH, KW * Cx) if stack else (KH * KW, Nc)
            groups = KT * -(-units // (512 // n_mma))
        splits = max(1, 148 // groups)
    sk = -1 if pair else S
    ws = torch.empty(L.lr_conv3d_wgrad_workspace(KT, KH, KW, Nc, splits, sk), dtype=torch.uint8, device=x.device)
    out = torch.empty(L.lr_conv3d_wgrad_out_floats(KT, KH, KW, Nc, sk), dtype=torch.float32, device=x.device)
    rec = KERNEL_TIMING
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        macs = algo_macs if algo_macs is not None else B * T * H * W * Cx * Cy * Gy * taps
        rec.append((tag, e0, e1, 2.0 * macs))
    N.check(L.lr_conv3d_wgrad(N.ptr(x), N.ptr(dy), N.ptr(out), N.ptr(ws), ws.numel(), B, T, H, W, Hp, Wp, Cx, Cy,
                              Gy, dy_off, KT, KH, KW, m_is_x, int(stack), sk, int(fuse_kt), splits, N.stream()),
            "lr_conv3d_wgrad")
    if rec is not None:
        e1.record()
    if pair:                        # [KT][P][2][64][Nc], flat spatial tap = 2u + b -> [taps][64][Nc]
        return out.reshape(KT, P * 2, 64, Nc)[:, :KH * KW].reshape(taps, 64, Nc)
    if S > 1:                       # [KT][U][S(b)][Cy][KW][Cx], ky = u*S + S-1-b -> [taps][Cy][Cx]
        out = out.reshape(KT, U, S, Cy, KW, Cx).flip(2).reshape(KT, U * S, Cy, KW, Cx)[:, :KH]
        return out.permute(0, 1, 3, 2, 4).reshape(taps, Cy, Cx)
    out = out.reshape(taps, 64, Nc)
    if stack:                       # [KT*KH][64][KW][Cx] -> [taps][64][Cx]
        out = out.reshape(KT * KH, 64, KW, Cx).permute(0, 2, 1, 3).reshape(taps, 64, Cx)
    return out


def feature_dim(H, W):
    h, w = H // 2, W // 2          # conv1 stride 2 (k5,p2)
    for _ in range(3):
        h, w = h // 2, w // 2      # three 2x2 pools
    return 96 * h * w


# dtype of the per-frame features handed to the recurrent layer: float32 (parity path) or bfloat16 — what the conv3
# epilogue writes, so the throughput path skips a widen + narrow round trip in front of the bf16 input GEMM
OUT_DTYPE = torch.float32

# bench.py sets this to a list to get per-launch CUDA-event timings of the conv kernel:
# entries are (tag, start_event, end_event, algorithmic_flops)
KERNEL_TIMING = None


# conv orientation (see lr_b200.h): 0 = one MMA per tap, 1 = channels on the MMA M lanes, 3 = the KT kt-taps of a
# spatial tap stacked on N (input plane c x [W(kt=KT-1);..;W(kt=0)] -> accumulators of frames c-KT+1..c)
SWAP = 3
# False (default): the MMA-issuing warps of the conv kernel share one chunk list per work item (without the LR_CONV_SINGLE_WRITER flag,
# ~15 % fewer tensor-pipe cycles on conv2's input gradient); the frames at a hand-over point are then summed in a
# timing-dependent order.  True: every accumulator has a single writer -> bit-reproducible runs.
DETERMINISTIC = False
FUSE_UNPOOL = True      # dgrad epilogue writes the un-pooled gradient of the layer below (+ its bias gradient) directly
DGRAD_KX_STACK = False  # conv2 dgrad (Cout = 32): the 5 kx-taps of a filter row share one N = 160 MMA (orientation 2)


def conv3d_native(x, w, bias, y, argmax, B, T, H, W, Hp, Wp, Cin, CG, Cout, K, epi_mode, ovol, ooff, J=0,
                  tag="conv", algo_macs=None, swap=None, d_bias=None):
    """Thin call into lr_conv3d_fwd (see include/lr_b200.h).  w: bf16 [Cout][CG][taps][Cin] (packed here).
    epi_mode 2 = lr_conv3d_dgrad_unpool: `argmax` is the INPUT arg-max map, `y` the padded dY volume of the layer
    below, `d_bias` receives that layer's bias gradient."""
    taps = K[0] * K[1] * K[2]
    w = N.cont(w)
    assert w.dtype == torch.bfloat16 and w.numel() == Cout * CG * taps * Cin
    mode = int(SWAP if swap is None else swap)
    if mode == 3 and K[0] == 1:
        mode = 0                                   # nothing to stack
    wp = torch.empty_like(w)
    if mode == 3:
        N.check(N.lib().lr_pack_conv_weights_kt(N.ptr(w), N.ptr(wp), Cout, CG, K[0], K[1] * K[2], Cin, N.stream()),
                "lr_pack_conv_weights_kt")
    else:
        N.check(N.lib().lr_pack_conv_weights(N.ptr(w), N.ptr(wp), Cout, CG, taps, Cin, N.stream()), "lr_pack_conv_weights")
    w = wp
    rec = KERNEL_TIMING
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if DETERMINISTIC or torch.are_deterministic_algorithms_enabled():
        mode |= 0x100                              # LR_CONV_SINGLE_WRITER: per-call flag bit of `swap`
    if epi_mode == 2:
        N.check(N.lib().lr_conv3d_dgrad_unpool(N.ptr(x), N.ptr(w), N.ptr(argmax), N.ptr(y), N.ptr(d_bias), B, T, H, W,
                                               Hp, Wp, Cin, CG, Cout, K[0], K[1], K[2], ovol[0], ovol[1], ovol[2],
                                               ooff[0], ooff[1], ooff[2], J, mode, N.stream()),
                "lr_conv3d_dgrad_unpool")
    else:
        N.check(N.lib().lr_conv3d_fwd(N.ptr(x), N.ptr(w), N.ptr(bias), N.ptr(y), N.ptr(argmax), B, T, H, W, Hp, Wp,
                                      Cin, CG, Cout, K[0], K[1], K[2], epi_mode, ovol[0], ovol[1], ovol[2],
                                      ooff[0], ooff[1], ooff[2], J, mode, N.stream()),
                "lr_conv3d_fwd")
    if rec is not None:
        e1.record()
        macs = algo_macs if algo_macs is not None else B * T * H * W * Cout * Cin * CG * K[0] * K[1] * K[2]
        rec.append((tag, e0, e1, 2.0 * macs))


def s2d_weight(w1):
    """(32,3,3,5,5) stride-2 5x5 kernel -> (32, 3,3,3, 16) stride-1 kernel over the 2x2
    space-to-depth clip: W'[co,kt,a,b,(dy,dx,c)] = W[co,c,kt,2a+dy,2b+dx] (0 beyond 5)."""
    co = w1.shape[0]
    wp = torch.nn.functional.pad(w1, (0, 1, 0, 1))                       # (co,3,3,6,6)
    wp = wp.reshape(co, 3, 3, 3, 2, 3, 2)                                # co,c,kt,a,dy,b,dx
    wp = wp.permute(0, 2, 3, 5, 4, 6, 1).reshape(co, 3, 3, 3, 12)        # co,kt,a,b,(dy,dx,c)
    return torch.nn.functional.pad(wp, (0, 4))                           # 12 -> 16 channels


def s2d_weight_grad(dw16):
    """Inverse map for the gradient: (32,3,3,3,16) -> (32,3,3,5,5)."""
    co = dw16.shape[0]
    g = dw16[..., :12].reshape(co, 3, 3, 3, 2, 2, 3)                     # co,kt,a,b,dy,dx,c
    g = g.permute(0, 6, 1, 2, 4, 3, 5).reshape(co, 3, 3, 6, 6)           # co,c,kt,(a,dy),(b,dx)
    return g[:, :, :, :5, :5]


def gemm_weight(w):
    """(Cout,Cin,KT,KH,KW) -> (Cout, KT,KH,KW, Cin) K-major rows."""
    return w.permute(0, 2, 3, 4, 1).contiguous()


def dgrad_weight(w, cg):
    """(Cout,Cin,KT,KH,KW) -> flipped/transposed (Cin, CG, taps, Cout/CG) for the dgrad pass."""
    co, ci = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    wf = w.flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(ci, taps, co // cg, cg)
    return wf.permute(0, 2, 1, 3).contiguous()


class _VolumePool:
    """Zero-padded activation volumes are large (GBs at B=256) and only their INTERIOR is rewritten
    every step; re-zeroing them per step costs ~1 ms of pure memset.  The pool hands out ONE cached
    buffer per tag: borders are zeroed once at allocation and never written again.  The reference's
    length-sorted batches give a new T for almost every batch, so the pool must not keep one set of volumes
    per shape: a tag's storage is allocated for the largest volume seen so far and smaller volumes are carved
    from it (re-zeroed only when the geometry — and with it the border pattern — changes).  A generation
    counter catches the one unsafe pattern (two forwards before a backward)."""

    def __init__(self):
        self.bufs = {}
        self.generation = 0
        self.enabled = True

    def get(self, tag, shape, dtype, device):
        if not self.enabled:
            return torch.zeros(shape, dtype=dtype, device=device)
        key = (tag, dtype, str(device))
        n = 1
        for d in shape:
            n *= int(d)
        ent = self.bufs.get(key)
        if ent is None or ent[0].numel() < n:
            if ent is not None:
                self.bufs.pop(key)                   # release the smaller storage before asking for the larger
                del ent
            ent = [torch.zeros(n, dtype=dtype, device=device), tuple(shape)]
            self.bufs[key] = ent
        elif ent[1] != tuple(shape):
            ent[0][:n].zero_()                       # same storage, new geometry: old interior is in the new borders
            ent[1] = tuple(shape)
        return ent[0][:n].view(shape)


POOL = _VolumePool()


class _ConvStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clip, w1, b1, w2, b2, w3, b3):
        N.require_cuda(clip, w1, w2, w3)
        assert clip.dtype == torch.uint8 and clip.dim() == 5 and clip.shape[-1] == 3
        B, T, H, W, _ = clip.shape
        assert H % 16 == 0 or (H // 2) % 2 == 0, "frame height must survive one stride-2 conv + three pools"
        dev = clip.device
        bf = torch.bfloat16
        L = N.lib()
        H1, W1 = H // 2, W // 2                      # conv1 output (stride 2)
        H2, W2 = H1 // 2, W1 // 2                    # after pool1 = conv2 in/out
        H3, W3 = H2 // 2, W2 // 2                    # after pool2 = conv3 in/out
        H4, W4 = H3 // 2, W3 // 2                    # after pool3
        Wp1, Wp2, Wp3 = _pow2_at_least(W1 + 2), _pow2_at_least(W2 + 4), _pow2_at_least(W3 + 2)
        # zero-padded channels-last volumes (borders stay zero; interiors are fully overwritten)
        Hp1, Hp2, Hp3 = _plane_rows(H1, 3, Wp1), _plane_rows(H2, 5, Wp2), _plane_rows(H3, 3, Wp3)
        z = POOL.get("z", (B, T + 2, Hp1, Wp1, 16), bf, dev)
        a1 = POOL.get("a1", (B, T + 2, Hp2, Wp2, 32), bf, dev)
        a2 = POOL.get("a2", (B, T + 2, Hp3, Wp3, 64), bf, dev)
        POOL.generation += 1
        ctx.generation = POOL.generation
        feat = torch.empty((B, T, H4, W4, 96), dtype=bf, device=dev)
        am1 = torch.empty((B, T, H2, W2, 32), dtype=torch.uint8, device=dev)
        am2 = torch.empty((B, T, H3, W3, 64), dtype=torch.uint8, device=dev)
        am3 = torch.empty((B, T, H4, W4, 96), dtype=torch.uint8, device=dev)
        N.check(L.lr_clip_s2d(N.ptr(N.cont(clip)), N.ptr(z), B, T, H, W, Hp1, Wp1, N.stream()), "lr_clip_s2d")
        g1 = s2d_weight(w1.detach()).to(bf).contiguous()
        g2 = gemm_weight(w2.detach()).to(bf)
        g3 = gemm_weight(w3.detach()).to(bf)
        conv3d_native(z, g1, b1.detach().float(), a1, am1, B, T, H1, W1, Hp1, Wp1, 16, 1, 32, (3, 3, 3), 0,
                      (T + 2, Hp2, Wp2), (1, 2, 2), tag="conv1.fwd",
                      algo_macs=B * T * H1 * W1 * 32 * 3 * 75)       # the true 3x5x5x3 stride-2 conv
        conv3d_native(a1, g2, b2.detach().float(), a2, am2, B, T, H2, W2, Hp2, Wp2, 32, 1, 64, (3, 5, 5), 0,
                      (T + 2, Hp3, Wp3), (1, 1, 1), tag="conv2.fwd")
        conv3d_native(a2, g3, b3.detach().float(), feat, am3, B, T, H3, W3, Hp3, Wp3, 64, 1, 96, (3, 3, 3), 0,
                      (T, H4, W4), (0, 0, 0), tag="conv3.fwd")
        ctx.save_for_backward(z, a1, a2, am1, am2, am3, w1, w2, w3)
        ctx.geom = (B, T, H, W)
        feat = feat.reshape(B, T, H4 * W4 * 96)
        return feat if OUT_DTYPE == torch.bfloat16 else feat.float()

    @staticmethod
    def backward(ctx, d_feat):
        z, a1, a2, am1, am2, am3, w1, w2, w3 = ctx.saved_tensors
        if POOL.enabled and ctx.generation != POOL.generation:
            raise RuntimeError("conv front-end: another forward ran before this backward and reused the pooled "
                               "activation volumes; set lipreading_b200.conv_frontend.POOL.enabled = False")
        B, T, H, W = ctx.geom
        dev = z.device
        bf = torch.bfloat16
        L = N.lib()
        H1, W1 = H // 2, W // 2
        H2, W2 = H1 // 2, W1 // 2
        H3, W3 = H2 // 2, W2 // 2
        H4, W4 = H3 // 2, W3 // 2
        Hp1, Hp2, Hp3 = z.shape[2], a1.shape[2], a2.shape[2]
        Wp1, Wp2, Wp3 = z.shape[3], a1.shape[3], a2.shape[3]

        def unpool(dp, am, Hf, Wf, C, Cg, pad, Hp, Wp):
            """pooled gradient -> conv-output gradient inside a zero-padded, channel-grouped volume with
            the SAME plane geometry as the layer's input (so dgrad and wgrad can both read it)."""
            out = POOL.get("dy%d" % C, (C // Cg, B, T + 2, Hp, Wp, Cg), bf, dev)
            d_bias = torch.empty(C, dtype=torch.float32, device=dev)
            N.check(L.lr_unpool(N.ptr(dp), N.ptr(am), N.ptr(out), N.ptr(d_bias), B, T, Hf, Wf, C, Cg, T + 2, Hp,
                                Wp, pad[0], pad[1], pad[2], N.stream()), "lr_unpool")
            return out, d_bias

        # ---- layer 3: d_feat -> dY3 (3 groups x 32 ch, interior at (1,1,1)) ----
        dp3 = N.cont(d_feat.reshape(B, T, H4, W4, 96).to(bf))
        dy3, db3 = unpool(dp3, am3, H3, W3, 96, 32, (1, 1, 1), Hp3, Wp3)
        d3 = conv3d_wgrad_native(a2, dy3, B, T, H3, W3, Hp3, Wp3, 64, 32, 3, (Hp3 + 1) * Wp3 + 1, (3, 3, 3), 1,
                                 tag="conv3.wgrad")
        dw3 = d3.reshape(3, 3, 3, 64, 96).permute(4, 3, 0, 1, 2)       # [tap][ci][co] -> (co,ci,kt,ky,kx)

        def dgrad_unpool(dy, wflip, am, Hd, Wd, Hpd, Wpd, Cin, CG, C, K, pad, Hb, Wb, Hpo, Wpo, tag, swap=None):
            """dgrad of the layer whose output gradient is `dy` (pooled resolution Hd x Wd of the layer below),
            producing the dY volume (valid extent Hb x Wb, padded Hpo x Wpo) and bias gradient of the layer below"""
            if FUSE_UNPOOL:
                out = POOL.get("dy%d" % C, (1, B, T + 2, Hpo, Wpo, C), bf, dev)
                d_bias = torch.empty(C, dtype=torch.float32, device=dev)
                conv3d_native(dy, wflip, None, out, am, B, T, Hd, Wd, Hpd, Wpd, Cin, CG, C, K, 2, (T + 2, Hpo, Wpo),
                              pad, tag=tag, swap=swap, d_bias=d_bias)
                return out, d_bias
            da = torch.empty((B, T, Hd, Wd, C), dtype=bf, device=dev)
            conv3d_native(dy, wflip, None, da, None, B, T, Hd, Wd, Hpd, Wpd, Cin, CG, C, K, 1, (T, Hd, Wd), (0, 0, 0),
                          tag=tag, swap=swap)
            return unpool(da, am, Hb, Wb, C, C, pad, Hpo, Wpo)

        # ---- layer 2: dgrad of conv3 fused with the un-pooling of pool2 -> dY2 (+ conv2's bias gradient) ----
        dy2, db2 = dgrad_unpool(dy3, dgrad_weight(w3.detach(), 32).to(bf), am2, H3, W3, Hp3, Wp3, 32, 3, 64, (3, 3, 3),
                                (1, 2, 2), H2, W2, Hp2, Wp2, "conv3.dgrad")
        d2 = conv3d_wgrad_native(a1, dy2, B, T, H2, W2, Hp2, Wp2, 32, 64, 1, (Hp2 + 2) * Wp2 + 2, (3, 5, 5), 0,
                                 tag="conv2.wgrad")
        dw2 = d2.reshape(3, 5, 5, 64, 32).permute(3, 4, 0, 1, 2)       # [tap][co][ci]
        # ---- layer 1 (no input gradient: the clip is data); dY1 top-left aligned in z's geometry ----
        dy1, db1 = dgrad_unpool(dy2, dgrad_weight(w2.detach(), 64).to(bf), am1, H2, W2, Hp2, Wp2, 64, 1, 32, (3, 5, 5),
                                (0, 0, 0), H1, W1, Hp1, Wp1, "conv2.dgrad", swap=2 if DGRAD_KX_STACK else None)
        d1 = conv3d_wgrad_native(z, dy1, B, T, H1, W1, Hp1, Wp1, 16, 32, 1, 0, (3, 3, 3), 0, tag="conv1.wgrad",
                                 algo_macs=B * T * H1 * W1 * 32 * 3 * 75)     # the true 3x5x5x3 stride-2 conv
        dw1_16 = d1[:, :32, :].permute(1, 0, 2).reshape(32, 3, 3, 3, 16)
        dw1 = s2d_weight_grad(dw1_16)
        return None, dw1.contiguous(), db1, dw2.contiguous(), db2, dw3.contiguous(), db3


class ConvFrontEnd(nn.Module):
    """clip (B,T,H,W,3) uint8 -> per-frame features (B,T,feature_dim(H,W)) float32."""

    def __init__(self, frame_hw=(100, 50)):
        super().__init__()
        for name, ci, co, k, s, p in LAYERS:
            setattr(self, name, nn.Conv3d(ci, co, k, stride=s, padding=p))   # parameter container + init
        self.out_features = feature_dim(*frame_hw)

    def forward(self, clip):
        return _ConvStack.apply(clip, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                self.conv3.weight, self.conv3.bias)
