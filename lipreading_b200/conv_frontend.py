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
N.register("lr_clip_s2d", _i, [_vp, _vp, _i, _i, _i, _i, _i, _vp])
N.register("lr_unpool", _i, [_vp, _vp, _vp] + [_i] * 12 + [_vp])
N.register("lr_conv3d_fwd", _i, [_vp, _vp, _vp, _vp, _vp] + [_i] * 19 + [_vp])

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


def feature_dim(H, W):
    h, w = H // 2, W // 2          # conv1 stride 2 (k5,p2)
    for _ in range(3):
        h, w = h // 2, w // 2      # three 2x2 pools
    return 96 * h * w


# bench.py sets this to a list to get per-launch CUDA-event timings of the conv kernel:
# entries are (tag, start_event, end_event, algorithmic_flops)
KERNEL_TIMING = None


def conv3d_native(x, w, bias, y, argmax, B, T, H, W, Wp, Cin, CG, Cout, K, epi_mode, ovol, ooff, J=0,
                  tag="conv", algo_macs=None):
    """Thin call into lr_conv3d_fwd (see include/lr_b200.h)."""
    rec = KERNEL_TIMING
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    N.check(N.lib().lr_conv3d_fwd(N.ptr(x), N.ptr(w), N.ptr(bias), N.ptr(y), N.ptr(argmax), B, T, H, W, Wp,
                                  Cin, CG, Cout, K[0], K[1], K[2], epi_mode, ovol[0], ovol[1], ovol[2],
                                  ooff[0], ooff[1], ooff[2], J, N.stream()), "lr_conv3d_fwd")
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
    wf = w.flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(ci, -1, co // cg, cg) if False else None
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    wf = w.flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(ci, taps, co // cg, cg)
    return wf.permute(0, 2, 1, 3).contiguous()


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
        z = torch.zeros((B, T + 2, H1 + 2, Wp1, 16), dtype=bf, device=dev)
        a1 = torch.zeros((B, T + 2, H2 + 4, Wp2, 32), dtype=bf, device=dev)
        a2 = torch.zeros((B, T + 2, H3 + 2, Wp3, 64), dtype=bf, device=dev)
        feat = torch.empty((B, T, H4, W4, 96), dtype=bf, device=dev)
        am1 = torch.empty((B, T, H2, W2, 32), dtype=torch.uint8, device=dev)
        am2 = torch.empty((B, T, H3, W3, 64), dtype=torch.uint8, device=dev)
        am3 = torch.empty((B, T, H4, W4, 96), dtype=torch.uint8, device=dev)
        N.check(L.lr_clip_s2d(N.ptr(N.cont(clip)), N.ptr(z), B, T, H, W, Wp1, N.stream()), "lr_clip_s2d")
        g1 = s2d_weight(w1.detach()).to(bf).contiguous()
        g2 = gemm_weight(w2.detach()).to(bf)
        g3 = gemm_weight(w3.detach()).to(bf)
        conv3d_native(z, g1, b1.detach().float(), a1, am1, B, T, H1, W1, Wp1, 16, 1, 32, (3, 3, 3), 0,
                      (T + 2, H2 + 4, Wp2), (1, 2, 2), tag="conv1.fwd",
                      algo_macs=B * T * H1 * W1 * 32 * 3 * 75)       # the true 3x5x5x3 stride-2 conv
        conv3d_native(a1, g2, b2.detach().float(), a2, am2, B, T, H2, W2, Wp2, 32, 1, 64, (3, 5, 5), 0,
                      (T + 2, H3 + 2, Wp3), (1, 1, 1), tag="conv2.fwd")
        conv3d_native(a2, g3, b3.detach().float(), feat, am3, B, T, H3, W3, Wp3, 64, 1, 96, (3, 3, 3), 0,
                      (T, H4, W4), (0, 0, 0), tag="conv3.fwd")
        ctx.save_for_backward(z, a1, a2, am1, am2, am3, w1, w2, w3)
        ctx.geom = (B, T, H, W)
        return feat.reshape(B, T, H4 * W4 * 96).float()

    @staticmethod
    def backward(ctx, d_feat):
        z, a1, a2, am1, am2, am3, w1, w2, w3 = ctx.saved_tensors
        B, T, H, W = ctx.geom
        dev = z.device
        bf = torch.bfloat16
        L = N.lib()
        H1, W1 = H // 2, W // 2
        H2, W2 = H1 // 2, W1 // 2
        H3, W3 = H2 // 2, W2 // 2
        H4, W4 = H3 // 2, W3 // 2
        Wp1, Wp2, Wp3 = z.shape[3], a1.shape[3], a2.shape[3]

        def unpool(dp, am, Hf, Wf, C, Cg, pad, Wp):
            out = torch.zeros((C // Cg, B, T + 2 * pad[0], Hf + 2 * pad[1], Wp, Cg), dtype=bf, device=dev)
            N.check(L.lr_unpool(N.ptr(dp), N.ptr(am), N.ptr(out), B, T, Hf, Wf, C, Cg, T + 2 * pad[0],
                                Hf + 2 * pad[1], Wp, pad[0], pad[1], pad[2], N.stream()), "lr_unpool")
            return out

        def interior(vol, pad, Hf, Wf):
            """(G,B,Tp,Hp,Wp,Cg) padded grouped volume -> (B,C,T,H,W) view-ish tensor for library calls."""
            G, _, _, _, _, Cg = vol.shape
            v = vol[:, :, pad[0]:pad[0] + T, pad[1]:pad[1] + Hf, pad[2]:pad[2] + Wf]
            return v.permute(1, 0, 5, 2, 3, 4).reshape(B, G * Cg, T, Hf, Wf)

        def act_interior(a, pad, Hf, Wf):
            return a[:, pad[0]:pad[0] + T, pad[1]:pad[1] + Hf, pad[2]:pad[2] + Wf].permute(0, 4, 1, 2, 3)

        # ---- layer 3 ----
        dp3 = N.cont(d_feat.reshape(B, T, H4, W4, 96).to(bf))
        dy3 = unpool(dp3, am3, H3, W3, 96, 32, (1, 1, 1), Wp3)                   # (3,B,T+2,H3+2,Wp3,32)
        dy3_n = interior(dy3, (1, 1, 1), H3, W3)
        # weight gradient: library call for now (cuDNN wgrad) — see DESIGN.md "interim"
        dw3 = torch.nn.grad.conv3d_weight(act_interior(a2, (1, 1, 1), H3, W3), w3.shape, dy3_n, padding=(1, 1, 1))
        db3 = dy3_n.float().sum((0, 2, 3, 4))
        da2 = torch.empty((B, T, H3, W3, 64), dtype=bf, device=dev)
        conv3d_native(dy3, dgrad_weight(w3.detach(), 32).to(bf), None, da2, None, B, T, H3, W3, Wp3, 32, 3, 64,
                      (3, 3, 3), 1, (T, H3, W3), (0, 0, 0), tag="conv3.dgrad")
        # ---- layer 2 ----
        dy2 = unpool(da2, am2, H2, W2, 64, 64, (1, 2, 2), Wp2)                   # (1,B,T+2,H2+4,Wp2,64)
        dy2_n = interior(dy2, (1, 2, 2), H2, W2)
        dw2 = torch.nn.grad.conv3d_weight(act_interior(a1, (1, 2, 2), H2, W2), w2.shape, dy2_n, padding=(1, 2, 2))
        db2 = dy2_n.float().sum((0, 2, 3, 4))
        da1 = torch.empty((B, T, H2, W2, 32), dtype=bf, device=dev)
        conv3d_native(dy2, dgrad_weight(w2.detach(), 64).to(bf), None, da1, None, B, T, H2, W2, Wp2, 64, 1, 32,
                      (3, 5, 5), 1, (T, H2, W2), (0, 0, 0), tag="conv2.dgrad")
        # ---- layer 1 (no input gradient: the clip is data) ----
        dy1 = unpool(da1, am1, H1, W1, 32, 32, (0, 0, 0), W1)                    # (1,B,T,H1,W1,32)
        dy1_n = interior(dy1, (0, 0, 0), H1, W1)
        dw1_16 = torch.nn.grad.conv3d_weight(act_interior(z, (1, 1, 1), H1, W1), (32, 16, 3, 3, 3), dy1_n,
                                             padding=(1, 1, 1))
        dw1 = s2d_weight_grad(dw1_16.permute(0, 2, 3, 4, 1))
        db1 = dy1_n.float().sum((0, 2, 3, 4))
        return None, dw1.float(), db1, dw2.float(), db2, dw3.float(), db3


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
