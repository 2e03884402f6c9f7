"""-m gpu parity tests of the tcgen05 conv front-end vs the fp32 conv3d oracle (oracle/conv3d.py).

Operands are bf16 on both sides (the oracle rounds where the CUDA path stores bf16), accumulation is
fp32 on both sides, so the bar is the bf16 output rounding: |err| <= 2^-7 relative to the tensor's
max (one bf16 ulp at the top of the range) for activations, 2e-2 for gradients that went through
two bf16 stages."""
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d as OC

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _padded(x_ndhwc, pad, Wp, Hp=None):
    """(B,T,H,W,C) -> zero padded (B,T+2pt,Hp,Wp,C) with the interior at (pt,ph,pw)."""
    B, T, H, W, C = x_ndhwc.shape
    Hp = Hp or H + 2 * pad[1]
    out = torch.zeros((B, T + 2 * pad[0], Hp, Wp, C), dtype=x_ndhwc.dtype, device=x_ndhwc.device)
    out[:, pad[0]:pad[0] + T, pad[1]:pad[1] + H, pad[2]:pad[2] + W] = x_ndhwc
    return out


@pytest.mark.parametrize("Cin,CG,Cout,K,H,W,Wp,T,B", [
    (32, 1, 64, (3, 5, 5), 25, 12, 16, 7, 2),      # conv2 geometry
    (64, 1, 96, (3, 3, 3), 12, 6, 8, 6, 3),        # conv3 geometry
    (16, 1, 32, (3, 3, 3), 50, 25, 32, 5, 1),      # conv1 (space-to-depth) geometry
    (32, 3, 64, (3, 3, 3), 12, 6, 8, 5, 2),        # grouped input channels (conv3 dgrad geometry)
    (64, 1, 32, (3, 5, 5), 25, 12, 16, 4, 1),      # conv2 dgrad geometry
    (32, 1, 32, (1, 1, 1), 8, 8, 8, 3, 1),         # 1x1x1: no shifts at all
    (32, 1, 32, (1, 1, 3), 8, 6, 8, 3, 1),         # x shifts only
    (32, 1, 32, (1, 3, 1), 8, 8, 8, 3, 1),         # y shifts only
    (64, 1, 32, (3, 5, 5), 25, 12, 16, 9, 3),      # conv2 dgrad geometry, several items per CTA
    (16, 1, 32, (3, 3, 3), 22, 25, 32, 7, 2),      # ragged last tile (H not a multiple of the tile rows)
    (16, 1, 32, (3, 3, 3), 12, 25, 32, 21, 2),     # several full frame groups + a tail group per clip
    (32, 1, 64, (3, 5, 5), 9, 12, 16, 11, 2),      # T = 2 full groups of 4 + tail of 3
    (64, 1, 96, (3, 3, 3), 12, 6, 8, 1, 2),        # single frame: every chunk is an edge chunk
    (32, 1, 64, (5, 3, 3), 8, 6, 8, 9, 1),         # KT = 5
])
@pytest.mark.parametrize("swap", [0, 1, 2, 3])    # 0 positions on M, 1 swapped, 2 kx-taps stacked on N (Cout = 32), 3 kt-taps stacked on N
def test_conv3d_plain_matches_torch(native_lib, cuda, Cin, CG, Cout, K, H, W, Wp, T, B, swap):
    from lipreading_b200.conv_frontend import conv3d_native, _plane_rows
    if swap == 2 and (Cout != 32 or K[2] < 2):
        pytest.skip("kx-stacking: Cout = 32 and KW >= 2 only")
    g = torch.Generator().manual_seed(1234)
    C = Cin * CG
    pad = tuple((k - 1) // 2 for k in K)
    Hp = _plane_rows(H, K[1], Wp)
    x = torch.randn(B, T, H, W, C, generator=g).to(BF)
    w = (torch.randn(Cout, C, *K, generator=g) / (C * K[0] * K[1] * K[2]) ** 0.5).to(BF)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), None, padding=pad).permute(0, 2, 3, 4, 1)
    xd = x.to(cuda)
    # channel-grouped padded volume [CG][B][Tp][Hp][Wp][Cin]
    vol = torch.stack([_padded(xd[..., gi * Cin:(gi + 1) * Cin], pad, Wp, Hp) for gi in range(CG)], 0).contiguous()
    # weights [Cout][CG][taps][Cin]
    wk = w.to(cuda).permute(0, 2, 3, 4, 1).reshape(Cout, -1, CG, Cin).permute(0, 2, 1, 3).contiguous()
    y = torch.full((B, T, H, W, Cout), float("nan"), dtype=BF, device=cuda)
    conv3d_native(vol, wk, None, y, None, B, T, H, W, Hp, Wp, Cin, CG, Cout, K, 1, (T, H, W), (0, 0, 0), swap=swap)
    torch.cuda.synchronize()
    err = (y.float().cpu() - ref).abs().max() / ref.abs().max()
    assert torch.isfinite(y.float()).all()
    assert float(err) < 2 ** -7, float(err)


@pytest.mark.parametrize("swap", [0, 1, 3])
@pytest.mark.parametrize("J", [0, 1, 2])
def test_conv3d_relu_pool_epilogue(native_lib, cuda, J, swap):
    from lipreading_b200.conv_frontend import conv3d_native, _plane_rows
    g = torch.Generator().manual_seed(77)
    B, T, H, W, Cin, Cout, K, Wp = 2, 5, 25, 12, 32, 64, (3, 5, 5), 16
    x = torch.randn(B, T, H, W, Cin, generator=g).to(BF)
    w = (torch.randn(Cout, Cin, *K, generator=g) / (Cin * 75) ** 0.5).to(BF)
    bias = torch.randn(Cout, generator=g) * 0.1
    conv = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), bias, padding=(1, 2, 2))
    act = F.relu(conv).to(BF).float()
    ref, idx = F.max_pool3d(act, (1, 2, 2), return_indices=True)
    ref = ref.permute(0, 2, 3, 4, 1)                                       # B,T,12,6,C
    Hp = _plane_rows(H, 5, Wp)
    vol = _padded(x.to(cuda), (1, 2, 2), Wp, Hp).unsqueeze(0).contiguous()
    wk = w.to(cuda).permute(0, 2, 3, 4, 1).contiguous()
    PH, PW = H // 2, W // 2
    out = torch.zeros((B, T + 2, PH + 2, 8, Cout), dtype=BF, device=cuda)   # next layer's padded volume
    am = torch.full((B, T, PH, PW, Cout), 255, dtype=torch.uint8, device=cuda)
    conv3d_native(vol, wk, bias.to(cuda), out, am, B, T, H, W, Hp, Wp, Cin, 1, Cout, K, 0, (T + 2, PH + 2, 8), (1, 1, 1), J, swap=swap)
    torch.cuda.synchronize()
    got = out[:, 1:1 + T, 1:1 + PH, 1:1 + PW].float().cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) < 2 ** -7
    # borders untouched
    assert float(out[:, 0].abs().max()) == 0 and float(out[:, :, 0].abs().max()) == 0 and float(out[:, :, :, 0].abs().max()) == 0
    assert float(out[:, :, :, 1 + PW:].abs().max()) == 0
    # arg-max bytes: value at the recorded position equals the pooled max (ties may pick either)
    am = am.cpu()
    assert int(am.max()) <= 4
    a = act.permute(0, 2, 3, 4, 1)                                         # B,T,H,W,C
    dead = am == 4
    assert bool(((ref <= 0) == dead).all())
    yy = (torch.arange(PH) * 2).view(1, 1, PH, 1, 1) + (am.clamp(max=3) // 2)
    xx = (torch.arange(PW) * 2).view(1, 1, 1, PW, 1) + (am.clamp(max=3) % 2)
    bb = torch.arange(B).view(B, 1, 1, 1, 1).expand_as(am)
    tt = torch.arange(T).view(1, T, 1, 1, 1).expand_as(am)
    cc = torch.arange(Cout).view(1, 1, 1, 1, Cout).expand_as(am)
    picked = a[bb, tt, yy, xx, cc]
    assert bool((picked[~dead] == ref[~dead]).all())


@pytest.mark.parametrize("Cin,CG,Cout,K,H,W,Wp", [
    (16, 1, 32, (3, 3, 3), 50, 25, 32),            # conv1
    (32, 1, 64, (3, 5, 5), 25, 12, 16),            # conv2
    (64, 1, 96, (3, 3, 3), 12, 6, 8),              # conv3
    (32, 3, 64, (3, 3, 3), 12, 6, 8),              # conv3 dgrad
    (64, 1, 32, (3, 5, 5), 25, 12, 16),            # conv2 dgrad
])
@pytest.mark.parametrize("deterministic", [True, False])
def test_conv3d_kt_stacking_agrees_at_scale(native_lib, cuda, Cin, CG, Cout, K, H, W, Wp, deterministic):
    """Orientation 3 issues differently shaped MMAs onto overlapping accumulator columns; a mis-ordered or lost
    update would show up as a whole missing tap.  Compare with the one-MMA-per-tap orientation on a batch that
    keeps every SM busy for many work items (same operands, fp32 accumulation in a different order)."""
    from lipreading_b200 import conv_frontend as CF
    from lipreading_b200.conv_frontend import conv3d_native, _plane_rows
    g = torch.Generator(device="cuda").manual_seed(99)
    B, T = 24, 75
    pad = tuple((k - 1) // 2 for k in K)
    Hp = _plane_rows(H, K[1], Wp)
    vol = torch.zeros((CG, B, T + 2 * pad[0], Hp, Wp, Cin), dtype=BF, device=cuda)
    vol[:, :, pad[0]:pad[0] + T, pad[1]:pad[1] + H, pad[2]:pad[2] + W] = torch.randn(
        (CG, B, T, H, W, Cin), generator=g, device=cuda).to(BF)
    taps = K[0] * K[1] * K[2]
    wk = (torch.randn((Cout, CG, taps, Cin), generator=g, device=cuda) / (Cin * CG * taps) ** 0.5).to(BF)
    ys = []
    CF.DETERMINISTIC = deterministic
    try:
        for mode in (0, 3, 3):
            y = torch.full((B, T, H, W, Cout), float("nan"), dtype=BF, device=cuda)
            conv3d_native(vol, wk, None, y, None, B, T, H, W, Hp, Wp, Cin, CG, Cout, K, 1, (T, H, W), (0, 0, 0), swap=mode)
            ys.append(y.float())
        torch.cuda.synchronize()
    finally:
        CF.DETERMINISTIC = False
    assert torch.isfinite(ys[1]).all()
    if deterministic:
        assert torch.equal(ys[1], ys[2])                   # every accumulator has one writer: run-to-run identical
    else:
        # shared chunk list: hand-over frames are summed in a timing-dependent order -> at most a bf16 rounding flip
        assert float((ys[1] - ys[2]).abs().max() / ys[1].abs().max()) < 2 ** -7
    assert float((ys[0] - ys[1]).abs().max() / ys[0].abs().max()) < 2 ** -8


@pytest.mark.parametrize("name,Cx,Cy,Gy,K,H,W,Wp,m_is_x,B,T", [
    ("conv2", 32, 64, 1, (3, 5, 5), 25, 12, 16, 0, 2, 5),
    ("conv3", 64, 32, 3, (3, 3, 3), 12, 6, 8, 1, 3, 4),
    ("conv1", 16, 32, 1, (3, 3, 3), 50, 25, 32, 0, 1, 3),
])
@pytest.mark.parametrize("stack", [0, 1, 2])     # 0: one MMA per tap; 1: kx on N; 2: + ky on M=128 (+ fused kt)
def test_conv3d_wgrad_matches_autograd(native_lib, cuda, name, Cx, Cy, Gy, K, H, W, Wp, m_is_x, B, T, stack):
    from lipreading_b200.conv_frontend import conv3d_wgrad_native, _plane_rows
    g = torch.Generator().manual_seed(4321)
    pad = tuple((k - 1) // 2 for k in K)
    Co = Cy * Gy
    Hp = _plane_rows(H, K[1], Wp)
    x = torch.randn(B, T, H, W, Cx, generator=g).to(BF)
    dy = torch.randn(B, T, H, W, Co, generator=g).to(BF)
    w = torch.zeros(Co, Cx, *K, requires_grad=True)
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, None, padding=pad)
    y.backward(dy.float().permute(0, 4, 1, 2, 3))
    ref = w.grad                                                        # (Co,Cx,kt,ky,kx)
    xv = _padded(x.to(cuda), pad, Wp, Hp).contiguous()
    if name == "conv1":            # top-left aligned gradient volume (no interior offset)
        dyv = torch.zeros((1, B, T + 2, Hp, Wp, Co), dtype=BF, device=cuda)
        dyv[0, :, :T, :H, :W] = dy.to(cuda)
        off = 0
    else:
        dyd = dy.to(cuda)
        dyv = torch.stack([_padded(dyd[..., gi * Cy:(gi + 1) * Cy], pad, Wp, Hp) for gi in range(Gy)], 0).contiguous()
        off = (pad[0] * Hp + pad[1]) * Wp + pad[2]
    out = conv3d_wgrad_native(xv, dyv, B, T, H, W, Hp, Wp, Cx, Cy, Gy, off, K, m_is_x, stack_kx=stack >= 1,
                              stack_ky=stack == 2)
    torch.cuda.synchronize()
    out = out.cpu().reshape(K[0], K[1], K[2], out.shape[1], -1)
    got = out.permute(4, 3, 0, 1, 2) if m_is_x else out[:, :, :, :Co].permute(3, 4, 0, 1, 2)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-3


def test_conv_stack_forward_backward_matches_oracle(native_lib, cuda):
    from lipreading_b200.conv_frontend import ConvFrontEnd, feature_dim
    torch.manual_seed(123456)
    B, T, H, W = 2, 6, 100, 50
    front = ConvFrontEnd((H, W)).to(cuda)
    assert front.out_features == feature_dim(H, W) == 1728
    clip = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in front.state_dict().items()}
    feat_ref, _ = OC.stcnn_forward(clip, params, quantize=True)
    up = torch.randn(feat_ref.shape)
    (feat_ref * up).sum().backward()
    feat = front(clip.to(cuda))
    (feat * up.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    assert feat.shape == (B, T, 1728)
    assert float((feat.cpu() - feat_ref.detach()).abs().max() / feat_ref.abs().max()) < 2 ** -6
    for name, p in front.named_parameters():
        r = params[name].grad
        err = float((p.grad.cpu() - r).abs().max() / (r.abs().max() + 1e-12))
        assert err < 3e-2, (name, err)


@pytest.mark.parametrize("B,T", [(2, 6), (3, 11)])
def test_fused_dgrad_unpool_equals_two_pass(native_lib, cuda, B, T):
    """lr_conv3d_dgrad_unpool (dgrad epilogue routes each pooled gradient to its arg-max slot and sums the bias
    gradient) vs lr_conv3d_fwd(epi_mode 1) + lr_unpool: the dY volumes are bit-identical, so the weight gradients of
    the layers below are too; bias gradients differ only by fp32 summation order."""
    from lipreading_b200 import conv_frontend as CF
    torch.manual_seed(7)
    front = CF.ConvFrontEnd((100, 50)).to(cuda)
    clip = torch.randint(0, 256, (B, T, 100, 50, 3), dtype=torch.uint8).to(cuda)
    up = torch.randn(B, T, 1728, device=cuda)
    grads, vols = {}, {}
    for fused in (False, True):
        CF.FUSE_UNPOOL = fused
        CF.DETERMINISTIC = True             # bit-for-bit comparison: single-writer accumulators
        try:
            front.zero_grad()
            (front(clip) * up).sum().backward()
            torch.cuda.synchronize()
        finally:
            CF.FUSE_UNPOOL = True
            CF.DETERMINISTIC = False
        grads[fused] = {k: v.grad.clone() for k, v in front.named_parameters()}
        vols[fused] = {k: v[0][: int(torch.tensor(v[1]).prod())].clone() for k, v in CF.POOL.bufs.items()
                       if k[0] in ("dy32", "dy64") and v[1][1] == B and v[1][2] == T + 2}
    assert len(vols[True]) == 2
    for k in vols[True]:
        assert torch.equal(vols[True][k], vols[False][k]), k[0]
        assert float(vols[True][k].abs().max()) > 0
    for k in ("conv1.weight", "conv2.weight", "conv3.weight"):
        assert torch.equal(grads[True][k], grads[False][k]), k
    for k in ("conv1.bias", "conv2.bias", "conv3.bias"):          # atomically accumulated in both paths
        ref = grads[False][k]
        assert float((grads[True][k] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-6, k


@pytest.mark.parametrize("H,W,C,Cg,pad", [(50, 25, 32, 32, (0, 0, 0)), (25, 12, 64, 64, (1, 2, 2)), (12, 6, 96, 32, (1, 1, 1))])
def test_unpool_routes_gradient_to_argmax(native_lib, cuda, H, W, C, Cg, pad):
    """lr_unpool vs direct indexing: the arg-max position of each 2x2 window receives the pooled gradient,
    everything else (incl. ReLU-dead units, code 4, and the odd last row/column) stays zero; d_bias = sum."""
    from lipreading_b200 import native as N
    g = torch.Generator().manual_seed(77)
    B, T = 2, 5
    PH, PW = H // 2, W // 2
    Wp = 8
    while Wp < W + 2 * pad[2]:
        Wp *= 2
    Hp = H + 2 * pad[1] + 1
    dp = torch.randn(B, T, PH, PW, C, generator=g).to(BF)
    am = torch.randint(0, 5, (B, T, PH, PW, C), generator=g).to(torch.uint8)
    out = torch.zeros(C // Cg, B, T + 2, Hp, Wp, Cg, dtype=BF, device=cuda)
    db = torch.empty(C, dtype=torch.float32, device=cuda)
    dpc, amc = dp.to(cuda), am.to(cuda)
    N.check(N.lib().lr_unpool(N.ptr(dpc), N.ptr(amc), N.ptr(out), N.ptr(db), B, T, H, W, C, Cg, T + 2, Hp, Wp,
                              pad[0], pad[1], pad[2], N.stream()), "lr_unpool")
    torch.cuda.synchronize()
    ref = torch.zeros(B, T, H, W, C)
    for w in range(4):
        sel = (am == w)
        ref[:, :, (w >> 1):2 * PH:2, (w & 1):2 * PW:2, :] = torch.where(sel, dp.float(), torch.zeros(()))
    full = torch.zeros(C // Cg, B, T + 2, Hp, Wp, Cg)
    for gi in range(C // Cg):
        full[gi, :, pad[0]:pad[0] + T, pad[1]:pad[1] + H, pad[2]:pad[2] + W] = ref[..., gi * Cg:(gi + 1) * Cg]
    assert torch.equal(out.float().cpu(), full)
    ref_b = torch.where(am < 4, dp.float(), torch.zeros(())).sum((0, 1, 2, 3))
    assert float((db.cpu() - ref_b).abs().max()) < 1e-3
