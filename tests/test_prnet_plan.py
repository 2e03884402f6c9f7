"""CPU tests of the launch plan `prnet_tc5.compile_plan` builds for the position-map CNN (rows a5 / f4): every launch
is replayed by oracle/tapgemm.py (the CPU statement of lr_tapgemm) and each stem / resBlock / transposed-conv output is
held to `prnet.ResFcn256.forward` — the restatement of the reference's resfcn256 (src/models/face/prnet.py:211-280).
Tolerance: bf16 storage of every activation (2^-8 relative per layer, accumulating over 53 launches): 2e-2 of the
layer's largest value."""
import pytest
import torch

from lipreading_b200 import prnet_tc5
from lipreading_b200.prnet import ResFcn256
from oracle import tapgemm as OT


def randomized_net(seed=0):
    torch.manual_seed(seed)
    net = ResFcn256().eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):          # non-trivial inference statistics: the folding is exercised
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                m.weight.mul_(2.0)                             # keep activations O(1) through 28 layers
    return net


def layer_outputs(net, x_nhwc):
    acts = {}

    def hook(name):
        def f(mod, inp, out):
            acts[name] = out.detach().permute(0, 2, 3, 1)
        return f
    hs = [net.stem.register_forward_hook(hook("stem"))]
    hs += [b.register_forward_hook(hook("enc%d" % i)) for i, b in enumerate(net.enc)]
    hs += [d.register_forward_hook(hook("dec%d" % i)) for i, d in enumerate(net.dec)]
    with torch.no_grad():
        y = net(x_nhwc.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    for h in hs:
        h.remove()
    return acts, y


# (PACK_POSITIONS, FUSE_TAPS, RESIDENT); "streamed": the packed layout with every tap's tiles streamed through the ring
LAYOUTS = {"packed": (True, True, True), "fused": (False, True, True), "plain": (False, False, True),
           "streamed": (True, True, False)}


def compile_with_layout(layout, net, B, R, device):
    """the three ways a C < 64 layer can be laid out for the kernel (prnet_tc5._groups); packed is the default"""
    saved = prnet_tc5.PACK_POSITIONS, prnet_tc5.FUSE_TAPS, prnet_tc5.RESIDENT
    prnet_tc5.PACK_POSITIONS, prnet_tc5.FUSE_TAPS, prnet_tc5.RESIDENT = LAYOUTS[layout]
    try:
        return prnet_tc5.compile_plan(net, B, R, device)
    finally:
        prnet_tc5.PACK_POSITIONS, prnet_tc5.FUSE_TAPS, prnet_tc5.RESIDENT = saved


@pytest.mark.parametrize("layout", ["packed", "fused", "plain"])
def test_plan_replayed_on_cpu_matches_resfcn256(layout):
    net = randomized_net()
    B, R = 2, 32
    x = torch.rand(B, R, R, 3, generator=torch.Generator().manual_seed(1))
    plan = compile_with_layout(layout, net, B, R, "cpu")
    assert any(s["pack"] > 1 for s in plan.specs) == (layout == "packed")
    assert len(plan.specs) == 53                       # 1 stem + 5*4 + 5*3 resBlock launches + 17 transposed convs
    y = OT.run_plan(plan, x)
    acts, ref = layer_outputs(net, x)
    assert [n for n, _ in plan.marks] == list(acts)[:len(plan.marks)]
    for name, vol in plan.marks:
        r = acts[name]
        got = vol.interior()[..., :r.shape[-1]].float()
        assert float((got - r).abs().max()) <= 2e-2 * float(r.abs().max()), name
        assert float(vol.interior()[..., r.shape[-1]:].abs().max() if vol.C > r.shape[-1] else 0.0) == 0.0     # padded channels
    assert float((y - ref * R * 1.1).abs().max()) <= 2e-2 * R * 1.1
    # borders of every volume stay zero (they ARE the conv padding)
    for _, vol in plan.marks:
        v = vol.t[:vol.rows].view(vol.B, vol.Hp, vol.Wp, vol.C).float()
        inner = torch.zeros_like(v, dtype=torch.bool)
        inner[:, prnet_tc5.P:prnet_tc5.P + vol.H, prnet_tc5.P:prnet_tc5.P + vol.W] = True
        assert float(v[~inner].abs().max()) == 0.0


def test_fused_tap_groups_cover_each_tap_once():
    """C = 16: the four x-taps of a filter row share one K = 64 tile; C = 32: two tiles per row; C >= 64: one per tap."""
    w = torch.randn(16, 16, 4, 4)
    taps = prnet_tc5._conv_taps(w, 4, False)
    saved = prnet_tc5.PACK_POSITIONS
    prnet_tc5.PACK_POSITIONS = False
    try:
        offs, mats, Kg, Kt, pack = prnet_tc5._groups(taps, 16, 100, 16)
        assert (Kg, Kt, len(offs), pack) == (64, 64, 4, 1) and offs == [(ky - 1) * 100 - 1 for ky in range(4)]
        for ky in range(4):
            for kx in range(4):
                assert torch.equal(mats[ky][:, kx * 16:(kx + 1) * 16], w[:, :, ky, kx])
        w32 = torch.randn(32, 32, 4, 4)
        offs, mats, Kg, Kt, pack = prnet_tc5._groups(prnet_tc5._conv_taps(w32, 4, False), 32, 100, 32)
        assert len(offs) == 8 and Kg == 64
    finally:
        prnet_tc5.PACK_POSITIONS = saved
    # packed: rows of 4 positions; taps dx = -1..2 reach the row groups -1, 0, +1 -> 3 aligned K tiles per filter row, and
    # block (j, i) of group go holds W[dx = 4*go + i - j]
    offs, mats, Kg, Kt, pack = prnet_tc5._groups(taps, 16, 100, 16)
    assert (Kg, Kt, pack, len(offs)) == (64, 64, 4, 12) and offs[:3] == [-25 - 1, -25, -25 + 1]
    for ky in range(4):
        for g, go in enumerate((-1, 0, 1)):
            m = mats[3 * ky + g]
            for j in range(4):
                for i in range(4):
                    dx = 4 * go + i - j
                    blk = m[j * 16:(j + 1) * 16, i * 16:(i + 1) * 16]
                    assert torch.equal(blk, w[:, :, ky, dx + 1]) if -1 <= dx <= 2 else float(blk.abs().max()) == 0.0
    offs, mats, Kg, Kt, pack = prnet_tc5._groups(prnet_tc5._conv_taps(torch.randn(64, 64, 4, 4), 4, False), 64, 100, 64)
    assert len(offs) == 16 and Kg == 64 and Kt == 64 and pack == 1
    # transposed stride-2 conv: 4 phases x 4 taps, each kernel element used exactly once
    wt = torch.randn(32, 16, 4, 4)
    phases = prnet_tc5._deconv_phases(wt, 2)
    assert len(phases) == 4 and all(len(p) == 4 for p in phases)
    total = sum(float(m.abs().sum()) for p in phases for _, _, m in p)
    assert abs(total - float(wt.abs().sum())) < 1e-3 * total


def test_predict_batch_runs_any_batch_on_power_of_two_plans():
    """PosPrediction (tcgen05 engine) compiles plans for power-of-two batches only and runs any other batch as its binary
    decomposition, so ragged tail batches do not each compile (and keep) a plan of their own."""
    from lipreading_b200 import prnet as P
    pred = P.PosPrediction(device="cpu", engine="tcgen05", resolution_inp=32, resolution_op=32)
    seen = []

    class FakePlan:
        def __init__(self, b):
            self.b = b

        def run(self, x):
            assert x.shape[0] == self.b and x.is_contiguous()
            seen.append(self.b)
            return x[..., :3] * 2.0

    pred.plan = lambda b: FakePlan(b)
    for n, want in ((1, [1]), (64, [64]), (37, [32, 4, 1]), (24, [16, 8]), (100, [64, 32, 4]), (63, [32, 16, 8, 4, 2, 1])):
        del seen[:]
        x = torch.rand(n, 32, 32, 3)
        y = pred.predict_batch(x)
        assert seen == want and torch.equal(y, x * 2.0)
    assert pred.predict_batch(torch.zeros(0, 32, 32, 3)).shape == (0, 32, 32, 3)
    out = pred.predict(torch.rand(32, 32, 3).numpy())                  # numpy in -> numpy out, like the reference
    assert out.shape == (32, 32, 3) and out.dtype.name == "float32"
