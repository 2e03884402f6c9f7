"""Position-map CNN body on the tcgen05 "tap GEMM" kernel (reference
src/models/face/prnet.py:211-280, called per frame at :305-309).

`compile_plan(net, batch, resolution, device)` turns a `prnet.ResFcn256` (weights + inference batch-norm statistics)
into a list of `lr_tapgemm` launches over zero-padded channels-last bf16 volumes:

  * every activation is a volume (B, H+4, W+4, C) whose 2-pixel border stays zero — TF 'SAME' padding of the 4x4
    convs (1 before / 2 after) and of the stride-1 transposed convs (2 before / 1 after) is then a row offset;
  * batch-norm is folded into the epilogue (alpha = gamma/sqrt(var+eps), beta = beta - mean*alpha, kept in fp32);
  * resBlock: shortcut (1x1, raw) and main path meet in the epilogue of the last 1x1 conv:
    relu(alpha*acc + beta + alpha*shortcut)  ==  relu(bn(conv + shortcut))   (prnet.py:222-226);
  * a stride-2 4x4 conv reads the space-to-depth volume its 1x1 predecessor wrote (store mode 2): out(y,x) =
    sum_{ky,kx} w[ky,kx] in(2y+ky-1, 2x+kx-1) and pixel row r lives in block (r+1)>>1, slot (r+1)&1, so the conv
    is 2x2 taps over 4C channels; the stride-2 1x1 shortcut reads the even-subsampled copy the PREVIOUS layer's
    epilogue wrote next to its normal output;
  * a stride-2 transposed conv (Y = 2y - 1 + ky) is four output phases of 2x2 taps: phase 0 uses (ky=1, dy=0),
    (ky=3, dy=-1); phase 1 uses (ky=0, dy=+1), (ky=2, dy=0); same along x;
  * layers with C < 64 channels fuse 64/C horizontally adjacent taps into one K = 64 tile (see tapgemm_sm100.cu).

The plan is plain data (`spec` dicts): tests/ replay it on the CPU with oracle/tapgemm.py to hold the plan logic to
`ResFcn256.forward`; on the GPU `Plan.run` launches the kernels (no torch ops in between).
"""
import ctypes

import torch

from . import native

P = 2                                    # border of every normal volume
FUSE_TAPS = True                         # C < 64: fuse 64/C adjacent x-taps into one 128-byte K tile
RESIDENT = True                          # let the kernel keep a layer's weights + one input chunk per tile in shared memory
PACK_POSITIONS = True                    # C < 64: a matrix row = 64/C adjacent positions (block-Toeplitz weights), see _groups


_Desc = native.TapGemmDesc


def _pad16(c):
    return max(16, (c + 15) // 16 * 16)


class Volume:
    """Zero-initialised channels-last bf16 volume (B, Hp, Wp, C) stored as [rows + slack][C]; `s2d`: the odd-phase
    space-to-depth grid (H/2+1, W/2+1) of an (H, W) image with 4*c channels and no border."""

    def __init__(self, B, H, W, C, device, s2d=False):
        self.B, self.H, self.W, self.C, self.s2d = B, H, W, C, s2d
        if s2d:
            self.Hp, self.Wp, self.pad = H // 2 + 1, W // 2 + 1, 0
        else:
            self.Hp, self.Wp, self.pad = H + 2 * P, W + 2 * P, P
        self.rows = B * self.Hp * self.Wp
        self.t = torch.zeros(self.rows + 8, C, dtype=torch.bfloat16, device=device)     # slack: fused-tap rows overrun

    def interior(self):
        v = self.t[:self.rows].view(self.B, self.Hp, self.Wp, self.C)
        return v[:, P:P + self.H, P:P + self.W] if not self.s2d else v


def _groups(taps, C, Wp, cout_pad):
    """taps: [(dy, dx, W[cout, cin<=C] f32)] -> (row offsets, weight [n_groups, pack*cout_pad, Kg] f32, Kg, Kt, pack).

    C >= 64: one group per tap, K over the channels of one position.
    C < 64, packed (default): a matrix row holds pack = 64/C horizontally adjacent positions p*pack .. p*pack+pack-1 and the
    accumulator row their pack*cout_pad outputs.  Output j of a row needs input position p*pack + j + dx, which lives in
    row p + go at slot i with go*pack + i = j + dx: per filter row one group per needed `go`, whose weight block
    [(j, n), (i, c)] is W[dx = go*pack + i - j] (zero when that is no tap) — every K tile is an ALIGNED 128-byte row
    (one TMA request per row; the overlapping-row fusion below costs two), and a tile covers 128*pack positions.
    C < 64, fused: the 128-byte row of position q runs over positions q .. q+64/C-1 (row pitch = position pitch).
    """
    def padded(w):
        out = torch.zeros(cout_pad, C)
        out[:w.shape[0], :w.shape[1]] = w
        return out
    pk = 64 // C if C < 64 else 1
    if C < 64 and PACK_POSITIONS and Wp % pk == 0 and pk * cout_pad <= 256:
        offs, mats = [], []
        for dy in sorted({t[0] for t in taps}):
            row = {dx: padded(w) for y, dx, w in taps if y == dy}
            for go in sorted({(dx + j) // pk for dx in row for j in range(pk)}):
                m = torch.zeros(pk * cout_pad, 64)
                for j in range(pk):
                    for i in range(pk):
                        dx = go * pk + i - j
                        if dx in row:
                            m[j * cout_pad:(j + 1) * cout_pad, i * C:(i + 1) * C] = row[dx]
                offs.append(dy * (Wp // pk) + go)
                mats.append(m)
        return offs, torch.stack(mats), 64, 64, pk
    single = len({(t[0]) for t in taps}) == len(taps)          # one tap per row (1x1 convs): nothing to fuse, and a
    if C >= 64 or not FUSE_TAPS or single:                    # K = 64 tile would fetch 64/C times the bytes it needs
        Kt = 64 if C >= 64 else C
        offs = [dy * Wp + dx for dy, dx, _ in taps]
        return offs, torch.stack([padded(w) for _, _, w in taps]), C, Kt, 1
    f = 64 // C
    offs, mats = [], []
    for dy in sorted({t[0] for t in taps}):
        row = {dx: w for y, dx, w in taps if y == dy}
        todo = sorted(row)
        while todo:
            dx0 = todo[0]
            m = torch.zeros(cout_pad, 64)
            for j in range(f):
                if dx0 + j in row:
                    m[:, j * C:(j + 1) * C] = padded(row[dx0 + j])
            todo = [d for d in todo if d >= dx0 + f]
            offs.append(dy * Wp + dx0)
            mats.append(m)
    return offs, torch.stack(mats), 64, 64, 1


class Plan:
    def __init__(self, B, R, device):
        self.B, self.R, self.device = B, R, torch.device(device)
        self.specs = []                  # launch descriptions (plain dicts of tensors / ints)
        self._descs = None
        self.vin = Volume(B, R, R, 16, device)
        self.out = torch.zeros(B, R, R, 3, dtype=torch.float32, device=device)
        self.marks = []                  # (name, Volume) of the stem / resBlock / transposed-conv outputs, in network order
        self.model_flops = 0             # 2 * MACs of the network itself (no padding, no zero-weight fused taps), per batch
        self.flops = 0                   # MMA work actually issued (2 * rows_valid * Kg * groups * Cout_pad), per batch

    # ---- building -------------------------------------------------------------------------------
    def add(self, vin, phases, cout, mode, out, alpha=None, beta=None, gamma=None, act=0, res=None, aux=None,
            out_scale=1.0):
        """phases: list (1, or 4 in mode 1) of tap lists [(dy, dx, W[cout, cin])] in the input volume's grid."""
        dev = self.device
        cp = _pad16(cout)
        packed = [_groups(t, vin.C, vin.Wp, cp) for t in phases]
        n_groups = max(len(pk[0]) for pk in packed)
        Kg, Kt, pack = packed[0][2], packed[0][3], packed[0][4]
        w = torch.zeros(len(phases), n_groups, pack * cp, Kg)
        offs = []
        for i, (o, m, _, _, _) in enumerate(packed):
            w[i, :len(o)] = m
            offs += o + [0] * (n_groups - len(o))

        def vec(v, fill):
            if v is None:
                return None
            o = torch.full((cp,), fill, dtype=torch.float32)
            o[:cout] = v.detach().float().cpu()
            return o.to(dev)
        spec = {"a": vin, "w": w.reshape(-1, Kg).to(torch.bfloat16).to(dev), "Kg": Kg, "Kt": Kt, "n_phases": len(phases),
                "n_groups": n_groups, "tap_off": offs, "pack": pack, "flags": 0 if RESIDENT else 1, "Cout_pad": cp, "Cout": cout, "alpha": vec(alpha, 0.0),
                "beta": vec(beta, 0.0), "gamma": vec(gamma, 0.0), "act": act, "mode": mode, "out": out, "res": res,
                "aux": aux, "out_scale": out_scale,
                "valid": (0, 0, vin.H // 2, vin.W // 2) if vin.s2d else (P, P, vin.H, vin.W)}
        # (padded channels: zero weights, alpha = beta = 0 -> ReLU gives 0; the sigmoid layer stores only Cout channels)
        self.specs.append(spec)
        self.model_flops += 2 * vin.B * spec["valid"][2] * spec["valid"][3] * sum(int((m != 0).sum()) for t in phases for _, _, m in t)
        self.flops += 2 * vin.B * spec["valid"][2] * spec["valid"][3] * Kg * n_groups * len(phases) * cp     # (pack cancels)
        return out

    # ---- running on the GPU ---------------------------------------------------------------------
    def _desc(self, s):
        d = _Desc()
        a = s["a"]
        d.a, d.rows, d.C, d.Hp, d.Wp = a.t.data_ptr(), a.rows // s["pack"], a.C, a.Hp, a.Wp
        d.pack, d.flags = s["pack"], s["flags"]
        d.vy0, d.vx0, d.H, d.W = s["valid"]
        d.w, d.w_pitch, d.Kg, d.Kt = s["w"].data_ptr(), s["Kg"], s["Kg"], s["Kt"]
        d.n_phases, d.n_groups = s["n_phases"], s["n_groups"]
        s["_off"] = (ctypes.c_int32 * len(s["tap_off"]))(*s["tap_off"])
        d.tap_off = ctypes.cast(s["_off"], ctypes.POINTER(ctypes.c_int32))
        d.Cout_pad, d.Cout = s["Cout_pad"], s["Cout"]
        for k in ("alpha", "beta", "gamma"):
            setattr(d, k, s[k].data_ptr() if s[k] is not None else None)
        d.act, d.mode = s["act"], s["mode"]
        o = s["out"]
        if isinstance(o, Volume):
            d.out, d.oHp, d.oWp, d.oC, d.opy, d.opx = o.t.data_ptr(), o.Hp, o.Wp, o.C, o.pad, o.pad
        else:
            d.out, d.oC = o.data_ptr(), o.shape[-1]
        if s["res"] is not None:
            d.res, d.resC = s["res"].t.data_ptr(), s["res"].C
        if s["aux"] is not None:
            x = s["aux"]
            d.aux, d.aHp, d.aWp, d.aC, d.apad = x.t.data_ptr(), x.Hp, x.Wp, x.C, x.pad
        d.out_scale = s["out_scale"]
        return d

    def run(self, images):
        """images (B,R,R,3) f32 CUDA in [0,1] -> position map (B,R,R,3) f32 (already * MaxPos)."""
        native.require_cuda(images)
        lib = native.lib()
        assert tuple(images.shape) == (self.B, self.R, self.R, 3) and images.dtype == torch.float32
        if self._descs is None:
            self._descs = [self._desc(s) for s in self.specs]
        st = native.stream()
        native.check(lib.lr_pack_image16(native.ptr(native.cont(images)), native.ptr(self.vin.t), self.B, self.R, self.R,
                                         P, st), "lr_pack_image16")
        for d in self._descs:
            native.check(lib.lr_tapgemm(ctypes.byref(d), st), "lr_tapgemm")
        return self.out


# ------------------------------------------------------------------------------------------------
def _bn(bn):
    a = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return a, bn.bias.detach().float() - bn.running_mean.detach().float() * a


def _conv_taps(w, k, s2d):
    """torch conv weight (cout, cin, k, k), TF 'SAME' -> tap list.  s2d: the stride-2 4x4 conv over the odd-phase
    space-to-depth grid (2x2 taps, channel slot (ky&1)*2 + (kx&1))."""
    w = w.detach().float().cpu()
    cout, cin = w.shape[:2]
    if k == 1:
        return [(0, 0, w[:, :, 0, 0])]
    if not s2d:
        return [(ky - 1, kx - 1, w[:, :, ky, kx]) for ky in range(4) for kx in range(4)]
    taps = []
    for dY in range(2):
        for dX in range(2):
            cp = _pad16(cin)                              # slot pitch = the producing layer's padded channel count
            m = torch.zeros(cout, 4 * cp)
            for sy in range(2):
                for sx in range(2):
                    slot = sy * 2 + sx
                    m[:, slot * cp:slot * cp + cin] = w[:, :, 2 * dY + sy, 2 * dX + sx]
            taps.append((dY, dX, m))
    return taps


_PHASE_TAPS = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}     # output parity -> ((k index, input offset), ...)


def _deconv_phases(w, stride):
    """torch ConvTranspose2d weight (cin, cout, 4, 4) -> phases of tap lists (TF conv2d_transpose 'SAME')."""
    w = w.detach().float().cpu()
    if stride == 1:
        wf = w.flip(2, 3).transpose(0, 1)                # (cout, cin, 4, 4): plain conv, padded (2 before, 1 after)
        return [[(ky - 2, kx - 2, wf[:, :, ky, kx]) for ky in range(4) for kx in range(4)]]
    phases = []
    for py in range(2):
        for px in range(2):
            phases.append([(dy, dx, w[:, :, ky, kx].t()) for ky, dy in _PHASE_TAPS[py] for kx, dx in _PHASE_TAPS[px]])
    return phases


def compile_plan(net, batch, resolution=256, device="cuda", max_pos=None):
    """ResFcn256 (eval) -> Plan for `batch` images of `resolution`^2 (a multiple of 32)."""
    assert resolution % 32 == 0
    dev = torch.device(device)
    plan = Plan(batch, resolution, dev)
    B, R = batch, resolution
    max_pos = resolution * 1.1 if max_pos is None else max_pos

    def vol(H, C, s2d=False):
        return Volume(B, H, H, C, dev, s2d=s2d)

    strides = [b.c1.stride for b in net.enc]
    # stem (prnet.py:244)
    a, b_ = _bn(net.stem.bn)
    x = vol(R, 16)
    x_sub = vol(R // 2, 16) if strides[0] == 2 else None
    plan.add(plan.vin, [_conv_taps(net.stem.conv.weight, 4, False)], 16, 0, x, alpha=a, beta=b_, act=1, aux=x_sub)
    plan.marks.append(("stem", x))
    H = R
    for i, blk in enumerate(net.enc):
        cin, cout, s = blk.c0.conv.in_channels, blk.c2.conv.out_channels, blk.c1.stride
        half = cout // 2
        a0, b0 = _bn(blk.c0.bn)
        a1, b1 = _bn(blk.c1.bn)
        a2, b2 = _bn(blk.bn)
        Ho = H // s
        nxt_sub = vol(Ho // 2, _pad16(cout)) if i + 1 < len(net.enc) and strides[i + 1] == 2 else None
        if s == 2:
            c0 = vol(H, 4 * _pad16(half), s2d=True)
            plan.add(x, [_conv_taps(blk.c0.conv.weight, 1, False)], half, 2, c0, alpha=a0, beta=b0, act=1)
            c1 = vol(Ho, _pad16(half))
            plan.add(c0, [_conv_taps(blk.c1.conv.weight, 4, True)], half, 0, c1, alpha=a1, beta=b1, act=1)
            sc = vol(Ho, _pad16(cout))
            plan.add(x_sub, [_conv_taps(blk.shortcut.conv.weight, 1, False)], cout, 0, sc)
        else:
            c0 = vol(H, _pad16(half))
            plan.add(x, [_conv_taps(blk.c0.conv.weight, 1, False)], half, 0, c0, alpha=a0, beta=b0, act=1)
            c1 = vol(Ho, _pad16(half))
            plan.add(c0, [_conv_taps(blk.c1.conv.weight, 4, False)], half, 0, c1, alpha=a1, beta=b1, act=1)
            if blk.shortcut is not None:
                sc = vol(Ho, _pad16(cout))
                plan.add(x, [_conv_taps(blk.shortcut.conv.weight, 1, False)], cout, 0, sc)
            else:
                sc = x
        y = vol(Ho, _pad16(cout))
        plan.add(c1, [_conv_taps(blk.c2.conv.weight, 1, False)], cout, 0, y, alpha=a2, beta=b2, gamma=a2, act=1, res=sc,
                 aux=nxt_sub)
        x, x_sub, H = y, nxt_sub, Ho
        plan.marks.append(("enc%d" % i, y))
    for i, d in enumerate(net.dec):
        cout, s = d.conv.out_channels, d.stride
        a, b_ = _bn(d.bn)
        Ho = H * s
        phases = _deconv_phases(d.conv.weight, s)
        if d.final:
            plan.add(x, phases, cout, 4, plan.out, alpha=a, beta=b_, act=2, out_scale=float(max_pos))
        else:
            y = vol(Ho, _pad16(cout))
            plan.add(x, phases, cout, 1 if s == 2 else 0, y, alpha=a, beta=b_, act=1)
            x, H = y, Ho
            plan.marks.append(("dec%d" % i, y))
    return plan
