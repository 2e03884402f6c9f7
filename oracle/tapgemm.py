"""CPU statement of what one `lr_tapgemm` launch computes (include/lr_b200.h), over the same launch descriptions
`lipreading_b200.prnet_tc5.Plan` hands to the CUDA kernel.  TEST INFRASTRUCTURE: tests/ replay a compiled plan
through this file to hold the PLAN (tap offsets, space-to-depth slots, transposed-conv phases, folded batch-norm,
residual wiring) to `prnet.ResFcn256.forward` — the restatement of the reference's resfcn256
(src/models/face/prnet.py:211-280) — and the GPU tests hold the kernel to this file launch by launch.

Arithmetic: fp32 accumulation of bf16 operands (the kernel's tcgen05 kind::f16 with fp32 accumulators), fp32
epilogue, bf16 rounding where the kernel stores bf16.
"""
import torch


def _windows(vol, Kg, pack=1):
    """A matrix the kernel's TMA boxes see: row q = the Kg values starting at position q's channel 0 (runs over the
    following positions when C < Kg; rows outside the volume read as zero).  pack > 1: row q = positions q*pack ..
    q*pack+pack-1 (aligned, pack*C == Kg)."""
    flat = vol.t.reshape(-1).float()
    C = vol.C
    if pack > 1:
        return flat[:vol.rows * C].reshape(vol.rows // pack, pack * C)
    if C >= Kg:
        return vol.t[:vol.rows, :Kg].float()
    idx = torch.arange(vol.rows)[:, None] * C + torch.arange(Kg)[None, :]
    return flat[idx]


def run_launch(s):
    a = s["a"]
    rows, Kg, cp, G, NP = a.rows, s["Kg"], s["Cout_pad"], s["n_groups"], s["n_phases"]
    pack = s.get("pack", 1)
    mrows = rows // pack                                                 # matrix rows
    A = _windows(a, Kg, pack)                                            # (mrows, Kg)
    W = s["w"].float().reshape(NP, G, pack * cp, Kg)
    qm = torch.arange(mrows)
    q = torch.arange(rows)
    b = q // (a.Hp * a.Wp)
    r = q % (a.Hp * a.Wp)
    y, x = r // a.Wp, r % a.Wp
    vy0, vx0, H, Wd = s["valid"]
    yy, xx = y - vy0, x - vx0
    valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < Wd)
    out = s["out"]
    for ph in range(NP):
        acc = torch.zeros(mrows, pack * cp)
        for g in range(G):
            off = s["tap_off"][ph * G + g]
            src = qm + off
            ok = (src >= 0) & (src < mrows)
            Ag = torch.zeros(mrows, Kg)
            Ag[ok] = A[src[ok]]
            acc += Ag @ W[ph, g].t()
        v = acc.reshape(rows, cp)                                         # columns [j*cp, (j+1)*cp) -> position q*pack + j
        if s["alpha"] is not None:
            v = v * s["alpha"].float().cpu()
        if s["beta"] is not None:
            v = v + s["beta"].float().cpu()
        mode = s["mode"]
        if mode in (0, 1, 2):
            if mode == 0:
                orow = (b * out.Hp + yy + out.pad) * out.Wp + xx + out.pad
                coff = torch.zeros_like(q)
            elif mode == 1:
                orow = (b * out.Hp + 2 * yy + (ph >> 1) + out.pad) * out.Wp + 2 * xx + (ph & 1) + out.pad
                coff = torch.zeros_like(q)
            else:
                orow = (b * out.Hp + ((yy + 1) >> 1)) * out.Wp + ((xx + 1) >> 1)
                coff = ((((yy + 1) & 1) << 1) | ((xx + 1) & 1)) * cp
            if s["res"] is not None:
                res = s["res"].t[orow.clamp(0, s["res"].rows - 1)][:, :cp].float()
                g_ = s["gamma"].float().cpu() if s["gamma"] is not None else 1.0
                v = v + g_ * res
            if s["act"] == 1:
                v = v.clamp_min(0)
            elif s["act"] == 2:
                v = torch.sigmoid(v)
            vb = v.to(torch.bfloat16)
            rows_ok = torch.nonzero(valid)[:, 0]
            cols = coff[rows_ok][:, None] + torch.arange(cp)[None, :]
            out.t[orow[rows_ok][:, None], cols] = vb[rows_ok]
            if s["aux"] is not None:
                ax = s["aux"]
                sel = torch.nonzero(valid & (yy % 2 == 0) & (xx % 2 == 0))[:, 0]
                arow = (b[sel] * ax.Hp + yy[sel] // 2 + ax.pad) * ax.Wp + xx[sel] // 2 + ax.pad
                ax.t[arow, :cp] = vb[sel]
        else:
            if s["act"] == 1:
                v = v.clamp_min(0)
            elif s["act"] == 2:
                v = torch.sigmoid(v)
            if mode == 3:
                out[:rows, :s["Cout"]] = v[:, :s["Cout"]]
            else:
                sel = torch.nonzero(valid)[:, 0]
                out.view(-1, s["Cout"])[(b[sel] * H + yy[sel]) * Wd + xx[sel]] = v[sel, :s["Cout"]] * s["out_scale"]


def run_plan(plan, images):
    """images (B,R,R,3) f32 in [0,1] -> position map (B,R,R,3) f32, every launch replayed on the CPU."""
    from lipreading_b200.prnet_tc5 import P
    B, R = plan.B, plan.R
    v = plan.vin.t[:plan.vin.rows].view(B, R + 2 * P, R + 2 * P, 16)
    v.zero_()
    v[:, P:P + R, P:P + R, :3] = images.to(torch.bfloat16)
    for s in plan.specs:
        run_launch(s)
    return plan.out
