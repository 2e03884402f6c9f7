"""-m gpu parity tests of the per-frame vision kernels vs oracle/vision.py (numpy float64).

Bar: integer outputs (padded rect, crop size, gather indices) bit-exact; warp <= 1e-6 abs on [0,1]
pixels; restored landmarks/vertices <= 1e-9 relative (float64 out), z bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import vision as V

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fixtures():
    uv = np.loadtxt(os.path.join(GOLD, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(GOLD, "face_ind.npy"))
    return uv, face


def _rects(rng, n, H, W):
    out = []
    for _ in range(n):
        w = int(rng.integers(40, min(H, W) // 2))
        h = w + int(rng.integers(-10, 10))
        l = int(rng.integers(-20, W - w + 20))          # may poke outside the frame
        t = int(rng.integers(-20, H - h + 20))
        out.append((l, l + w, t, t + h))
    return out


def test_rect_geometry_bit_exact(native_lib, cuda):
    from lipreading_b200 import functional as LF
    rng = np.random.default_rng(123456)
    H, W = 720, 1280
    rects = _rects(rng, 500, H, W) + [(0, 0, 0, 0), (10, 11, 10, 11), (0, W, 0, H), (300, 1999 + 300, 5, 1999 + 5)]
    rp, crop = LF.rect_geometry(torch.tensor(rects, dtype=torch.int32, device=cuda), H, W)
    rp, crop = rp.cpu().numpy(), crop.cpu().numpy()
    for i, r in enumerate(rects):
        assert tuple(rp[i]) == V.apply_padding((H, W, 3), r, 0.3), r
        c, s = V.crop_box(r)
        assert crop[i, 2] == s and crop[i, 0] == int(round(2 * c[0])) and crop[i, 1] == int(round(2 * c[1]))


def test_warp256_matches_skimage_restatement(native_lib, cuda):
    from lipreading_b200 import functional as LF
    rng = np.random.default_rng(7)
    H, W, n = 240, 320, 5
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    rects = _rects(rng, n - 1, H, W) + [(250, 330, 180, 250)]      # last crop hangs off the frame
    r = torch.tensor(rects, dtype=torch.int32, device=cuda)
    _, crop = LF.rect_geometry(r, H, W)
    out = LF.warp256(torch.from_numpy(frames).to(cuda), crop).cpu().numpy()
    for i, rect in enumerate(rects):
        c, s = V.crop_box(rect)
        T = V.crop_transform(c, s)                                   # Umeyama fit, like skimage
        ref = V.warp_bilinear_constant(frames[i], np.linalg.inv(T))
        assert np.abs(out[i] - ref).max() < 1e-6, i
    assert out.min() >= 0.0 and out.max() <= 1.0


def test_posmap_gather_landmarks_and_vertices(native_lib, cuda):
    from lipreading_b200 import functional as LF
    uv, face = _fixtures()
    assert uv.shape == (2, 68) and face.shape == (43867,)
    rng = np.random.default_rng(3)
    H, W, n = 720, 1280, 3
    rects = _rects(rng, n, H, W)
    pos = (rng.random((n, 256, 256, 3), dtype=np.float32) * 281.6).astype(np.float32)
    r = torch.tensor(rects, dtype=torch.int32, device=cuda)
    rp, crop = LF.rect_geometry(r, H, W)
    kidx = torch.from_numpy(V.flat_kpt_index(uv)).to(cuda)
    lmk, vtx = LF.posmap_gather(torch.from_numpy(pos).to(cuda), crop, rp, kidx, torch.from_numpy(face).to(cuda))
    lmk, vtx = lmk.cpu().numpy(), vtx.cpu().numpy()
    for i, rect in enumerate(rects):
        l_ref, v_ref, rp_ref = V.frame_landmarks((H, W, 3), rect, pos[i], uv, face)
        assert tuple(rp.cpu().numpy()[i]) == rp_ref
        assert np.abs(lmk[i][:, :2] - l_ref[:, :2]).max() < 1e-9 * max(H, W)
        assert np.abs(vtx[i][:, :2] - v_ref[:, :2]).max() < 1e-9 * max(H, W)
        assert np.array_equal(lmk[i][:, 2], l_ref[:, 2]) and np.array_equal(vtx[i][:, 2], v_ref[:, 2])   # z bit-exact
    lmk_only = LF.posmap_gather(torch.from_numpy(pos).to(cuda), crop, rp, kidx).cpu().numpy()
    assert np.array_equal(lmk_only, lmk)


def test_collate_pad(native_lib, cuda):
    from lipreading_b200 import functional as LF
    rng = np.random.default_rng(0)
    lens = [3, 7, 7, 12]
    rows = [rng.standard_normal((t, 68, 3)) for t in lens]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda)
    src = torch.from_numpy(np.concatenate(rows, 0).reshape(-1, 204)).to(cuda)
    out = LF.collate_pad(src, offs, len(lens), max(lens), 204).cpu()
    from oracle.sequence import collate
    ref, ref_lens, _, _ = collate([(r, np.array([1, 5, 2])) for r in rows])
    assert torch.equal(out.reshape(ref.shape), ref)


@pytest.mark.parametrize("H,W,out_h,out_w,wide", [(360, 480, 50, 100, False), (181, 243, 100, 50, False),
                                                  (720, 1280, 100, 50, True), (97, 131, 50, 100, False)])
def test_mouth_crop_matches_spec(native_lib, cuda, H, W, out_h, out_w, wide):
    """bit-exact against the numpy spec: staged-row path, odd frame sizes (unaligned rows, partial last chunk of the
    buffer), ROIs hanging off every frame edge, and ROIs wider than the staging rows (direct path)."""
    from lipreading_b200 import functional as LF
    rng = np.random.default_rng(5)
    n = 6
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    lmk = rng.random((n, 68, 3)) * 100.0
    span = 0.55 * W if wide else 0.15 * W
    lmk[:, 48:68, 0] = 0.12 * W + rng.random((n, 20)) * span
    lmk[:, 48:68, 1] = 0.25 * H + rng.random((n, 20)) * 0.08 * H
    rp = np.zeros((n, 4), dtype=np.int32)
    rp[:, 0] = [int(0.2 * W), 0, int(0.62 * W), 5, -int(0.2 * W), int(0.3 * W)]       # left pad: ROI off the right / left edge
    rp[:, 2] = [int(0.1 * H), 0, int(0.55 * H), 5, -int(0.3 * H), int(0.7 * H)]       # top pad: off the bottom / top edge
    rp[:, 1], rp[:, 3] = rp[:, 0] + W // 2, rp[:, 2] + H // 2
    out, roi = LF.mouth_crop(torch.from_numpy(frames).to(cuda), torch.from_numpy(lmk).to(cuda),
                             torch.from_numpy(rp).to(cuda), out_h, out_w)
    out, roi = out.cpu().numpy(), roi.cpu().numpy()
    for i in range(n):
        r_ref = V.mouth_roi(lmk[i], rp[i], out_h, out_w)
        assert tuple(roi[i]) == r_ref
        ref = V.mouth_crop(frames[i], r_ref, out_h, out_w)
        assert np.array_equal(out[i], ref), (i, int(np.abs(out[i].astype(int) - ref.astype(int)).max()))
