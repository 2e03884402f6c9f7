"""CPU tests of oracle/vision.py: integer identities the kernels rely on, the closed-form crop
transform vs the Umeyama fit skimage would compute, the reference's own index fixtures, and a
cross-check of the bilinear warp restatement against OpenCV."""
import os

import numpy as np
import pytest

from oracle import vision as V

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_integer_identities_used_by_rect_geometry_kernel():
    for bw in range(0, 2000):
        assert int(0.3 * bw) == (3 * bw) // 10
    for k in range(0, 6000):
        assert int((k / 2) * 1.6) == (4 * k) // 5


def test_padding_and_crop_box_examples():
    assert V.apply_padding((720, 1280, 3), (400, 700, 150, 460), 0.3) == (310, 790, 57, 553)
    assert V.apply_padding((720, 1280, 3), (10, 500, 5, 700), 0.3) == (0, 647, 0, 720)      # clamped
    c, s = V.crop_box((400, 700, 150, 460))
    assert s == 488 and tuple(c) == (550.0, 305.0)


def test_closed_form_similarity_equals_umeyama_fit():
    rng = np.random.default_rng(0)
    for _ in range(200):
        l, t = rng.integers(0, 1000, 2)
        w, h = rng.integers(20, 600, 2)
        c, s = V.crop_box((int(l), int(l + w), int(t), int(t + h)))
        T1, T2 = V.crop_transform(c, s), V.crop_transform_closed_form(c, s)
        assert np.abs(T1 - T2).max() < 1e-9 * max(1.0, np.abs(T2).max())


def test_reference_index_fixtures():
    uv = np.loadtxt(os.path.join(GOLD, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(GOLD, "face_ind.npy"))
    assert uv.shape == (2, 68) and uv.min() >= 15 and uv.max() <= 240
    assert face.shape == (43867,) and face.min() >= 0 and face.max() < 65536 and (np.diff(face) > 0).all()
    flat = V.flat_kpt_index(uv)
    pos = np.arange(256 * 256 * 3, dtype=np.float64).reshape(256, 256, 3)
    assert np.array_equal(V.get_landmarks(pos, uv), pos.reshape(-1, 3)[flat])


def test_warp_restatement_cross_checked_with_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (200, 260, 3), dtype=np.uint8)
    c, s = V.crop_box((60, 200, 40, 170))
    T = V.crop_transform(c, s)
    mine = V.warp_bilinear_constant(img, np.linalg.inv(T))
    ref = cv2.warpAffine(img.astype(np.float64) / 255.0, T[:2], (256, 256), flags=cv2.INTER_LINEAR,
                         borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    # OpenCV quantises the interpolation weights to 1/32 px -> loose bound; interior only
    assert np.abs(mine[4:-4, 4:-4] - ref[4:-4, 4:-4]).mean() < 2e-2


def test_restore_uses_float32_division_for_z():
    rng = np.random.default_rng(2)
    pos = (rng.random((256, 256, 3), dtype=np.float32) * 281.6).astype(np.float32)
    c, s = V.crop_box((400, 700, 150, 460))
    T = V.crop_transform_closed_form(c, s)
    out = V.restore_posmap(pos, T)
    z32 = (pos[..., 2] / np.float32(T[0, 0])).astype(np.float32)
    assert np.array_equal(out[..., 2], z32.astype(np.float64))
    x = pos[..., 0].astype(np.float64) * (s / 255.0) + (c[0] - s / 2)
    assert np.abs(out[..., 0] - x).max() < 1e-9
