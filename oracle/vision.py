"""CPU restatement (numpy, float64) of the reference's per-frame vision geometry.  TEST INFRA.

Follows SURVEY Appendix B step by step; every function cites the reference lines it restates.
skimage 0.14.1 (`estimate_transform`, `warp`) and dlib are absent from the reference tree and this
image, so `umeyama_similarity` and `warp_bilinear_constant` restate their published algorithms
(Umeyama 1991 as implemented in skimage.transform._geometric._umeyama; skimage `_warp_fast`
bilinear, mode='constant', cval=0).  PARITY UNPINNED beyond this restatement — tolerances in tests.
"""
import math

import numpy as np


# ---- a2: face.py:76-90 ---------------------------------------------------------------------------
def apply_padding(dims, rect, padding=0.3):
    img_h, img_w = dims[0], dims[1]
    left, right, top, bottom = rect
    box_h, box_w = bottom - top, right - left
    left = max(0, left - int(padding * box_w))
    right = min(img_w, right + int(padding * box_w))
    top = max(0, top - int(padding * box_h))
    bottom = min(img_h, bottom + int(padding * box_h))
    return left, right, top, bottom


# ---- a3: prnet.py:112-119 -------------------------------------------------------------------------
def crop_box(rect):
    left, right, top, bottom = rect
    old_size = (right - left + bottom - top) / 2
    center = np.array([right - (right - left) / 2.0, bottom - (bottom - top) / 2.0])
    size = int(old_size * 1.6)
    return center, size


def umeyama_similarity(src, dst):
    """Least-squares similarity (scale, rotation, translation) src->dst, 3x3 homogeneous, float64.
    Restates skimage.transform._geometric._umeyama(estimate_scale=True)."""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    num, dim = src.shape
    src_mean, dst_mean = src.mean(0), dst.mean(0)
    src_d, dst_d = src - src_mean, dst - dst_mean
    A = dst_d.T @ src_d / num
    d = np.ones(dim)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    if rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    scale = 1.0 / src_d.var(axis=0).sum() * (S @ d)
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


def crop_transform(center, size, resolution=256):
    """prnet.py:137-140: the three crop corners -> [[0,0],[0,255],[255,0]] similarity."""
    src = np.array([[center[0] - size / 2, center[1] - size / 2],
                    [center[0] - size / 2, center[1] + size / 2],
                    [center[0] + size / 2, center[1] - size / 2]])
    dst = np.array([[0, 0], [0, resolution - 1], [resolution - 1, 0]])
    return umeyama_similarity(src, dst)


def crop_transform_closed_form(center, size, resolution=256):
    """What the kernels evaluate: the fit above is an exact axis-aligned scale + shift."""
    s = (resolution - 1) / size
    T = np.eye(3)
    T[0, 0] = T[1, 1] = s
    T[0, 2] = -s * (center[0] - size / 2)
    T[1, 2] = -s * (center[1] - size / 2)
    return T


# ---- a4: prnet.py:142-143 -------------------------------------------------------------------------
def warp_bilinear_constant(image_u8, T_inv, out_shape=(256, 256)):
    """`warp(image/255., tform.inverse, output_shape)`: out[v,u] = bilinear(image, T^-1 (u,v)),
    taps outside the image contribute cval=0, floor/ceil taps as in skimage's _warp_fast."""
    img = image_u8.astype(np.float64) / 255.0
    rows, cols = img.shape[:2]
    vv, uu = np.meshgrid(np.arange(out_shape[0], dtype=np.float64), np.arange(out_shape[1], dtype=np.float64), indexing="ij")
    x = T_inv[0, 0] * uu + T_inv[0, 1] * vv + T_inv[0, 2]
    y = T_inv[1, 0] * uu + T_inv[1, 1] * vv + T_inv[1, 2]
    minc, minr = np.floor(x), np.floor(y)
    maxc, maxr = np.ceil(x), np.ceil(y)
    dc, dr = x - minc, y - minr

    def px(r, c):
        ok = (r >= 0) & (r < rows) & (c >= 0) & (c < cols)
        ri, ci = np.clip(r, 0, rows - 1).astype(np.int64), np.clip(c, 0, cols - 1).astype(np.int64)
        return img[ri, ci] * ok[..., None]

    top = (1 - dc)[..., None] * px(minr, minc) + dc[..., None] * px(minr, maxc)
    bot = (1 - dc)[..., None] * px(maxr, minc) + dc[..., None] * px(maxr, maxc)
    return (1 - dr)[..., None] * top + dr[..., None] * bot


# ---- a6: prnet.py:151-156 -------------------------------------------------------------------------
def restore_posmap(cropped_pos_f32, T):
    """cropped_pos (256,256,3) float32 (CNN output * 281.6).  NumPy-1.x value-based casting (the
    reference's era) keeps `float32_array / float64_scalar` in float32; restated explicitly because
    NumPy 2 would promote."""
    cv = np.reshape(cropped_pos_f32.astype(np.float32), [-1, 3]).T.copy()
    z = (cv[2, :] / np.float32(T[0, 0])).astype(np.float32)
    cv[2, :] = 1
    vertices = np.linalg.inv(T) @ cv.astype(np.float64)
    vertices = np.vstack((vertices[:2, :], z.astype(np.float64)))
    return np.reshape(vertices.T, [256, 256, 3])


# ---- a7/a8: prnet.py:162-182 ------------------------------------------------------------------------
def get_landmarks(pos, uv_kpt_ind):
    return pos[uv_kpt_ind[1, :], uv_kpt_ind[0, :], :]


def get_vertices(pos, face_ind):
    return np.reshape(pos, [256 * 256, -1])[face_ind, :]


# ---- a9: face.py:164-175 ----------------------------------------------------------------------------
def get_face(inp, rect):
    left, _, top, _ = rect
    res = inp.copy()
    res[:, 0] -= left
    res[:, 1] -= top
    return res


def frame_landmarks(frame_shape, rect, cropped_pos_f32, uv_kpt_ind, face_ind=None, closed_form=False):
    """generate_dataview._gen_data (:58-76) minus the detector and the CNN: rect + position map ->
    face-relative landmarks (68,3) [and vertices]."""
    rect_pad = apply_padding(frame_shape, rect, 0.3)
    center, size = crop_box(rect)
    T = crop_transform_closed_form(center, size) if closed_form else crop_transform(center, size)
    pos = restore_posmap(cropped_pos_f32, T)
    lmk = get_face(get_landmarks(pos, uv_kpt_ind), rect_pad)
    if face_ind is None:
        return lmk, rect_pad
    return lmk, get_face(get_vertices(pos, face_ind), rect_pad), rect_pad


def flat_kpt_index(uv_kpt_ind):
    """(2,68) file layout -> flat row*256+col (row = second file row, col = first; prnet.py:169)."""
    return (uv_kpt_ind[1, :].astype(np.int64) * 256 + uv_kpt_ind[0, :].astype(np.int64)).astype(np.int32)


# ---- N2 (extension, spec lives here) -----------------------------------------------------------------
def mouth_roi(lmk_face_rel, rect_pad, out_h=50, out_w=100):
    pts = np.asarray(lmk_face_rel, np.float64)[48:68]
    x = pts[:, 0] + float(rect_pad[0])
    y = pts[:, 1] + float(rect_pad[2])
    aspect = out_w / out_h
    rw = max(1.2 * max(x.max() - x.min(), (y.max() - y.min()) * aspect), 2.0)
    rh = max(rw / aspect, 2.0)
    cx, cy = 0.5 * (x.min() + x.max()), 0.5 * (y.min() + y.max())
    return (int(math.floor(cx - 0.5 * rw)), int(math.floor(cy - 0.5 * rh)), int(math.ceil(rw)), int(math.ceil(rh)))


def mouth_crop(frame_u8, roi, out_h=50, out_w=100):
    H, W = frame_u8.shape[:2]
    x_lo, y_lo, rw, rh = roi
    f32 = np.float32
    sxs, sys_ = f32(rw) / f32(out_w), f32(rh) / f32(out_h)
    ox = np.arange(out_w, dtype=np.float32)
    oy = np.arange(out_h, dtype=np.float32)
    sx = ((ox + f32(0.5)) * sxs + f32(-0.5)) + f32(x_lo)
    sy = ((oy + f32(0.5)) * sys_ + f32(-0.5)) + f32(y_lo)
    fx, fy = np.floor(sx), np.floor(sy)
    ax, ay = (sx - fx).astype(np.float32), (sy - fy).astype(np.float32)
    x0 = np.clip(fx.astype(np.int64), 0, W - 1)
    x1 = np.clip(fx.astype(np.int64) + 1, 0, W - 1)
    y0 = np.clip(fy.astype(np.int64), 0, H - 1)
    y1 = np.clip(fy.astype(np.int64) + 1, 0, H - 1)
    img = frame_u8.astype(np.float32)
    axb, ayb = ax[None, :, None], ay[:, None, None]
    top = (f32(1) - axb) * img[y0][:, x0] + axb * img[y0][:, x1]
    bot = (f32(1) - axb) * img[y1][:, x0] + axb * img[y1][:, x1]
    val = (f32(1) - ayb) * top + ayb * bot
    return np.clip(np.rint(val), 0, 255).astype(np.uint8)
