"""Per-frame vision boundary of the reference (src/utils/data/face.py:65-175 and
src/models/face/prnet.py:80-182), batched on the GPU.

What runs where
  * face box: dlib's HOG detector is an un-vendored third-party model (not in the reference tree,
    not in this image) -> the box is an INPUT here: pass `rects`, or plug any detector through
    `set_detector(callable(frames)->(N,4) left,right,top,bottom)`;
  * position-map CNN (PRNet resfcn256): `lipreading_b200.prnet.PosPrediction` (architecture pinned by the
    reference's checkpoint index; on CUDA the body runs as 53 launches of the tcgen05 tap-GEMM kernel, prnet_tc5.py).
    The reference ships no weights (`.gitignore:5`):
    the default predictor restores `data/weights/prnet/net/256_256_resfcn256_weight` when the data shard is
    there (prnet.py:42-45) and asserts like the reference when it is not; any other predictor plugs in through
    `set_posmap_predictor`.  SURVEY §8 row a5 / f4;
  * everything between and after (pad rect, crop box, /255 + similarity warp to 256x256, restore,
    68-landmark / 43 867-vertex gather, translate to the padded face frame): sm_100a kernels.
"""
import os

import numpy as np
import torch

from . import functional as LF

_HERE = os.path.dirname(os.path.abspath(__file__))
_mouth = slice(48, 68)
_face_pts = 68

_detector = None
_predictor = None
_prn = None


def set_detector(fn):
    global _detector
    _detector = fn


def set_posmap_predictor(fn):
    global _predictor, _prn
    _predictor = fn
    _prn = None


class PRN:
    """Geometry half of the reference's PRN (prnet.py:18-182): crop, warp, restore, gathers.
    `predict_batch(images (N,256,256,3) f32 in [0,1]) -> (N,256,256,3) f32 * 281.6` is the plug for the
    CNN (PosPrediction.predict_batch, prnet.py:311-314)."""

    resolution_inp = 256
    resolution_op = 256
    MaxPos = 256 * 1.1

    def __init__(self, predict_batch=None, uv_kpt_ind=None, face_ind=None, device="cuda"):
        self.device = torch.device(device)
        self.predict_batch = predict_batch
        uv_dir = None
        if uv_kpt_ind is None or face_ind is None:
            from . import workspace as _ws
            uv_dir = os.path.dirname(_ws.getRelWeightsPath("prnet", "uv", "x"))   # data/weights/prnet/uv
        if uv_kpt_ind is None:
            uv_kpt_ind = np.loadtxt(os.path.join(uv_dir, "uv_kpt_ind.txt")).astype(np.int32)   # prnet.py:48
        if face_ind is None:
            face_ind = np.loadtxt(os.path.join(uv_dir, "face_ind.txt")).astype(np.int32)       # prnet.py:50
        self.uv_kpt_ind = np.asarray(uv_kpt_ind, dtype=np.int32)
        self.face_ind = np.asarray(face_ind, dtype=np.int32)
        flat = self.uv_kpt_ind[1, :].astype(np.int64) * 256 + self.uv_kpt_ind[0, :].astype(np.int64)
        self._kpt_flat = torch.from_numpy(flat.astype(np.int32)).to(self.device)
        self._face_flat = torch.from_numpy(self.face_ind).to(self.device)
        self.pos = None                      # cached like the reference (prnet.py:159)
        self._geom = None

    # -- batched device path ---------------------------------------------------------------------
    def crop_batch(self, frames, rects):
        """frames (N,H,W,3) u8 cuda, rects (N,4) int -> cropped (N,256,256,3) f32, (rect_pad, crop)."""
        n, H, W, _ = frames.shape
        rect_pad, crop = LF.rect_geometry(rects.to(self.device, torch.int32), H, W)
        return LF.warp256(frames, crop), (rect_pad, crop)

    def landmarks_batch(self, posmap, geom, with_vertices=False):
        """posmap (N,256,256,3) f32 (CNN output * MaxPos) -> face-relative landmarks (N,68,3) f64
        [, vertices (N,43867,3) f64]."""
        rect_pad, crop = geom
        if with_vertices:
            return LF.posmap_gather(posmap, crop, rect_pad, self._kpt_flat, self._face_flat)
        return LF.posmap_gather(posmap, crop, rect_pad, self._kpt_flat)

    def process_batch(self, frames, rects, with_vertices=False):
        cropped, geom = self.crop_batch(frames, rects)
        if self.predict_batch is None:
            raise RuntimeError("no position-map predictor installed: the reference ships no PRNet weights "
                               "(set_posmap_predictor / PRN(predict_batch=...))")
        posmap = self.predict_batch(cropped)
        return self.landmarks_batch(posmap, geom, with_vertices), geom

    # -- single-frame API, numpy in / numpy out (prnet.py:80-182) ---------------------------------
    def process(self, input, image_info=None):
        assert image_info is not None, "bounding box required (dlib CNN detector path is out of scope)"
        frame = torch.from_numpy(np.ascontiguousarray(input)).to(self.device)[None]
        rects = torch.tensor([list(image_info)], dtype=torch.int32)
        cropped, geom = self.crop_batch(frame, rects)
        posmap = self.predict_batch(cropped)
        # the reference caches the RESTORED map; we cache what the gathers need and restore lazily
        self.pos = posmap
        self._geom = geom
        return posmap, input / 255.0

    def get_landmarks(self, pos):
        rect_pad = torch.zeros_like(self._geom[0])          # raw (un-translated) coordinates
        return LF.posmap_gather(pos, self._geom[1], rect_pad, self._kpt_flat)[0].cpu().numpy()

    def get_vertices(self, pos):
        rect_pad = torch.zeros_like(self._geom[0])
        return LF.posmap_gather(pos, self._geom[1], rect_pad, self._kpt_flat, self._face_flat)[1][0].cpu().numpy()


def _default_predictor():
    """PosPrediction restored from the workspace weights, exactly where the reference looks (prnet.py:42-45)."""
    from . import workspace as _ws
    from .prnet import PosPrediction
    prn_path = _ws.getRelWeightsPath("prnet", "net/256_256_resfcn256_weight")
    assert os.path.isfile(prn_path + ".data-00000-of-00001"), "please download PRN trained model first."
    pred = PosPrediction()
    pred.restore(prn_path)
    return pred.predict_batch


def _getSharedPrn():
    global _prn
    if _prn is None:
        _prn = PRN(predict_batch=_predictor if _predictor is not None else _default_predictor())
    return _prn


def detectMaxFaceRect(img, times_to_upsample=2):
    """(left, right, top, bottom) of the first detected face (face.py:65-74)."""
    assert _detector is not None, "no face detector installed (dlib is not available): set_detector(...)"
    rects = _detector(img[None] if img.ndim == 3 else img)
    assert len(rects) > 0
    r = rects[0]
    return int(r[0]), int(r[1]), int(r[2]), int(r[3])


def _applyPadding(dims, rect, padding):
    """face.py:76-90 on the host (the batched path uses lr_rect_geometry)."""
    img_h, img_w = dims[0], dims[1]
    left, right, top, bottom = rect
    bw, bh = right - left, bottom - top
    return (max(0, left - int(padding * bw)), min(img_w, right + int(padding * bw)),
            max(0, top - int(padding * bh)), min(img_h, bottom + int(padding * bh)))


def extractFace(img, rect, padding=None):
    assert len(img.shape) == 3 and isinstance(rect, tuple) and len(rect) == 4
    if padding is not None:
        assert 0 < padding <= 0.5
        rect = _applyPadding(img.shape, rect, padding)
    left, right, top, bottom = rect
    res = img[top:bottom, left:right, :]
    assert all(x > 0 for x in res.shape)
    return res, rect


def detect3dLandmarks(img, rect=None):
    prn = _getSharedPrn()
    pos, inp = prn.process(img, image_info=rect)
    return prn.get_landmarks(pos), inp


def get3dVertices():
    prn = _getSharedPrn()
    assert prn.pos is not None
    return prn.get_vertices(prn.pos)


def getFace(inp, rect):
    assert len(inp.shape) == 2 and inp.shape[1] == 3 and isinstance(rect, tuple) and len(rect) == 4
    res = inp.copy()
    res[:, 0] -= rect[0]
    res[:, 1] -= rect[2]
    return res
