"""generate_dataview with the reference's signature and on-disk format
(src/scripts/generate_dataview.py:51-239), frames processed in GPU batches.

Per caption window the reference loops frame by frame (HOG box -> pad rect -> PRNet -> 68 landmarks ->
translate).  Here a whole window of frames goes through the batched kernels at once; the
failure semantics are kept: the first frame that fails truncates the caption's sequence
(`break`, :127-131) and a caption is kept iff at least one frame succeeded (:134-143).

Columns written per video (np.save, :229-233): s_e.npy (N,2) f64, face_lmk_seq.npy object array of
(T_i,68,3) f64 in padded-face coordinates, cap.npy strings [, face_vtx_seq.npy].

Extension (rows N2 -> N1 of SURVEY §8, the chain the north star names): with `gen_mouth=True` a fifth column
`mouth_clip_seq.npy` — object array of (T_i,H,W,3) u8 mouth crops cut by `lr_mouth_crop` from the same frames
around landmarks 48:68 (`face.py:21`) — is written next to the others; it is what
`FrameCaptionDataset(frame_type='mouth_clip_seq')` and the conv front-end (`--frame_processing=conv3d`) read."""
import collections
import glob
import os
import time

import numpy as np
import torch

from . import face as _face
from . import workspace as _ws

_log = _ws.getLogger("generate_dataview")
FPS = 29.97        # video.py:55-56: get_frame_idx = int(seconds * 29.97)


MOUTH_HW = (100, 50)     # (y, x) of a mouth-clip frame: the north star's dataview shape (n, y, x, c) = (75,100,50,3)


def landmarks_for_frames(frames_u8, rects, prn, gen_vtx=False, batch=64, gen_mouth=False, mouth_hw=MOUTH_HW):
    """frames (N,H,W,3) u8 (numpy or tensor), rects list of (l,r,t,b) or None per frame ->
    list of (68,3) f64 arrays up to the first failure [, list of (43867,3)] [, list of (h,w,3) u8 mouth crops]."""
    from . import functional as LF
    n_ok = 0
    for r in rects:
        if r is None:
            break
        n_ok += 1
    lm_out, vt_out, mc_out = [], [], []
    for i in range(0, n_ok, batch):
        fr = torch.as_tensor(np.ascontiguousarray(frames_u8[i:i + batch])).to(prn.device, non_blocking=True)
        rc = torch.tensor([list(r) for r in rects[i:min(i + batch, n_ok)]], dtype=torch.int32)
        res, geom = prn.process_batch(fr[: rc.shape[0]], rc, with_vertices=gen_vtx)
        lmk = res[0] if gen_vtx else res
        if gen_mouth:
            # the frames are still on the device: cut the mouth window now, with the landmarks just computed
            clips, _ = LF.mouth_crop(fr[: rc.shape[0]], lmk, geom[0], mouth_hw[0], mouth_hw[1])
            mc_out += list(clips.cpu().numpy())
        lm_out += list(lmk.cpu().numpy())
        if gen_vtx:
            vt_out += list(res[1].cpu().numpy())
    out = (lm_out,) + ((vt_out,) if gen_vtx else ()) + ((mc_out,) if gen_mouth else ())
    return out if len(out) > 1 else lm_out


def _generate_dataview(video_reader, captions, prn, detector, gen_vtx=False, timedelay=0, gen_mouth=False,
                       mouth_hw=MOUTH_HW):
    """video_reader: object with genFrames(lo, hi) -> list/array of (H,W,3) u8 frames
    (src/utils/data/video.py:61-70); captions: OrderedDict {(start_s, end_s): text}."""
    assert isinstance(captions, collections.OrderedDict) and len(captions) > 0
    cols = ("s_e", "face_lmk_seq") + (("face_vtx_seq",) if gen_vtx else ()) + \
        (("mouth_clip_seq",) if gen_mouth else ()) + ("cap",)
    dataview = collections.OrderedDict((c, []) for c in cols)
    for (start, end), cap in captions.items():
        frames = video_reader.genFrames(int(start * FPS), int(end * FPS))
        if len(frames) == 0:
            continue
        rects = []
        for f in frames:
            try:
                rects.append(tuple(detector(f)))
            except Exception as e:                       # no face -> this and all later frames dropped
                _log.error("\tUnexpected exception '%s', skipping rest of caption...", e)
                rects.append(None)
                break
        rects += [None] * (len(frames) - len(rects))
        res = landmarks_for_frames(np.stack(frames), rects, prn, gen_vtx=gen_vtx, gen_mouth=gen_mouth,
                                   mouth_hw=mouth_hw)
        res = res if isinstance(res, tuple) else (res,)
        lmks = np.array(res[0])
        if lmks.ndim == 3:
            dataview["s_e"].append((start, end))
            dataview["cap"].append(cap)
            dataview["face_lmk_seq"].append(lmks)
            if gen_vtx:
                dataview["face_vtx_seq"].append(np.array(res[1]))
            if gen_mouth:
                dataview["mouth_clip_seq"].append(np.array(res[-1], dtype=np.uint8))
    return dataview


def _object_column(rows):
    col = np.empty(len(rows), dtype=object)
    for i, r in enumerate(rows):
        col[i] = r
    return col


def save_dataview(dst_dir, dataview, out_ext=".npy", force=False):
    """One .npy per column (generate_dataview.py:51-56,229-233).  Ragged landmark rows are stored as
    an object array, which is what np.array(list_of_ragged) produced under the reference's numpy."""
    _ws.mkdirP(dst_dir)
    for col, rows in dataview.items():
        path = os.path.join(dst_dir, col + out_ext)
        if not force and os.path.isfile(path):
            continue
        assert isinstance(rows, list) and len(rows) > 0
        if col in ("face_lmk_seq", "face_vtx_seq", "mouth_clip_seq"):
            arr = _object_column(rows)
        elif col == "s_e":
            arr = np.array(rows, dtype=np.float64)
        else:
            arr = np.array(rows)
        np.save(path, arr, allow_pickle=True)


def generate_dataview(inp="StephenColbert/nano2", vid_ext=".mp4", cap_ext=".vtt", out_ext=".npy", timedelay=0,
                      gen_vtx=False, force=False, seed=123456, gen_mouth=False, video_reader_cls=None,
                      caption_reader=None, detector=None, prn=None):
    """Generates dataviews for the given input directory of video/caption pairs (same flags as the
    reference; the trailing keyword arguments are the plugs for the out-of-scope decoders/detector)."""
    from .media import VideoReader, extract_captions, prune_and_filter_captions
    rand = np.random.RandomState(seed=seed)
    inp_dir, outp_dir = _ws.getRelRawPath(inp), _ws.getRelDatasetsPath(inp)
    vid_paths = sorted(glob.glob(os.path.join(inp_dir, "*" + vid_ext)))
    cap_paths = sorted(glob.glob(os.path.join(inp_dir, "*" + cap_ext)))
    assert len(vid_paths) == len(cap_paths) > 0
    order = np.arange(len(vid_paths), dtype=np.int64)
    rand.shuffle(order)
    prn = prn if prn is not None else _face._getSharedPrn()
    detector = detector if detector is not None else (lambda f: _face.detectMaxFaceRect(f, times_to_upsample=1))
    ts = time.time()
    for i in order:
        vid_path, cap_path = vid_paths[i], cap_paths[i]
        base = os.path.basename(vid_path).split(".")[0]
        assert base == os.path.basename(cap_path).split(".")[0]
        dst = os.path.join(outp_dir, base)
        if not force and os.path.isdir(dst):
            _log.warning("\tSkipping existing file: '%s'...", dst)
            continue
        caps = (caption_reader or (lambda p: prune_and_filter_captions(extract_captions(p))))(cap_path)
        view = _generate_dataview((video_reader_cls or VideoReader)(vid_path), caps, prn, detector,
                                  gen_vtx=gen_vtx, timedelay=timedelay, gen_mouth=gen_mouth)
        save_dataview(dst, view, out_ext=out_ext, force=force)
    _log.info("Done writing dataviews! Took %0.3f seconds", time.time() - ts)
