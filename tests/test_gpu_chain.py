"""-m gpu: the chain the north star names, end to end through the reference's own entry points
(BASELINE config "nano"): `generate_dataview(**flags)` (frames -> boxes -> landmarks -> mouth clips, written as
dataview columns) -> `python -m src.scripts.train $(cat config/train/<cfg>.txt)` for one epoch, on a synthetic
workspace.  The video / caption decoders and the face detector + position-map CNN are plugs (out of scope or
un-vendored): deterministic stand-ins are installed here.
Reference call sites: src/scripts/generate_dataview.py:151-239, src/scripts/train.py:134-354,
src/data/data_loader.py:154-257."""
import collections
import os
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

H, W = 180, 240
N_VIDEOS = 6


class _FakeVideo:
    """genFrames(lo, hi) of a synthetic talking-head video: smooth random texture, deterministic per path."""

    def __init__(self, path):
        self.rng = np.random.default_rng(zlib.crc32(os.path.basename(path).encode()))
        self.base = self.rng.integers(0, 256, (H, W, 3), dtype=np.uint8)

    def genFrames(self, lo, hi):
        out = []
        for i in range(lo, hi):
            out.append(np.roll(self.base, i % 7, axis=1))
        return out


def _captions(path):
    rng = np.random.default_rng(zlib.crc32(os.path.basename(path).encode()) + 1)
    caps = collections.OrderedDict()
    t = 0.0
    for i in range(4):
        dur = float(rng.uniform(1.0, 1.6))                      # 30-48 frames per caption window
        caps[(t, t + dur)] = "".join(rng.choice(list("abcdefgh "), size=int(rng.integers(6, 12)))).strip() or "abc"
        t += dur
    return caps


def _detector(frame):
    return (70, 170, 40, 140)                                     # (left, right, top, bottom)


def _fake_cnn(cropped):
    """stand-in position map: a smooth function of the crop, scaled like PRNet's output (prnet.py:292)"""
    n = cropped.shape[0]
    v, u = torch.meshgrid(torch.arange(256.0, device=cropped.device), torch.arange(256.0, device=cropped.device),
                          indexing="ij")
    pm = torch.stack([u, v, 30 + 0 * u], -1)[None].expand(n, -1, -1, -1) + cropped * 2.0
    return pm.contiguous()


@pytest.fixture()
def workspace(tmp_path, monkeypatch):
    monkeypatch.setenv("LIP_READING_WS_PATH", str(tmp_path))
    raw = tmp_path / "data" / "raw" / "Fake" / "nano"
    raw.mkdir(parents=True)
    for v in range(N_VIDEOS):
        (raw / ("vid%02d.mp4" % v)).write_bytes(b"")
        (raw / ("vid%02d.vtt" % v)).write_text("")
    return tmp_path


def _generate(cuda):
    from lipreading_b200 import dataview
    from lipreading_b200.face import PRN
    uv = np.loadtxt(os.path.join(GOLD, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(GOLD, "face_ind.npy"))
    prn = PRN(predict_batch=_fake_cnn, uv_kpt_ind=uv, face_ind=face, device=cuda)
    dataview.generate_dataview(inp="Fake/nano", gen_mouth=True, video_reader_cls=_FakeVideo, caption_reader=_captions,
                               detector=_detector, prn=prn)


def test_generate_dataview_writes_landmark_and_mouth_columns(native_lib, cuda, workspace):
    from oracle import vision as V
    _generate(cuda)
    d = os.path.join(str(workspace), "data", "datasets", "Fake", "nano", "vid00")
    assert sorted(os.listdir(d)) == ["cap.npy", "face_lmk_seq.npy", "mouth_clip_seq.npy", "s_e.npy"]
    lm = np.load(os.path.join(d, "face_lmk_seq.npy"), allow_pickle=True)
    mc = np.load(os.path.join(d, "mouth_clip_seq.npy"), allow_pickle=True)
    se = np.load(os.path.join(d, "s_e.npy"))
    assert len(lm) == len(mc) == len(se) == 4
    for (s, e), l, m in zip(se, lm, mc):
        n = int(e * 29.97) - int(s * 29.97)
        assert l.shape == (n, 68, 3) and l.dtype == np.float64
        assert m.shape == (n, 100, 50, 3) and m.dtype == np.uint8
    # the mouth column equals the numpy spec applied to the landmark column (same frames, same boxes)
    vid = _FakeVideo(os.path.join(str(workspace), "data", "raw", "Fake", "nano", "vid00.mp4"))
    s, e = se[1]
    frames = vid.genFrames(int(s * 29.97), int(e * 29.97))
    rp = np.array(V.apply_padding((H, W, 3), _detector(None), 0.3), dtype=np.int32)
    for t in (0, len(frames) - 1):
        roi = V.mouth_roi(lm[1][t], rp, 100, 50)
        assert np.array_equal(mc[1][t], V.mouth_crop(frames[t], roi, 100, 50))


@pytest.mark.parametrize("cfg,extra", [
    ("stcnn_bigru256_ctc.txt", ["--batch_size=4"]),
    ("bigru256_ctc.txt", ["--batch_size=4"]),
])
def test_cli_train_one_epoch_from_config_file(native_lib, cuda, workspace, cfg, extra, capsys):
    """`python -m src.scripts.train $(cat config/train/<cfg>)` with the dataset flag pointed at the synthetic
    workspace: dataset build (filter, sort, pickle cache), device collate / prefetch, initial eval, one epoch of
    train (decoder + CTC), eval, checkpoints."""
    from lipreading_b200.cli import read_config
    from lipreading_b200 import train_script
    _generate(cuda)
    argv = read_config(os.path.join(ROOT, "config", "train", cfg)) + ["--data=Fake/nano", "--max_epochs=1",
                                                                     "--refresh", "-v", "0"] + extra
    from lipreading_b200.cli import parseArgsForClassOrScript
    args = vars(parseArgsForClassOrScript(train_script.train, argv))
    args.pop("verbosity", None)
    out = train_script.train(**args)
    assert len(out["val_cers"]) == 1 and 0.0 <= out["val_cers"][0] <= 1.0
    assert np.isfinite(out["dec_losses"][0]) and np.isfinite(out["ctc_losses"][0]) and out["ctc_losses"][0] > 0
    assert os.path.isfile(os.path.join(out["weights_dir"], "best_encoder.pth")) or out["val_cers"][0] >= 1.0
    frame_type = "mouth_clip_seq" if "stcnn" in cfg else None
    pk = os.path.join(str(workspace), "data", "pickles", "Fake", "nano", "non-sentence", "train")
    assert os.path.isfile(os.path.join(pk, frame_type or "", "frames.pkl"))


def test_gpu_batch_loader_matches_host_collate(native_lib, cuda, workspace):
    from lipreading_b200 import data
    _generate(cuda)
    rand = np.random.RandomState(seed=123456)
    tr, va, te = data.split_dataset("Fake/nano", train_split=0.8, rand=rand)
    for frame_type in ("face_lmk_seq", "mouth_clip_seq"):
        ds = data.FrameCaptionDataset("Fake/nano", "train", tr, refresh=True, frame_type=frame_type)
        ld = data.GpuBatchLoader(ds, 5, cuda)
        assert len(ld) == -(-len(ds) // 5)
        full, k = [], 0
        for frames, lens, chars, char_lens in ld:           # compare while iterating: the prefetcher recycles its slots
            rows = [ds[i] for i in range(5 * k, min(len(ds), 5 * k + 5))]
            ref = data._collate_fn(rows)
            assert frames.is_cuda and torch.equal(lens, ref[1]) and torch.equal(chars, ref[2])
            want = ref[0] if frame_type == "face_lmk_seq" else ref[0].to(torch.uint8)
            assert frames.dtype == want.dtype and torch.equal(frames.cpu(), want)
            full.append((frames.shape[0], lens.clone()))
            k += 1
        assert k == len(ld)
        # two ranks: contiguous slices of the same global batches, the same number of batches on both
        halves = [list(data.GpuBatchLoader(ds, 5, cuda, rank=r, world=2, prefetch=False)) for r in (0, 1)]
        assert len(halves[0]) == len(halves[1])
        for a, b, (n, lens) in zip(halves[0], halves[1], full):
            assert a[0].shape[0] + b[0].shape[0] == n
            assert torch.equal(torch.cat([a[1], b[1]]), lens)
