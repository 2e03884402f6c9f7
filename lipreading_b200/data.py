"""Dataview -> dataset -> padded batches, with the reference's on-disk formats and semantics
(src/data/data_loader.py:29-290).

Formats kept: `<datasets>/<name>/<video_id>/{s_e,face_lmk_seq,cap}.npy` in, pickle cache
`<pickles>/<name>/{sentence,non-sentence}/<split>/{char2idx,frames,captions}.pkl` out.
`collate_gpu` is the device version of `_collate_fn` (one H2D of the ragged rows + the pad kernel)."""
import glob
import json
import os
import pickle

import numpy as np
import torch
import torch.utils.data as _data

from . import workspace as _ws
from .vocab import BOS, EOS, MARKERS2ID, UNK, build_char2idx

_log = _ws.getLogger("data_loader")


def gen_vid_ids(dataset_name, rand=None):
    dataset_dir = _ws.getRelDatasetsPath(dataset_name)
    vid_ids = sorted(glob.glob(os.path.join(dataset_dir, "*/")))
    assert len(vid_ids) > 0, f"No video ids found: '{dataset_dir}'"
    (rand if rand is not None else np.random).shuffle(vid_ids)
    return vid_ids


def split_dataset(dataset_name, train_split=0.8, rand=None):
    """Split by video (data_loader.py:49-62): train | val | test with val = half of the remainder."""
    vid_ids = gen_vid_ids(dataset_name, rand=rand)
    n_train = int(train_split * len(vid_ids))
    n_val = n_train + (len(vid_ids) - n_train) // 2
    return vid_ids[:n_train], vid_ids[n_train:n_val], vid_ids[n_val:]


def filter_occlusions(frames, captions, start_ends, fps=29.97, threshold=0.8):
    """Keep a caption iff >= threshold of its time window produced landmarks and the frame count
    exceeds len(caption)+2 (room for BOS/EOS) (data_loader.py:80-91)."""
    keep_f, keep_c = [], []
    for f, c, (start, end) in zip(frames, captions, start_ends):
        if (end - start) * fps * threshold <= len(f) and len(c) + 2 < len(f):
            keep_f.append(f)
            keep_c.append(c)
    return keep_f, keep_c


def sort_by_seqlen(frames, captions):
    """Ascending frame count (np.argsort default kind, like the reference :93-98)."""
    order = np.argsort([x.shape[0] for x in frames])
    f = np.empty(len(frames), dtype=object)
    for i, x in enumerate(frames):
        f[i] = x
    return f[order], np.array(captions)[order]


def build_vocab(dataset_name, labels):
    path = os.path.join(_ws.getRelRawPath(dataset_name), labels)
    try:
        with open(path) as fh:
            chars = str("".join(json.load(fh)))
    except Exception:
        chars = None
        _log.warning("Could not open '%s'...\n\tUsing hardcoded labels", path)
    return build_char2idx(chars)


def parse_caption(cap, char2idx):
    ids = [MARKERS2ID[BOS]] + [char2idx.get(ch, MARKERS2ID[UNK]) for ch in cap] + [MARKERS2ID[EOS]]
    assert len(ids) > 2
    return np.array(ids)


def _pad(seqs, dtype):
    lens = torch.LongTensor([len(x) for x in seqs])
    out = torch.zeros((len(seqs), int(lens.max())) + tuple(np.shape(seqs[0])[1:]), dtype=dtype)
    for i, s in enumerate(seqs):
        out[i, : lens[i]] = torch.as_tensor(np.asarray(s)).to(dtype)
    return out, lens


def _collate_fn(batch):
    """Host collate, same output contract as the reference (:117-152)."""
    assert all(len(x) == 2 for x in batch)
    frames, captions = zip(*batch)
    src, src_lens = _pad(frames, torch.float32)
    tgt, tgt_lens = _pad(captions, torch.long)
    return src, src_lens, tgt, tgt_lens


def collate_gpu(batch, device):
    """Same contract, frames padded on the device by lr_collate_pad_f64: the ragged float64 rows go up
    in one pinned copy and are cast + zero-padded by the kernel."""
    from . import functional as LF
    frames, captions = zip(*batch)
    lens = torch.LongTensor([len(x) for x in frames])
    feat = int(np.prod(np.shape(frames[0])[1:]))
    rows = np.concatenate([np.asarray(f, dtype=np.float64).reshape(len(f), feat) for f in frames], 0)
    src = torch.from_numpy(rows).pin_memory().to(device, non_blocking=True)
    offs = torch.zeros(len(frames) + 1, dtype=torch.int64)
    offs[1:] = lens.cumsum(0)
    out = LF.collate_pad(src, offs.to(device), len(frames), int(lens.max()), feat)
    tgt, tgt_lens = _pad(captions, torch.long)
    return out.reshape((len(frames), int(lens.max())) + tuple(np.shape(frames[0])[1:])), lens, tgt, tgt_lens


def collate_clips_pinned(batch):
    """Mouth-clip rows (T_i,H,W,3) u8 -> zero-padded (B,Tmax,H,W,3) u8 in PINNED host memory (the DevicePrefetcher
    copies it to the GPU one step ahead), captions like `_collate_fn`."""
    frames, captions = zip(*batch)
    lens = torch.LongTensor([len(x) for x in frames])
    shape = (len(frames), int(lens.max())) + tuple(np.shape(frames[0])[1:])
    out = torch.zeros(shape, dtype=torch.uint8)
    if torch.cuda.is_available():
        out = out.pin_memory()
    for i, f in enumerate(frames):
        out[i, : len(f)] = torch.from_numpy(np.ascontiguousarray(f, dtype=np.uint8))
    tgt, tgt_lens = _pad(captions, torch.long)
    return out, lens, tgt, tgt_lens


class GpuBatchLoader:
    """The loader `train(**flags)` iterates: consecutive `batch_size` rows of the (length-sorted) dataset, no
    shuffling (train.py:209-211), each batch collated FOR the device —
      * landmark rows (T,68,3) f64: one pinned copy of the ragged rows + `lr_collate_pad_f64` (cast + zero pad);
      * mouth-clip rows (T,H,W,3) u8: zero-padded in pinned memory, copied one batch ahead on a side stream;
    lengths and captions stay on the host (the trainer wants them there).  Under data parallelism `batch_size` is
    the GLOBAL batch and every rank collates only its contiguous slice (`dist.shard_slice`); a tail batch with
    fewer rows than ranks is dropped on every rank."""

    def __init__(self, dataset, batch_size, device, rank=0, world=1, prefetch=True):
        self.dataset, self.batch_size, self.device = dataset, int(batch_size), torch.device(device)
        self.rank, self.world, self.prefetch = rank, world, prefetch

    def _batches(self):
        n = len(self.dataset)
        out = []
        for lo in range(0, n, self.batch_size):
            hi = min(n, lo + self.batch_size)
            if hi - lo < self.world:
                continue
            out.append((lo, hi))
        return out

    def __len__(self):
        return len(self._batches())

    def _host_iter(self, clips):
        from .dist import shard_slice
        for lo, hi in self._batches():
            a, b = shard_slice(hi - lo, self.rank, self.world)
            rows = [self.dataset[i] for i in range(lo + a, lo + b)]
            yield collate_clips_pinned(rows) if clips else collate_gpu(rows, self.device)

    def __iter__(self):
        clips = len(self.dataset) > 0 and np.ndim(self.dataset[0][0]) == 4
        it = self._host_iter(clips)
        if clips and self.prefetch and self.device.type == "cuda":
            return iter(DevicePrefetcher(_Sized(it, len(self)), self.device))
        return it


class _Sized:
    def __init__(self, it, n):
        self.it, self.n = it, n

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter(self.it)


class FrameCaptionDataset(_data.Dataset):
    """Rows = (landmark sequence (T,68,3), parsed caption ids) (data_loader.py:154-257)."""

    def __init__(self, dataset_name, split_name, vid_ids, labels="labels.json", start_end="s_e",
                 threshold=0.8, fps=29.97, cap="cap", frame_type="face_lmk_seq", sentence_dataset=False,
                 in_ext=".npy", out_ext=".pkl", refresh=False):
        super().__init__()
        assert all(os.path.isdir(x) for x in vid_ids)
        assert frame_type in ("face_lmk_seq", "face_vtx_seq", "mouth_clip_seq")     # mouth clips: row N2 -> N1
        assert not sentence_dataset, "--sentence_dataset needs spaCy (out of scope, SURVEY §2 row 6)"
        self.frame_type = frame_type
        pickle_dir = _ws.getRelPicklesPath(dataset_name, "non-sentence", split_name)
        if frame_type != "face_lmk_seq":
            pickle_dir = os.path.join(pickle_dir, frame_type)     # one cache per column (the reference has one column)
        if refresh or not os.path.isdir(pickle_dir):
            self.char2idx, self.frames, self.captions = self.construct_dataset(
                dataset_name, pickle_dir, vid_ids, labels=labels, start_end=start_end, cap=cap,
                frame_type=frame_type, in_ext=in_ext, out_ext=out_ext, fps=fps, threshold=threshold)
        else:
            self.char2idx, self.frames, self.captions = load_dataset(pickle_dir, out_ext)
        assert len(self.frames) == len(self.captions) > 0
        self.idx2char = {v: k for k, v in self.char2idx.items()}
        self.num_elements = len(self.captions)

    def __len__(self):
        return self.num_elements

    def __getitem__(self, index):
        frames = self.frames[index]
        assert len(frames.shape) == (4 if self.frame_type == "mouth_clip_seq" else 3)
        return frames, parse_caption(self.captions[index], self.char2idx)

    def parse_caption(self, cap):
        return parse_caption(cap, self.char2idx)

    @staticmethod
    def construct_dataset(dataset_name, pickle_dir, vid_ids, labels="labels.json", start_end="s_e", cap="cap",
                          frame_type="face_lmk_seq", in_ext=".npy", out_ext=".pkl", fps=29.97, threshold=0.8):
        def col(name):
            paths = [os.path.join(v, name + in_ext) for v in vid_ids]
            assert all(os.path.isfile(p) for p in paths)
            return [np.load(p, allow_pickle=True) for p in paths]     # object arrays need allow_pickle
        frames = [x for arr in col(frame_type) for x in arr]
        captions = [str(x) for arr in col(cap) for x in arr]
        start_ends = [x for arr in col(start_end) for x in arr]
        assert len(frames) == len(captions) == len(start_ends)
        assert all(len(x.shape) == (4 if frame_type == "mouth_clip_seq" else 3) for x in frames)
        frames, captions = filter_occlusions(frames, captions, start_ends, fps=fps, threshold=threshold)
        frames, captions = sort_by_seqlen(frames, captions)
        char2idx = build_vocab(dataset_name, labels)
        _ws.mkdirP(pickle_dir)
        for name, obj in (("char2idx", char2idx), ("frames", frames), ("captions", captions)):
            with open(os.path.join(pickle_dir, name + out_ext), "wb") as fh:
                pickle.dump(obj, fh)
        return char2idx, frames, captions


def load_dataset(pickle_dir, out_ext=".pkl"):
    out = []
    for name in ("char2idx", "frames", "captions"):
        path = os.path.join(pickle_dir, name + out_ext)
        assert os.path.isfile(path), "File not found: '{}'".format(path)
        with open(path, "rb") as fh:
            out.append(pickle.load(fh))
    return tuple(out)


COPY_STREAMS = 1      # host->device copy streams of DevicePrefetcher (LR_H2D_STREAMS overrides).  Measured on the B200 box:
                      # 1 -> 1.286 M frames/s end to end, 2 -> 1.285 M, 4 -> 1.108 M: the link (~19 GB/s), not the engine, limits


class DevicePrefetcher:
    """Iterate a loader of HOST batches `depth - 1` steps ahead: the (large) frames tensor of a later batch is
    copied host->device on a side stream while earlier batches compute.  Lengths and captions stay on the
    host (the trainer wants them there).  Use pinned host tensors for truly asynchronous copies.

    The device side is a ring of `depth` persistent buffers (re-allocated only when a batch's shape
    changes), so the steady state makes no allocator calls: the copy into slot k waits, on the side
    stream, for the event recorded when the consumer asked for the batch after the one that last
    used slot k (i.e. all work reading the slot has been enqueued).

    depth = 3 (two batches in flight): a copy can only START once the work that last read its slot has finished
    on the GPU, so with two slots its window is one step — enough on an idle host link (55 GB/s: 5 ms for a
    288 MB batch of clips against a 13 ms step), not when eight ranks share the host (24 GB/s measured on the
    four GPUs behind the busier root port: 12 ms plus contention stretched the step to 21.7 ms).  A third slot
    doubles the window."""

    def __init__(self, loader, device, copy_streams=None, depth=None):
        self.loader, self.device = loader, torch.device(device)
        # the big tensor is split along the batch axis over this many copy streams: one cudaMemcpyAsync keeps a single
        # copy engine busy, several in flight let the link's other engines work too
        if copy_streams is None:
            copy_streams = int(os.environ.get("LR_H2D_STREAMS", COPY_STREAMS))
        self.copy_streams = max(1, int(copy_streams))
        if depth is None:
            depth = int(os.environ.get("LR_PREFETCH_DEPTH", 3))
        self.depth = max(2, int(depth))
        self.trace = None               # a list: (start, end) timing events of every copy are appended (diagnostics)
        self.kick_mode = os.environ.get("LR_H2D_KICK", "0") == "1"
        self._kick = None

    def kick(self, wait_current=True):
        """Consumer hook (kick mode): enqueue the pending host->device copies now; with `wait_current` they start only
        after the work enqueued so far on the current stream (e.g. the forward pass) has run."""
        f, self._kick = self._kick, None
        if f is None:
            return
        if wait_current:
            self._after = torch.cuda.Event()
            self._after.record(torch.cuda.current_stream(self.device))
        else:
            self._after = None
        f()
        self._after = None

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        import collections
        sides = [torch.cuda.Stream(self.device) for _ in range(self.copy_streams)]
        depth = self.depth
        slots = [None] * depth          # device buffers
        released = [None] * depth       # event: consumer is done enqueueing work on the slot

        def stage(batch, k):
            src = batch[0]
            buf = slots[k]
            if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
                buf = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                slots[k] = buf
            n = src.shape[0] if src.dim() > 0 else 1
            parts = min(len(sides), max(1, n))
            step = -(-n // parts)
            evs = []
            for p_i in range(parts):
                lo, hi = p_i * step, min(n, (p_i + 1) * step)
                if lo >= hi:
                    break
                side = sides[p_i]
                if released[k] is not None:
                    side.wait_event(released[k])
                if getattr(self, "_after", None) is not None:
                    side.wait_event(self._after)
                trace = self.trace
                if trace is not None:
                    t_a = torch.cuda.Event(enable_timing=True)
                    t_a.record(side)
                with torch.cuda.stream(side):
                    (buf[lo:hi] if src.dim() > 0 else buf).copy_(src[lo:hi] if src.dim() > 0 else src, non_blocking=True)
                ev = torch.cuda.Event(enable_timing=trace is not None)
                ev.record(side)
                if trace is not None:
                    trace.append((t_a, ev))
                evs.append(ev)
            return buf, evs, batch[1:]

        it = iter(self.loader)
        queue = collections.deque()     # staged batches, oldest first
        n_staged = 0

        def fill():
            nonlocal n_staged
            while len(queue) < depth - 1:
                try:
                    batch = next(it)
                except StopIteration:
                    return
                queue.append((n_staged % depth, stage(batch, n_staged % depth)))
                n_staged += 1

        fill()
        while queue:
            k, (frames, ev, rest) = queue.popleft()
            if self.kick_mode:
                # the next copy is enqueued when the consumer calls kick() (e.g. after its forward pass): it then starts
                # behind the work enqueued so far instead of at the start of the step
                self._kick = fill
            else:
                fill()                   # keep depth-1 copies in flight behind the batch handed out now
            cur = torch.cuda.current_stream(self.device)
            for e in ev:
                cur.wait_event(e)
            yield (frames,) + tuple(rest)
            # the consumer came back for the next batch: everything that reads `frames` is enqueued
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            released[k] = done
            if self._kick is not None:   # the consumer never kicked during the step
                self.kick(wait_current=False)
