"""End-to-end inference stream frames -> characters (BASELINE config 5).

The reference has no single frames->chars path (its two halves meet on disk).  Here one call takes a
batch of clips (u8 frames for the conv front-end, or landmark tensors for the reference's 'flatten'
encoder), runs the encoder and decodes on the device with the greedy CTC kernel; only the token ids
come back to the host."""
import torch

from . import functional as LF


class Recognizer:
    def __init__(self, encoder, char2idx):
        assert encoder.enable_ctc
        self.encoder = encoder.eval()
        self.idx2char = {v: k for k, v in char2idx.items()}

    @torch.no_grad()
    def tokens(self, frames, frame_lens):
        """frames on the encoder's device -> (tokens (B,T) int32 CTC classes minus 1 = char ids, lens)."""
        log_probs, _, _ = self.encoder(frames, frame_lens)
        tok, n = LF.ctc_greedy_decode(log_probs, frame_lens.clamp(max=log_probs.shape[1]))
        return tok - 1, n            # class c (blank = 0) is char id c-1 (labels+1, ctc_loss.py:80)

    @torch.no_grad()
    def __call__(self, frames, frame_lens):
        tok, n = self.tokens(frames, frame_lens)
        tok, n = tok.cpu(), n.cpu()
        out = []
        for b in range(tok.shape[0]):
            # drop the four markers by ID (PAD, BOS, EOS, UNK = 0..3); '<' and '>' are ordinary characters
            out.append("".join(self.idx2char.get(int(i), "") for i in tok[b, : int(n[b])] if int(i) >= 4))
        return out


class FrameRecognizer:
    """Raw frames -> characters on the device (BASELINE config 5): the two halves the reference
    joins through files on disk (generate_dataview.py -> train.py) chained in memory:

        frames (N,H,W,3) u8 + face boxes -> [rect geometry -> /255 + warp 256x256 -> position map -> 68 landmarks]
        -> mouth crop (N,h,w,3) u8 -> clips (N/T,T,h,w,3) -> conv front-end -> BiGRU -> greedy CTC -> token ids

    `prn` is a lipreading_b200.face.PRN (its `predict_batch` is the position-map CNN plug); boxes come from the
    caller or from `detector(frames) -> (N,4) (left,right,top,bottom)` (row a1: dlib's HOG model is un-vendored).
    Only the token ids leave the GPU."""

    def __init__(self, encoder, char2idx, prn, detector=None, mouth_hw=(100, 50), batch=64):
        assert getattr(encoder, "frame_processing", "") == "conv3d", "the frame stream feeds the conv front-end"
        self.rec = Recognizer(encoder, char2idx)
        self.prn, self.detector, self.mouth_hw, self.batch = prn, detector, mouth_hw, batch

    @torch.no_grad()
    def mouth_clips(self, frames, rects=None):
        """frames (N,H,W,3) u8 on the device -> (N,h,w,3) u8 mouth crops (+ the landmarks (N,68,3) f64)."""
        if rects is None:
            assert self.detector is not None, "no face boxes and no detector installed"
            rects = self.detector(frames)
        rects = torch.as_tensor(rects, dtype=torch.int32)
        crops, lmks = [], []
        for i in range(0, frames.shape[0], self.batch):
            fr = frames[i:i + self.batch]
            lmk, geom = self.prn.process_batch(fr, rects[i:i + self.batch])
            clip, _ = LF.mouth_crop(fr, lmk, geom[0], self.mouth_hw[0], self.mouth_hw[1])
            crops.append(clip)
            lmks.append(lmk)
        return torch.cat(crops, 0), torch.cat(lmks, 0)

    @torch.no_grad()
    def tokens(self, frames, clip_len, rects=None):
        """frames of consecutive clips, `clip_len` frames each -> (token ids (B,T) int32, lens (B))."""
        n = frames.shape[0]
        assert n % clip_len == 0
        crops, _ = self.mouth_clips(frames, rects)
        clips = crops.reshape(n // clip_len, clip_len, *crops.shape[1:])
        lens = torch.full((clips.shape[0],), clip_len, dtype=torch.int64, device=clips.device)
        return self.rec.tokens(clips, lens)

    @torch.no_grad()
    def __call__(self, frames, clip_len, rects=None):
        tok, n = self.tokens(frames, clip_len, rects)
        tok, n = tok.cpu(), n.cpu()
        return ["".join(self.rec.idx2char.get(int(i), "") for i in tok[b, : int(n[b])] if int(i) >= 4)
                for b in range(tok.shape[0])]
