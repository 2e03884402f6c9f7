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
