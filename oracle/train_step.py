"""fp32 CPU port of ONE north-star training step (what bench.py times on the GPU).  TEST INFRASTRUCTURE.

    u8 clips (B,T,H,W,3) -> STCNN front-end (oracle/conv3d.py, F.conv3d fp32) -> packed nn.{GRU,LSTM,RNN}
    (the reference's own route, better_model.py:64-89, through oracle/sequence.rnn_packed) -> Linear + masked
    log-softmax (better_model.py:91-94) -> ctc_loss wrapper 'mean' (src/train/ctc_loss.py:28-114) -> backward ->
    clip_grad_norm_ -> Adam          [loop shape: src/scripts/archive/train_model.py, SURVEY §8 row a16]

Used by tests/test_gpu_bench_config.py as the checker of the exact configuration bench.py runs, and by bench.py's
`cpu_baseline` / `--impl reference` leg as the timed CPU path.  The landmark ('flatten') variant of the same step
is `landmark_step`; it is the oracle restatement of a CTC-only epoch of the reference's encoder."""
import torch

from . import conv3d as OC
from . import sequence as O

CONV_SHAPES = {"conv1": (3, 32, (3, 5, 5), (1, 2, 2), (1, 2, 2)),
               "conv2": (32, 64, (3, 5, 5), (1, 1, 1), (1, 2, 2)),
               "conv3": (64, 96, (3, 3, 3), (1, 1, 1), (1, 1, 1))}


class CpuStep:
    """Holds fp32 parameters under the SAME names as lipreading_b200.model.VideoEncoder's state_dict
    (front.conv{1,2,3}.{weight,bias}, rnn.*, output_proj.*), so weights move between the two verbatim."""

    def __init__(self, state, rnn_type, char2idx, lr=1e-4, grad_norm=50, quantize=False):
        self.params = {k: torch.nn.Parameter(v.detach().clone().float()) for k, v in state.items()}
        self.rnn_type, self.char2idx, self.grad_norm, self.quantize = rnn_type, char2idx, grad_norm, quantize
        self.bidirectional = "rnn.weight_ih_l0_reverse" in self.params
        self.conv = any(k.startswith("front.") for k in self.params)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)
        self.log_mask = O.log_mask_vector(len(char2idx), char2idx)

    def forward(self, frames, lens):
        if self.conv:
            cp = {k[len("front."):]: p for k, p in self.params.items() if k.startswith("front.")}
            feat, _ = OC.stcnn_forward(frames, cp, quantize=self.quantize)
        else:
            feat = frames.reshape(frames.shape[0], frames.shape[1], -1).float()
        weights = {k[len("rnn."):]: p for k, p in self.params.items() if k.startswith("rnn.")}
        hidden, _ = O.rnn_packed(feat, lens, weights, self.rnn_type, self.bidirectional)
        logits = hidden @ self.params["output_proj.weight"].t() + self.params["output_proj.bias"]
        return O.masked_log_softmax(logits, self.log_mask)

    def loss(self, batch):
        frames, lens, chars, char_lens = batch
        lp = self.forward(frames, lens)
        return lp, O.ctc_loss_wrapper(lp, chars[:, 1:], lens, char_lens - 1, "mean")

    def step(self, batch):
        """one optimizer step; returns (loss, log_probs, {name: grad before clipping})"""
        lp, loss = self.loss(batch)
        self.opt.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in self.params.items()}
        if self.grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(list(self.params.values()), self.grad_norm)
        self.opt.step()
        return float(loss), lp.detach(), grads

    def state(self):
        return {k: p.detach().clone() for k, p in self.params.items()}


def random_state(hidden, rnn_type, char2idx, conv=True, seed=123456, frame_dim=None):
    """Seeded fp32 parameters with VideoEncoder's names (torch default initialisers)."""
    torch.manual_seed(seed)
    state = {}
    if conv:
        for name, (ci, co, k, s, p) in CONV_SHAPES.items():
            m = torch.nn.Conv3d(ci, co, k, s, p)
            state["front.%s.weight" % name], state["front.%s.bias" % name] = m.weight.detach(), m.bias.detach()
    rnn = getattr(torch.nn, rnn_type)(frame_dim or (1728 if conv else 204), hidden, bidirectional=True, batch_first=True)
    for k, v in rnn.named_parameters():
        state["rnn." + k] = v.detach()
    proj = torch.nn.Linear(2 * hidden, len(char2idx) + 1)
    state["output_proj.weight"], state["output_proj.bias"] = proj.weight.detach(), proj.bias.detach()
    return state


class CpuReferenceTrain:
    """The reference's `train()` epoch body (src/train/train_better_model.py:7-87) restated on the oracle's
    encoder / decoder for ONE pre-collated ref-shape batch (B,T,68,3): encoder -> CTC 'mean' (aux) -> teacher-forced
    decoder loop with NLL 'sum' / n_tokens -> decoder backward (retain graph) -> CTC backward -> separate clips ->
    optimizer step.  Pinned by tests/test_oracle_golden.py against the golden `train()` step of the unmodified
    reference.  bench.py times it as the CPU side of its `ref_shape` block."""

    def __init__(self, enc_state, dec_module, rnn_type, char2idx, lr=1e-4, grad_norm=50):
        self.enc = {k: torch.nn.Parameter(v.detach().clone().float()) for k, v in enc_state.items()}
        self.dec, self.rnn_type, self.char2idx, self.grad_norm = dec_module, rnn_type, char2idx, grad_norm
        self.bidirectional = "rnn.weight_ih_l0_reverse" in self.enc
        self.opt = torch.optim.Adam(list(self.enc.values()) + list(dec_module.parameters()), lr=lr)

    def step(self, batch, teacher_forcing_ratio=1.0):
        frames, lens, chars, char_lens = batch
        pad, bos = self.char2idx["<PAD>"], self.char2idx["<BOS>"]
        labels, label_lens = chars[:, 1:], char_lens - 1
        lp, enc_h, state = O.encoder_forward(self.enc, frames, lens, self.rnn_type, self.bidirectional, self.char2idx)
        ctc = O.ctc_loss_wrapper(lp, labels, lens, label_lens, "mean")
        if ctc is None:
            return None
        prev = torch.full((frames.shape[0],), bos, dtype=torch.long)
        dec_loss = 0
        for i in range(int(label_lens.max())):
            tf = bool(torch.rand(1) < teacher_forcing_ratio)
            logp, state = self.dec(chars[:, i] if tf else prev, state, lens, enc_h)
            dec_loss = dec_loss + torch.nn.functional.nll_loss(logp, labels[:, i], ignore_index=pad, reduction="sum")
            prev = logp.exp().multinomial(1).squeeze(-1)
        dec_loss = dec_loss / (labels != pad).sum()
        self.opt.zero_grad()
        dec_loss.backward(retain_graph=True)
        ctc.backward()
        if self.grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(list(self.enc.values()), self.grad_norm)
            torch.nn.utils.clip_grad_norm_(list(self.dec.parameters()), self.grad_norm)
        self.opt.step()
        return float(dec_loss.detach()), float(ctc.detach())
