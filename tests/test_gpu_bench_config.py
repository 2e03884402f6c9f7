"""-m gpu: parity of EXACTLY what bench.py times, and of the BASELINE.json shapes the small-case tests do not reach.

1. The throughput configuration (`bench.configure_throughput_path()`: bf16 GEMM operands, persistent cluster
   recurrence, bf16 conv features, fused Adam, linear-space CTC at B >= 64) — one `trainer.train_ctc` step on the
   bench's own synthetic batch (`bench.synth_batch`, T=75, 100x50 clips, L in [10,30]) against the fp32 CPU port of
   the same step (`oracle/train_step.py`): log-probs <= 2e-2, CTC loss <= 1e-2 relative, gradients by relative
   Frobenius error, updated weights within Adam's first-step bound.  SURVEY §8d: "bf16 path 2e-2".
2. The fp32 parity path at the BASELINE shapes: ref-shape (B,T,68,3) BiGRU-256 at B=256 and BiLSTM-768 at B=128
   with equal and mixed T in [40,75] (log-probs and CTC loss <= 1e-4; better_model.py:53-96, ctc_loss.py:28-114),
   and a mixed-length batch through the conv front-end.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sequence as O          # noqa: E402
from oracle import train_step as TS       # noqa: E402

pytestmark = pytest.mark.gpu
SEED = 123456


def _rel_fro(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def throughput_step_parity(B, cuda, hidden=256, rnn="GRU", lr=1e-4):
    """Runs one bench-configured step on the GPU and the fp32 port on the CPU; returns the parity numbers.
    bench.py calls this too (its `parity` block), with the same thresholds."""
    import bench
    from lipreading_b200 import trainer
    from lipreading_b200.ctc import ctc_loss
    from lipreading_b200.model import VideoEncoder
    c2i = O.build_char2idx()
    restore = bench.configure_throughput_path()
    try:
        torch.manual_seed(SEED)
        enc = VideoEncoder(1728, hidden, frame_processing="conv3d", rnn_type=rnn, bidirectional=True, enable_ctc=True,
                           vocab_size=len(c2i), char2idx=c2i, device=cuda).to(cuda)
        state0 = {k: v.detach().cpu().clone() for k, v in enc.state_dict().items()}
        batch = bench.synth_batch(B, SEED, c2i)
        clips, lens, chars, char_lens = batch
        dev_batch = tuple(t.to(cuda) for t in batch)
        # pass 1: forward/backward only, for log-probs and raw gradients
        enc.train()
        lp, _, _ = enc(dev_batch[0], dev_batch[1])
        loss = ctc_loss(lp, dev_batch[2][:, 1:], dev_batch[1], char_lens - 1, "mean", cuda, host_lens=(lens, char_lens - 1))
        loss.backward()
        grads = {k: p.grad.detach().float().cpu().clone() for k, p in enc.named_parameters()}
        lp_gpu, loss_gpu = lp.detach().float().cpu(), float(loss)
        enc.zero_grad(set_to_none=True)
        # pass 2: the timed call itself — trainer.train_ctc with fused Adam and clip 50
        opt = torch.optim.Adam(enc.parameters(), lr=lr, fused=True)
        avg = trainer.train_ctc(enc, [dev_batch], opt, cuda, c2i, grad_norm=50)
        state1 = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
    finally:
        restore()
    ref = TS.CpuStep(state0, rnn, c2i, lr=lr, grad_norm=50, quantize=False)
    loss_ref, lp_ref, grads_ref = ref.step(batch)
    state_ref = ref.state()
    out = {"clips": B, "logprob_max_abs_err": float((lp_gpu - lp_ref).abs().max()),
           "ctc_loss": loss_gpu, "ctc_loss_ref_fp32": loss_ref, "ctc_loss_rel_err": abs(loss_gpu - loss_ref) / abs(loss_ref),
           "train_ctc_loss": avg,
           "grad_rel_fro_err": {k: _rel_fro(grads[k], grads_ref[k]) for k in grads_ref},
           "weight_update_max_over_lr": max(float((state1[k] - state_ref[k]).abs().max()) for k in state_ref) / lr,
           "weight_update_mean_disagreement_over_lr":
               max(float((state1[k] - state_ref[k]).abs().mean()) for k in state_ref if state_ref[k].numel() > 1000) / lr}
    return out


@pytest.mark.parametrize("B", [32, 256])
def test_throughput_configuration_matches_fp32_port(native_lib, cuda, B):
    r = throughput_step_parity(B, cuda)
    print(r)
    assert r["logprob_max_abs_err"] <= 2e-2, r
    assert r["ctc_loss_rel_err"] <= 1e-2, r
    assert abs(r["train_ctc_loss"] - r["ctc_loss"]) <= 1e-3 * abs(r["ctc_loss"])      # the step sees the same loss
    for k, e in r["grad_rel_fro_err"].items():
        assert e <= 8e-2, (k, e, r)                 # bf16 operands through three conv stages + 75 recurrent steps
    # Adam's first step moves every weight by lr * g/(|g|+eps): two runs can differ by at most 2 lr per weight, and
    # do so only where bf16 noise flips the sign of a near-zero gradient
    assert r["weight_update_max_over_lr"] <= 2.02, r
    assert r["weight_update_mean_disagreement_over_lr"] <= 0.25, r


def _ref_shape_batch(B, T, mixed, g, c2i):
    if mixed:
        lens = torch.randint(40, T + 1, (B,), generator=g).sort().values
        lens[-1] = T
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    frames = torch.randn(B, T, 68, 3, generator=g)
    for b in range(B):
        frames[b, int(lens[b]):] = 0
    L = torch.randint(10, 31, (B,), generator=g)
    chars = torch.zeros(B, int(L.max()) + 2, dtype=torch.long)
    for b in range(B):
        n = int(L[b])
        chars[b, 0] = c2i["<BOS>"]
        chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
        chars[b, 1 + n] = c2i["<EOS>"]
    return frames, lens, chars, L + 2


@pytest.mark.parametrize("rnn,H,B", [("GRU", 256, 256), ("LSTM", 768, 128)])
@pytest.mark.parametrize("mixed", [False, True])
def test_fp32_path_at_baseline_shapes(native_lib, cuda, rnn, H, B, mixed):
    """BASELINE.md §3.1 inputs: randn(B,75,68,3), labels L in [10,30], (i) all T=75, (ii) ascending mixed T in [40,75]."""
    from lipreading_b200.ctc import ctc_loss
    from lipreading_b200.model import VideoEncoder
    c2i = O.build_char2idx()
    torch.manual_seed(SEED)
    enc = VideoEncoder(204, H, rnn_type=rnn, bidirectional=True, enable_ctc=True, vocab_size=len(c2i), char2idx=c2i,
                       device=cuda).to(cuda)
    g = torch.Generator().manual_seed(SEED + (1 if mixed else 0))
    frames, lens, chars, char_lens = _ref_shape_batch(B, 75, mixed, g, c2i)
    lp, hidden, final = enc(frames.to(cuda), lens.to(cuda))
    loss = ctc_loss(lp, chars[:, 1:].to(cuda), lens, char_lens - 1, "mean", cuda)
    loss.backward()
    state = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    params = {k: torch.nn.Parameter(v.clone()) for k, v in state.items()}
    lp_r, hidden_r, final_r = O.encoder_forward(params, frames, lens, rnn, True, c2i)
    loss_r = O.ctc_loss_wrapper(lp_r, chars[:, 1:], lens, char_lens - 1, "mean")
    loss_r.backward()
    assert float((lp.detach().cpu() - lp_r.detach()).abs().max()) <= 1e-4
    assert float((hidden.detach().cpu() - hidden_r.detach()).abs().max()) <= 1e-4
    assert abs(float(loss) - float(loss_r)) <= 1e-4 * max(1.0, abs(float(loss_r)))
    for k, p in enc.named_parameters():
        ref = params[k].grad
        assert float((p.grad.cpu() - ref).abs().max()) <= 2e-4 * max(1e-3, float(ref.abs().max())), k


def test_mixed_length_clips_through_conv_front_end(native_lib, cuda):
    """Padded mouth clips of mixed length (T in [40,75]) through conv front-end -> BiGRU (fp32 GEMMs, per-step
    kernels) -> CTC against the oracle with bf16-rounded conv operands."""
    from lipreading_b200.ctc import ctc_loss
    from lipreading_b200.model import VideoEncoder
    c2i = O.build_char2idx()
    torch.manual_seed(SEED)
    enc = VideoEncoder(1728, 64, frame_processing="conv3d", rnn_type="GRU", bidirectional=True, enable_ctc=True,
                       vocab_size=len(c2i), char2idx=c2i, device=cuda).to(cuda)
    g = torch.Generator().manual_seed(5)
    B, T = 12, 75
    lens = torch.randint(40, T + 1, (B,), generator=g).sort().values
    lens[-1] = T
    clips = torch.randint(0, 256, (B, T, 100, 50, 3), dtype=torch.uint8, generator=g)
    for b in range(B):
        clips[b, int(lens[b]):] = 0                        # collate pads with zero frames
    L = torch.randint(10, 31, (B,), generator=g)
    chars = torch.zeros(B, int(L.max()) + 2, dtype=torch.long)
    for b in range(B):
        n = int(L[b])
        chars[b, 0], chars[b, 1 + n] = 1, 2
        chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
    char_lens = L + 2
    lp, _, _ = enc(clips.to(cuda), lens.to(cuda))
    loss = ctc_loss(lp, chars[:, 1:].to(cuda), lens, char_lens - 1, "mean", cuda)
    loss.backward()
    state = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    ref = TS.CpuStep(state, "GRU", c2i, quantize=True)
    lp_r, loss_r = ref.loss((clips, lens, chars, char_lens))
    loss_r.backward()
    assert float((lp.detach().cpu() - lp_r.detach()).abs().max()) <= 5e-3
    assert abs(float(loss) - float(loss_r)) <= 1e-3 * abs(float(loss_r))
    for k, p in enc.named_parameters():
        assert _rel_fro(p.grad.cpu(), ref.params[k].grad) <= 3e-2, k
