"""CPU tests: oracle/sequence.py is pinned against (a) the golden vectors generated from the
UNMODIFIED reference (tests/golden/make_golden.py) and (b) the live reference when its checkout
is present (build container only)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness
from oracle import sequence as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "seq_*.npz")))


def _load(path):
    z = np.load(path)
    meta = z["meta"]
    return z, str(meta[0]), int(meta[1]), bool(int(meta[2])), str(meta[3])


def test_golden_files_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_encoder_and_ctc_match_golden(path):
    z, rnn_type, H, bi, attn = _load(path)
    c2i = O.build_char2idx()
    state = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc.")}
    frames, lens = torch.from_numpy(z["frames"]), torch.from_numpy(z["frame_lens"])
    chars, char_lens = torch.from_numpy(z["chars"]), torch.from_numpy(z["char_lens"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    lp, hidden, final = O.encoder_forward(leaves, frames, lens, rnn_type, bi, c2i)
    assert np.abs(lp.detach().numpy() - z["log_probs"]).max() < 1e-5
    assert np.abs(hidden.detach().numpy() - z["hidden"]).max() < 1e-5
    fh = final[0] if isinstance(final, tuple) else final
    assert np.abs(fh.detach().numpy() - z["final_h"]).max() < 1e-5
    if isinstance(final, tuple):
        assert np.abs(final[1].detach().numpy() - z["final_c"]).max() < 1e-5
    labels, label_lens = chars[:, 1:], char_lens - 1
    for red in ("mean", "sum"):
        loss = O.ctc_loss_wrapper(lp, labels, lens, label_lens, red)
        assert abs(float(loss) - float(z["ctc_" + red])) < 1e-4 * max(1.0, abs(float(z["ctc_" + red])))
    loss = O.ctc_loss_wrapper(lp, labels, lens, label_lens, "mean")
    loss.backward()
    for k, leaf in leaves.items():
        ref = z["grad_ctc_mean." + k]
        assert np.abs(leaf.grad.numpy() - ref).max() <= 1e-4 * max(1e-3, np.abs(ref).max()), k
    # the masking formulation (what the kernels do) == the packed formulation (what the reference does)
    weights = {k[4:]: v for k, v in state.items() if k.startswith("rnn.")}
    hm, _ = O.rnn_masked(frames.reshape(frames.shape[0], frames.shape[1], -1), lens, weights, rnn_type, bi)
    assert np.abs(hm.numpy()[:, : z["hidden"].shape[1]] - z["hidden"]).max() < 1e-5


def test_ctc_pure_numpy_matches_torch():
    g = torch.Generator().manual_seed(3)
    lp = torch.randn(1, 14, 9, generator=g).log_softmax(-1)
    tgt = torch.tensor([[3, 3, 5, 1, 1, 2]])
    lpr = lp.clone().double().requires_grad_(True)
    nll = O.ctc_nll_torch(lpr, tgt, torch.tensor([14]), torch.tensor([6]))
    nll.sum().backward()
    n0, g0 = O.ctc_alpha_beta(lp[0].numpy(), tgt[0].numpy())
    assert abs(n0 - float(nll)) < 1e-9
    assert np.abs(g0 - lpr.grad[0].numpy()).max() < 1e-9


def test_ctc_wrapper_edge_semantics():
    g = torch.Generator().manual_seed(5)
    lp = torch.randn(4, 10, 65, generator=g).log_softmax(-1)
    fl = torch.tensor([6, 6, 10, 10])
    # all labels too long -> None (ctc_loss.py:49-51)
    assert O.ctc_loss_wrapper(lp, torch.zeros(4, 300, dtype=torch.long), fl, torch.full((4,), 300), "mean") is None
    # one infeasible sample (label longer than frames): dropped from its run, run re-weighted
    lab = torch.randint(4, 60, (4, 8), generator=g)
    ll = torch.tensor([3, 8, 4, 4])
    loss = O.ctc_loss_wrapper(lp, lab, fl, ll, "mean")
    assert loss is not None and torch.isfinite(loss)
    # decreasing frame_lens trip the reference's assertion (ctc_loss.py:39)
    with pytest.raises(AssertionError):
        O.ctc_loss_wrapper(lp, lab, torch.tensor([10, 6, 6, 6]), ll, "mean")


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present (GPU box)")
def test_oracle_matches_live_reference_including_inf_paths():
    ref = ref_harness.load()
    g = torch.Generator().manual_seed(17)
    B, T, C = 8, 16, 65
    fl = torch.tensor([8, 8, 8, 12, 12, 16, 16, 16])
    lab = torch.randint(4, 60, (B, 10), generator=g)
    ll = torch.tensor([3, 4, 10, 5, 5, 6, 4, 3])          # sample 2 infeasible (10 labels, 8 frames)
    lp = torch.randn(B, T, C, generator=g).log_softmax(-1)
    for red in ("mean", "sum"):
        a = O.ctc_loss_wrapper(lp, lab, fl, ll, red)
        b = ref.ctc_loss.ctc_loss(lp, lab, fl, ll, red, "cpu")
        assert abs(float(a) - float(b)) < 1e-5 * max(1.0, abs(float(b)))
    # whole first run infeasible -> the reference does not advance prev_change_point (ctc_loss.py:91-93)
    ll2 = torch.tensor([10, 10, 10, 5, 5, 6, 4, 3])
    for red in ("mean", "sum"):
        a = O.ctc_loss_wrapper(lp, lab, fl, ll2, red)
        b = ref.ctc_loss.ctc_loss(lp, lab, fl, ll2, red, "cpu")
        assert abs(float(a) - float(b)) < 1e-5 * max(1.0, abs(float(b)))
    # decoder restatement == reference CharDecodingStep for every attention type
    c2i = O.build_char2idx()
    for attn in ("none", "dot", "general", "1_layer_nn", "concat"):
        torch.manual_seed(0)
        enc = ref.better_model.VideoEncoder(204, 6, rnn_type="GRU", bidirectional=True, enable_ctc=True,
                                            vocab_size=64, char2idx=c2i)
        dec = ref.better_model.CharDecodingStep(enc, 5, 64, c2i, attention_type=attn, attn_hidden_size=7)
        mine = O.OracleDecoder(12, "GRU", 5, 64, c2i, attention_type=attn, attn_hidden_size=7)
        mine.load_state_dict(dec.state_dict())
        eh = torch.randn(3, 9, 12, generator=g)
        st = torch.randn(1, 3, 12, generator=g)
        inp = torch.tensor([1, 7, 20])
        el = torch.tensor([9, 4, 6])
        a, sa = mine(inp, st, el, eh)
        b, sb = dec(inp, st, el, eh)
        assert float((a - b).abs().max()) < 1e-5 and float((sa - sb).abs().max()) < 1e-6, attn


@pytest.mark.parametrize("T,L,seed", [(20, 6, 0), (75, 30, 1), (75, 1, 2), (300, 31, 3), (9, 0, 4)])
def test_linear_rescaled_ctc_recursion_equals_log_space(T, L, seed):
    """The recursion ctc_linear_warp_kernel runs (probabilities, even-frame rescaling, constant normaliser rho,
    gradient without a division by p) is the same function as the published log-space recursion and torch's native CTC:
    float64 restatement against both; in float32 arithmetic it stays inside the 1e-4 parity bar."""
    g = torch.Generator().manual_seed(100 + seed)
    C = 65
    lp = torch.randn(T, C, generator=g).log_softmax(-1)
    lp[:, 1:3] -= 100.0                                   # the two masked classes of the encoder (log(1e-45))
    tgt = torch.randint(3, C, (L,), generator=g)
    if L >= 4:
        tgt[2] = tgt[1]                                   # a repeat: no s-2 -> s transition there
        tgt[-1] = tgt[0]                                  # the same class twice, apart
    n_log, g_log = O.ctc_alpha_beta(lp.numpy(), tgt.numpy())
    n_lin, g_lin, spread = O.ctc_linear_rescaled(lp.numpy(), tgt.numpy())
    assert abs(n_lin - n_log) < 1e-9 * max(1.0, abs(n_log))
    assert np.abs(g_lin - g_log).max() < 1e-9
    assert spread < 1e-9
    lp_t = lp.double().clone().requires_grad_(True)
    nll_t = O.ctc_nll_torch(lp_t[None], tgt[None], torch.tensor([T]), torch.tensor([L]))
    nll_t.sum().backward()
    assert abs(float(nll_t) - n_lin) < 1e-8 * max(1.0, abs(n_lin))
    assert np.abs(lp_t.grad.numpy() - g_lin).max() < 1e-8
    n32, g32, spread32 = O.ctc_linear_rescaled(lp.numpy(), tgt.numpy(), dtype=np.float32)
    assert abs(n32 - n_log) < 1e-4 * max(1.0, abs(n_log))
    assert np.abs(g32 - g_log).max() < 1e-4 and spread32 < 1e-4


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
@pytest.mark.parametrize("tfr", [1.0, 0.5])
def test_oracle_train_step_port_matches_golden(path, tfr):
    """oracle/train_step.CpuReferenceTrain (bench.py's CPU arm for the ref-shape block) reproduces the unmodified
    reference's train() step: teacher forcing 1 from seq_<case>.npz, teacher forcing 0.5 — with the reference's RNG
    consumption (rand(1) + one multinomial per label position) — from eval_<case>.npz."""
    from oracle import train_step as TS
    z, rnn_type, H, bi, attn = _load(path)
    ze = np.load(path.replace("seq_", "eval_"))
    c2i = O.build_char2idx()
    pre, seed = ("enc.", 123456 + 1) if tfr == 1.0 else ("enc_after.", 123456 + 3)
    dpre = "dec." if tfr == 1.0 else "dec_after."
    enc_state = {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    dec = O.OracleDecoder(H * (2 if bi else 1), rnn_type, 10, 64, c2i, attention_type=attn)
    dec.load_state_dict({k[len(dpre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(dpre)})
    batch = tuple(torch.from_numpy(z[k]) for k in ("frames", "frame_lens", "chars", "char_lens"))
    ref = TS.CpuReferenceTrain(enc_state, dec, rnn_type, c2i, lr=1e-3, grad_norm=50)
    torch.manual_seed(seed)
    d_loss, c_loss = ref.step(batch, teacher_forcing_ratio=tfr)
    want_d, want_c = (z["train_dec_loss"], z["train_ctc_loss"]) if tfr == 1.0 else (ze["tfr_dec_loss"], ze["tfr_ctc_loss"])
    assert abs(d_loss - float(want_d)) < 1e-5 and abs(c_loss - float(want_c)) < 1e-4
    after = (lambda k: z["enc_after." + k]) if tfr == 1.0 else (lambda k: ze["enc_tfr." + k])
    for k, p in ref.enc.items():
        assert np.abs(p.detach().numpy() - after(k)).max() < 2e-5, k
    dafter = (lambda k: z["dec_after." + k]) if tfr == 1.0 else (lambda k: ze["dec_tfr." + k])
    for k, p in dec.state_dict().items():
        assert np.abs(p.numpy() - dafter(k)).max() < 2e-5, k
