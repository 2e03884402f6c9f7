"""-m gpu: the vectorised teacher-forced decoder — the fused attention kernel vs the reference's
formulation (better_model.py:195-223 with allennlp's masked_softmax), and CharDecodingStep.forward_sequence vs the
reference's per-step loop (forward chained through final_state), values and gradients.  fp32 both sides: 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from oracle import sequence as O

pytestmark = pytest.mark.gpu


def _ref_attention(q, enc, lens):
    T = enc.shape[1]
    scores = torch.einsum("bth,blh->blt", enc, q)
    m = (torch.arange(T, device=q.device).unsqueeze(0) < lens.unsqueeze(1)).float().unsqueeze(1)
    w = F.softmax(scores * m, dim=-1) * m
    w = w / (w.sum(dim=-1, keepdim=True) + 1e-13)
    return torch.bmm(w, enc), w


@pytest.mark.parametrize("B,L,T,H", [(5, 7, 20, 64), (3, 32, 75, 512), (4, 1, 9, 30), (2, 6, 300, 256), (3, 18, 40, 128)])
def test_attn_context_matches_reference_formula(native_lib, cuda, B, L, T, H):
    from lipreading_b200 import functional as LF
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(B, L, H, generator=g) * 0.3).double()
    enc = (torch.randn(B, T, H, generator=g) * 0.5).double()
    lens = torch.randint(max(1, T // 3), T + 1, (B,), generator=g)
    lens[0] = T
    up = torch.randn(B, L, H, generator=g).double()
    qr, er = q.clone().requires_grad_(True), enc.clone().requires_grad_(True)
    c_ref, w_ref = _ref_attention(qr, er, lens)
    (c_ref * up).sum().backward()
    qd = q.float().to(cuda).requires_grad_(True)
    ed = enc.float().to(cuda).requires_grad_(True)
    c, w = LF.attn_context(qd, ed, lens.to(cuda))
    (c * up.float().to(cuda)).sum().backward()
    torch.cuda.synchronize()
    assert float((w.cpu().double() - w_ref).abs().max()) < 1e-5
    assert float((c.cpu().double() - c_ref).abs().max()) < 1e-5 * max(1.0, float(c_ref.abs().max()))
    for got, ref in ((qd.grad, qr.grad), (ed.grad, er.grad)):
        assert float((got.cpu().double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    # masked encoder positions get no gradient, masked weights are exactly zero
    for b in range(B):
        assert float(w[b, :, int(lens[b]):].abs().max() if int(lens[b]) < T else 0.0) == 0.0
        assert float(ed.grad[b, int(lens[b]):].abs().max() if int(lens[b]) < T else 0.0) == 0.0


@pytest.mark.parametrize("rnn_type", ["GRU", "LSTM"])
@pytest.mark.parametrize("attn", ["none", "dot", "general", "1_layer_nn", "concat"])
def test_forward_sequence_equals_step_loop(native_lib, cuda, rnn_type, attn):
    from lipreading_b200.model import CharDecodingStep, VideoEncoder
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False      # both sides run torch's RNN cell on the device: keep it fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    c2i = O.build_char2idx()
    H, B, T, L = 32, 6, 17, 9
    enc_m = VideoEncoder(204, H, rnn_type=rnn_type, bidirectional=True, enable_ctc=False, vocab_size=64, char2idx=c2i,
                         device=cuda).to(cuda)
    dec = CharDecodingStep(enc_m, char_dim=12, vocab_size=64, char2idx=c2i, attention_type=attn, attn_hidden_size=20,
                           device=cuda).to(cuda)
    g = torch.Generator().manual_seed(3)
    enc_h0 = torch.randn(B, T, 2 * H, generator=g).to(cuda)
    lens = torch.tensor([17, 17, 12, 9, 17, 5]).to(cuda)
    chars = torch.randint(3, 64, (B, L), generator=g).to(cuda)
    chars[2, 6:] = 0                                                    # PAD inputs past a short label
    up = torch.randn(B, L, 64, generator=g).to(cuda)

    def state():
        h = torch.randn(1, B, 2 * H, generator=torch.Generator().manual_seed(9)).to(cuda).requires_grad_(True)
        if rnn_type == "LSTM":
            c = torch.randn(1, B, 2 * H, generator=torch.Generator().manual_seed(10)).to(cuda).requires_grad_(True)
            return (h, c), [h, c]
        return h, [h]

    # reference structure: one forward() per position, chained through the returned state
    e1 = enc_h0.clone().requires_grad_(True)
    st, leaves1 = state()
    outs = []
    for i in range(L):
        lp, st = dec(chars[:, i], st, lens, e1)
        outs.append(lp)
    ref = torch.stack(outs, 1)
    dec.zero_grad()
    (ref * up).sum().backward()
    g_ref = {k: v.grad.clone() for k, v in dec.named_parameters() if v.grad is not None}
    # vectorised pass
    e2 = enc_h0.clone().requires_grad_(True)
    st2, leaves2 = state()
    got, fin = dec.forward_sequence(chars, st2, lens, e2)
    dec.zero_grad()
    (got * up).sum().backward()
    torch.cuda.synchronize()
    # (the two masked classes sit at ~-107, where one fp32 ulp is 7.6e-6)
    assert torch.allclose(got, ref, rtol=2e-7, atol=1e-5), float((got - ref).abs().max())
    fin_ref = st if isinstance(st, tuple) else (st,)
    fin_got = fin if isinstance(fin, tuple) else (fin,)
    for a, b in zip(fin_got, fin_ref):
        assert float((a - b).abs().max()) < 1e-5
    if attn != "none":
        assert float((e2.grad - e1.grad).abs().max()) < 2e-5 * max(1.0, float(e1.grad.abs().max()))
    for a, b in zip(leaves2, leaves1):
        assert float((a.grad - b.grad).abs().max()) < 2e-5 * max(1.0, float(b.grad.abs().max()))
    for k, v in dec.named_parameters():
        if k in g_ref:
            assert float((v.grad - g_ref[k]).abs().max()) < 1e-4 * max(1.0, float(g_ref[k].abs().max())), k


def test_train_and_eval_sequence_decode_equals_step_loop(native_lib, cuda):
    """trainer.train / eval with the vectorised decode vs the reference's step loop on the same weights and batch."""
    from lipreading_b200 import trainer
    from lipreading_b200.model import CharDecodingStep, VideoEncoder
    c2i = O.build_char2idx()
    res = {}
    for seq in (False, True):
        torch.manual_seed(77)
        enc = VideoEncoder(204, 32, rnn_type="GRU", bidirectional=True, enable_ctc=True, vocab_size=64, char2idx=c2i,
                           device=cuda).to(cuda)
        dec = CharDecodingStep(enc, char_dim=10, vocab_size=64, char2idx=c2i, attention_type="dot", device=cuda).to(cuda)
        g = torch.Generator().manual_seed(1)
        B, T = 6, 30
        frames = torch.randn(B, T, 68, 3, generator=g)
        frame_lens = torch.tensor([22, 24, 26, 28, 30, 30])
        for b in range(B):
            frames[b, int(frame_lens[b]):] = 0
        Ls = [5, 7, 4, 8, 6, 8]
        chars = torch.zeros(B, max(Ls) + 2, dtype=torch.long)
        for b, n in enumerate(Ls):
            chars[b, 0] = c2i["<BOS>"]
            chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
            chars[b, 1 + n] = c2i["<EOS>"]
        batch = (frames, frame_lens, chars, torch.tensor(Ls) + 2)
        opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-3)
        trainer.SEQUENCE_DECODE = seq
        try:
            losses = trainer.train(enc, dec, [batch], opt, cuda, c2i, teacher_forcing_ratio=1, grad_norm=50)
            ev = trainer.eval(enc, dec, [batch], cuda, c2i)
        finally:
            trainer.SEQUENCE_DECODE = True
        res[seq] = (losses, float(ev[0]), float(ev[2]), {k: v.detach().clone() for k, v in dec.state_dict().items()})
    assert abs(res[True][0][0] - res[False][0][0]) < 1e-5 and abs(res[True][0][1] - res[False][0][1]) < 1e-5
    assert abs(res[True][1] - res[False][1]) < 1e-4 and res[True][2] == res[False][2]
    for k in res[True][3]:
        assert float((res[True][3][k] - res[False][3][k]).abs().max()) < 2.1e-3, k      # one Adam step of lr 1e-3
