"""Generate the committed golden vectors by running the UNMODIFIED reference (sequence half) in the
build container.  Usage (from the repo root, with /root/reference present):

    python tests/golden/make_golden.py

Writes tests/golden/seq_<case>.npz: seeded inputs, the reference modules' state_dict, and their
outputs (log-probs, hidden states, final state, CTC 'mean'/'sum' incl. the mixed-length weighting
quirk, gradients w.r.t. every encoder parameter, one full train() step's losses and updated
weights).  The reference is imported through oracle/ref_harness.py (two shims: allennlp.nn.util,
spacy).  Also copies the reference's own index fixtures (uv_kpt_ind.txt, face_ind.txt, the PRNet checkpoint index).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.dirname(os.path.abspath(__file__))
SEED = 123456                     # the reference's default seed (src/scripts/train.py:165)

CASES = {
    # name: rnn_type, hidden, bidirectional, B, T, mixed lengths, attention
    "gru_bi_equal": ("GRU", 16, True, 6, 20, False, "none"),
    "gru_bi_mixed": ("GRU", 16, True, 8, 24, True, "none"),
    "lstm_bi_mixed": ("LSTM", 12, True, 7, 22, True, "1_layer_nn"),
    "lstm_uni_mixed": ("LSTM", 12, False, 5, 18, True, "dot"),
    "rnn_bi_mixed": ("RNN", 8, True, 5, 16, True, "none"),
}


def make_batch(g, B, T, mixed, char2idx):
    if mixed:
        lens = torch.randint(max(T // 2, 10), T + 1, (B,), generator=g).sort().values
        lens[-1] = T
        if B > 3:
            lens[1] = lens[0]                      # at least one run of equal lengths (quirk weighting)
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    frames = torch.randn(B, T, 68, 3, generator=g)
    for b in range(B):
        frames[b, int(lens[b]):] = 0
    cap_lens = torch.randint(3, 8, (B,), generator=g)
    Lmax = int(cap_lens.max()) + 2
    chars = torch.zeros(B, Lmax, dtype=torch.long)
    for b in range(B):
        n = int(cap_lens[b])
        chars[b, 0] = char2idx["<BOS>"]
        chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
        chars[b, 1 + n] = char2idx["<EOS>"]
    return frames, lens, chars, cap_lens + 2


def main():
    ref = ref_harness.load()
    dl, bm, tb, cl = ref.data_loader, ref.better_model, ref.train_better_model, ref.ctc_loss
    char2idx = dict(dl._markers2Id)
    for ch in dl._labels:
        char2idx[ch] = len(char2idx)
    assert len(char2idx) == 64
    for name, (rnn_type, H, bi, B, T, mixed, attn) in CASES.items():
        torch.manual_seed(SEED)
        g = torch.Generator().manual_seed(SEED)
        enc = bm.VideoEncoder(204, H, rnn_type=rnn_type, bidirectional=bi, enable_ctc=True,
                              vocab_size=len(char2idx), char2idx=char2idx, device="cpu")
        dec = bm.CharDecodingStep(enc, char_dim=10, vocab_size=len(char2idx), char2idx=char2idx,
                                  attention_type=attn, device="cpu")
        frames, lens, chars, char_lens = make_batch(g, B, T, mixed, char2idx)
        out = {"frames": frames.numpy(), "frame_lens": lens.numpy(), "chars": chars.numpy(),
               "char_lens": char_lens.numpy()}
        for k, v in enc.state_dict().items():
            out["enc." + k] = v.numpy().copy()
        for k, v in dec.state_dict().items():
            out["dec." + k] = v.numpy().copy()
        enc.eval()
        lp, hidden, final = enc(frames, lens)
        out["log_probs"], out["hidden"] = lp.detach().numpy(), hidden.detach().numpy()
        if isinstance(final, tuple):
            out["final_h"], out["final_c"] = final[0].detach().numpy(), final[1].detach().numpy()
        else:
            out["final_h"] = final.detach().numpy()
        labels, label_lens = chars[:, 1:], char_lens - 1
        for red in ("mean", "sum"):
            enc.zero_grad()
            lp, _, _ = enc(frames, lens)
            loss = cl.ctc_loss(lp, labels, lens, label_lens, red, "cpu")
            loss.backward()
            out["ctc_" + red] = np.float64(loss.item())
            if red == "mean":
                for k, p in enc.named_parameters():
                    out["grad_ctc_mean." + k] = p.grad.numpy().copy()
        # one full reference train() step (decoder + CTC, clip 50, Adam 1e-3), teacher forcing 1
        torch.manual_seed(SEED + 1)
        opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-3)
        d_loss, c_loss = tb.train(enc, dec, [(frames, lens, chars, char_lens)], opt, torch.device("cpu"),
                                  char2idx, teacher_forcing_ratio=1, grad_norm=50)
        out["train_dec_loss"], out["train_ctc_loss"] = np.float64(d_loss), np.float64(c_loss)
        for k, v in enc.state_dict().items():
            out["enc_after." + k] = v.numpy().copy()
        for k, v in dec.state_dict().items():
            out["dec_after." + k] = v.numpy().copy()
        out["meta"] = np.array([rnn_type, str(H), str(int(bi)), attn])
        np.savez_compressed(os.path.join(GOLD, "seq_%s.npz" % name), **out)
        print(name, "ctc_mean=%.6f ctc_sum=%.6f dec=%.6f" % (out["ctc_mean"], out["ctc_sum"], d_loss))
    # the reference's own index fixtures
    src = os.path.join(ref_harness.REFERENCE_ROOT, "src/models/extern/prnet/Data/uv")
    import shutil
    shutil.copyfile(os.path.join(src, "uv_kpt_ind.txt"), os.path.join(GOLD, "uv_kpt_ind.txt"))
    np.save(os.path.join(GOLD, "face_ind.npy"), np.loadtxt(os.path.join(src, "face_ind.txt")).astype(np.int32))
    # the PRNet checkpoint's index (variable names / shapes / offsets; the data shard is not in the reference tree)
    shutil.copyfile(os.path.join(ref_harness.REFERENCE_ROOT,
                                 "src/models/extern/prnet/Data/net-data/256_256_resfcn256_weight.index"),
                    os.path.join(GOLD, "prnet_256_256_resfcn256_weight.index"))


if __name__ == "__main__":
    main()
