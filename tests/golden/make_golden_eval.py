"""Golden vectors for the CER protocol and for a teacher-forcing-ratio < 1 train step,
produced by the UNMODIFIED reference in the build container:

    python tests/golden/make_golden_eval.py

For every seq_<case>.npz (made by make_golden.py) the reference modules are rebuilt with the weights the case
holds after its train step (`enc_after.*`, `dec_after.*`), then
  * `torch.manual_seed(SEED + 2); train_better_model.eval(...)` -> decoder loss, `correct` (multinomial samples
    drawn from the CPU generator, one `(B,V)` draw per label position: train_better_model.py:130), `count`, plus
    the deterministic arg-max `correct`;
  * `torch.manual_seed(SEED + 3); train_better_model.train(..., teacher_forcing_ratio=0.5)` (Adam 1e-3, clip 50)
    -> both losses and every updated weight.  RNG consumption per label position: `torch.rand(1)` then one
    multinomial draw (train_better_model.py:56-63).
Writes tests/golden/eval_<case>.npz.
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.dirname(os.path.abspath(__file__))
SEED = 123456


def main():
    ref = ref_harness.load()
    dl, bm, tb = ref.data_loader, ref.better_model, ref.train_better_model
    char2idx = dict(dl._markers2Id)
    for ch in dl._labels:
        char2idx[ch] = len(char2idx)
    for path in sorted(glob.glob(os.path.join(GOLD, "seq_*.npz"))):
        name = os.path.basename(path)[4:-4]
        z = np.load(path)
        rnn_type, H, bi, attn = z["meta"]
        enc = bm.VideoEncoder(204, int(H), rnn_type=str(rnn_type), bidirectional=bool(int(bi)), enable_ctc=True,
                              vocab_size=len(char2idx), char2idx=char2idx, device="cpu")
        dec = bm.CharDecodingStep(enc, char_dim=10, vocab_size=len(char2idx), char2idx=char2idx,
                                  attention_type=str(attn), device="cpu")
        enc.load_state_dict({k[10:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc_after.")})
        dec.load_state_dict({k[10:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dec_after.")})
        batch = tuple(torch.from_numpy(z[k]) for k in ("frames", "frame_lens", "chars", "char_lens"))
        out = {}
        torch.manual_seed(SEED + 2)
        d_loss, correct, count = tb.eval(enc, dec, [batch], torch.device("cpu"), char2idx)
        out["eval_dec_loss"], out["eval_correct"], out["eval_count"] = np.float64(d_loss), np.int64(correct), np.int64(count)
        # deterministic arg-max protocol on the same teacher-forced distributions
        with torch.no_grad():
            frames, lens, chars, char_lens = batch
            _, enc_h, state = enc(frames, lens)
            labels = chars[:, 1:]
            hit = 0
            for i in range(int((char_lens - 1).max())):
                lp, state = dec(chars[:, i], state, lens, enc_h)
                hit += int(((lp.argmax(-1) == labels[:, i]) & (labels[:, i] != char2idx["<PAD>"])).sum())
        out["eval_correct_argmax"] = np.int64(hit)
        # the sampled characters themselves (eval() only returns their hit count): same seed, same draws
        torch.manual_seed(SEED + 2)
        with torch.no_grad():
            _, enc_h, state = enc(frames, lens)
            samples, hit_s = [], 0
            for i in range(int((char_lens - 1).max())):
                lp, state = dec(chars[:, i], state, lens, enc_h)
                s_i = lp.exp().multinomial(1).squeeze(-1)
                samples.append(s_i)
                hit_s += int(((s_i == labels[:, i]) & (labels[:, i] != char2idx["<PAD>"])).sum())
        assert hit_s == int(correct), "restated sampling loop does not consume the RNG like eval()"
        out["eval_samples"] = torch.stack(samples, 1).numpy()
        torch.manual_seed(SEED + 3)
        opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-3)
        d2, c2 = tb.train(enc, dec, [batch], opt, torch.device("cpu"), char2idx, teacher_forcing_ratio=0.5, grad_norm=50)
        out["tfr_dec_loss"], out["tfr_ctc_loss"] = np.float64(d2), np.float64(c2)
        for k, v in enc.state_dict().items():
            out["enc_tfr." + k] = v.numpy().copy()
        for k, v in dec.state_dict().items():
            out["dec_tfr." + k] = v.numpy().copy()
        np.savez_compressed(os.path.join(GOLD, "eval_%s.npz" % name), **out)
        print(name, "eval: loss %.6f correct %d (argmax %d) of %d; tfr=0.5 step: dec %.6f ctc %.6f"
              % (out["eval_dec_loss"], out["eval_correct"], hit, out["eval_count"], d2, c2))


if __name__ == "__main__":
    main()
