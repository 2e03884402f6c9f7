"""CPU restatement of the reference's sequence half (torch CPU fp32 / numpy).  TEST INFRASTRUCTURE.

Each function cites the reference lines it follows.  The RNN cell and CTC arithmetic live in torch
(an un-vendored dependency of the reference: torch==0.4.1 in requirements.macos.txt:80); they are
restated twice here: once through the same torch entry points the reference calls (nn.GRU/LSTM/RNN
on a PackedSequence, F.ctc_loss) and once from the published equations with explicit loops
(`rnn_masked`, `ctc_alpha_beta`), so the masking formulation the CUDA kernels use is itself checked
against the packed formulation the reference uses.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ---- vocabulary: src/data/data_loader.py:29-35,100-115 ------------------------------------------
BOS, EOS, PAD, UNK = "<BOS>", "<EOS>", "<PAD>", "<UNK>"
MARKERS = {PAD: 0, BOS: 1, EOS: 2, UNK: 3}
FALLBACK_LABELS = list(" !\"#$%&'()*+,-./0123456789:;<>?@[]abcdefghijklmnopqrstuvwxyz")


def build_char2idx(labels=None):
    """data_loader.build_vocab (:100-115) with the hard-coded fallback label list (:35)."""
    char2idx = dict(MARKERS)
    for ch in (labels if labels is not None else FALLBACK_LABELS):
        char2idx[ch] = len(char2idx)
    return char2idx


def parse_caption(cap, char2idx):
    """FrameCaptionDataset.parse_caption (data_loader.py:283-290)."""
    return np.array([MARKERS[BOS]] + [char2idx.get(c, MARKERS[UNK]) for c in cap] + [MARKERS[EOS]])


def collate(batch):
    """_collate_fn (data_loader.py:117-152): zero-pad frames to (B,Tmax,68,3) f32, captions to i64."""
    frames, caps = zip(*batch)

    def pad(seqs, dtype):
        lens = torch.tensor([len(s) for s in seqs], dtype=torch.long)
        out = torch.zeros((len(seqs), int(lens.max())) + tuple(np.asarray(seqs[0]).shape[1:]), dtype=dtype)
        for i, s in enumerate(seqs):
            out[i, : len(s)] = torch.as_tensor(np.asarray(s)).to(dtype)
        return out, lens

    f, fl = pad(frames, torch.float32)
    c, cl = pad(caps, torch.long)
    return f, fl, c, cl


# ---- allennlp masked_log_softmax (better_model.py:93) --------------------------------------------
def log_mask_vector(vocab_size, char2idx):
    """VideoEncoder.output_mask (better_model.py:43-45) -> additive term log(mask + 1e-45), fp32."""
    mask = torch.ones(vocab_size + 1)
    mask[char2idx[PAD] + 1] = 0
    mask[char2idx[BOS] + 1] = 0
    return (mask + 1e-45).log()


def masked_log_softmax(logits, log_mask):
    return F.log_softmax(logits + log_mask, dim=-1)


# ---- recurrent layer ------------------------------------------------------------------------------
_GATES = {"RNN": 1, "GRU": 3, "LSTM": 4}


def rnn_packed(x, lens, weights, rnn_type, bidirectional):
    """The reference's own route (better_model.py:64-89): sort desc, pack, nn.<RNN>, unpack, unsort.
    weights: dict with torch-default names weight_ih_l0[_reverse], weight_hh_l0..., bias_*."""
    H = weights["weight_hh_l0"].shape[1]
    I = weights["weight_ih_l0"].shape[1]
    with torch.random.fork_rng(devices=[]):        # the throw-away module's initialiser must not advance the generator
        rnn = getattr(torch.nn, rnn_type)(I, H, num_layers=1, bidirectional=bidirectional, batch_first=True)
    sorted_lens, perm = lens.sort(0, descending=True)
    _, inv = perm.sort(0)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x.index_select(0, perm), sorted_lens.cpu(), batch_first=True)
    out, final = torch.func.functional_call(rnn, {k: torch.as_tensor(v) for k, v in weights.items()}, (packed,))
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    out = out.index_select(0, inv)
    if isinstance(final, tuple):
        final = tuple(s.index_select(1, inv) for s in final)
    else:
        final = final.index_select(1, inv)
    return out, final


def rnn_masked(x, lens, weights, rnn_type, bidirectional, dtype=torch.float32):
    """Same layer from the cell equations with length masking instead of packing (what the CUDA
    kernels implement).  PyTorch gate order: GRU r,z,n ; LSTM i,f,g,o."""
    x = x.to(dtype)
    B, T, _ = x.shape
    H = weights["weight_hh_l0"].shape[1]
    dirs = ["", "_reverse"] if bidirectional else [""]
    outs, h_fin, c_fin = [], [], []
    for sfx in dirs:
        w_ih = torch.as_tensor(weights["weight_ih_l0" + sfx]).to(dtype)
        w_hh = torch.as_tensor(weights["weight_hh_l0" + sfx]).to(dtype)
        b_ih = torch.as_tensor(weights["bias_ih_l0" + sfx]).to(dtype)
        b_hh = torch.as_tensor(weights["bias_hh_l0" + sfx]).to(dtype)
        h = torch.zeros(B, H, dtype=dtype)
        c = torch.zeros(B, H, dtype=dtype)
        out = torch.zeros(B, T, H, dtype=dtype)
        order = range(T) if sfx == "" else range(T - 1, -1, -1)
        for t in order:
            act = (t < lens).to(dtype).unsqueeze(1)
            gi = x[:, t] @ w_ih.t() + b_ih
            gh = h @ w_hh.t() + b_hh
            if rnn_type == "GRU":
                r = torch.sigmoid(gi[:, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                hn = (1 - z) * n + z * h
            elif rnn_type == "LSTM":
                g = gi + gh
                i_, f_, g_, o_ = g[:, :H].sigmoid(), g[:, H:2 * H].sigmoid(), g[:, 2 * H:3 * H].tanh(), g[:, 3 * H:].sigmoid()
                cn = f_ * c + i_ * g_
                hn = o_ * cn.tanh()
                c = act * cn + (1 - act) * c
            else:
                hn = torch.tanh(gi + gh)
            h = act * hn + (1 - act) * h
            out[:, t] = act * hn
        outs.append(out)
        h_fin.append(h)
        c_fin.append(c)
    hidden = torch.cat(outs, dim=2)
    h_n = torch.stack(h_fin, 0)
    if rnn_type == "LSTM":
        return hidden, (h_n, torch.stack(c_fin, 0))
    return hidden, h_n


def cat_directions(final):
    """VideoEncoder._cat_directions (better_model.py:98-112): (L*2,B,H) -> (L,B,2H)."""
    def cat(s):
        return torch.cat([s[0::2], s[1::2]], dim=2)
    return tuple(cat(s) for s in final) if isinstance(final, tuple) else cat(final)


def encoder_forward(state, frames, frame_lens, rnn_type, bidirectional, char2idx, enable_ctc=True):
    """VideoEncoder.forward (better_model.py:53-96).  state: the module's state_dict (torch-default
    names `rnn.*`, `output_proj.*`).  Returns (log_probs, hidden, final) or (hidden, final)."""
    x = frames.reshape(frames.shape[0], frames.shape[1], -1)
    weights = {k[len("rnn."):]: v for k, v in state.items() if k.startswith("rnn.")}
    hidden, final = rnn_packed(x, frame_lens, weights, rnn_type, bidirectional)
    if bidirectional:
        final = cat_directions(final)
    if not enable_ctc:
        return hidden, final
    logits = hidden @ torch.as_tensor(state["output_proj.weight"]).t() + torch.as_tensor(state["output_proj.bias"])
    lp = masked_log_softmax(logits, log_mask_vector(len(char2idx), char2idx))
    return lp, hidden, final


# ---- CTC -------------------------------------------------------------------------------------------
def ctc_alpha_beta(lp, target, want_grad=True):
    """Log-space CTC for ONE sample from the published recursion (Graves 2006), float64 numpy.
    lp (T,C) log-probs, target (L,) classes (blank=0).  Returns nll and d nll / d lp with torch's
    native-ctc convention (exp(lp) - occupancy/prob), cf. aten/src/ATen/native/LossCTC.cpp."""
    lp = np.asarray(lp, dtype=np.float64)
    T, C = lp.shape
    L = len(target)
    S = 2 * L + 1
    ext = np.zeros(S, dtype=np.int64)
    ext[1::2] = target
    ninf = -np.inf

    def lse(*a):
        m = max(a)
        if m == ninf:
            return ninf
        return m + math.log(sum(math.exp(v - m) for v in a))

    alpha = np.full((T, S), ninf)
    beta = np.full((T, S), ninf)
    alpha[0, 0] = lp[0, 0]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            a = [alpha[t - 1, s]]
            if s >= 1:
                a.append(alpha[t - 1, s - 1])
            if s >= 2 and (s & 1) and ext[s] != ext[s - 2]:
                a.append(alpha[t - 1, s - 2])
            alpha[t, s] = lse(*a) + lp[t, ext[s]]
    ll = lse(alpha[T - 1, S - 1], alpha[T - 1, S - 2]) if S > 1 else alpha[T - 1, 0]
    nll = -ll
    if not want_grad:
        return nll, None
    beta[T - 1, S - 1] = lp[T - 1, 0]
    if S > 1:
        beta[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            a = [beta[t + 1, s]]
            if s + 1 < S:
                a.append(beta[t + 1, s + 1])
            if s + 2 < S and (s & 1) and ext[s + 2] != ext[s]:
                a.append(beta[t + 1, s + 2])
            beta[t, s] = lse(*a) + lp[t, ext[s]]
    grad = np.exp(lp)
    for t in range(T):
        acc = {}
        for s in range(S):
            v = alpha[t, s] + beta[t, s]
            c = int(ext[s])
            acc[c] = lse(acc[c], v) if c in acc else v
        for c, v in acc.items():
            if v != ninf:
                grad[t, c] -= math.exp(v + nll - lp[t, c])
    return nll, grad


def ctc_linear_rescaled(lp, target, dtype=np.float64):
    """The SAME quantity as ctc_alpha_beta, through the recursion ctc_linear_warp_kernel runs
    (lipreading_b200/csrc/ctc.cu): probabilities with Rabiner-style rescaling instead of log-sum-exp.
        alpha^_t = (M alpha^_{t-1}) . p_t . k_t ,   k_t = 1 on odd frames, 1 / (lattice mass at frame t-2) on even ones
        beta~_t  =  M' beta^_{t+1} ,  beta^_t = beta~_t . p_t . k_t
        rho = alpha^_T(S-1) + alpha^_T(S-2) = sum_s alpha^_t(s) beta~_t(s) for every t
        nll = sum_t log k_t - log rho ;  grad[t,c] = p_t(c) - sum_{s: l'_s = c} alpha^_t(s) beta~_t(s) / rho
    Returns (nll, grad, spread) where spread = max_t |sum_s alpha^_t beta~_t / rho - 1| (the invariant the kernel relies
    on when it divides by the constant rho).  `dtype=np.float32` shows the rounding profile of the kernel's arithmetic."""
    lp = np.asarray(lp, dtype=np.float64)
    T, C = lp.shape
    L = len(target)
    S = 2 * L + 1
    ext = np.zeros(S, dtype=np.int64)
    ext[1::2] = target
    f = dtype
    p = np.exp(lp).astype(f)
    skip = np.zeros(S, dtype=bool)                       # s-2 -> s allowed
    for s in range(3, S, 2):
        skip[s] = ext[s] != ext[s - 2]

    def fwd(a):                                          # M a
        out = a.copy()
        out[1:] += a[:-1]
        out[2:] += np.where(skip[2:], a[:-2], f(0))
        return out

    def bwd(b):                                          # M' b
        out = b.copy()
        out[:-1] += b[1:]
        out[:-2] += np.where(skip[2:], b[2:], f(0))
        return out

    alpha = np.zeros((T, S), dtype=f)
    kf = np.ones(T, dtype=f)
    a = np.zeros(S, dtype=f)
    a[0] = 1                                             # "alpha_-1": frame 0 is then the generic step
    first = np.zeros(S, dtype=f)
    first[0] = p[0, 0]
    if S > 1:
        first[1] = p[0, ext[1]]
    mass = {}
    for t in range(T):
        new = first if t == 0 else fwd(a) * p[t, ext]
        if t >= 2 and t % 2 == 0:
            kf[t] = f(1) / mass[t - 2]
        a = (new * kf[t]).astype(f)
        alpha[t] = a
        if t % 2 == 0:
            mass[t] = a.sum(dtype=f)
    rho = alpha[T - 1, S - 1] + (alpha[T - 1, S - 2] if S > 1 else f(0))
    nll = float(np.sum(np.log(kf.astype(np.float64))) - math.log(float(rho)))
    grad = p.astype(np.float64).copy()
    b = np.zeros(S, dtype=f)                             # "beta_T"
    spread = 0.0
    for t in range(T - 1, -1, -1):
        if t == T - 1:
            bt = np.zeros(S, dtype=f)
            bt[S - 1] = 1
            if S > 1:
                bt[S - 2] = 1
        else:
            bt = bwd(b)
        e = alpha[t] * bt
        spread = max(spread, abs(float(e.sum(dtype=np.float64)) / float(rho) - 1.0))
        for s in range(S):
            grad[t, ext[s]] -= float(e[s]) / float(rho)
        b = (bt * p[t, ext] * kf[t]).astype(f)
    return nll, grad, spread


def ctc_nll_torch(log_probs_btc, targets_padded, in_lens, tgt_lens):
    """Per-sample nll through the same torch entry point the reference calls (ctc_loss.py:85)."""
    concat = torch.cat([targets_padded[i, : int(tgt_lens[i])] for i in range(len(tgt_lens))])
    return F.ctc_loss(log_probs_btc.transpose(0, 1), concat.int(), in_lens.int(), tgt_lens.int(),
                      blank=0, reduction="none")


def ctc_loss_wrapper(log_probs, labels, frame_lens, label_lens, reduction, per_sample_nll=ctc_nll_torch):
    """src/train/ctc_loss.py:28-114 restated, including:
      * the non-decreasing frame_lens assertion (:39);
      * dropping samples with label_len > 256 (:46-56), None when nothing is left;
      * splitting the batch into runs of equal frame length (:64-77);
      * labels + 1 so that class 0 is the blank (:80);
      * inf handling per run (:87-101): drop the infeasible samples of the run, recompute;
      * the run-weighting quirk for 'mean' (:74,103-106): `minibatch_size` is read BEFORE the batch is
        re-sliced, so run k is weighted by the size of run k-1 (run 0 by the size of the kept batch);
        F.ctc_loss 'mean' itself is mean_b(nll_b / max(label_len_b, 1)).
    Returns a scalar tensor (differentiable w.r.t. log_probs) or None."""
    frame_lens = frame_lens.to(torch.int64)
    label_lens = label_lens.to(torch.int64)
    assert bool((frame_lens[1:] - frame_lens[:-1] >= 0).all())
    keep = (label_lens <= 256).nonzero().squeeze(-1)
    if keep.numel() == 0:
        return None
    if keep.numel() < len(label_lens):
        log_probs, labels = log_probs.index_select(0, keep), labels.index_select(0, keep)
        frame_lens, label_lens = frame_lens.index_select(0, keep), label_lens.index_select(0, keep)
    n = len(frame_lens)
    cuts = ((frame_lens[1:] - frame_lens[:-1]).nonzero().squeeze(-1) + 1).tolist() + [n]
    total, count, prev, prev_slice_len = 0, 0, 0, n
    for cut in cuts:
        weight = prev_slice_len                 # the quirk: len(frame_lens) of the *previous* slice
        sl = slice(prev, cut)
        nll = per_sample_nll(log_probs[sl], labels[sl] + 1, frame_lens[sl], label_lens[sl])
        ll = label_lens[sl]
        prev_slice_len = cut - prev
        if bool(torch.isinf(nll).any()):
            ok = (~torch.isinf(nll)).nonzero().squeeze(-1)
            if ok.numel() == 0:
                continue                        # (:91-93) NB: prev_change_point is NOT advanced
            nll, ll = nll.index_select(0, ok), ll.index_select(0, ok)
            weight = prev_slice_len = int(ok.numel())
        if reduction == "mean":
            loss = (nll / ll.clamp(min=1).to(nll.dtype)).mean()
            total = total + loss * weight
            count += weight
        else:
            total = total + nll.sum()
        prev = cut
    if isinstance(total, int) or float(total.detach()) == 0:
        return None
    return total / count if reduction == "mean" else total


# ---- attention decoder step (better_model.py:124-235), functional ---------------------------------
def masked_softmax(v, mask):
    mask = mask.float()
    r = F.softmax(v * mask, dim=-1) * mask
    return r / (r.sum(dim=-1, keepdim=True) + 1e-13)


class OracleDecoder(torch.nn.Module):
    """CharDecodingStep restated (same parameter names, so reference checkpoints load)."""

    def __init__(self, enc_hidden, rnn_type, char_dim, vocab_size, char2idx, attention_type="none", attn_hidden_size=-1):
        super().__init__()
        H = enc_hidden
        self.H, self.rnn_type, self.attention_type, self.vocab_size = H, rnn_type, attention_type, vocab_size
        self.embedding = torch.nn.Embedding(vocab_size, char_dim, padding_idx=char2idx[PAD])
        self.rnn = getattr(torch.nn, rnn_type)(char_dim, H, num_layers=1, batch_first=True)
        if attention_type == "1_layer_nn":
            self.attn_proj_1_layer_nn = torch.nn.Linear(2 * H, 1)
        elif attention_type == "general":
            self.attn_proj_general = torch.nn.Linear(H, H)
        elif attention_type == "concat":
            self.attn_proj_layer1 = torch.nn.Linear(2 * H, attn_hidden_size)
            self.attn_proj_layer2 = torch.nn.Linear(attn_hidden_size, 1)
        self.concat_layer = torch.nn.Linear(2 * H, H)
        self.output_proj = torch.nn.Linear(H, vocab_size)
        mask = torch.ones(vocab_size)
        mask[char2idx[PAD]] = 0
        mask[char2idx[BOS]] = 0
        self.register_buffer("log_mask", (mask + 1e-45).log(), persistent=False)

    def forward(self, input_, prev_state, enc_lens, enc_h):
        B, Te, _ = enc_h.shape
        h, state = self.rnn(self.embedding(input_).unsqueeze(1), prev_state)
        q = h.squeeze(1)
        if self.attention_type != "none":
            if self.attention_type == "dot":
                logits = (enc_h * q.unsqueeze(1)).sum(-1)
            elif self.attention_type == "general":
                logits = (enc_h * self.attn_proj_general(q).unsqueeze(1)).sum(-1)
            elif self.attention_type == "1_layer_nn":
                logits = self.attn_proj_1_layer_nn(torch.cat([enc_h, q.unsqueeze(1).expand_as(enc_h)], 2)).squeeze(-1)
            else:
                logits = self.attn_proj_layer2(self.attn_proj_layer1(torch.cat([enc_h, q.unsqueeze(1).expand_as(enc_h)], 2)).tanh()).squeeze(-1)
            emask = torch.arange(Te).expand(B, Te) < enc_lens.unsqueeze(1)
            ctx = masked_softmax(logits, emask).unsqueeze(1).bmm(enc_h).squeeze(1)
            q = self.concat_layer(torch.cat([ctx, q], 1)).tanh()
        return F.log_softmax(self.output_proj(q) + self.log_mask, dim=-1), state


def greedy_ctc_decode(log_probs, lens):
    """Greedy CTC decode (argmax, collapse repeats, drop blank) — semantics of the reference's
    GreedyDecoder (src/models/lipreader/decoder.py:165-197) on (B,T,C) log-probs."""
    out = []
    arg = log_probs.argmax(-1)
    for b in range(arg.shape[0]):
        prev, seq = -1, []
        for t in range(int(lens[b])):
            c = int(arg[b, t])
            if c != prev and c != 0:
                seq.append(c)
            prev = c
        out.append(seq)
    return out
