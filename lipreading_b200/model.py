"""VideoEncoder / CharDecodingStep with the reference's constructor, forward signature and
state_dict keys (src/models/lipreader/better_model.py:13-245), running on the sm_100a kernels.

Checkpoint compatibility: parameters are registered under torch's default names
(`rnn.weight_ih_l0[_reverse]`, ..., `output_proj.weight/bias`), drawn from the RNG in the same order
as `nn.{LSTM,GRU,RNN}` + `nn.Linear`, so a reference `best_encoder.pth` loads unchanged and the same
seed gives the same initial weights.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as LF
from .vocab import BOS, PAD

_ALLOWED_RNN_TYPES = {"LSTM", "GRU", "RNN"}
_ALLOWED_FRAME_PROCESSING = {"flatten", "conv3d"}
_ALLOWED_ATTENTION_TYPES = {"none", "dot", "general", "1_layer_nn", "concat"}


class NativeRNN(nn.Module):
    """Parameter container + forward for a stack of (bi)directional recurrent layers.  Same
    parameter names / shapes / init order as torch.nn.{RNN,GRU,LSTM}(batch_first=True)."""

    def __init__(self, rnn_type, input_size, hidden_size, num_layers=1, bidirectional=False, dropout=0.0):
        super().__init__()
        self.mode, self.input_size, self.hidden_size = rnn_type, input_size, hidden_size
        self.num_layers, self.bidirectional, self.dropout = num_layers, bidirectional, dropout
        template = getattr(nn, rnn_type)(input_size, hidden_size, num_layers=num_layers,
                                         bidirectional=bidirectional, batch_first=True, dropout=dropout)
        self._names = list(template._flat_weights_names)
        for name in self._names:
            self.register_parameter(name, nn.Parameter(getattr(template, name).detach().clone()))

    def _layer_weights(self, layer):
        out = []
        for sfx in (["", "_reverse"] if self.bidirectional else [""]):
            out += [getattr(self, "%s_l%d%s" % (n, layer, sfx)) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return out

    def forward(self, x, lens):
        """x (B,T,I) zero padded, lens (B).  Returns (hidden (B,T,D*H), final) with final laid out
        like torch: (L*D,B,H) (a tuple (h,c) for LSTM)."""
        h_all, c_all = [], []
        inp = x
        for layer in range(self.num_layers):
            res = LF.rnn_layer(inp, lens, self.mode, self._layer_weights(layer))
            inp = res[0]
            h_all.append(res[1])
            if self.mode == "LSTM":
                c_all.append(res[2])
            if self.dropout > 0 and self.training and layer + 1 < self.num_layers:
                inp = F.dropout(inp, self.dropout, True)
        h_n = torch.cat(h_all, 0)
        return inp, ((h_n, torch.cat(c_all, 0)) if self.mode == "LSTM" else h_n)


_FUSED_PROJ_MAX_CLASSES = 68          # kMaxC of csrc/proj_logsoftmax.cu (4 class groups x 17)


def _proj_log_softmax(hidden, proj, log_mask):
    """Linear + masked log-softmax (better_model.py:91-94, 231-233).  Up to 68 classes (the reference's 64-entry
    vocabulary + blank and then some) this is the fused kernel; a larger labels.json takes the wide-vocabulary
    route: a plain library GEMM followed by torch's row log-softmax, on the same device (the reference accepts any
    vocabulary, so must this)."""
    LF.N.require_cuda(hidden)
    if proj.out_features <= _FUSED_PROJ_MAX_CLASSES:
        return LF.proj_masked_log_softmax(hidden, proj.weight, proj.bias, log_mask)
    return F.log_softmax(F.linear(hidden, proj.weight, proj.bias) + log_mask, dim=-1)


class VideoEncoder(nn.Module):
    """Drop-in for better_model.VideoEncoder (:13-122)."""

    def __init__(self, frame_dim, hidden_size, frame_processing="flatten",
                 rnn_type="LSTM", num_layers=1, bidirectional=True, rnn_dropout=0,
                 enable_ctc=False, vocab_size=-1, char2idx=None, device="cpu"):
        super().__init__()
        assert frame_processing in _ALLOWED_FRAME_PROCESSING
        assert rnn_type in _ALLOWED_RNN_TYPES
        if enable_ctc:
            assert vocab_size > 0 and char2idx is not None
        self.frame_dim, self.hidden_size, self.frame_processing = frame_dim, hidden_size, frame_processing
        self.rnn_type, self.num_layers, self.bidirectional = rnn_type, num_layers, bidirectional
        self.rnn_dropout, self.enable_ctc, self.best_error = rnn_dropout, enable_ctc, 1
        self.num_dirs = 2 if bidirectional else 1
        if frame_processing == "conv3d":
            from .conv_frontend import ConvFrontEnd          # north-star extension (row N1)
            self.front = ConvFrontEnd()
            assert frame_dim == self.front.out_features, \
                "frame_processing='conv3d' needs frame_dim=%d" % self.front.out_features
        self.rnn = NativeRNN(rnn_type, frame_dim, hidden_size, num_layers=num_layers,
                             bidirectional=bidirectional, dropout=rnn_dropout)
        if enable_ctc:
            self.vocab_size, self.adj_vocab_size, self.char2idx = vocab_size, vocab_size + 1, char2idx
            mask = torch.ones(self.adj_vocab_size, device=device)
            mask[char2idx[PAD] + 1] = 0
            mask[char2idx[BOS] + 1] = 0
            self.output_mask = mask                      # plain attribute, like the reference (:43-45)
            self.output_proj = nn.Linear(self.num_dirs * hidden_size, self.adj_vocab_size)

    def _log_mask(self, device):
        lm = getattr(self, "_log_mask_cache", None)
        if lm is None or lm.device != device:
            lm = (self.output_mask.to(device=device, dtype=torch.float32) + 1e-45).log()
            self._log_mask_cache = lm
        return lm

    def forward(self, frames, frame_lens):
        """frames (B,T,68,3) f32 [or (B,T,H,W,3) u8 clips for 'conv3d'], frame_lens (B,) ->
        (log_probs (B,T,V+1), hidden (B,T,D*H), final_state) or (hidden, final_state)."""
        if self.frame_processing == "flatten":
            x = frames.reshape(frames.shape[0], frames.shape[1], -1)
        else:
            x = self.front(frames)
        hidden, final = self.rnn(x, frame_lens)
        # pad_packed_sequence trims the time axis to the longest clip of the batch (:78)
        t_max = getattr(self, "_t_max_hint", None)
        if t_max is None:
            t_max = x.shape[1] if frame_lens.numel() == 0 else int(frame_lens.max())
        if t_max < hidden.shape[1]:
            hidden = hidden[:, :t_max].contiguous()
        if self.bidirectional:
            final = self._cat_directions(final)
        if not self.enable_ctc:
            return hidden, final
        log_probs = _proj_log_softmax(hidden, self.output_proj, self._log_mask(hidden.device))
        return log_probs, hidden, final

    def _cat_directions(self, final_state):
        def cat(s):
            return torch.cat([s[0::2], s[1::2]], dim=2)
        return tuple(cat(s) for s in final_state) if isinstance(final_state, tuple) else cat(final_state)

    def save_best_model(self, error, file_path):
        if error < self.best_error:
            self.best_error = error
            os.makedirs(os.path.dirname(file_path) or ".", exist_ok=True)
            torch.save(self.state_dict(), file_path)
            print("\tSaving best error '{}' to '{}'".format(self.best_error, file_path))


class CharDecodingStep(nn.Module):
    """Attention decoder step, same parameters / semantics as better_model.CharDecodingStep
    (:124-245).  SURVEY §8f row f1 ("next"): runs on stock PyTorch ops for now."""

    def __init__(self, encoder, char_dim, vocab_size, char2idx, rnn_dropout=0, attention_type="none",
                 attn_hidden_size=-1, device="cpu"):
        super().__init__()
        assert attention_type in _ALLOWED_ATTENTION_TYPES
        if attention_type == "concat":
            assert attn_hidden_size > 0
        self.hidden_size = encoder.hidden_size * (2 if encoder.bidirectional else 1)
        self.rnn_type, self.num_layers, self.rnn_dropout = encoder.rnn_type, encoder.num_layers, rnn_dropout
        self.char_dim, self.vocab_size, self.char2idx, self.attention_type = char_dim, vocab_size, char2idx, attention_type
        mask = torch.ones(vocab_size, device=device)
        mask[char2idx[PAD]] = 0
        mask[char2idx[BOS]] = 0
        self.output_mask = mask
        H = self.hidden_size
        self.embedding = nn.Embedding(vocab_size, char_dim, padding_idx=char2idx[PAD])
        self.rnn = getattr(nn, self.rnn_type)(char_dim, H, num_layers=self.num_layers, batch_first=True,
                                              dropout=rnn_dropout)
        if attention_type == "1_layer_nn":
            self.attn_proj_1_layer_nn = nn.Linear(2 * H, 1)
        elif attention_type == "general":
            self.attn_proj_general = nn.Linear(H, H)
        elif attention_type == "concat":
            self.attn_proj_layer1 = nn.Linear(2 * H, attn_hidden_size)
            self.attn_proj_layer2 = nn.Linear(attn_hidden_size, 1)
        self.concat_layer = nn.Linear(2 * H, H)
        self.output_proj = nn.Linear(H, vocab_size)
        self.best_error = 1

    def forward(self, input_, previous_state, encoder_lens, encoder_hidden_states):
        B, Te = encoder_hidden_states.shape[0], encoder_hidden_states.shape[1]
        h, final_state = self.rnn(self.embedding(input_).unsqueeze(1), previous_state)
        q = h.squeeze(1)
        enc = encoder_hidden_states
        if self.attention_type != "none":
            if self.attention_type == "dot":
                scores = torch.einsum("bth,bh->bt", enc, q)
            elif self.attention_type == "general":
                scores = torch.einsum("bth,bh->bt", enc, self.attn_proj_general(q))
            else:
                both = torch.cat([enc, q.unsqueeze(1).expand(-1, Te, -1)], dim=2)
                if self.attention_type == "1_layer_nn":
                    scores = self.attn_proj_1_layer_nn(both).squeeze(-1)
                else:
                    scores = self.attn_proj_layer2(self.attn_proj_layer1(both).tanh()).squeeze(-1)
            valid = (torch.arange(Te, device=input_.device).unsqueeze(0) < encoder_lens.unsqueeze(1)).float()
            w = F.softmax(scores * valid, dim=-1) * valid            # allennlp masked_softmax
            w = w / (w.sum(dim=-1, keepdim=True) + 1e-13)
            context = torch.bmm(w.unsqueeze(1), enc).squeeze(1)
            q = self.concat_layer(torch.cat([context, q], dim=1)).tanh()
        logits = self.output_proj(q)
        log_mask = (self.output_mask.to(logits.device) + 1e-45).log()
        return F.log_softmax(logits + log_mask, dim=-1), final_state

    def forward_sequence(self, inputs, previous_state, encoder_lens, encoder_hidden_states):
        """All decode steps of a teacher-forced pass at once: inputs (B,L) are the characters fed at steps 0..L-1.
        Returns (log_probs (B,L,V), final_state) — the same numbers as L calls of `forward` chained through
        `final_state`, because the recurrent state never depends on the attention output (better_model.py:184,
        223-229): the L-step recurrence is one RNN call, the attention of all positions is one launch of the fused
        kernel (`LF.attn_context`: the clip's encoder states are staged in shared memory once instead of being
        re-read from HBM twice per position), the output projection + masked log-softmax is the encoder's kernel."""
        B, L = inputs.shape
        h_seq, final_state = self.rnn(self.embedding(inputs), previous_state)          # (B,L,H)
        q = h_seq
        enc = encoder_hidden_states
        Te = enc.shape[1]
        if self.attention_type != "none":
            H = self.hidden_size
            if self.attention_type in ("dot", "general"):
                qq = q if self.attention_type == "dot" else self.attn_proj_general(q)
                context, _ = LF.attn_context(qq, enc, encoder_lens)
            else:
                # score functions over [enc ; q] (better_model.py:204-221): the Linear over the concatenation splits
                # into an encoder part — evaluated ONCE per clip, not once per label position — and a query part
                if self.attention_type == "1_layer_nn":
                    w = self.attn_proj_1_layer_nn.weight                       # (1, 2H)
                    s_enc = enc @ w[0, :H]                                       # (B,Te)
                    s_q = q @ w[0, H:] + self.attn_proj_1_layer_nn.bias          # (B,L)
                    scores = s_enc.unsqueeze(1) + s_q.unsqueeze(2)
                else:
                    w1 = self.attn_proj_layer1.weight                            # (A, 2H)
                    p_enc = enc @ w1[:, :H].t()                                  # (B,Te,A)
                    p_q = q @ w1[:, H:].t() + self.attn_proj_layer1.bias         # (B,L,A)
                    hid = (p_enc.unsqueeze(1) + p_q.unsqueeze(2)).tanh()         # (B,L,Te,A)
                    scores = (hid @ self.attn_proj_layer2.weight[0]) + self.attn_proj_layer2.bias
                context, _ = LF.attn_context_scores(scores, enc, encoder_lens)
            q = self.concat_layer(torch.cat([context, q], dim=2)).tanh()
        log_mask = (self.output_mask.to(q.device) + 1e-45).log()
        return _proj_log_softmax(q, self.output_proj, log_mask), final_state

    def save_best_model(self, error, file_path):
        if error < self.best_error:
            self.best_error = error
            os.makedirs(os.path.dirname(file_path) or ".", exist_ok=True)
            torch.save(self.state_dict(), file_path)
            print("\tSaving best error '{}' to '{}'".format(self.best_error, file_path))
