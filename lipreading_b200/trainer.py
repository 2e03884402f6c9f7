"""One epoch of training / evaluation with the reference's step structure
(src/train/train_better_model.py:7-143): encoder -> CTC (aux) -> teacher-forced attention decoder
loop -> decoder backward (retain graph) -> CTC backward -> separate grad clips -> optimizer step.

Differences are confined to where the arithmetic runs (sm_100a kernels for encoder / projection /
CTC) and to host<->device traffic: the log-probs never leave the GPU, lengths are kept on the host
from the loader, and the per-batch loss scalars are accumulated on the device and read back once
per epoch.  Under data parallelism (`dist`), gradients are all-reduced (one flat bucket) between
backward and clipping so every rank clips and steps on identical gradients.
"""
import torch
import torch.nn.functional as F

from .ctc import ctc_loss
from .vocab import BOS, EOS, PAD


# True: a fully teacher-forced decode runs as one vectorised pass (CharDecodingStep.forward_sequence) instead of a
# Python loop over label positions; False keeps the reference's step loop (used by the parity tests as the oracle path)
SEQUENCE_DECODE = True


def _check_batch(chars, char_lens, frame_lens, char2idx, use_ctc):
    assert bool((chars[:, 0] == char2idx[BOS]).all())
    assert bool((chars.gather(1, (char_lens - 1).unsqueeze(1)).squeeze(1) == char2idx[EOS]).all())
    if use_ctc:
        assert bool((frame_lens >= char_lens).all())             # otherwise ctc loss will produce inf


def train(encoder, decoding_step, data_loader, opt, device, char2idx,
          teacher_forcing_ratio=1, grad_norm=None, dist=None):
    """Same contract as the reference `train`: returns (avg_decoder_loss, avg_ctc_loss)."""
    use_ctc = encoder.enable_ctc
    dec_sum = torch.zeros((), device=device)
    ctc_sum = torch.zeros((), device=device)
    encoder.train()
    decoding_step.train()
    pad = char2idx[PAD]
    for frames, frame_lens, chars, char_lens in data_loader:
        fl_h, cl_h, chars_h = frame_lens.cpu(), char_lens.cpu(), chars.cpu()
        _check_batch(chars_h, cl_h, fl_h, char2idx, use_ctc)
        ll_h = cl_h - 1
        n_tokens = int(ll_h.sum())
        assert int((chars_h[:, 1:] != pad).sum()) == n_tokens
        frames = frames.to(device, non_blocking=True)
        chars = chars.to(device, non_blocking=True)
        frame_lens_d = fl_h.to(device, non_blocking=True)
        labels = chars[:, 1:]
        batch_size = frames.shape[0]
        max_label_len = int(ll_h.max())

        encoder._t_max_hint = int(fl_h.max())
        if use_ctc:
            enc_out, enc_h, prev_state = encoder(frames, frame_lens_d)
            cur_ctc = ctc_loss(enc_out, labels, frame_lens_d, ll_h, "mean", device, host_lens=(fl_h, ll_h))
            if cur_ctc is None:
                continue
        else:
            enc_h, prev_state = encoder(frames, frame_lens_d)
        prev_output = torch.full((batch_size,), char2idx[BOS], dtype=torch.long, device=device)

        if teacher_forcing_ratio >= 1 and SEQUENCE_DECODE and hasattr(decoding_step, "forward_sequence"):
            # always teacher forcing: no step depends on a sampled character, so the whole decode loop is one
            # vectorised pass (the reference's per-step torch.rand / multinomial draws have no effect on the loss)
            log_probs, prev_state = decoding_step.forward_sequence(chars[:, :max_label_len], prev_state, frame_lens_d, enc_h)
            decoder_loss = F.nll_loss(log_probs.reshape(-1, log_probs.shape[-1]), labels[:, :max_label_len].reshape(-1),
                                      ignore_index=pad, reduction="sum")
        else:
            decoder_loss = 0
            for i in range(max_label_len):
                teacher_forcing = bool(torch.rand(1) < teacher_forcing_ratio)
                input_ = chars[:, i] if teacher_forcing else prev_output
                log_probs, prev_state = decoding_step(input_, prev_state, frame_lens_d, enc_h)
                decoder_loss = decoder_loss + F.nll_loss(log_probs, labels[:, i], ignore_index=pad, reduction="sum")
                prev_output = log_probs.exp().multinomial(1).squeeze(-1)
        decoder_loss = decoder_loss / n_tokens

        opt.zero_grad()
        decoder_loss.backward(retain_graph=use_ctc)
        dec_sum += decoder_loss.detach()
        if use_ctc:
            cur_ctc.backward()
            ctc_sum += cur_ctc.detach()
        if dist is not None:
            dist.allreduce_grads(list(encoder.parameters()) + list(decoding_step.parameters()))
        if grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(encoder.parameters(), grad_norm)
            torch.nn.utils.clip_grad_norm_(decoding_step.parameters(), grad_norm)
        opt.step()
    encoder._t_max_hint = None

    n_batches = max(len(data_loader), 1)
    avg_decoder_loss = float(dec_sum) / n_batches
    print(f"\tTraining decoder_loss: {avg_decoder_loss}")
    avg_ctc_loss = 0
    if use_ctc:
        avg_ctc_loss = float(ctc_sum) / n_batches
        print(f"\tTraining ctc_loss: {avg_ctc_loss}")
    return avg_decoder_loss, avg_ctc_loss


def eval(encoder, decoding_step, data_loader, device, char2idx):
    """Teacher-forced decode; returns (decoder_loss, correct, count) like the reference (:89-143).
    `correct` comes from a multinomial sample of the decoder distribution, as in the reference, so
    CER is stochastic unless the generator is seeded."""
    use_ctc = encoder.enable_ctc
    encoder.eval()
    decoding_step.eval()
    pad = char2idx[PAD]
    decoder_loss = torch.zeros((), device=device)
    correct = torch.zeros((), dtype=torch.long, device=device)
    count = 0
    with torch.no_grad():
        for frames, frame_lens, chars, char_lens in data_loader:
            fl_h, cl_h = frame_lens.cpu(), char_lens.cpu()
            _check_batch(chars.cpu(), cl_h, fl_h, char2idx, False)
            frames, chars = frames.to(device, non_blocking=True), chars.to(device, non_blocking=True)
            frame_lens_d = fl_h.to(device, non_blocking=True)
            labels = chars[:, 1:]
            ll_h = cl_h - 1
            batch_size = frames.shape[0]
            encoder._t_max_hint = int(fl_h.max())
            if use_ctc:
                enc_out, enc_h, prev_state = encoder(frames, frame_lens_d)
                cur = ctc_loss(enc_out, labels, frame_lens_d, ll_h, "sum", device, host_lens=(fl_h, ll_h))
                if cur is None:
                    continue
            else:
                enc_h, prev_state = encoder(frames, frame_lens_d)
            Lm = int(ll_h.max())
            if SEQUENCE_DECODE and hasattr(decoding_step, "forward_sequence"):
                log_probs, prev_state = decoding_step.forward_sequence(chars[:, :Lm], prev_state, frame_lens_d, enc_h)
                flat, lab = log_probs.reshape(-1, log_probs.shape[-1]), labels[:, :Lm].reshape(-1)
                decoder_loss += F.nll_loss(flat, lab, ignore_index=pad, reduction="sum")
                sampled = flat.exp().multinomial(1).squeeze(-1)      # same distribution as the per-step draws
                correct += ((sampled == lab) & (lab != pad)).sum()
            else:
                for i in range(Lm):
                    log_probs, prev_state = decoding_step(chars[:, i], prev_state, frame_lens_d, enc_h)
                    decoder_loss += F.nll_loss(log_probs, labels[:, i], ignore_index=pad, reduction="sum")
                    sampled = log_probs.exp().multinomial(1).squeeze(-1)
                    correct += ((sampled == labels[:, i]) & (labels[:, i] != pad)).sum()
            count += int(ll_h.sum())
    encoder._t_max_hint = None
    count_t = torch.tensor(float(count), device=device)
    return decoder_loss / count_t, correct, count_t


def train_ctc(encoder, data_loader, opt, device, char2idx=None, grad_norm=None, dist=None, on_step=None):
    """CTC-only epoch: encoder -> CTC 'mean' -> backward -> [all-reduce] -> clip -> step.  This is the
    loop shape of the reference's archived `train_model.py` (src/scripts/archive/train_model.py:
    model -> CTCLoss -> backward -> clip -> optimizer.step) on top of the live VideoEncoder, i.e. the
    north-star hot path without the attention decoder.  Returns the average CTC loss.
    `data_loader` yields (frames|clips, frame_lens, chars, char_lens) with tensors on the host or the
    device; BOS-prefixed / EOS-terminated chars as everywhere else."""
    assert encoder.enable_ctc
    encoder.train()
    total = torch.zeros((), device=device)
    n = 0
    for frames, frame_lens, chars, char_lens in data_loader:
        fl_h, cl_h = frame_lens.cpu(), char_lens.cpu()
        ll_h = cl_h - 1
        frames = frames.to(device, non_blocking=True)
        labels = chars.to(device, non_blocking=True)[:, 1:]
        fl_d = fl_h.to(device, non_blocking=True)
        encoder._t_max_hint = int(fl_h.max())
        log_probs, _, _ = encoder(frames, fl_d)
        loss = ctc_loss(log_probs, labels, fl_d, ll_h, "mean", device, host_lens=(fl_h, ll_h))
        if loss is None:
            continue
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if dist is not None:
            dist.allreduce_grads(list(encoder.parameters()))
        if grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(encoder.parameters(), grad_norm)
        opt.step()
        total += loss.detach()
        n += 1
        if on_step is not None:
            on_step(loss)
    encoder._t_max_hint = None
    return float(total) / max(n, 1)
