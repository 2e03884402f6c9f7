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

# Where the decoder's characters are sampled (train: the non-teacher-forced input; eval: the CER hits).
#   "device": one multinomial on the GPU (fast; its random stream is the CUDA generator's, so CER is only
#             statistically comparable with a reference run);
#   "cpu":    the CER-parity protocol of SURVEY §7 — the log-probs computed on the GPU are brought to the host and
#             sampled there with the reference's calls in the reference's order (per label position one
#             `torch.rand(1)` in train, one `(B,V)` `multinomial(1)` in train and eval:
#             train_better_model.py:56-63,130), so after `torch.manual_seed(seed)` the draws — and with them
#             `correct`, CER and the fed-back characters — equal those of the reference running on its CPU device.
SAMPLING = "device"


def _check_batch(chars, char_lens, frame_lens, char2idx, use_ctc):
    assert bool((chars[:, 0] == char2idx[BOS]).all())
    assert bool((chars.gather(1, (char_lens - 1).unsqueeze(1)).squeeze(1) == char2idx[EOS]).all())
    if use_ctc:
        assert bool((frame_lens >= char_lens).all())             # otherwise ctc loss will produce inf


def train(encoder, decoding_step, data_loader, opt, device, char2idx,
          teacher_forcing_ratio=1, grad_norm=None, dist=None, sampling=None):
    """Same contract as the reference `train`: returns (avg_decoder_loss, avg_ctc_loss)."""
    use_ctc = encoder.enable_ctc
    sampling = sampling or SAMPLING
    assert sampling in ("device", "cpu")
    dec_sum = torch.zeros((), device=device)
    ctc_sum = torch.zeros((), device=device)
    encoder.train()
    decoding_step.train()
    pad = char2idx[PAD]
    params_all = list(encoder.parameters()) + list(decoding_step.parameters())
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
            cur_ctc = ctc_loss(enc_out, labels, frame_lens_d, ll_h, "mean", device, host_lens=(fl_h, ll_h),
                               host_labels=chars_h[:, 1:])
            if cur_ctc is None:
                # the reference skips the batch (train_better_model.py:41-42).  Under data parallelism the other
                # ranks are about to all-reduce: join them with zero gradients so the collectives stay paired
                if dist is not None:
                    opt.zero_grad()
                    dist.allreduce_grads(params_all, valid=False)
                    if grad_norm is not None:
                        torch.nn.utils.clip_grad_norm_(encoder.parameters(), grad_norm)
                        torch.nn.utils.clip_grad_norm_(decoding_step.parameters(), grad_norm)
                    opt.step()
                continue
        else:
            enc_h, prev_state = encoder(frames, frame_lens_d)
        prev_output = torch.full((batch_size,), char2idx[BOS], dtype=torch.long, device=device)

        if SEQUENCE_DECODE and hasattr(decoding_step, "forward_sequence"):
            # Segmented decode.  Which positions are teacher forced is decided up front (one `torch.rand(1)` per
            # position, as in the reference); a maximal run of positions whose INPUTS are known — a sampled
            # character followed by teacher-forced ones — is one vectorised pass (`forward_sequence`: one RNN call,
            # one attention launch, one projection launch), chained through the recurrent state.  With every step
            # teacher forced that is a single pass; at the reference's default ratio of 0.9 about L/10 + 1 passes
            # instead of L.  Same arithmetic as the step loop: the recurrent state never depends on the attention
            # output (better_model.py:184,223-229), and a position's log-probs depend on earlier positions only
            # through that state and the fed character.
            forced, noise = [], []
            for i in range(max_label_len):
                forced.append(bool(torch.rand(1) < teacher_forcing_ratio))
                if sampling == "cpu":
                    # the reference's multinomial(1) of position i = argmax(p / q) with q ~ Exp(1) drawn here, in
                    # the reference's order (ATen's single-sample path), so the host generator stays in step
                    noise.append(torch.empty(batch_size, decoding_step.vocab_size).exponential_(1))
            parts, i = [], 0
            while i < max_label_len:
                j = i + 1
                while j < max_label_len and forced[j]:
                    j += 1
                inputs = chars[:, i:j]
                if not forced[i]:
                    inputs = torch.cat([prev_output.unsqueeze(1), inputs[:, 1:]], dim=1)
                log_probs, prev_state = decoding_step.forward_sequence(inputs, prev_state, frame_lens_d, enc_h)
                parts.append(log_probs)
                if j < max_label_len:                              # position j is fed the character sampled at j-1
                    last = log_probs[:, -1].detach()
                    if sampling == "cpu":
                        prev_output = (last.cpu().exp() / noise[j - 1]).argmax(-1).to(device)
                    else:
                        prev_output = last.exp().multinomial(1).squeeze(-1)
                i = j
            log_probs = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            decoder_loss = F.nll_loss(log_probs.reshape(-1, log_probs.shape[-1]), labels[:, :max_label_len].reshape(-1),
                                      ignore_index=pad, reduction="sum")
        else:
            decoder_loss = 0
            for i in range(max_label_len):
                teacher_forcing = bool(torch.rand(1) < teacher_forcing_ratio)
                input_ = chars[:, i] if teacher_forcing else prev_output
                log_probs, prev_state = decoding_step(input_, prev_state, frame_lens_d, enc_h)
                decoder_loss = decoder_loss + F.nll_loss(log_probs, labels[:, i], ignore_index=pad, reduction="sum")
                if sampling == "cpu":
                    prev_output = log_probs.detach().cpu().exp().multinomial(1).squeeze(-1).to(device)
                else:
                    prev_output = log_probs.detach().exp().multinomial(1).squeeze(-1)
        decoder_loss = decoder_loss / n_tokens

        opt.zero_grad()
        decoder_loss.backward(retain_graph=use_ctc)
        dec_sum += decoder_loss.detach()
        if use_ctc:
            cur_ctc.backward()
            ctc_sum += cur_ctc.detach()
        if dist is not None:
            dist.allreduce_grads(params_all)
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


def eval(encoder, decoding_step, data_loader, device, char2idx, sampling=None, details=None):
    """Teacher-forced decode; returns (decoder_loss, correct, count) like the reference (:89-143).
    `correct` comes from a multinomial sample of the decoder distribution, as in the reference, so
    CER is stochastic unless the generator is seeded: `sampling="cpu"` (or trainer.SAMPLING) draws the samples
    on the host exactly as the reference does (see SAMPLING).  `details`, when a dict, additionally receives the
    deterministic protocol: 'correct_argmax' (hits of the arg-max character) and 'samples' (list of (B,L) tensors)."""
    use_ctc = encoder.enable_ctc
    sampling = sampling or SAMPLING
    assert sampling in ("device", "cpu")
    hit_argmax = torch.zeros((), dtype=torch.long, device=device)
    all_samples = []
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
                if sampling == "cpu":
                    lp_h = log_probs.cpu()
                    sampled = torch.stack([lp_h[:, i].exp().multinomial(1).squeeze(-1) for i in range(Lm)], 1)
                    sampled = sampled.to(device).reshape(-1)
                else:
                    sampled = flat.exp().multinomial(1).squeeze(-1)      # same distribution as the per-step draws
                correct += ((sampled == lab) & (lab != pad)).sum()
                hit_argmax += ((flat.argmax(-1) == lab) & (lab != pad)).sum()
                all_samples.append(sampled.reshape(batch_size, Lm))
            else:
                step_samples = []
                for i in range(Lm):
                    log_probs, prev_state = decoding_step(chars[:, i], prev_state, frame_lens_d, enc_h)
                    decoder_loss += F.nll_loss(log_probs, labels[:, i], ignore_index=pad, reduction="sum")
                    if sampling == "cpu":
                        sampled = log_probs.cpu().exp().multinomial(1).squeeze(-1).to(device)
                    else:
                        sampled = log_probs.exp().multinomial(1).squeeze(-1)
                    correct += ((sampled == labels[:, i]) & (labels[:, i] != pad)).sum()
                    hit_argmax += ((log_probs.argmax(-1) == labels[:, i]) & (labels[:, i] != pad)).sum()
                    step_samples.append(sampled)
                all_samples.append(torch.stack(step_samples, 1))
            count += int(ll_h.sum())
    encoder._t_max_hint = None
    count_t = torch.tensor(float(count), device=device)
    if details is not None:
        details["correct_argmax"] = hit_argmax
        details["samples"] = all_samples
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
    params = list(encoder.parameters())
    # with a conv front-end, the recurrent layer's and the projection's gradients are complete while the front-end's
    # backward (most of the step) still runs: they form the early all-reduce bucket
    early = (list(encoder.rnn.parameters()) + list(encoder.output_proj.parameters())) \
        if (dist is not None and hasattr(dist, "arm") and getattr(encoder, "frame_processing", "") == "conv3d") else None
    for frames, frame_lens, chars, char_lens in data_loader:
        fl_h, cl_h, chars_h = frame_lens.cpu(), char_lens.cpu(), chars.cpu()
        ll_h = cl_h - 1
        frames = frames.to(device, non_blocking=True)
        labels = chars.to(device, non_blocking=True)[:, 1:]
        fl_d = fl_h.to(device, non_blocking=True)
        encoder._t_max_hint = int(fl_h.max())
        log_probs, _, _ = encoder(frames, fl_d)
        if hasattr(data_loader, "kick"):
            data_loader.kick()                               # prefetcher in kick mode: next copy starts behind the forward pass
        # (labels and lengths are on the host already: the wrapper decides feasibility there and never reads the
        # device, so the host keeps enqueueing a step ahead of the GPU)
        loss = ctc_loss(log_probs, labels, fl_d, ll_h, "mean", device, host_lens=(fl_h, ll_h), host_labels=chars_h[:, 1:])
        opt.zero_grad(set_to_none=True)
        if loss is None:
            if dist is None:
                continue
            dist.allreduce_grads(params, valid=False)        # keep the collectives paired across ranks
            if grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(params, grad_norm)
            opt.step()
            continue
        if dist is not None and early:
            dist.arm(early)                                  # recurrent + projection gradients reduce during conv backward
        loss.backward()
        if dist is not None:
            dist.allreduce_grads(params)
        if grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(params, grad_norm)
        opt.step()
        total += loss.detach()
        n += 1
        if on_step is not None:
            on_step(loss)
    encoder._t_max_hint = None
    return float(total) / max(n, 1)
