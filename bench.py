#!/usr/bin/env python
"""bench.py — frames/sec of the north-star training step on synthetic clips.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port) on host cores

One "step" = one CTC-only training step of the hot path on one batch of synthetic video:
    u8 clips (B,75,100,50,3) -> STCNN conv front-end (tcgen05) -> BiGRU-256 -> Linear+masked
    log-softmax -> CTC loss -> backward -> [NCCL grad all-reduce] -> clip 50 -> Adam
with B = 256 clips per GPU (BASELINE.json configs[2]/[3]; weak scaling: per-GPU batch fixed).
`value` = frames/s with the batch resident in HBM; `e2e` = the same step through the public API
(`lipreading_b200.trainer.train_ctc`) fed from pinned host memory (H2D of every batch inside the
timed region, loss read back every step).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES, H_FRAME, W_FRAME = 75, 100, 50
SEED = 123456


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU (weak scaling) / in total (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --batch is the GLOBAL batch, split evenly over the ranks (BASELINE config 4 reads 'batch=256 clips at 1/2/4/8')")
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--rnn", default="GRU")
    ap.add_argument("--cpu-batch", type=int, default=8, help="clips per step of the CPU baseline sample")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel micro rooflines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity columns (GPU step vs the fp32 CPU port)")
    ap.add_argument("--no-ref-shape", action="store_true", help="skip the reference-shape (B,T,68,3) train() block")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def synth_batch(B, seed, char2idx, L_lo=10, L_hi=30):
    """Seeded synthetic batch of the dataview/clip shape: u8 clips + BOS/EOS-wrapped labels."""
    g = torch.Generator().manual_seed(seed)
    clips = torch.randint(0, 256, (B, T_FRAMES, H_FRAME, W_FRAME, 3), dtype=torch.uint8, generator=g)
    lens = torch.full((B,), T_FRAMES, dtype=torch.long)
    L = torch.randint(L_lo, L_hi + 1, (B,), generator=g)
    chars = torch.zeros(B, int(L.max()) + 2, dtype=torch.long)
    for b in range(B):
        n = int(L[b])
        chars[b, 0] = char2idx["<BOS>"]
        chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
        chars[b, 1 + n] = char2idx["<EOS>"]
    return clips, lens, chars, L + 2


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU path: the oracle port of the same step (torch CPU fp32: conv3d -> nn.GRU packed -> CTC)
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(hidden, rnn, char2idx):
    """fp32 CPU port of the step (oracle/train_step.py: F.conv3d -> packed nn.GRU -> masked log-softmax -> the
    ctc_loss wrapper -> clip -> Adam)"""
    from oracle import train_step as TS
    port = TS.CpuStep(TS.random_state(hidden, rnn, char2idx, conv=True, seed=SEED), rnn, char2idx, lr=1e-4, grad_norm=50)
    return lambda batch: port.step(batch)[0]


def configure_throughput_path():
    """The three switches of the configuration this benchmark times (BASELINE config 3: "bf16"); returns a function
    that restores the parity-path defaults.  tests/test_gpu_bench_config.py holds exactly this configuration to the
    fp32 CPU port."""
    from lipreading_b200 import conv_frontend, functional as LF
    saved = (LF.GEMM_DTYPE, LF.RNN_CLUSTER, conv_frontend.OUT_DTYPE, LF.PROJ_VARIANT, LF.GEMM_TCGEN05)
    LF.GEMM_DTYPE = torch.bfloat16              # plain GEMMs: bf16 operands, fp32 accumulate
    LF.RNN_CLUSTER = True                       # persistent cluster recurrence (bf16 operands, fp32 state)
    conv_frontend.OUT_DTYPE = torch.bfloat16    # the conv3 epilogue's bf16 features feed the bf16 input GEMM directly
    LF.PROJ_VARIANT = 2                         # projection + log-softmax on tcgen05 kind::tf32
    # the four plain GEMMs around the recurrent kernel stay on cuBLAS (measured: 12.81 ms/step against 13.46 with them on
    # lr_tapgemm — their K-major transposes and a one-CTA tile without multicast cost more than the library call);
    # LR_GEMM_TCGEN05=1 routes them through this repo's kernel
    LF.GEMM_TCGEN05 = os.environ.get("LR_GEMM_TCGEN05", "0") == "1"

    def restore():
        LF.GEMM_DTYPE, LF.RNN_CLUSTER, conv_frontend.OUT_DTYPE, LF.PROJ_VARIANT, LF.GEMM_TCGEN05 = saved
    return restore


def usable_cores():
    """Cores this process may really use: affinity mask, capped by the cgroup CPU quota (a container that
    sees 128 logical CPUs but is limited to a few would otherwise oversubscribe its OpenMP pool)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:                       # cgroup v2
            quota, period = fh.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        try:                                                              # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fh:
                quota = int(fh.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fh:
                period = int(fh.read())
            if quota > 0:
                n = max(1, min(n, quota // period))
        except Exception:
            pass
    return n


def run_cpu(args, char2idx, steps, warmup, budget_s=150.0):
    """Times the oracle port on the host cores.  The per-step sample (clips per step) is sized from a
    one-clip calibration step so that warmup+steps fit in `budget_s`."""
    step = cpu_step_fn(args.hidden, args.rnn, char2idx)
    one = synth_batch(1, SEED, char2idx)
    # all the host threads it can use -- but no more than help: pick the fastest pool size on one clip
    # (a 128-thread pool on a quota-limited container ran 100x slower than 8 threads)
    limit = usable_cores()
    best = None
    for n in sorted({limit, min(limit, 64), min(limit, 32), min(limit, 16), min(limit, 8)}, reverse=True):
        torch.set_num_threads(n)
        step(one)                                               # untimed: allocator / thread-pool warm-up
        t0 = time.perf_counter()
        step(one)
        t = time.perf_counter() - t0
        if best is None or t < best[0]:
            best = (t, n)
        if t > 20.0:
            continue
    t1, cores = best
    torch.set_num_threads(cores)
    n_clips = int(max(1, min(args.cpu_batch, budget_s / (max(steps + warmup, 1) * t1))))
    batch = synth_batch(n_clips, SEED, char2idx)
    for _ in range(warmup):
        step(batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(batch)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_clips * T_FRAMES / dt, dt, cores, n_clips


# ------------------------------------------------------------------------------------------------
# parity columns and the reference-shape block, rank 0 at N=1 only
# ------------------------------------------------------------------------------------------------
def parity_block(dev, clips=64):
    """The timed configuration held to the fp32 CPU port on `clips` clips of the bench's own synthetic batch, in
    this very run: the same function tests/test_gpu_bench_config.py asserts on (B = 32 and 256 there)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_bench_config import throughput_step_parity
    r = throughput_step_parity(clips, dev)
    g = r.pop("grad_rel_fro_err")
    r["grad_rel_fro_err_max"] = max(g.values())
    r["bars"] = {"logprob_max_abs_err": 2e-2, "ctc_loss_rel_err": 1e-2, "grad_rel_fro_err": 8e-2}
    r["within_bars"] = bool(r["logprob_max_abs_err"] <= 2e-2 and r["ctc_loss_rel_err"] <= 1e-2 and
                            r["grad_rel_fro_err_max"] <= 8e-2)
    r["also"] = ("fp32 path: log-probs / CTC loss / gradients <= 1e-4 vs the unmodified reference's golden vectors and at "
                 "B=256 BiGRU-256 / B=128 BiLSTM-768 (tests/test_gpu_train_step.py, test_gpu_bench_config.py); padded rect, "
                 "crop size, gather indices bit-exact (tests/test_gpu_vision.py); CER under the seeded host-sampling "
                 "protocol equal to the reference's eval() (tests/test_gpu_train_step.py)")
    return r


def _ref_shape_batch(B, T, mixed, seed, char2idx):
    """BASELINE.md §3.1 inputs: randn(B,T,68,3) f32, labels [BOS]+randint(4,64,(L,))+[EOS], L in [10,30];
    all T=75 or ascending mixed T in [40,75]."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(40, T + 1, (B,), generator=g).sort().values if mixed else torch.full((B,), T, dtype=torch.long)
    lens[-1] = T
    frames = torch.randn(B, T, 68, 3, generator=g)
    for b in range(B):
        frames[b, int(lens[b]):] = 0
    L = torch.randint(10, 31, (B,), generator=g)
    chars = torch.zeros(B, int(L.max()) + 2, dtype=torch.long)
    for b in range(B):
        n = int(L[b])
        chars[b, 0], chars[b, 1 + n] = char2idx["<BOS>"], char2idx["<EOS>"]
        chars[b, 1:1 + n] = torch.randint(4, 64, (n,), generator=g)
    return frames, lens, chars, L + 2


def ref_shape_block(dev, char2idx, steps=3):
    """The reference's own `train()` (encoder + CTC aux + teacher-forced decoder, clip 50, Adam) on ref-shape batches:
    GPU arm = lipreading_b200.trainer.train on pinned host batches (fp32 parity path for the kernels, H2D inside the
    timed region); CPU arm = the oracle port of train() (oracle/train_step.CpuReferenceTrain, pinned against the
    unmodified reference's golden train step) on a bounded sample of the same batch."""
    from lipreading_b200 import functional as LF, trainer
    from lipreading_b200.model import CharDecodingStep, VideoEncoder
    from oracle import sequence as O
    from oracle import train_step as TS
    saved = (LF.GEMM_DTYPE, LF.RNN_CLUSTER)
    out = []
    try:
        for rnn, H, B, cpu_B in (("GRU", 256, 256, 64), ("LSTM", 768, 128, 16)):
            for mixed in (False, True):
                LF.GEMM_DTYPE, LF.RNN_CLUSTER = torch.float32, False
                torch.manual_seed(SEED)
                enc = VideoEncoder(204, H, rnn_type=rnn, bidirectional=True, enable_ctc=True, vocab_size=len(char2idx),
                                   char2idx=char2idx, device=dev).to(dev)
                dec = CharDecodingStep(enc, char_dim=256, vocab_size=len(char2idx), char2idx=char2idx,
                                       attention_type="none", device=dev).to(dev)
                batch = tuple(t.pin_memory() for t in _ref_shape_batch(B, T_FRAMES, mixed, SEED, char2idx))
                opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-4)
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):
                    trainer.train(enc, dec, [batch], opt, dev, char2idx, teacher_forcing_ratio=1, grad_norm=50)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dl, cl = trainer.train(enc, dec, [batch] * steps, opt, dev, char2idx, teacher_forcing_ratio=1, grad_norm=50)
                    e1.record()
                    torch.cuda.synchronize()
                s_gpu = e0.elapsed_time(e1) * 1e-3 / steps
                n_frames = int(batch[1].sum())
                row = {"config": "bi%s%d B=%d T=%s ref-shape (B,T,68,3), train() with decoder, attention none"
                                 % (rnn.lower(), H, B, "40..75 mixed" if mixed else "75"),
                       "gpu_frames_per_s": n_frames / s_gpu, "gpu_ms_per_step": s_gpu * 1e3, "gpu_path": "fp32 parity path",
                       "decoder_loss": dl, "ctc_loss": cl}
                # the same call on the throughput path: persistent recurrent kernels (8-CTA cluster for H = 256, the
                # grid-persistent kernels for H = 768) with bf16 operands, bf16 library GEMMs around them
                LF.GEMM_DTYPE, LF.RNN_CLUSTER = torch.bfloat16, True
                with contextlib.redirect_stdout(io.StringIO()):
                    trainer.train(enc, dec, [batch], opt, dev, char2idx, teacher_forcing_ratio=1, grad_norm=50)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    trainer.train(enc, dec, [batch] * steps, opt, dev, char2idx, teacher_forcing_ratio=1, grad_norm=50)
                    e1.record()
                    torch.cuda.synchronize()
                s_thr = e0.elapsed_time(e1) * 1e-3 / steps
                row["gpu_frames_per_s_persistent_bf16"] = n_frames / s_thr
                row["gpu_ms_per_step_persistent_bf16"] = s_thr * 1e3
                LF.GEMM_DTYPE, LF.RNN_CLUSTER = torch.float32, False
                if not mixed:
                    # CPU arm on a bounded sample (first cpu_B clips of the same batch), all usable host threads
                    sample = tuple(t[:cpu_B].clone() for t in batch)
                    sample = (sample[0], sample[1], sample[2][:, : int(sample[3].max())], sample[3])
                    enc_state = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
                    dec_o = O.OracleDecoder(2 * H, rnn, 256, len(char2idx), char2idx, attention_type="none")
                    dec_o.load_state_dict({k: v.detach().cpu() for k, v in dec.state_dict().items()})
                    port = TS.CpuReferenceTrain(enc_state, dec_o, rnn, char2idx, lr=1e-4, grad_norm=50)
                    torch.set_num_threads(min(usable_cores(), 32))
                    port.step(sample)
                    t0 = time.perf_counter()
                    port.step(sample)
                    s_cpu = time.perf_counter() - t0
                    row["cpu_frames_per_s"] = int(sample[1].sum()) / s_cpu
                    row["cpu_sample"] = "%d clips/step, 1 warm-up + 1 timed, %d threads, oracle port of train()" % (
                        cpu_B, torch.get_num_threads())
                out.append(row)
                del enc, dec, opt
    finally:
        LF.GEMM_DTYPE, LF.RNN_CLUSTER = saved
    return out


def frame_stream_block(dev, enc, char2idx, n_clips=8):
    """Raw 720p frames + boxes -> characters in one pass (infer.FrameRecognizer): rect geometry -> warp256 -> position
    map CNN (tcgen05 tap-GEMM plan, random weights: the reference ships none) -> 68 landmarks -> mouth crop -> conv
    front-end -> BiGRU -> greedy CTC.  The CNN (8.26 GFLOP/frame) dominates; the per-stage kernels are in `kernels`."""
    import numpy as np
    from lipreading_b200.face import PRN
    from lipreading_b200.infer import FrameRecognizer
    from lipreading_b200.prnet import PosPrediction
    gold = os.path.join(ROOT, "tests", "golden")
    uv = np.loadtxt(os.path.join(gold, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(gold, "face_ind.npy"))
    pred = PosPrediction(device=dev)
    prn = PRN(predict_batch=pred.predict_batch, uv_kpt_ind=uv, face_ind=face, device=dev)
    stream = FrameRecognizer(enc, char2idx, prn, batch=64)
    n = n_clips * T_FRAMES
    frames = torch.randint(0, 256, (n, 720, 1280, 3), dtype=torch.uint8, device=dev)
    rects = torch.tensor([[400, 700, 150, 450]] * n, dtype=torch.int32)
    s = time_cuda(lambda: stream.tokens(frames, T_FRAMES, rects), iters=3, warm=2)
    enc.train()
    del frames, stream, prn, pred
    torch.cuda.empty_cache()
    return {"value": n / s, "unit": "frames/s", "ms": s * 1e3, "frames": n,
            "what": "%d raw 720p frames (%d clips) + boxes -> token ids, incl. the position-map CNN" % (n, n_clips)}


# ------------------------------------------------------------------------------------------------
# per-kernel micro rooflines (rank 0, N=1 only): the HBM-bound kernels of the path
# ------------------------------------------------------------------------------------------------
def time_cuda(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()                       # write a buffer larger than L2 between iterations
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e-3


def kernel_rooflines(dev, pk, char2idx):
    import numpy as np
    from lipreading_b200 import functional as LF
    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm = pk["hbm_gbs"] * 1e9
    g = torch.Generator().manual_seed(SEED)
    # CTC alpha/beta + grad: 2*T*C*4 B per clip (39 000 B at T=75, C=65)
    B, T, C = 4096, 75, 65
    lp = torch.randn(B, T, C, generator=g).log_softmax(-1).to(dev).requires_grad_(True)
    tg = torch.randint(5, 65, (B, 30), generator=g).to(dev).int()
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.randint(10, 31, (B,), generator=g).to(dev).int()
    s = time_cuda(lambda: LF.ctc_nll(lp, tg, il, tl), flush=flush)
    out.append({"kernel": "ctc fwd+grad (linear-space warp kernel)", "bound": "hbm", "unit": "GB/s", "achieved": B * 2 * T * C * 4 / s / 1e9,
                "frac": B * 2 * T * C * 4 / s / hbm, "shape": "B=4096,T=75,C=65,L<=30", "ms": s * 1e3})
    # proj + masked log-softmax fwd: reads M*K*4, writes M*C*4
    M, K = 256 * 75, 512
    h = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(C, K, generator=g) / 22).to(dev)
    b = torch.zeros(C, device=dev)
    mask = torch.ones(C)
    mask[1] = mask[2] = 0                       # PAD+1, BOS+1 (better_model.py:43-45)
    lm = (mask + 1e-45).log().to(dev)
    s = time_cuda(lambda: LF.proj_masked_log_softmax(h, w, b, lm), flush=flush)
    byts = M * (K + C) * 4
    out.append({"kernel": "proj_logsoftmax_fwd", "bound": "hbm", "unit": "GB/s", "achieved": byts / s / 1e9,
                "frac": byts / s / hbm, "shape": "M=19200,K=512,C=65", "ms": s * 1e3})
    # warp256: reads size^2*3 u8 window, writes 256*256*3 f32
    n, H, W = 384, 720, 1280           # 1 GB of frames, 0.3 GB of output: far larger than the 126 MB L2
    frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
    rects = torch.tensor([[400, 700, 150, 450]] * n, dtype=torch.int32, device=dev)
    rp, crop = LF.rect_geometry(rects, H, W)
    s = time_cuda(lambda: LF.warp256(frames, crop), flush=flush)
    size = int(crop[0, 2])
    byts = n * (size * size * 3 + 256 * 256 * 3 * 4)
    out.append({"kernel": "warp256", "bound": "hbm", "unit": "GB/s", "achieved": byts / s / 1e9, "frac": byts / s / hbm,
                "shape": "n=%d,720p,size=%d" % (n, size), "ms": s * 1e3})
    # posmap gather with vertices: reads 43867*12 B (+68*12), writes (43867+68)*24 B per frame
    gold = os.path.join(ROOT, "tests", "golden")
    uv = np.loadtxt(os.path.join(gold, "uv_kpt_ind.txt")).astype(np.int64)
    kidx = torch.from_numpy((uv[1] * 256 + uv[0]).astype(np.int32)).to(dev)
    fidx = torch.from_numpy(np.load(os.path.join(gold, "face_ind.npy"))).to(dev)
    pos = (torch.rand(n, 256, 256, 3, generator=g) * 281.6).to(dev)
    s = time_cuda(lambda: LF.posmap_gather(pos, crop, rp, kidx, fidx), flush=flush)
    byts = n * (43867 + 68) * (12 + 24)
    out.append({"kernel": "posmap_gather(lmk+vtx)", "bound": "hbm", "unit": "GB/s", "achieved": byts / s / 1e9,
                "frac": byts / s / hbm, "shape": "n=%d" % n, "ms": s * 1e3})
    # N2 mouth crop: reads the mouth ROI of each frame, writes (100,50,3) u8 — the dataview clip shape.  The position
    # map is the identity map of the crop (+ noise), i.e. a frontal face: the 20 mouth landmarks then span ~1/6 of the
    # face box, as on real faces (a uniformly random map makes the "mouth" as large as the whole face).  n = 4096
    # 720p frames (11 GB): far beyond L2, 28 CTAs per SM.
    nm = 4096
    vv, uu = torch.meshgrid(torch.arange(256.0), torch.arange(256.0), indexing="ij")
    pos_id = (torch.stack([uu, vv, torch.full_like(uu, 40.0)], -1)[None] + torch.randn(8, 256, 256, 3, generator=g)).to(dev)
    lmk8 = LF.posmap_gather(pos_id, crop[:8], rp[:8], kidx)
    lmk_m = lmk8.repeat(nm // 8, 1, 1)
    rp_m = rp[:1].repeat(nm, 1)
    frames_m = torch.randint(0, 256, (nm, H, W, 3), dtype=torch.uint8, device=dev)
    s = time_cuda(lambda: LF.mouth_crop(frames_m, lmk_m, rp_m, 100, 50), flush=flush, iters=10)
    _, roi = LF.mouth_crop(frames_m, lmk_m, rp_m, 100, 50)
    roi_h = roi.cpu().long()
    # roi = (x_lo, y_lo, w, h); algorithmic bytes = the ROI's pixels (every one is a tap when the ROI is about the
    # output's size) + the written clip frame
    roi_px = int((roi_h[:, 2].clamp(min=0) * roi_h[:, 3].clamp(min=0)).sum())
    byts = roi_px * 3 + nm * 100 * 50 * 3
    out.append({"kernel": "mouth_crop", "bound": "hbm", "unit": "GB/s", "achieved": byts / s / 1e9, "frac": byts / s / hbm,
                "shape": "n=%d 720p, ROI %dx%d px read + 15 kB written per frame" % (nm, int(roi_h[0, 2]), int(roi_h[0, 3])),
                "ms": s * 1e3})
    del frames_m, lmk_m, rp_m
    # position-map CNN body alone (rows a5 / f4): 53 lr_tapgemm launches + the image packing per batch of 64 frames,
    # 4.13 GMAC/frame of model work (the plan's own count, padding and zero-weight fused taps excluded)
    try:
        from lipreading_b200.prnet import PosPrediction
        pred = PosPrediction(device=dev)
        nb = 64
        plan = pred.plan(nb)
        img = torch.rand(nb, 256, 256, 3, generator=g).to(dev)
        s = time_cuda(lambda: plan.run(img), iters=10, warm=3, flush=flush)
        out.append({"kernel": "prnet_body (tapgemm_kernel x53, tcgen05 bf16)", "bound": "tensor", "unit": "TFLOP/s",
                    "achieved": plan.model_flops / s / 1e12, "frac": plan.model_flops / s / 1e12 / pk["bf16_tflops"],
                    "issued_tflops": plan.flops / s / 1e12, "frames_per_s": nb / s, "shape": "batch %d x 256x256x3" % nb,
                    "ms": s * 1e3})
        # whole per-frame vision path (BASELINE config 2): rect geometry -> warp256 -> position-map CNN -> restore +
        # 68-landmark gather, every stage a kernel of this repo (random CNN weights: the reference ships none)

        def vision():
            for i in range(0, n, nb):
                r2, c2 = LF.rect_geometry(rects[i:i + nb], H, W)
                pm = pred.predict_batch(LF.warp256(frames[i:i + nb], c2))
                LF.posmap_gather(pm, c2, r2, kidx)
        s = time_cuda(vision, iters=3, warm=2)
        out.append({"kernel": "vision_stream(frames->68 landmarks, incl. the position-map CNN on tcgen05)", "bound": "tensor",
                    "unit": "frames/s", "achieved": n / s, "frac": n * plan.model_flops / nb / s / 1e12 / pk["bf16_tflops"],
                    "shape": "n=%d,720p, batches of %d" % (n, nb), "ms": s * 1e3})
        ref = PosPrediction(device=dev, dtype=torch.bfloat16, engine="torch")
        s = time_cuda(lambda: ref.predict_batch(img), iters=3, warm=2)
        out.append({"kernel": "prnet_body via torch/cuDNN bf16 channels-last (library baseline)", "bound": "tensor",
                    "unit": "TFLOP/s", "achieved": plan.model_flops / s / 1e12, "frac": plan.model_flops / s / 1e12 / pk["bf16_tflops"],
                    "frames_per_s": nb / s, "shape": "batch %d" % nb, "ms": s * 1e3})
        del pred, ref, plan
    except Exception as e:
        out.append({"kernel": "vision_stream", "error": repr(e)})
    # recurrent layer fwd (BiGRU-256, B=256, T=75): latency-bound; report FLOP/s of the recurrent GEMMs
    from lipreading_b200.model import NativeRNN
    rnn = NativeRNN("GRU", 1728, 256, bidirectional=True).to(dev)
    x = torch.randn(256, 75, 1728, generator=g).to(dev)
    lens = torch.full((256,), 75, dtype=torch.int32, device=dev)
    with torch.no_grad():
        s = time_cuda(lambda: rnn(x, lens), iters=5)
    fl = 2 * 75 * 256 * 2 * 3 * 256 * (1728 + 256)
    out.append({"kernel": "bigru256_fwd(layer incl. input GEMM)", "bound": "tensor", "unit": "TFLOP/s",
                "achieved": fl / s / 1e12, "frac": fl / s / 1e12 / pk["bf16_tflops"], "shape": "B=256,T=75,I=1728,H=256 fp32",
                "ms": s * 1e3, "us_per_step": s / 75 * 1e6})
    return out


# ------------------------------------------------------------------------------------------------
def main():
    # NCCL's own log lines (NCCL_DEBUG=INFO when the driver sets it) must not mix with the ONE JSON line on stdout:
    # send them to stderr instead of silencing them
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    args = parse()
    from lipreading_b200.vocab import build_char2idx
    char2idx = build_char2idx()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.scaling == "strong":
        assert args.batch % world == 0, "--scaling strong: the global batch must divide over the ranks"
        args.batch //= world
    cfg = {"workload": "stcnn+bi%s%d+ctc train step, synthetic u8 clips (B,75,100,50,3), B=%d clips/GPU, L in [10,30]"
                       % (args.rnn.lower(), args.hidden, args.batch),
           "clips_per_gpu": args.batch, "clip_shape": [T_FRAMES, H_FRAME, W_FRAME, 3], "optimizer": "adam lr=1e-4, clip 50",
           "parallelism": "dp%d" % world, "l2": "inputs larger than L2 (288 MB u8 clips per step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        value, dt, cores, n_clips = run_cpu(args, char2idx, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "frames/sec end-to-end (3Dconv+BiGRU+CTC train step)", "value": value,
                "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "sample_clips_per_step": n_clips,
                "note": "CPU arm = oracle/train_step.py (fp32 port of the same step; the reference itself has no conv "
                        "front-end) on a bounded sample of the workload: %d clips per step instead of %d" % (n_clips, args.batch),
                "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d clips/step (same shapes), %d steps" % (n_clips, args.steps)},
                "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    from lipreading_b200 import conv_frontend, dist as ldist, functional as LF, native, trainer
    configure_throughput_path()
    from lipreading_b200.model import VideoEncoder
    rank, local_rank, world = ldist.init()
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    native.lib()
    pk = peaks()
    torch.manual_seed(SEED)
    enc = VideoEncoder(1728, args.hidden, frame_processing="conv3d", rnn_type=args.rnn, bidirectional=True,
                       enable_ctc=True, vocab_size=len(char2idx), char2idx=char2idx, device=dev).to(dev)
    opt = torch.optim.Adam(enc.parameters(), lr=1e-4, fused=True)
    reducer = ldist.GradAllReducer(world) if world > 1 else None
    host = [synth_batch(args.batch, SEED + 17 * rank + i, char2idx) for i in range(2)]
    host = [tuple(t.pin_memory() for t in b) for b in host]
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(loader_fn, steps, on_step=None):
        trainer.train_ctc(enc, loader_fn(steps), opt, dev, char2idx, grad_norm=50, dist=reducer, on_step=on_step)

    def resident_loader(n):
        return [resident[i % 2] for i in range(n)]

    # ---- device-resident timing ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # nvidia-smi needs ~0.5 s to come up: start before the warm-up
    run(resident_loader, args.warmup)
    barrier()
    if rank == 0:
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 3.0:
            time.sleep(0.05)
        sampler.rows.clear()                # keep only samples taken during the timed region
    conv_frontend.KERNEL_TIMING = []
    n0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run(resident_loader, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = native.launch_count() - n0
    ktimes = conv_frontend.KERNEL_TIMING
    conv_frontend.KERNEL_TIMING = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t)
    value = world * args.batch * T_FRAMES * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host batches, H2D inside the timed region, loss read back each step ----
    from lipreading_b200.data import DevicePrefetcher
    trace = [] if os.environ.get("LR_E2E_TRACE") else None
    copy_trace = [] if trace is not None else None

    def host_loader(n):
        pf = DevicePrefetcher([host[i % 2] for i in range(n)], dev)
        pf.trace = copy_trace
        return pf
    # every step's loss is read back to the host inside the timed region, through a pinned staging
    # slot and an event (the value is collected two steps later, so the read does not stall the queue)
    losses, pending = [], []
    slots = torch.zeros(args.steps + 4, dtype=torch.float32).pin_memory()

    LAG = 2         # the loss of step i is collected while step i + 2 is being enqueued

    def read_back(loss):
        i = len(pending)
        t_in = time.perf_counter()
        slots[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        # blocking event: the waiting host thread sleeps instead of spinning (eight ranks share the host's cores with
        # their autograd and NCCL threads)
        ev = torch.cuda.Event(enable_timing=trace is not None, blocking=True)
        ev.record()
        pending.append(ev)
        if i >= LAG:
            pending[i - LAG].synchronize()
            losses.append(float(slots[i - LAG]))
        if trace is not None:
            trace.append((t_in, time.perf_counter()))
    run(host_loader, 3)
    pending.clear()
    if trace is not None:
        del trace[:], copy_trace[:]
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    run(host_loader, args.steps, on_step=read_back)
    for j in range(max(0, len(pending) - LAG), len(pending)):       # the last LAG losses: still inside the timed region
        pending[j].synchronize()
        losses.append(float(slots[j]))
    t1.record()
    barrier()
    if trace is not None:
        # per step: host time from one read_back entry to the next (enqueue of a whole step), time blocked in the
        # read-back synchronize, GPU completion time of the step and duration of each H2D copy (ms)
        torch.cuda.synchronize()
        t_host0 = trace[0][0]
        rec = {"rank": rank, "host_enter_ms": [(a - t_host0) * 1e3 for a, _ in trace],
               "host_blocked_ms": [(b - a) * 1e3 for a, b in trace],
               "gpu_step_end_ms": [t0.elapsed_time(e) for e in pending],
               "copy_ms": [a.elapsed_time(b) for a, b in copy_trace],
               "copy_start_ms": [t0.elapsed_time(a) for a, _ in copy_trace],
               "total_ms": t0.elapsed_time(t1)}
        with open(os.path.join(ROOT, "gpurun_out", "e2e_trace_rank%d.json" % rank), "w") as fh:
            json.dump(rec, fh)
    te = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * args.batch * T_FRAMES * args.steps / (float(te) * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- BASELINE config 5: the inference stream clips -> characters on every rank at once -----------------------
    # (conv front-end -> BiGRU -> greedy CTC on the device, token ids only come back; max over ranks, whole-job rate)
    inference = None
    try:
        from lipreading_b200.infer import Recognizer
        rec = Recognizer(enc, char2idx)
        clips_d, lens_d = resident[0][0], resident[0][1]
        for _ in range(2):
            rec.tokens(clips_d, lens_d)
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        n_inf = 5
        for i in range(n_inf):
            tok, tok_n = rec.tokens(resident[i % 2][0], resident[i % 2][1])
        tok_n = tok_n.cpu()                                   # the result reaches the host inside the timed region
        i1.record()
        barrier()
        ti = torch.tensor([i0.elapsed_time(i1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ti, op=torch.distributed.ReduceOp.MAX)
        inference = {"value": world * args.batch * T_FRAMES * n_inf / (float(ti) * 1e-3), "unit": "frames/s",
                     "ms_per_batch": float(ti) / n_inf, "n_gpus": world,
                     "what": "u8 clips resident -> token ids (greedy CTC), %d batches of %d clips per GPU" % (n_inf, args.batch)}
        enc.train()
    except Exception as e:
        inference = {"error": repr(e)}

    # ---- BASELINE config 5, from RAW frames: every rank runs the whole vision + sequence chain on its own frames ----
    # (720p frames + boxes -> landmarks via the position-map CNN -> mouth crops -> conv front-end -> BiGRU -> greedy CTC;
    #  max over ranks, whole-job rate).  Every rank reaches the all-reduce whether or not its own run succeeded.
    frame_stream = None
    if not args.no_kernels:
        fs_local, fs_err = None, None
        try:
            fs_local = frame_stream_block(dev, enc, char2idx)
        except Exception as e:
            fs_err = repr(e)
        tf = torch.tensor([fs_local["ms"] if fs_local is not None else float("inf")], device=dev)
        if world > 1:
            torch.distributed.all_reduce(tf, op=torch.distributed.ReduceOp.MAX)
        if fs_local is not None and float(tf) != float("inf"):
            n_fr = fs_local["frames"]
            frame_stream = {"value": world * n_fr / (float(tf) * 1e-3), "unit": "frames/s", "ms": float(tf), "n_gpus": world,
                            "what": fs_local["what"] + " per GPU"}
        else:
            frame_stream = {"error": fs_err or "another rank failed"}
    if world > 1:
        torch.distributed.barrier()
        if rank != 0:
            torch.distributed.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernels: ALL EIGHT tensor-core conv launches of a step ------------
    # (conv1-3 fwd, conv3/conv2 dgrad: conv3d_tcgen05_kernel; conv1-3 wgrad: conv3d_wgrad_tcgen05_kernel), each timed
    # with CUDA events on the launching stream inside the timed region; achieved = algorithmic FLOPs / time.
    torch.cuda.synchronize()
    per = {}
    for tag, a, b, fl in ktimes:
        d = per.setdefault(tag, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b) * 1e-3
        d[1] += fl
        d[2] += 1
    tot_s = sum(v[0] for v in per.values())
    tot_f = sum(v[1] for v in per.values())
    # the timed region is a fraction of a second at boost clocks: the burst figure is the honest denominator
    peak = pk["bf16_tflops"]
    # DRAM traffic per launch from the committed `ncu --set full` capture of the same command (profiles/), if any
    traffic, traffic_src = None, None
    for name in ("r2_ncu_full_step_b256.json", "r1_ncu_full_step_b256_v4.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                cap = json.load(fh)
            if args.batch == 256:
                traffic = {k: (v["dram_read_MB"] + v["dram_write_MB"]) * 1e6 for k, v in cap.items()
                           if isinstance(v, dict) and "dram_read_MB" in v}
                traffic_src = "dram bytes read+written per launch, profiles/" + name + " (ncu --set full; not measured in this run)"
            break
        except Exception:
            continue
    roofline = {"kernel": "conv3d_tcgen05_kernel + conv3d_wgrad_tcgen05_kernel (8 launches/step: conv1-3 fwd, conv3/conv2 "
                          "dgrad, conv1-3 wgrad)",
                "bound": "tensor", "achieved": tot_f / tot_s / 1e12 if tot_s else None, "peak": peak,
                "unit": "TFLOP/s", "frac": (tot_f / tot_s / 1e12 / peak) if tot_s else None,
                "frac_of_sustained": (tot_f / tot_s / 1e12 / pk["bf16_tflops_sustained"]) if tot_s else None,
                "traffic": (traffic or {}).get("conv2.fwd"), "traffic_per_launch": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"] + " (burst bf16 cuBLAS; sustained %.1f)" % pk["bf16_tflops_sustained"],
                "share_of_step": tot_s / (ms * 1e-3),
                "whole_step_tflops": tot_f / args.steps / (ms / args.steps * 1e-3) / 1e12 if tot_s else None,
                "per_launch": {k: {"ms": v[0] / v[2] * 1e3, "tflops": v[1] / v[0] / 1e12,
                                   "frac": v[1] / v[0] / 1e12 / peak} for k, v in per.items()}}
    line = {"metric": "frames/sec end-to-end (3Dconv+BiGRU+CTC train step)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfg, "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "final_loss": losses[-1] if losses else None},
            "roofline": roofline, "inference_stream": inference, "frame_stream": frame_stream}
    if world == 1 and not args.no_cpu_baseline:
        v, dt, cores, n_clips = run_cpu(args, char2idx, 2, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d clips/step (same shapes; the GPU arm runs %d), 1 warm-up + 2 timed steps"
                                          % (n_clips, args.batch)}
    if world == 1 and not args.no_parity:
        try:
            line["parity"] = parity_block(dev)
        except Exception as e:
            line["parity_error"] = repr(e)
    if world == 1 and not args.no_ref_shape:
        try:
            line["ref_shape"] = ref_shape_block(dev, char2idx)
        except Exception as e:
            line["ref_shape_error"] = repr(e)
    if world == 1 and not args.no_kernels:
        try:
            line["kernels"] = kernel_rooflines(dev, pk, char2idx)
        except Exception as e:           # micro-benches must never lose the headline
            line["kernels_error"] = repr(e)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
