#!/usr/bin/env python
"""Recurrent layer timing: persistent path (cluster / grid kernels) vs the fp32 per-step kernels.
usage: python tools/rnn_time.py [LSTM|GRU] [H] [B] [T]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lipreading_b200 import functional as LF, native  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "LSTM"
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 75
    I, D, G = 204, 2, native.RNN_GATES[mode]
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, I, generator=g).to(dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    ws = []
    for _ in range(D):
        ws += [(torch.randn(G * H, I, generator=g) / I ** 0.5).to(dev).requires_grad_(True),
               (torch.randn(G * H, H, generator=g) / H ** 0.5).to(dev).requires_grad_(True),
               torch.zeros(G * H, device=dev, requires_grad=True), torch.zeros(G * H, device=dev, requires_grad=True)]
    out = {"mode": mode, "H": H, "B": B, "T": T, "cluster_supported": int(native.lib().lr_rnn_cluster_supported(native.RNN_MODES[mode], H)),
           "grid_supported": int(native.lib().lr_rnn_grid_supported(native.RNN_MODES[mode], H, D))}
    up = torch.randn(B, T, D * H, generator=g).to(dev)
    for name, persistent, dt in (("per_step_fp32", False, torch.float32), ("persistent_bf16", True, torch.bfloat16)):
        LF.RNN_CLUSTER, LF.GEMM_DTYPE = persistent, dt

        def fwd():
            with torch.no_grad():
                return LF.rnn_layer(x, lens, mode, ws)

        def fwd_bwd():
            for w in ws:
                w.grad = None
            res = LF.rnn_layer(x, lens, mode, ws)
            (res[0] * up).sum().backward()
        f = timeit(fwd)
        fb = timeit(fwd_bwd)
        out[name] = {"fwd_ms": f, "fwd_bwd_ms": fb, "fwd_us_per_step": f / T * 1e3, "bwd_us_per_step": (fb - f) / T * 1e3}
    LF.RNN_CLUSTER, LF.GEMM_DTYPE = False, torch.float32
    print(json.dumps(out))


if __name__ == "__main__":
    main()
