#!/usr/bin/env python
"""Position-map CNN throughput: tcgen05 tap-GEMM plan vs the torch (cuDNN bf16, channels-last) module, plus the time
of every launch of the plan (CUDA events).   python tools/prnet_time.py [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lipreading_b200 import native, prnet as P  # noqa: E402


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
    return ts[len(ts) // 2] * 1e-3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.rand(B, 256, 256, 3, device=dev)
    out = {"batch": B}
    pred = P.PosPrediction(device=dev)
    plan = pred.plan(B)
    s = timeit(lambda: plan.run(x))
    out["tcgen05"] = {"frames_per_s": B / s, "ms": s * 1e3, "issued_tflops": plan.flops / s / 1e12,
                      "model_tflops": B * 2 * 4.13e9 / s / 1e12}
    # per launch
    import ctypes
    lib = native.lib()
    st = native.stream()
    rows = []
    for i, (spec, d) in enumerate(zip(plan.specs, plan._descs)):
        t = timeit(lambda: native.check(lib.lr_tapgemm(ctypes.byref(d), st), "lr_tapgemm"), iters=3, warm=1)
        a = spec["a"]
        fl = 2 * a.B * spec["valid"][2] * spec["valid"][3] * spec["Kg"] * spec["n_groups"] * spec["n_phases"] * spec["Cout_pad"]
        o = spec["out"]
        byts = a.rows * a.C * 2 + (o.rows * o.C * 2 if hasattr(o, "rows") else o.numel() * 4)
        rows.append({"i": i, "C": a.C, "HW": a.H, "s2d": a.s2d, "Kg": spec["Kg"], "groups": spec["n_groups"],
                     "phases": spec["n_phases"], "Cout": spec["Cout_pad"], "mode": spec["mode"], "us": t * 1e6,
                     "tflops": fl / t / 1e12, "GBps": byts / t / 1e9})
    out["launches"] = rows
    out["sum_launch_ms"] = sum(r["us"] for r in rows) / 1e3
    try:
        ref = P.PosPrediction(device=dev, dtype=torch.bfloat16, engine="torch")
        s = timeit(lambda: ref.predict_batch(x), iters=3)
        out["torch_cudnn_bf16"] = {"frames_per_s": B / s, "ms": s * 1e3}
    except Exception as e:
        out["torch_cudnn_bf16"] = {"error": repr(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
