"""CTC kernel timings on the GPU box: log-space warp kernel (select 2) vs linear-space first (select 3) vs CTA-per-clip
(select 1).   python tools/ctc_bench.py [B ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import functional as LF, native as N  # noqa: E402

dev = torch.device("cuda")
T, C = 75, 65
g = torch.Generator().manual_seed(123456)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [256, 4096, 16384]:
    lp = torch.randn(B, T, C, generator=g).log_softmax(-1).to(dev).requires_grad_(True)
    tg = torch.randint(5, 65, (B, 30), generator=g).to(dev).int()
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.randint(10, 31, (B,), generator=g).to(dev).int()
    res = {}
    for name, sel in (("cta", 1), ("log_warp", 2), ("linear_warp", 3)):
        LF.CTC_KERNEL = sel
        for _ in range(3):
            nll = LF.ctc_nll(lp, tg, il, tl)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nll = LF.ctc_nll(lp, tg, il, tl)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res[name] = (ts[len(ts) // 2], nll.detach().clone())
    LF.CTC_KERNEL = 0
    byts = B * 2 * T * C * 4
    print("B=%d " % B + "  ".join("%s %.1f us (%.0f GB/s)" % (k, v[0] * 1e3, byts / v[0] / 1e6) for k, v in res.items()) +
          "  max|nll_lin - nll_log| = %.2e" % float((res["linear_warp"][1] - res["log_warp"][1]).abs().max()))
