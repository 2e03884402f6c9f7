"""proj + masked log-softmax forward: time and error vs float64 (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lipreading_b200 import functional as LF
dev = torch.device("cuda")
g = torch.Generator().manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
C = 65
for M, K in ((256 * 75, 512), (128 * 75, 1536), (1000, 250)):
    h = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(C, K, generator=g) / 22).to(dev)
    b = torch.randn(C, generator=g).to(dev) * 0.1
    mask = torch.ones(C); mask[1] = mask[2] = 0
    lm = (mask + 1e-45).log().to(dev)
    s = bench.time_cuda(lambda: LF.proj_masked_log_softmax(h, w, b, lm), flush=flush)
    byts = M * (K + C) * 4
    ref = torch.log_softmax(h.double() @ w.double().t() + b.double() + lm.double(), -1)
    err = float((LF.proj_masked_log_softmax(h, w, b, lm).double() - ref).abs().max())
    print("M=%d K=%d: %.4f ms  %.0f GB/s  frac %.3f  max err %.2e" % (M, K, s * 1e3, byts / s / 1e9, byts / s / 6555.5e9, err))
