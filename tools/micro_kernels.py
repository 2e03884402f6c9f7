"""Launch each HBM-bound kernel of the path a few times at the bench shapes (for ncu captures on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:"warp256|posmap|ctc|proj_logsoftmax_fwd|mouth_crop" \
        -o gpurun_out/micro python tools/micro_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from lipreading_b200 import functional as LF  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(123456)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, T, C = 4096, 75, 65
lp = torch.randn(B, T, C, generator=g).log_softmax(-1).to(dev).requires_grad_(True)
tg = torch.randint(5, 65, (B, 30), generator=g).to(dev).int()
il = torch.full((B,), T, dtype=torch.int32, device=dev)
tl = torch.randint(10, 31, (B,), generator=g).to(dev).int()
M, K = 256 * 75, 512
h = torch.randn(M, K, generator=g).to(dev)
w = (torch.randn(C, K, generator=g) / 22).to(dev)
b = torch.zeros(C, device=dev)
mask = torch.ones(C)
mask[1] = mask[2] = 0
lm = (mask + 1e-45).log().to(dev)
n, H, W = 384, 720, 1280
frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
rects = torch.tensor([[400, 700, 150, 450]] * n, dtype=torch.int32, device=dev)
rp, crop = LF.rect_geometry(rects, H, W)
gold = os.path.join(ROOT, "tests", "golden")
uv = np.loadtxt(os.path.join(gold, "uv_kpt_ind.txt")).astype(np.int64)
kidx = torch.from_numpy((uv[1] * 256 + uv[0]).astype(np.int32)).to(dev)
fidx = torch.from_numpy(np.load(os.path.join(gold, "face_ind.npy"))).to(dev)
pos = (torch.rand(n, 256, 256, 3, generator=g) * 281.6).to(dev)
vv, uu = torch.meshgrid(torch.arange(256.0), torch.arange(256.0), indexing="ij")
pos_id = (torch.stack([uu, vv, torch.full_like(uu, 40.0)], -1)[None] + torch.randn(8, 256, 256, 3, generator=g)).to(dev)
lmk_id = LF.posmap_gather(pos_id, crop[:8], rp[:8], kidx).repeat(n // 8, 1, 1)     # frontal-face landmarks: mouth ROI ~ 82x163 px
lp256 = lp[:256].detach().clone().requires_grad_(True)
for _ in range(reps):
    LF.warp256(frames, crop)
    LF.posmap_gather(pos, crop, rp, kidx, fidx)
    LF.ctc_nll(lp, tg, il, tl)                       # linear-space warp kernel (+ the log-space pass on flagged clips)
    LF.ctc_nll(lp256, tg[:256], il[:256], tl[:256])   # the same at the training batch size
    LF.proj_masked_log_softmax(h, w, b, lm)
    LF.mouth_crop(frames, lmk_id, rp, 100, 50)
torch.cuda.synchronize()
print("done")
