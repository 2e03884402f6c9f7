import torch, sys
sys.path.insert(0, ".")
import bench
from lipreading_b200 import functional as LF
dev = torch.device("cuda")
g = torch.Generator().manual_seed(1)
n, H, W = 384, 720, 1280
frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rect in ([400, 700, 150, 450], [300, 500, 200, 400], [100, 900, 0, 700]):
    rects = torch.tensor([rect] * n, dtype=torch.int32, device=dev)
    rp, crop = LF.rect_geometry(rects, H, W)
    s = bench.time_cuda(lambda: LF.warp256(frames, crop), flush=flush)
    size = int(crop[0, 2]); byts = n * (size * size * 3 + 256 * 256 * 3 * 4)
    print("size", size, "ms %.4f" % (s * 1e3), "GB/s %.0f" % (byts / s / 1e9), "frac %.3f" % (byts / s / 6555.5e9))
