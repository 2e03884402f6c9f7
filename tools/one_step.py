"""One forward+backward of the conv front-end at B clips (for ncu captures on the GPU box):
    ncu --set full --clock-control none -k regex:conv3d -c 8 -o gpurun_out/conv python tools/one_step.py 256"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import conv_frontend as CF  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda")
torch.manual_seed(1)
front = CF.ConvFrontEnd().to(dev)
clip = torch.randint(0, 256, (B, 75, 100, 50, 3), dtype=torch.uint8, device=dev)
for _ in range(reps):
    front(clip).sum().backward()
torch.cuda.synchronize()
print("done", B)
