"""Time the weight-gradient kernel of each conv layer with and without ky-stacking (run on the GPU box).
    python tools/wgrad_time.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import conv_frontend as CF  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = 75
dev = torch.device("cuda")
bf = torch.bfloat16
LAYERS = [  # name, Cx, Cy, Gy, K, H, W, Wp, m_is_x, interior offset (t,y,x) of dY
    ("wgrad3", 64, 32, 3, (3, 3, 3), 12, 6, 8, 1, (1, 1, 1)),
    ("wgrad2", 32, 64, 1, (3, 5, 5), 25, 12, 16, 0, (1, 2, 2)),
    ("wgrad1", 16, 32, 1, (3, 3, 3), 50, 25, 32, 0, (0, 0, 0)),
]
for name, Cx, Cy, Gy, K, H, W, Wp, m_is_x, off in LAYERS:
    Hp = CF._plane_rows(H, K[1], Wp)
    x = torch.randn(B, T + 2, Hp, Wp, Cx, device=dev).to(bf)
    dy = torch.randn(Gy, B, T + 2, Hp, Wp, Cy, device=dev).to(bf)
    dy_off = (off[0] * Hp + off[1]) * Wp + off[2]
    flops = 2.0 * B * T * H * W * Cx * Cy * Gy * K[0] * K[1] * K[2]
    for mode, kw in (("per-tap", dict(stack_kx=False, stack_ky=False)), ("kx on N", dict(stack_kx=True, stack_ky=False)),
                     ("kx on N + ky on M=128", dict(stack_kx=True, stack_ky=True))):
        if m_is_x and mode != "per-tap":
            continue
        for _ in range(2):
            CF.conv3d_wgrad_native(x, dy, B, T, H, W, Hp, Wp, Cx, Cy, Gy, dy_off, K, m_is_x, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            CF.conv3d_wgrad_native(x, dy, B, T, H, W, Hp, Wp, Cx, Cy, Gy, dy_off, K, m_is_x, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%s B=%d %-24s %.3f ms  %.0f TFLOP/s (algorithmic)" % (name, B, mode, ms, flops / ms / 1e9), flush=True)
