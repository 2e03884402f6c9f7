"""Where do the conv kernel's warp roles wait?  (run on the GPU box)  python tools/conv_waits.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import native as N  # noqa: E402
N.use_diag_lib()          # the hooks live in liblr_b200_diag.so only
from lipreading_b200 import conv_frontend as CF  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SKIPS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
if len(sys.argv) > 3:
    CF.SWAP = int(sys.argv[3])
dev = torch.device("cuda")
front = CF.ConvFrontEnd().to(dev)
clip = torch.randint(0, 256, (B, 75, 100, 50, 3), dtype=torch.uint8, device=dev)
names = ["prod.a_empty", "prod.w_empty", "mma.acc_empty", "mma.a_full", "mma.w_full", "epi.acc_full", "-", "mma.total"]
feat = front(clip)
feat.sum().backward()
torch.cuda.synchronize()
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
CF.KERNEL_TIMING = []
orig = CF.conv3d_native


def wrapped(*a, **k):
    dbg.zero_()
    N.lib().lr_conv3d_set_debug(N.ptr(dbg))
    orig(*a, **k)
    torch.cuda.synchronize()
    N.lib().lr_conv3d_set_debug(None)
    d = dbg.view(148, 8).double()
    tot = d[:, 7].mean().item()
    print("%-12s total %.0f kcyc | " % (k.get("tag", "?"), tot / 1e3) +
          "  ".join("%s %.0f%%" % (n, 100 * d[:, i].mean().item() / tot) for i, n in enumerate(names) if n not in ("-", "mma.total")))


CF.conv3d_native = wrapped
import ctypes
skipf = ctypes.CDLL(N.DIAG_LIB_PATH).lr_conv3d_set_debug_skip
for sk in SKIPS:
    print("== skip mask %d (1 epilogue, 2 weight reloads, 4 input reloads), orientation %s" % (sk, CF.SWAP))
    skipf(sk)
    feat = front(clip)
    feat.sum().backward()
    torch.cuda.synchronize()
skipf(0)
