"""What one issuing thread sustains when the MMA descriptors are advanced inside the loop (run on the GPU box)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import native  # noqa: E402

native.lib()
f = ctypes.CDLL(native.DIAG_LIB_PATH).lr_umma_issue_bench
f.restype = ctypes.c_longlong
f.argtypes = [ctypes.c_int] * 8 + [ctypes.c_void_p]
torch.zeros(1, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
tiles = 400
print("N KS n_ops threads variant(0 loop-carried adds, 1 constant) -> cycles/MMA")
for N in (32, 64, 96, 192):
    for KS in (1, 2):
        for threads in (128, 288):
            for variant in (0, 1):
                n_ops = 6
                c = f(N, KS, tiles, n_ops, 13312, 32 if N <= 192 else 0, threads, variant, st)
                print(N, KS, n_ops, threads, variant, "-> %.1f" % (c / (tiles * n_ops * KS)))
