"""Measure cycles per tcgen05.mma for the shapes the conv kernels can choose from (run on the GPU box):
    python tools/umma_table.py > gpurun_out/umma_table.txt
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import native  # noqa: E402

L = native.lib()
f = ctypes.CDLL(native.DIAG_LIB_PATH).lr_umma_microbench
f.restype = ctypes.c_longlong
f.argtypes = [ctypes.c_int] * 10 + [ctypes.c_void_p]
torch.zeros(1, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 2000
print("M N rowA rowB majA majB n_acc a_tiles -> cycles/MMA  MAC/clk")
for (M, N, ra, rb, ma, mb, nacc, at) in [
    (128, 32, 32, 32, 0, 0, 8, 8), (128, 32, 32, 32, 0, 0, 1, 1), (128, 64, 64, 64, 0, 0, 4, 4), (128, 64, 64, 64, 0, 0, 1, 1),
    (128, 96, 128, 128, 0, 0, 2, 2), (128, 96, 128, 128, 0, 0, 1, 1), (128, 128, 128, 128, 0, 0, 2, 2),
    (128, 256, 128, 128, 0, 0, 2, 1), (128, 256, 128, 128, 0, 0, 1, 1),
    (64, 128, 64, 64, 0, 0, 2, 1), (64, 128, 64, 64, 0, 0, 1, 1), (64, 128, 32, 32, 0, 0, 2, 1), (64, 128, 128, 128, 0, 0, 2, 1),
    (64, 256, 64, 64, 0, 0, 2, 1), (64, 256, 128, 128, 0, 0, 1, 1), (128, 128, 64, 64, 0, 0, 2, 1), (128, 128, 32, 32, 0, 0, 2, 1),
    (64, 32, 128, 64, 1, 1, 8, 1), (64, 160, 128, 64, 1, 1, 3, 1), (64, 96, 128, 64, 1, 1, 5, 1), (64, 16, 64, 32, 1, 1, 8, 1),
    (64, 48, 64, 32, 1, 1, 8, 1), (128, 96, 128, 64, 1, 1, 2, 1), (128, 192, 128, 64, 1, 1, 2, 1), (128, 256, 128, 128, 1, 1, 2, 1),
]:
    c = f(M, N, ra, rb, ma, mb, nacc, at, iters, 0, st)
    per = c / iters if c > 0 else float("nan")
    print(M, N, ra, rb, ma, mb, nacc, at, "->", "%.1f" % per, "%.0f" % (M * N * 16 / per if c > 0 else 0))

print("row-shifted A operand (the conv kernels' tap shifts): M N rowA shift_rows -> cycles/MMA")
for (M, N, ra, rb) in [(128, 64, 64, 64), (128, 96, 128, 128), (128, 32, 32, 32), (128, 32, 128, 128)]:
    for sh in (0, 1, 2, 3, 4, 8, 16, 17):
        c = f(M, N, ra, rb, 0, 0, 4, 1, iters, sh, st)
        print(M, N, ra, sh, "->", "%.1f" % (c / iters))
