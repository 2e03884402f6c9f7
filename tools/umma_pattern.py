"""Cost of back-to-back tcgen05.mma with mixed N / overlapping accumulator windows (run on the GPU box):
    python tools/umma_pattern.py > gpurun_out/umma_pattern.txt
Each line: the 8-MMA trip (N@column ...) -> average cycles per MMA, and the sum of the stand-alone costs
max(N/2, 32 + N/4) for comparison."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import native  # noqa: E402

native.lib()
f = ctypes.CDLL(native.DIAG_LIB_PATH).lr_umma_pattern_bench
f.restype = ctypes.c_longlong
I8 = ctypes.c_int * 8
f.argtypes = [I8, I8, I8, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
torch.zeros(1, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 4000


def run(name, n, d, b=None, commit_every=0):
    b = b or [0] * 8
    c = f(I8(*n), I8(*d), I8(*b), iters, commit_every, st)
    model = sum(max(x / 2, 32 + x / 4) for x in n) / 8
    print("%-46s %s -> %.1f cyc/MMA (stand-alone model %.1f)" % (
        name, " ".join("%d@%d" % (x, y) for x, y in zip(n, d)), c / iters, model))


run("uniform N=192 same D", [192] * 8, [0] * 8)
run("uniform N=192, two disjoint D", [192] * 8, [0, 256] * 4)
run("uniform N=192, sliding D by 64", [192] * 8, [0, 64, 128, 192, 256, 0, 64, 128])
run("uniform N=192, sliding D, pairs (KS=2)", [192] * 8, [0, 0, 64, 64, 128, 128, 192, 192])
run("uniform N=64 sliding D by 64", [64] * 8, [0, 64, 128, 192, 256, 320, 384, 448])
run("uniform N=64 same D", [64] * 8, [0] * 8)
run("mixed 64/128/192 same D", [64, 128, 192, 192, 128, 64, 64, 128], [0] * 8)
run("mixed 64/128/192 disjoint D", [64, 128, 192, 64, 128, 192, 64, 128], [0, 64, 192, 384, 0, 128, 320, 384])
run("conv2 J=4 op order (KS=1)", [64, 128, 192, 192, 128, 64, 64, 128], [0, 0, 0, 64, 128, 192, 256, 256], [16, 8, 0, 0, 0, 0, 16, 8])
run("conv2 J=4 op order, pairs (KS=2)", [64, 64, 128, 128, 192, 192, 192, 192], [0, 0, 0, 0, 0, 0, 64, 64], [16, 16, 8, 8, 0, 0, 0, 0])
run("uniform N=96 sliding D by 32 (conv1)", [96] * 8, [0, 32, 64, 96, 128, 160, 192, 224])
run("uniform N=96 same D", [96] * 8, [0] * 8)
run("N=192 far-apart D (0 / 320)", [192] * 8, [0, 320] * 4)
run("N=128 sliding D by 64", [128] * 8, [0, 64, 128, 192, 256, 320, 384, 0])
run("N=128 sliding D by 128 (disjoint)", [128] * 8, [0, 128, 256, 384] * 2)
run("N=64 alternating 2 D", [64] * 8, [0, 64] * 4)
run("N=64 4 D round robin", [64] * 8, [0, 64, 128, 192] * 2)

print("tcgen05.commit every k MMAs (the per-stage release of the conv kernels):")
for ce in (0, 8, 16, 32, 64):
    run("N=64 4 D round robin, commit/%d" % ce, [64] * 8, [0, 64, 128, 192] * 2, commit_every=ce)
for ce in (0, 8, 16, 32):
    run("conv2 J=4 pairs, commit/%d" % ce, [64, 64, 128, 128, 192, 192, 192, 192], [0, 0, 0, 0, 0, 0, 64, 64], [16, 16, 8, 8, 0, 0, 0, 0], commit_every=ce)
