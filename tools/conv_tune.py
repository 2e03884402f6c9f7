"""Time each conv launch of a training step under different (J, TMEM sets) choices (run on the GPU box).
    python tools/conv_tune.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lipreading_b200 import conv_frontend as CF  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
front = CF.ConvFrontEnd().to(dev)
clip = torch.randint(0, 256, (B, 75, 100, 50, 3), dtype=torch.uint8, device=dev)
orig = CF.conv3d_native


def run(label, J_by_tag, sets):
    if sets:
        os.environ["LR_CONV_SETS"] = str(sets)
    else:
        os.environ.pop("LR_CONV_SETS", None)

    def wrapped(*a, **k):
        tag = k.get("tag", "")
        if tag in J_by_tag:
            k["J"] = J_by_tag[tag]
        return orig(*a, **k)
    CF.conv3d_native = wrapped
    for _ in range(2):
        front(clip).sum().backward()
    CF.KERNEL_TIMING = []
    for _ in range(3):
        front(clip).sum().backward()
    torch.cuda.synchronize()
    per = {}
    for tag, a, b, fl in CF.KERNEL_TIMING:
        per.setdefault(tag, []).append(a.elapsed_time(b))
    CF.KERNEL_TIMING = None
    print(label, "  ".join("%s %.3f" % (t, sum(v) / len(v)) for t, v in per.items()), flush=True)


run("default            ", {}, 0)
run("all sets=1 Jmax    ", {}, 1)
run("conv3/dgrad3 J=1 s2", {"conv3.fwd": 1, "conv3.dgrad": 1}, 2)
run("conv1 J=4          ", {"conv1.fwd": 4}, 0)
run("conv2 J=2          ", {"conv2.fwd": 2, "conv2.dgrad": 2}, 0)
run("conv2 J=3          ", {"conv2.fwd": 3, "conv2.dgrad": 3}, 0)
os.environ.pop("LR_CONV_SETS", None)
