#!/usr/bin/env python
"""Run selected launches of the position-map CNN plan (for `ncu -k regex:tapgemm`).  usage: prnet_ncu.py B i j k ..."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lipreading_b200 import native, prnet as P  # noqa: E402

B = int(sys.argv[1])
sel = [int(a) for a in sys.argv[2:]]
dev = torch.device("cuda:0")
pred = P.PosPrediction(device=dev)
plan = pred.plan(B)
x = torch.rand(B, 256, 256, 3, device=dev)
plan.run(x)
torch.cuda.synchronize()
lib, st = native.lib(), native.stream()
for i in sel:
    native.check(lib.lr_tapgemm(ctypes.byref(plan._descs[i]), st), "lr_tapgemm")
torch.cuda.synchronize()
