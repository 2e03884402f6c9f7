"""Oracle for the conv front-end: torch.nn.functional.conv3d in fp32.
North-star extension — there is no reference code; the spec is DESIGN.md §N1.  TEST INFRASTRUCTURE.

`quantize=True` rounds the operands (input/255, weights, and each layer's output activation) to
bf16 exactly where the CUDA path stores bf16, so the remaining difference is fp32 accumulation order.
"""
import torch
import torch.nn.functional as F


def _q(t, on):
    return t.to(torch.bfloat16).to(torch.float32) if on else t


def stcnn_forward(clip_u8, params, quantize=True):
    """clip (B,T,H,W,3) uint8; params dict conv{1,2,3}.{weight,bias} (nn.Conv3d layout) ->
    features (B,T,96*h*w) in (h,w,c) order, plus the per-layer pooled activations."""
    x = _q(clip_u8.to(torch.float32) / 255.0, quantize).permute(0, 4, 1, 2, 3)      # B,C,T,H,W
    acts = []
    cfg = (("conv1", (1, 2, 2), (1, 2, 2)), ("conv2", (1, 1, 1), (1, 2, 2)), ("conv3", (1, 1, 1), (1, 1, 1)))
    for name, stride, pad in cfg:
        w = _q(params[name + ".weight"].float(), quantize)
        b = params[name + ".bias"].float()
        x = F.conv3d(x, w, b, stride=stride, padding=pad)
        x = _q(F.relu(x), quantize)
        x = F.max_pool3d(x, (1, 2, 2))
        acts.append(x)
    B, C, T, h, w_ = x.shape
    feat = x.permute(0, 2, 3, 4, 1).reshape(B, T, h * w_ * C)
    return feat, acts
