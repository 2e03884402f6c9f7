"""the position-map CNN (src/models/face/prnet.py:211-314).

What pins it: the reference ships the checkpoint INDEX (names, shapes, offsets of every variable) but not the data
shard, so the architecture is held to that index; TF-slim's 'SAME' conv / transposed-conv semantics are held to their
definitions (explicit sum / input-gradient of the forward conv).  Values are unpinned (no weights exist)."""
import os
import shutil

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lipreading_b200 import prnet as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = os.path.join(GOLD, "prnet_256_256_resfcn256_weight.index")


def test_architecture_matches_reference_checkpoint_index():
    idx = P.read_tf_index(INDEX)
    model_vars = {n: e for n, e in idx.items() if "Adam" not in n and "_power" not in n}
    assert len(idx) == 545 and len(model_vars) == 245
    net = P.ResFcn256()
    shapes = net.tf_shapes()
    assert set(shapes) == set(model_vars)                       # every variable, no extras
    for n, shp in shapes.items():
        assert shp == model_vars[n]["shape"], n
        assert model_vars[n]["size"] == 4 * int(np.prod(shp)) and model_vars[n]["dtype"] == 1
    assert sum(p.numel() for p in net.parameters()) == 13353618          # 13.3 M parameters


def test_restore_reads_the_data_shard_without_tensorflow(tmp_path):
    idx = P.read_tf_index(INDEX)
    prefix = str(tmp_path / "256_256_resfcn256_weight")
    shutil.copyfile(INDEX, prefix + ".index")
    with pytest.raises(FileNotFoundError):
        P.load_tf_checkpoint(prefix, {"resfcn256/Conv/weights"})
    end = max(e["offset"] + e["size"] for e in idx.values())
    rng = np.random.default_rng(0)
    blob = (rng.standard_normal(end // 4 + 1).astype(np.float32) * 0.05).tobytes()[:end]
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(blob)
    pred = P.PosPrediction(device="cpu")
    pred.restore(prefix)
    net = pred.network
    for name, t, kind in net.tf_variables():
        e = idx[name]
        ref = np.frombuffer(blob[e["offset"]:e["offset"] + e["size"]], dtype=np.float32).reshape(e["shape"])
        got = t.detach().numpy()
        if kind != "vec":
            got = got.transpose(2, 3, 1, 0)            # torch (a,b,kh,kw) -> TF (kh,kw,b,a)
        assert np.array_equal(got, ref), name


def _same_conv_by_definition(x, w, stride):
    """TF 'SAME' correlation written out: out[y,x] = sum_ij w[i,j] * in[y*s + i - pad_before]."""
    k = w.shape[-1]
    n, c, H, W = x.shape
    out_h = -(-H // stride)
    total = max((out_h - 1) * stride + k - H, 0)
    before = total // 2
    xp = torch.zeros(n, c, H + total, W + total, dtype=x.dtype)
    xp[:, :, before:before + H, before:before + W] = x
    out = torch.zeros(n, w.shape[0], out_h, out_h, dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            patch = xp[:, :, i:i + (out_h - 1) * stride + 1:stride, j:j + (out_h - 1) * stride + 1:stride]
            out += torch.einsum("nchw,oc->nohw", patch, w[:, :, i, j])
    return out


@pytest.mark.parametrize("stride", [1, 2])
def test_same_padding_conv_matches_definition(stride):
    torch.manual_seed(0)
    m = P._Conv(3, 5, 4, stride, norm=False, act=False).double()
    x = torch.randn(2, 3, 8, 8, dtype=torch.float64)
    ref = _same_conv_by_definition(x, m.conv.weight.detach(), stride)
    assert torch.allclose(m(x), ref, atol=1e-12)


@pytest.mark.parametrize("stride", [1, 2])
def test_transposed_conv_is_the_input_gradient_of_the_same_conv(stride):
    """tf.nn.conv2d_transpose(y, W(kh,kw,out,in)) := d/dx <conv2d_SAME(x, W), y>."""
    torch.manual_seed(1)
    cin, cout, hin = 4, 3, 6                       # transposed conv: cin -> cout, hin -> hin*stride
    m = P._Deconv(cin, cout, stride).double()
    m.bn = torch.nn.Identity()
    y = torch.randn(2, cin, hin, hin, dtype=torch.float64)
    got = F.relu(m.conv(y)[:, :, 1:-2, 1:-2]) if stride == 1 else F.relu(m.conv(y))
    assert torch.allclose(got, m(y), atol=1e-12)       # (stride 1 runs as the equivalent flipped-kernel forward conv)
    # forward conv it is the gradient of: x (cout channels, hin*stride) -> (cin channels, hin), kernel (cin,cout,4,4)
    x = torch.zeros(2, cout, hin * stride, hin * stride, dtype=torch.float64, requires_grad=True)
    fwd = _same_conv_by_definition(x, m.conv.weight.detach(), stride)
    (fwd * y).sum().backward()
    assert torch.allclose(F.relu(x.grad), got, atol=1e-12)


def test_forward_shape_and_range_cpu():
    torch.manual_seed(0)
    pred = P.PosPrediction(device="cpu")
    img = np.random.default_rng(0).random((1, 256, 256, 3), dtype=np.float32)
    pos = pred.predict_batch(img)
    assert pos.shape == (1, 256, 256, 3) and pos.dtype == np.float32
    assert 0.0 <= pos.min() and pos.max() <= 256 * 1.1
    assert np.array_equal(pred.predict(img[0]), pos[0])


@pytest.mark.gpu
def test_posmap_cnn_on_device_feeds_the_landmark_kernels(native_lib, cuda):
    """frames -> lr_rect_geometry -> lr_warp256 -> PosPrediction -> lr_posmap_gather; the torch engine on the device
    agrees with the same weights on the CPU (fp32: 2e-3 of MaxPos after 28 layers); the tcgen05 engine (bf16 volumes,
    the default on CUDA) is held to the same CPU result in tests/test_gpu_tapgemm.py."""
    from lipreading_b200.face import PRN
    torch.manual_seed(0)
    pred = P.PosPrediction(device=cuda, engine="torch")
    with torch.no_grad():
        for m in pred.network.modules():                 # non-trivial batch-norm statistics
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.05)
                m.running_var.uniform_(0.5, 1.5)
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (3, 240, 320, 3), dtype=torch.uint8, generator=g).to(cuda)
    rects = torch.tensor([[100, 200, 60, 160], [80, 220, 40, 180], [120, 210, 70, 165]], dtype=torch.int32)
    kpt = np.loadtxt(os.path.join(GOLD, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(GOLD, "face_ind.npy"))
    prn = PRN(predict_batch=pred.predict_batch, uv_kpt_ind=kpt, face_ind=face, device=cuda)
    (lmk, vtx), geom = prn.process_batch(frames, rects, with_vertices=True)
    assert lmk.shape == (3, 68, 3) and vtx.shape == (3, 43867, 3) and torch.isfinite(lmk).all()
    cropped, _ = prn.crop_batch(frames, rects)
    pos_dev = pred.predict_batch(cropped)
    cpu = P.PosPrediction(device="cpu")
    cpu.network.load_state_dict({k: v.cpu() for k, v in pred.network.state_dict().items()})
    pos_cpu = torch.from_numpy(cpu.predict_batch(cropped.cpu().numpy()))
    assert float((pos_dev.cpu() - pos_cpu).abs().max()) < 2e-3 * 281.6
