"""Position-map CNN of the reference's vision path: `resfcn256` + `PosPrediction`
(src/models/face/prnet.py:211-314), restated for the GPU, plus a TensorFlow-free reader for the checkpoint the
reference restores (`Data/net-data/256_256_resfcn256_weight`, prnet.py:61-69).

Status.  The reference ships the checkpoint's `.index` (variable names, shapes, offsets) but NOT its `.data` shard
(`.gitignore:5`), and TensorFlow 1 (`tf.contrib`) is not installable here: the ARCHITECTURE is pinned by the
reference's own index file (every variable this module expects exists there with the same shape —
tests/test_prnet.py against tests/golden/prnet_index.json), the VALUES are unpinned until someone supplies the data
shard.  With the shard present `PosPrediction.restore(prefix)` loads it directly (no TensorFlow).

Engines.  On a CUDA device `PosPrediction` runs the body on the hand-written tcgen05 "tap GEMM" kernel
(`prnet_tc5.compile_plan` -> `lr_tapgemm`: every conv / transposed conv with its batch-norm, activation and residual
add fused into the epilogue, bf16 volumes, fp32 accumulation).  `engine="torch"` keeps the `nn.Module` below on
cuDNN — the fp32 statement of the architecture the kernel path is tested against (and the only engine on a CPU).

TF-slim semantics restated here:
  * `tcl.conv2d(k=4, 'SAME')`: stride 1 pads (1 before, 2 after); stride 2 (even input) pads (1, 1);
  * `tcl.conv2d_transpose(k=4, 'SAME')` is the input-gradient of that conv: stride 2 == ConvTranspose2d(padding=1);
    stride 1 == full transposed conv cropped by (1 before, 2 after);
  * every conv is bias-free and followed by inference batch-norm (epsilon 1e-3, scale=True) + ReLU, except the
    resBlock shortcut / last 1x1 (no norm, no activation) and the final layer (batch-norm + sigmoid);
  * variable names: `resfcn256/Conv`, `resfcn256/resBlock[_k]/{shortcut,Conv,Conv_1,Conv_2,BatchNorm}`,
    `resfcn256/Conv2d_transpose[_k]`; conv weights (kh,kw,in,out), transposed-conv weights (kh,kw,out,in).
"""
import os
import struct

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_BN_EPS = 1e-3      # tf.contrib.layers.batch_norm default epsilon


# ------------------------------------------------------------------------------------------------
# TensorFlow "bundle" checkpoint reader (index = LevelDB-format table of BundleEntryProto)
# ------------------------------------------------------------------------------------------------
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _table_block(buf, offset, size):
    """key/value pairs of one table block (prefix-compressed keys, restart array at the end)."""
    assert buf[offset + size] == 0, "compressed checkpoint index blocks are not supported"
    block = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", block, size - 4)[0]
    end = size - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _proto_fields(buf):
    """minimal protobuf wire decoder -> list of (field number, value) (varints and length-delimited only)"""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.append((field, v))
    return out


_TF_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def read_tf_index(index_path):
    """`<prefix>.index` -> {variable name: {"dtype", "shape", "shard", "offset", "size"}} (BundleEntryProto)."""
    buf = open(index_path, "rb").read()
    assert len(buf) >= 48 and buf[-8:] == struct.pack("<Q", 0xdb4775248b80fb57), "not a TensorFlow checkpoint index"
    pos = len(buf) - 48
    _, pos = _varint(buf, pos)          # metaindex handle
    _, pos = _varint(buf, pos)
    idx_off, pos = _varint(buf, pos)
    idx_size, pos = _varint(buf, pos)
    entries = {}
    for _, handle in _table_block(buf, idx_off, idx_size):
        off, p2 = _varint(handle, 0)
        size, _ = _varint(handle, p2)
        for key, val in _table_block(buf, off, size):
            if not key:
                continue                 # BundleHeaderProto
            e = {"dtype": 1, "shape": [], "shard": 0, "offset": 0, "size": 0}
            for field, v in _proto_fields(val):
                if field == 1:
                    e["dtype"] = v
                elif field == 2:
                    e["shape"] = [dict(_proto_fields(d)).get(1, 0) for f, d in _proto_fields(v) if f == 2]
                elif field == 3:
                    e["shard"] = v
                elif field == 4:
                    e["offset"] = v
                elif field == 5:
                    e["size"] = v
            entries[key.decode()] = e
    return entries


def load_tf_checkpoint(prefix, names=None):
    """{name: ndarray} from `<prefix>.index` + `<prefix>.data-0000k-of-0000n` (no TensorFlow needed)."""
    index = read_tf_index(prefix + ".index")
    n_shards = 1 + max(e["shard"] for e in index.values())
    out = {}
    for name, e in index.items():
        if names is not None and name not in names:
            continue
        path = "%s.data-%05d-of-%05d" % (prefix, e["shard"], n_shards)
        if not os.path.exists(path):
            raise FileNotFoundError("checkpoint data shard %s is missing (the reference does not ship it)" % path)
        with open(path, "rb") as fh:
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
        out[name] = np.frombuffer(raw, dtype=_TF_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ------------------------------------------------------------------------------------------------
# resfcn256 (prnet.py:211-280)
# ------------------------------------------------------------------------------------------------
class _Conv(nn.Module):
    """tcl.conv2d: bias-free conv, TF 'SAME' padding, optional inference batch-norm, optional ReLU."""

    def __init__(self, cin, cout, k, stride, norm=True, act=True):
        super().__init__()
        self.k, self.stride, self.act = k, stride, act
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=0, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=_BN_EPS) if norm else None

    def forward(self, x):
        if self.k > 1:
            total = self.k - self.stride          # even inputs: out = in / stride
            x = F.pad(x, (total // 2, total - total // 2, total // 2, total - total // 2))
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return F.relu(x) if self.act else x


class _Deconv(nn.Module):
    """tcl.conv2d_transpose(k=4, 'SAME'): bias-free, inference batch-norm, ReLU or sigmoid."""

    def __init__(self, cin, cout, stride, final=False):
        super().__init__()
        self.stride, self.final = stride, final
        self.conv = nn.ConvTranspose2d(cin, cout, 4, stride=stride, padding=1 if stride == 2 else 0, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=_BN_EPS)

    def forward(self, x):
        if self.stride == 1:
            # the stride-1 transposed conv, cropped by (1 before, 2 after), IS a plain conv with the kernel flipped
            # and its channel axes swapped over the input padded (2 before, 1 after) — same sums, but the library
            # then runs a forward-conv kernel instead of an input-gradient one (10 of the 17 decoder layers)
            w = self.conv.weight.flip(2, 3).transpose(0, 1)
            y = F.conv2d(F.pad(x, (2, 1, 2, 1)), w)
        else:
            y = self.conv(x)
        y = self.bn(y)
        return torch.sigmoid(y) if self.final else F.relu(y)


class _ResBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.shortcut = _Conv(cin, cout, 1, stride, norm=False, act=False) if (stride != 1 or cin != cout) else None
        self.c0 = _Conv(cin, cout // 2, 1, 1)
        self.c1 = _Conv(cout // 2, cout // 2, 4, stride)
        self.c2 = _Conv(cout // 2, cout, 1, 1, norm=False, act=False)
        self.bn = nn.BatchNorm2d(cout, eps=_BN_EPS)

    def forward(self, x):
        s = x if self.shortcut is None else self.shortcut(x)
        return F.relu(self.bn(self.c2(self.c1(self.c0(x))) + s))


_ENC = [(32, 2), (32, 1), (64, 2), (64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1)]
_DEC = [(512, 1), (256, 2), (256, 1), (256, 1), (128, 2), (128, 1), (128, 1), (64, 2), (64, 1), (64, 1),
        (32, 2), (32, 1), (16, 2), (16, 1), (3, 1), (3, 1), (3, 1)]


class ResFcn256(nn.Module):
    """(N,3,256,256) in [0,1] -> (N,3,256,256) in (0,1): 1 conv + 10 resBlocks + 17 transposed convs."""

    def __init__(self):
        super().__init__()
        self.stem = _Conv(3, 16, 4, 1)
        blocks, c = [], 16
        for cout, s in _ENC:
            blocks.append(_ResBlock(c, cout, s))
            c = cout
        self.enc = nn.ModuleList(blocks)
        dec = []
        for i, (cout, s) in enumerate(_DEC):
            dec.append(_Deconv(c, cout, s, final=(i == len(_DEC) - 1)))
            c = cout
        self.dec = nn.ModuleList(dec)

    def forward(self, x):
        x = self.stem(x)
        for b in self.enc:
            x = b(x)
        for d in self.dec:
            x = d(x)
        return x

    # ---- TensorFlow variable names --------------------------------------------------------------
    def tf_variables(self):
        """[(tf variable name, parameter/buffer tensor, kind)] with kind in {'conv', 'deconv', 'vec'}: the shape TF
        holds is (kh,kw,in,out) for 'conv', (kh,kw,out,in) for 'deconv'."""
        out = []

        def bn(scope, m):
            out.extend([(scope + "/BatchNorm/gamma", m.weight, "vec"), (scope + "/BatchNorm/beta", m.bias, "vec"),
                        (scope + "/BatchNorm/moving_mean", m.running_mean, "vec"),
                        (scope + "/BatchNorm/moving_variance", m.running_var, "vec")])

        def conv(scope, m, kind="conv"):
            out.append((scope + "/weights", m.conv.weight, kind))
            if m.bn is not None:
                bn(scope, m.bn)

        conv("resfcn256/Conv", self.stem)
        for i, b in enumerate(self.enc):
            s = "resfcn256/resBlock" + ("_%d" % i if i else "")
            if b.shortcut is not None:
                conv(s + "/shortcut", b.shortcut)
            conv(s + "/Conv", b.c0)
            conv(s + "/Conv_1", b.c1)
            conv(s + "/Conv_2", b.c2)
            bn(s, b.bn)
        for i, d in enumerate(self.dec):
            conv("resfcn256/Conv2d_transpose" + ("_%d" % i if i else ""), d, "deconv")
        return out

    def tf_shapes(self):
        """{tf variable name: shape as stored in the checkpoint}"""
        shapes = {}
        for name, t, kind in self.tf_variables():
            if kind == "vec":
                shapes[name] = [t.shape[0]]
            elif kind == "conv":                      # torch (out,in,kh,kw) -> TF (kh,kw,in,out)
                shapes[name] = [t.shape[2], t.shape[3], t.shape[1], t.shape[0]]
            else:                                     # torch (in,out,kh,kw) -> TF (kh,kw,out,in)
                shapes[name] = [t.shape[2], t.shape[3], t.shape[1], t.shape[0]]
        return shapes

    def load_tf_variables(self, arrays):
        """arrays: {tf variable name: ndarray} (load_tf_checkpoint, or an .npz exported elsewhere)."""
        with torch.no_grad():
            for name, t, kind in self.tf_variables():
                a = torch.from_numpy(np.asarray(arrays[name], dtype=np.float32))
                if kind != "vec":
                    a = a.permute(3, 2, 0, 1)         # both layouts map with the same permutation (see tf_shapes)
                assert tuple(a.shape) == tuple(t.shape), (name, tuple(a.shape), tuple(t.shape))
                t.copy_(a)
        return self


class PosPrediction:
    """Drop-in for prnet.PosPrediction (:283-314): predict / predict_batch take NHWC float images in [0,1]
    (numpy or torch) and return the position map * MaxPos in the same layout."""
    MAX_PLAN_BATCH = 64                 # largest batch a tcgen05 launch plan is compiled for (a power of two)

    def __init__(self, resolution_inp=256, resolution_op=256, device="cuda", dtype=torch.float32, engine=None):
        self.resolution_inp, self.resolution_op = resolution_inp, resolution_op
        self.MaxPos = resolution_inp * 1.1
        self.device, self.dtype = torch.device(device), dtype
        self.engine = engine or ("tcgen05" if self.device.type == "cuda" else "torch")
        assert self.engine in ("tcgen05", "torch")
        self.network = ResFcn256().eval().to(self.device)
        self._plans = {}
        self._configure()

    def _configure(self):
        self._plans = {}                 # compiled launch plans hold packed copies of the weights
        if self.engine == "torch":
            self.network = self.network.to(dtype=self.dtype, memory_format=torch.channels_last)

    def plan(self, batch):
        """The compiled tcgen05 launch plan for `batch` frames (built on first use, one per batch size)."""
        if batch not in self._plans:
            from . import prnet_tc5
            self._plans[batch] = prnet_tc5.compile_plan(self.network.float(), batch, self.resolution_inp, self.device,
                                                        max_pos=self.MaxPos)
        return self._plans[batch]

    def restore(self, model_path):
        names = [n for n, _, _ in self.network.tf_variables()]
        self.network.float().load_tf_variables(load_tf_checkpoint(model_path, set(names)))
        self._configure()

    def predict_batch(self, images):
        as_numpy = isinstance(images, np.ndarray)
        x = torch.as_tensor(images, device=self.device)
        if self.engine == "tcgen05":
            # a compiled plan owns ~33 MB of activation volumes per frame: only power-of-two batch sizes up to
            # MAX_PLAN_BATCH are compiled (at most 7 plans), any other batch runs as its binary decomposition
            # (37 frames = 32 + 4 + 1) — the ragged tail batches of generate_dataview would otherwise compile, and
            # keep, one plan per distinct size
            x = x.float().contiguous()
            n, outs, i = x.shape[0], [], 0
            if n == 0:
                y = x.new_zeros((0, self.resolution_op, self.resolution_op, 3))
                return y.cpu().numpy() if as_numpy else y
            while i < n:
                b = 1 << min((n - i).bit_length() - 1, self.MAX_PLAN_BATCH.bit_length() - 1)
                outs.append(self.plan(b).run(x[i:i + b]).clone())
                i += b
            y = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
            return y.cpu().numpy() if as_numpy else y
        with torch.no_grad():
            x = x.permute(0, 3, 1, 2).to(dtype=self.dtype, memory_format=torch.channels_last)
            y = self.network(x).permute(0, 2, 3, 1).float() * self.MaxPos
        y = y.contiguous()
        return y.cpu().numpy() if as_numpy else y

    def predict(self, image):
        out = self.predict_batch(image[None])
        return out[0]
