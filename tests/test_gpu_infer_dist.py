"""-m gpu: the frames->characters inference stream (greedy CTC on the device) and, when the box has two
GPUs, the NCCL data-parallel step (identical weights on both ranks after a step; matches one process
over the concatenated batch for equal-length clips)."""
import os
import socket

import pytest
import torch

from oracle import sequence as O

pytestmark = pytest.mark.gpu


def test_recognizer_tokens_match_oracle_pipeline(native_lib, cuda):
    from lipreading_b200.infer import Recognizer
    from lipreading_b200.model import VideoEncoder
    c2i = O.build_char2idx()
    torch.manual_seed(7)
    enc = VideoEncoder(204, 32, rnn_type="GRU", bidirectional=True, enable_ctc=True, vocab_size=64,
                       char2idx=c2i, device=cuda).to(cuda)
    g = torch.Generator().manual_seed(3)
    B, T = 9, 40
    lens = torch.randint(10, T + 1, (B,), generator=g).sort().values
    lens[-1] = T
    frames = torch.randn(B, T, 68, 3, generator=g) * 3
    for b in range(B):
        frames[b, int(lens[b]):] = 0
    rec = Recognizer(enc, c2i)
    tok, n = rec.tokens(frames.to(cuda), lens.to(cuda))
    state = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    lp, _, _ = O.encoder_forward(state, frames, lens, "GRU", True, c2i)
    ref = O.greedy_ctc_decode(lp, lens)
    for b in range(B):
        assert (tok[b, : int(n[b])].cpu() + 1).tolist() == ref[b]
    texts = rec(frames.to(cuda), lens.to(cuda))
    assert len(texts) == B and all(isinstance(t, str) for t in texts)


def test_frame_stream_matches_stage_by_stage_oracle(native_lib, cuda):
    """frames -> characters in one pass (infer.FrameRecognizer) == the oracle run stage by stage: numpy restatement of
    the landmark geometry + mouth-crop spec per frame, then the fp32 conv / BiGRU / greedy-CTC port on the clips."""
    import numpy as np
    from lipreading_b200.face import PRN
    from lipreading_b200.infer import FrameRecognizer
    from lipreading_b200.model import VideoEncoder
    from oracle import train_step as TS
    from oracle import vision as V
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    uv = np.loadtxt(os.path.join(gold, "uv_kpt_ind.txt")).astype(np.int32)
    face = np.load(os.path.join(gold, "face_ind.npy"))
    c2i = O.build_char2idx()
    torch.manual_seed(3)
    enc = VideoEncoder(1728, 32, frame_processing="conv3d", rnn_type="GRU", bidirectional=True, enable_ctc=True,
                       vocab_size=64, char2idx=c2i, device=cuda).to(cuda)
    rng = np.random.default_rng(2)
    H, W, T, B = 180, 240, 6, 2
    frames = rng.integers(0, 256, (B * T, H, W, 3), dtype=np.uint8)
    rects = [(60 + i, 160 + i, 35, 140) for i in range(B * T)]

    def fake_cnn(cropped):                      # identity position map of the crop + a smooth image-dependent term
        n = cropped.shape[0]
        v, u = torch.meshgrid(torch.arange(256.0, device=cropped.device), torch.arange(256.0, device=cropped.device),
                              indexing="ij")
        return (torch.stack([u, v, 30 + 0 * u], -1)[None].expand(n, -1, -1, -1) + cropped * 2.0).contiguous()
    prn = PRN(predict_batch=fake_cnn, uv_kpt_ind=uv, face_ind=face, device=cuda)
    stream = FrameRecognizer(enc, c2i, prn, batch=5)
    fr_d = torch.from_numpy(frames).to(cuda)
    crops, lmk = stream.mouth_clips(fr_d, rects)
    tok, n = stream.tokens(fr_d, T, rects)
    # oracle, stage by stage
    crops_ref = []
    for i in range(B * T):
        c, s = V.crop_box(rects[i])
        cropped = V.warp_bilinear_constant(frames[i], np.linalg.inv(V.crop_transform(c, s))).astype(np.float32)
        vv, uu = np.meshgrid(np.arange(256, dtype=np.float32), np.arange(256, dtype=np.float32), indexing="ij")
        pos = (np.stack([uu, vv, np.full_like(uu, 30)], -1) + cropped * np.float32(2.0)).astype(np.float32)
        l_ref, rp = V.frame_landmarks((H, W, 3), rects[i], pos, uv, None)
        assert np.abs(lmk[i].cpu().numpy() - l_ref).max() < 2e-3
        # the crop is cut with the landmarks the device produced (a 1e-3 px landmark difference may move an integer ROI edge)
        roi = V.mouth_roi(lmk[i].cpu().numpy(), rp, 100, 50)
        crops_ref.append(V.mouth_crop(frames[i], roi, 100, 50))
    crops_ref = np.stack(crops_ref)
    assert np.array_equal(crops.cpu().numpy(), crops_ref)
    state = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    port = TS.CpuStep(state, "GRU", c2i, quantize=True)
    clips = torch.from_numpy(crops_ref).reshape(B, T, 100, 50, 3)
    lens = torch.full((B,), T, dtype=torch.long)
    lp = port.forward(clips, lens).detach()
    ref = O.greedy_ctc_decode(lp, lens)
    lp_gpu, _, _ = enc.eval()(clips.to(cuda), lens.to(cuda))
    assert float((lp_gpu.cpu() - lp).abs().max()) < 5e-3
    for b in range(B):
        assert (tok[b, : int(n[b])].cpu() + 1).tolist() == ref[b]
    assert len(stream(fr_d, T, rects)) == B


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out_dir):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world)})
    from lipreading_b200 import dist as ldist, trainer
    from lipreading_b200.model import VideoEncoder
    r, lr_, w = ldist.init()
    dev = torch.device("cuda", lr_)
    c2i = O.build_char2idx()
    torch.manual_seed(11)
    enc = VideoEncoder(204, 32, rnn_type="GRU", bidirectional=True, enable_ctc=True, vocab_size=64,
                       char2idx=c2i, device=dev).to(dev)
    g = torch.Generator().manual_seed(5)
    B, T = 8, 24
    frames = torch.randn(B, T, 68, 3, generator=g)
    lens = torch.full((B,), T)
    chars = torch.zeros(B, 8, dtype=torch.long)
    chars[:, 0] = 1
    chars[:, 1:6] = torch.randint(4, 64, (B, 5), generator=g)
    chars[:, 6] = 2
    char_lens = torch.full((B,), 7)
    batch = ldist.shard_batch((frames, lens, chars, char_lens), rank, world)
    opt = torch.optim.SGD(enc.parameters(), lr=0.1)
    trainer.train_ctc(enc, [batch], opt, dev, c2i, grad_norm=50, dist=ldist.GradAllReducer(world))
    torch.save({k: v.cpu() for k, v in enc.state_dict().items()}, os.path.join(out_dir, "w%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_gpu_nccl_step_keeps_replicas_identical(native_lib, cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from lipreading_b200 import trainer
    from lipreading_b200.model import VideoEncoder
    port = _free_port()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0 = torch.load(os.path.join(tmp_path, "w0.pt"))
    w1 = torch.load(os.path.join(tmp_path, "w1.pt"))
    for k in w0:
        assert torch.equal(w0[k], w1[k]), k
    # single process over the whole batch: equal-length clips => mean of rank means == global mean
    c2i = O.build_char2idx()
    torch.manual_seed(11)
    enc = VideoEncoder(204, 32, rnn_type="GRU", bidirectional=True, enable_ctc=True, vocab_size=64,
                       char2idx=c2i, device=cuda).to(cuda)
    g = torch.Generator().manual_seed(5)
    B, T = 8, 24
    frames = torch.randn(B, T, 68, 3, generator=g)
    chars = torch.zeros(B, 8, dtype=torch.long)
    chars[:, 0] = 1
    chars[:, 1:6] = torch.randint(4, 64, (B, 5), generator=g)
    chars[:, 6] = 2
    opt = torch.optim.SGD(enc.parameters(), lr=0.1)
    trainer.train_ctc(enc, [(frames, torch.full((B,), T), chars, torch.full((B,), 7))], opt, cuda, c2i, grad_norm=50)
    for k, v in enc.state_dict().items():
        assert float((v.cpu() - w0[k]).abs().max()) < 1e-5, k
