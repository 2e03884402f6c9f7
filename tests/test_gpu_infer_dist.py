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
