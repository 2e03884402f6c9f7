"""CPU tests of the data-parallel host logic with world_size 2 over gloo: contiguous
sharding of the length-sorted global batch and the flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lipreading_b200 import dist as ldist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world)})
    r, _, w = ldist.init(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                   # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(1)
    B = 7                                                  # odd: remainder goes to the low ranks
    frames = torch.randn(B, 9, 6, generator=g)
    frame_lens = torch.tensor([4, 4, 5, 6, 8, 9, 9])
    chars = torch.randint(0, 5, (B, 6), generator=g)
    char_lens = torch.tensor([3, 4, 4, 5, 6, 6, 5])
    part = ldist.shard_batch((frames, frame_lens, chars, char_lens), rank, world)
    lo, hi = ldist.shard_slice(B, rank, world)
    assert part[0].shape[0] == hi - lo and part[0].shape[1] == int(frame_lens[lo:hi].max())
    assert bool((part[1][1:] >= part[1][:-1]).all())       # stays non-decreasing per rank (ctc_loss.py:39)
    # sum-of-per-sample loss so that (sum of rank grads) == global grad
    loss = model(part[0]).pow(2).sum()
    loss.backward()
    red = ldist.GradAllReducer(world)
    params = list(model.parameters())
    red.allreduce_grads(params)
    torch.save([p.grad.clone() for p in params], os.path.join(out_dir, "g%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_slices_partition_the_batch():
    for n in (1, 2, 7, 256, 257):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = ldist.shard_slice(n, r, world)
                cover += list(range(lo, hi))
            assert cover == list(range(n))


def test_two_rank_gloo_allreduce_matches_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, "g0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "g1.pt"))
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                           # every rank holds identical gradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(1)
    frames = torch.randn(7, 9, 6, generator=g)
    frame_lens = torch.tensor([4, 4, 5, 6, 8, 9, 9])
    total = 0
    for r in range(world):
        lo, hi = ldist.shard_slice(7, r, world)
        total = total + model(frames[lo:hi, : int(frame_lens[lo:hi].max())]).pow(2).sum()
    total.backward()
    for p, a in zip(model.parameters(), g0):
        assert torch.allclose(p.grad / world, a, atol=1e-6)
