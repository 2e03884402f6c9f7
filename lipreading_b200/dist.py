"""Data parallelism over clips: one process per GPU, gradients only.

The reference is single-process (src/scripts/train.py:200-203 asks "how to use multiple GPUs?").
Here every rank takes a CONTIGUOUS slice of each global batch of the length-sorted dataset (keeps
frame_lens non-decreasing per rank, which ctc_loss asserts), runs forward/backward locally, and the
gradients are summed with ONE flat-bucket NCCL all-reduce over NVLink/NVSwitch, divided by the world
size, and only then clipped — so every rank clips and steps on identical gradients.
The model is 0.7-7 M parameters (<= 28 MB fp32): a single bucket is latency-optimal.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_slice(n, rank, world):
    """Contiguous [lo, hi) slice of a global batch of n clips for `rank` (remainder to the low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world):
    """batch = (frames, frame_lens, chars, char_lens) for the GLOBAL batch -> this rank's slice, time and
    label axes trimmed to the slice's own maxima (padding stays minimal)."""
    frames, frame_lens, chars, char_lens = batch
    lo, hi = shard_slice(frames.shape[0], rank, world)
    fl, cl = frame_lens[lo:hi], char_lens[lo:hi]
    if hi <= lo:
        return None
    return frames[lo:hi, : int(fl.max())], fl, chars[lo:hi, : int(cl.max())], cl


class GradAllReducer:
    """allreduce_grads(params): one flat fp32 bucket, sum over ranks, / world."""

    def __init__(self, world=None, group=None):
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.group = group
        self._flat = None

    def allreduce_grads(self, params):
        if self.world <= 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        off = 0
        for g in grads:
            self._flat[off:off + g.numel()].copy_(g.reshape(-1))
            off += g.numel()
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        self._flat.div_(self.world)
        off = 0
        for g in grads:
            g.copy_(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class ShardedLoader:
    """Wrap a loader of GLOBAL batches so that each rank iterates its own contiguous slices."""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            part = shard_batch(batch, self.rank, self.world)
            if part is not None:
                yield part
