"""Data parallelism over clips: one process per GPU, gradients only.

The reference is single-process (src/scripts/train.py:200-203 asks "how to use multiple GPUs?").
Here every rank takes a CONTIGUOUS slice of each global batch of the length-sorted dataset (keeps
frame_lens non-decreasing per rank, which ctc_loss asserts), runs forward/backward locally, and the
gradients are summed with a flat-bucket NCCL all-reduce over NVLink/NVSwitch, divided by the number
of ranks that contributed, and only then clipped — so every rank clips and steps on identical
gradients.  The model is 0.7-7 M parameters (<= 28 MB fp32): one bucket is latency-optimal; an
"early" bucket (the recurrent layer + projection, whose gradients are complete long before the conv
weight gradients) can be reduced on a side stream while the rest of backward still runs.

Collective discipline: every rank issues the SAME number of collectives per global batch —
  * a global batch with fewer clips than ranks is dropped on every rank (`ShardedLoader`);
  * a rank whose local loss is unusable (ctc_loss returned None) still joins the all-reduce with
    zero gradients and `valid=False`; the sum is divided by the number of valid ranks.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def _gpu_numa_node(local_rank):
    """NUMA node of the GPU `local_rank` drives (sysfs), or None when the platform does not say (one-node VMs)."""
    try:
        import subprocess
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id",
                                       "--format=csv,noheader"], text=True, timeout=20).strip().lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read())
        return node if node >= 0 else None
    except Exception:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(local_rank):
    """Pin this process (its threads and its future allocations — the pinned staging buffers of the loader above all)
    to the NUMA node of its GPU.  With eight ranks on a two-socket host the default first-touch placement puts half
    of the pinned batches on the far socket, and their host->device copies then cross the socket interconnect.
    Returns the node, or None when nothing was done (single node, no sysfs, LR_NUMA_BIND=0)."""
    if os.environ.get("LR_NUMA_BIND", "1") == "0":
        return None
    node = _gpu_numa_node(local_rank)
    if node is None:
        return None
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = _parse_cpulist(fh.read()) & os.sched_getaffinity(0)
        if not cpus or not os.path.isdir("/sys/devices/system/node/node1"):
            return None
        os.sched_setaffinity(0, cpus)
        import ctypes
        mask = ctypes.c_ulong(1 << node)
        # set_mempolicy(MPOL_PREFERRED = 1): allocate on the GPU's node, fall back elsewhere when it is full
        ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
        return node
    except Exception:
        return None


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and torch.cuda.is_available():
        bind_to_gpu_numa(local_rank)
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
    label axes trimmed to the slice's own maxima (padding stays minimal).  None when the slice is empty."""
    frames, frame_lens, chars, char_lens = batch
    lo, hi = shard_slice(frames.shape[0], rank, world)
    fl, cl = frame_lens[lo:hi], char_lens[lo:hi]
    if hi <= lo:
        return None
    return frames[lo:hi, : int(fl.max())], fl, chars[lo:hi, : int(cl.max())], cl


class GradAllReducer:
    """Sum of the gradients over ranks / number of contributing ranks, through flat fp32 buckets.

    `allreduce_grads(params, valid=True)` — everything in one bucket after backward (one collective).
    `arm(early_params)` + `allreduce_grads(...)` — two collectives per step on every rank: the early
    bucket is launched from a post-accumulate hook on a side stream as soon as the last of
    `early_params` has its gradient (i.e. while the conv front-end's backward still runs), the rest
    at the call.  After the call every p.grad is a VIEW into the reduced bucket (no unpack copies);
    a parameter without a gradient contributes zeros."""

    def __init__(self, world=None, group=None):
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.group = group
        self._bufs = {}
        self._early = None              # (params, pending set, handles)
        self._early_done = None         # (event, flat, params)
        self._side = None

    # -- packing -------------------------------------------------------------------------------
    def _flat_for(self, key, params):
        n = sum(p.numel() for p in params) + 1              # + one slot: this rank's `valid` flag
        dev = params[0].device
        buf = self._bufs.get(key)
        if buf is None or buf.numel() != n or buf.device != dev:
            buf = torch.empty(n, dtype=torch.float32, device=dev)
            self._bufs[key] = buf
        return buf

    @staticmethod
    def _views(flat, params):
        out, off = [], 0
        for p in params:
            out.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return out

    def _pack(self, flat, params, valid):
        views = self._views(flat, params)
        have = [(v, p.grad) for v, p in zip(views, params) if p.grad is not None and valid]
        none = [v for v, p in zip(views, params) if p.grad is None or not valid]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])     # a few multi-tensor launches
        for v in none:
            v.zero_()
        flat[-1:].fill_(1.0 if valid else 0.0)
        return views

    def _reduce(self, flat):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        # divide by the number of ranks whose gradients are real (>= 1 so an all-invalid step is a zero gradient)
        flat[:-1].div_(flat[-1:].clamp(min=1.0))

    # -- early bucket ----------------------------------------------------------------------------
    def arm(self, early_params):
        """Register (once) hooks that all-reduce `early_params` on a side stream as soon as they all have
        gradients in a backward pass.  Call before every backward (cheap: resets the pending set)."""
        if self.world <= 1:
            return
        early_params = [p for p in early_params if p.requires_grad]
        if not early_params or not early_params[0].is_cuda:
            return
        if self._early is None or [id(p) for p in self._early[0]] != [id(p) for p in early_params]:
            if self._early is not None:
                for h in self._early[2]:
                    h.remove()
            handles = [p.register_post_accumulate_grad_hook(self._hook) for p in early_params]
            self._early = (early_params, set(), handles)
        self._early[1].clear()
        self._early[1].update(id(p) for p in early_params)
        self._early_done = None

    def _hook(self, p):
        if self._early is None:
            return
        pending = self._early[1]
        pending.discard(id(p))
        if pending or self._early_done is not None:
            return
        params = self._early[0]
        dev = params[0].device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        flat = self._flat_for("early", params)
        self._pack(flat, params, True)                       # on the compute stream, right behind the producers
        ready = torch.cuda.Event()
        ready.record(cur)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            self._reduce(flat)
            done = torch.cuda.Event()
            done.record(self._side)
        self._early_done = (done, flat, params)

    # -- the call ----------------------------------------------------------------------------------
    def allreduce_grads(self, params, valid=True):
        """Collective: every rank must call it once per step (with valid=False and whatever gradients it has
        when its local batch was unusable)."""
        if self.world <= 1:
            return
        params = [p for p in params if p.requires_grad]
        if not params:
            return
        early = self._early_done
        self._early_done = None
        early_ids = set()
        if self._early is not None and valid and early is None:
            # hooks armed but backward never completed the early set (e.g. unused parameters): fall through to a
            # second collective below so that the per-step collective count stays the same on every rank
            flat = self._flat_for("early", self._early[0])
            self._pack(flat, self._early[0], True)
            self._reduce(flat)
            early = (None, flat, self._early[0])
        elif self._early is not None and not valid and early is None:
            flat = self._flat_for("early", self._early[0])
            self._pack(flat, self._early[0], False)
            self._reduce(flat)
            early = (None, flat, self._early[0])
        if early is not None:
            done, flat_e, params_e = early
            if done is not None:
                torch.cuda.current_stream(params_e[0].device).wait_event(done)
            for p, v in zip(params_e, self._views(flat_e, params_e)):
                p.grad = v
            early_ids = {id(p) for p in params_e}
        rest = [p for p in params if id(p) not in early_ids]
        if rest:
            flat = self._flat_for("late", rest)
            views = self._pack(flat, rest, valid)
            self._reduce(flat)
            for p, v in zip(rest, views):
                p.grad = v


class ShardedLoader:
    """Wrap a loader of GLOBAL batches so that each rank iterates its own contiguous slices.  A global batch with
    fewer clips than ranks (only the tail batch can be) is dropped on EVERY rank: no rank ever sits out a step,
    so the ranks always issue the same sequence of collectives."""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            if batch[0].shape[0] < self.world:
                continue
            yield shard_batch(batch, self.rank, self.world)


def allreduce_counts(*values, device=None):
    """Sum a few host/device scalars over the ranks (evaluation counters); returns float tensors on `device`."""
    t = torch.stack([torch.as_tensor(v, dtype=torch.float64, device=device).reshape(()) for v in values])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [x for x in t]


def broadcast_object(obj, src=0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]
    return obj


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
