#!/usr/bin/env python
"""Pinned host->device copy bandwidth per GPU, alone and with all GPUs copying at once (VERDICT r1 item 4:
separate link, NUMA and pinned-pool effects behind the end-to-end number).

    python tools/h2d_bench.py                 # one GPU: default placement, then pinned on every NUMA node in turn
    torchrun --nproc-per-node N tools/h2d_bench.py --concurrent   # N ranks copying simultaneously

Prints one JSON line per measurement (rank 0 gathers under torchrun)."""
import argparse
import ctypes
import glob
import json
import os
import time

import torch


def numa_nodes():
    out = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            with open(os.path.join(d, "cpulist")) as fh:
                out[int(d.rsplit("node", 1)[1])] = fh.read().strip()
        except Exception:
            pass
    return out


def parse_cpulist(s):
    cpus = set()
    for part in s.split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa(index):
    import subprocess
    try:
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True).strip()
    except Exception:
        return None, None
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            return bdf, int(fh.read())
    except Exception:
        return bdf, None


def set_mempolicy_bind(node):
    """MPOL_BIND to one node for this thread's future allocations (syscall 238 on x86-64); returns success."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        return libc.syscall(238, 2, ctypes.byref(mask), ctypes.c_ulong(64)) == 0
    except Exception:
        return False


def reset_mempolicy():
    try:
        ctypes.CDLL(None).syscall(238, 0, None, ctypes.c_ulong(0))
    except Exception:
        pass


def measure(dev, nbytes, iters=8, streams=1):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(7)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ss = [torch.cuda.Stream(dev) for _ in range(streams)]
    step = -(-nbytes // streams)

    def go():
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                dst[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    go()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        go()
    torch.cuda.synchronize(dev)
    return nbytes * iters / (time.perf_counter() - t0) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=288)
    ap.add_argument("--concurrent", action="store_true")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    nbytes = args.mb << 20
    nodes = numa_nodes()
    bdf, node = gpu_numa(local)
    base = {"rank": rank, "gpu": local, "pci": bdf, "gpu_numa_node": node, "numa_nodes": len(nodes),
            "affinity_cpus": len(os.sched_getaffinity(0)), "mb": args.mb}
    rows = []
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        dist.barrier()
    rows.append(dict(base, placement="default", streams=1, GBps=measure(dev, nbytes)))
    if not args.concurrent:
        rows.append(dict(base, placement="default", streams=2, GBps=measure(dev, nbytes, streams=2)))
        all_cpus = os.sched_getaffinity(0)
        for n, cpul in nodes.items():
            cpus = parse_cpulist(cpul) & all_cpus
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
            ok = set_mempolicy_bind(n)
            rows.append(dict(base, placement="node%d%s" % (n, "" if ok else " (affinity only)"), streams=1,
                             GBps=measure(dev, nbytes)))
            reset_mempolicy()
            os.sched_setaffinity(0, all_cpus)
    else:
        if node is not None and node >= 0 and node in nodes:
            cpus = parse_cpulist(nodes[node]) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                ok = set_mempolicy_bind(node)
                import torch.distributed as dist
                dist.barrier()
                rows.append(dict(base, placement="gpu-local node%d%s" % (node, "" if ok else " (affinity only)"), streams=1,
                                 GBps=measure(dev, nbytes)))
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, rows)
        if rank == 0:
            for rr in gathered:
                for r in rr:
                    print(json.dumps(r))
            by = {}
            for rr in gathered:
                for r in rr:
                    by.setdefault(r["placement"].split(" ")[0], []).append(r["GBps"])
            print(json.dumps({"world": world, "aggregate_GBps": {k: sum(v) for k, v in by.items()},
                              "min_GBps": {k: min(v) for k, v in by.items()}}))
        dist.destroy_process_group()
    else:
        for r in rows:
            print(json.dumps(r))


if __name__ == "__main__":
    main()
