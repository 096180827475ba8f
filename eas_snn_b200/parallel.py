"""Sequence sharding across ranks.

Every event window is independent (the sampler is stateless across calls and the backbone state
is reset per batch, ``yolox/core/trainer.py:115-117``), so the path shards by contiguous blocks of
windows with NO data-path collective for inference -- what the reference does with
``DistributedSampler`` (``yolox/exp/event_yolox_base.py:295-296, 490-494``).  Training adds exactly
one collective, the gradient all-reduce (``DistributedDataParallel``, ``yolox/core/trainer.py:176``),
here a flat-bucket NCCL all-reduce.
"""
from __future__ import annotations

import os
from typing import Iterable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: str | None = None):
    """One process per GPU; rendezvous from the torchrun environment (127.0.0.1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process (and so the first-touch placement of its pinned host buffers) to the NUMA node the GPU's
    PCIe root hangs off.  With one process per GPU the host <-> device copies of N ranks otherwise fight over one
    socket's memory and the inter-socket links; the staging path is the bound of the end-to-end number.  A no-op
    (returns why) when the topology files are missing or there is a single node."""
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        info.update(pci=bus, node=node)
        if node < 0:
            info["why"] = "no NUMA affinity reported for the device"
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            info["why"] = "single node or nothing to narrow"
            return info
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus))
    except (OSError, ValueError, AttributeError, RuntimeError) as e:   # topology not exposed (containers): stay put
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info


def shard_range(n_windows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of windows owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_events(x, y, t, p, offsets, rank: int, world: int):
    """Slice a host (numpy) batch of windows to this rank's block; offsets are re-based to 0."""
    B = len(offsets) - 1
    lo, hi = shard_range(B, rank, world)
    s, e = int(offsets[lo]), int(offsets[hi])
    return x[s:e], y[s:e], t[s:e], p[s:e], (np.asarray(offsets[lo:hi + 1]) - s).astype(np.int64)


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    tns = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(tns, op=dist.ReduceOp.MAX)
    return float(tns.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    tns = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(tns, op=dist.ReduceOp.SUM)
    return float(tns.item())


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20):
    """Average gradients over ranks with flat buckets (NCCL over NVLink 5 on the GPU box).

    The reference's only training collective (DDP, trainer.py:176).  NVSwitch makes the cost
    latency- not link-bound, so buckets are sized for launch count (64 MB), not for link count.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    grads = [q.grad for q in params if q.grad is not None]
    bucket: list[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        o = 0
        for g in bucket:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
        bucket, size = [], 0

    for g in grads:
        if bucket and (bucket[0].dtype != g.dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
