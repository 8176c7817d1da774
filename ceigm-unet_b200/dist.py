"""Multi-GPU plumbing for the hot path: the scan shards by batch with NO data-path collective —
every (batch, direction, channel) row is independent — so ranks run replicas on their batch shard. The only
collective around the operator is the data-parallel gradient all-reduce of its parameters (NCCL over NVLink on
GPUs, gloo in the CPU tests), plus a max-over-ranks for device timings.
"""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (no-op for 1 rank)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of this rank's samples: contiguous shards whose sizes differ by at most one."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's value (timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(tensors: Iterable[torch.Tensor]) -> None:
    """In-place mean over ranks of a set of gradient tensors, flattened into one bucket per dtype (one collective
    per dtype, as DDP's bucketing would)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    by_dtype = {}
    for t in tensors:
        if t is not None:
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for group in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in group])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for t in group:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


def allreduce_module_grads_(module: torch.nn.Module) -> None:
    allreduce_mean_([p.grad for p in module.parameters() if p.grad is not None])
