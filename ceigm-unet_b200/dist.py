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


class GradReducer:
    """Data-parallel gradient all-reduce overlapped with the backward pass — the one collective of the path
    (Lightning's default DDP strategy in train_synapse.py:211-214).

    Parameters are packed, last-registered first (the order in which backward produces their gradients), into flat
    buckets of about `bucket_bytes`; every `p.grad` is a VIEW into its bucket (what DDP calls gradient_as_bucket_view), so
    no gather / scatter copy is needed. A post-accumulate hook counts the bucket's gradients in; when the last one lands
    the bucket's all-reduce (mean over ranks; NCCL's own stream, so it overlaps the rest of the backward) is launched.
    With `defer = True` the hooks only count and `finish()` launches everything (used when the backward itself is replayed
    from a CUDA graph, where no hook runs). `finish()` — called after `loss.backward()` — launches the buckets whose parameters took no part in this step (their
    gradients are zeros) and waits for all of them. Use `zero_grad()` of this object instead of the optimizer's
    (`set_to_none=True` would detach the views).
    """

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 25 << 20):
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.avg = self.world > 1 and dist.get_backend() == "nccl"
        self.defer = False      # True: hooks only count; every bucket is launched by finish() (CUDA-graph captured backward)
        params = [p for p in module.parameters() if p.requires_grad]
        groups, cur, cur_bytes, key = [], [], 0, None
        for p in reversed(params):
            k = (p.dtype, p.device)
            if cur and (k != key or cur_bytes >= bucket_bytes):
                groups.append(cur)
                cur, cur_bytes = [], 0
            key = k
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
        if cur:
            groups.append(cur)
        self.buckets = []
        self._bucket_of = {}
        self._hooks = []
        for gi, group in enumerate(groups):
            total = sum((p.numel() + 3) // 4 * 4 for p in group)          # 16-byte aligned slots
            flat = torch.zeros(total, dtype=group[0].dtype, device=group[0].device)
            off = 0
            for p in group:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
                self._bucket_of[p] = gi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._ready))
            self.buckets.append({"flat": flat, "n": len(group), "pending": len(group), "work": None, "launched": False})
        self.bytes = sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)

    def _launch(self, b) -> None:
        b["launched"] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
            b["work"] = dist.all_reduce(b["flat"], op=op, async_op=True)

    def _ready(self, p) -> None:
        b = self.buckets[self._bucket_of[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"] and not self.defer:
            self._launch(b)

    def finish(self) -> None:
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
                if not self.avg:
                    b["flat"].div_(self.world)
            b["pending"], b["launched"] = b["n"], False

    def zero_grad(self) -> None:
        for b in self.buckets:
            b["flat"].zero_()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
