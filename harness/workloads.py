"""Model-level workloads of BASELINE.json (configs 1, 2, 3, 5) on top of the drop-in: a Lightning-free restatement of what
`train_synapse.py` / `train_acdc.py` / `inference.py` do per step, driving the UNMODIFIED reference model
(harness/refmodel.py) with synthetic tensors of the real shapes (there is no dataset and no lightning / monai in this image).

  train step   train_synapse.py:140-151 (`training_step`: H2D of image + label, forward, DiceCELoss(ce 0.4, dc 0.6) :90-93,
               `loss.item()` logged) + Lightning's backward / AdamW(lr 5e-4, wd 1e-3) :102-108 step / zero_grad;
               ACDC: train_acdc.py (4 classes, wd 1e-4). Under N > 1 ranks the gradients are averaged by
               ceigm_unet_b200.dist.GradReducer (NCCL all-reduce overlapped with backward) — Lightning's default DDP strategy.
  inference    eval.py:72-77 / inference.py:38-112 per slice: H2D, forward, argmax(softmax) -> label map, D2H;
               here batched (config 5: 512 x 512, batch 64 per GPU).
"""
from __future__ import annotations

import time

import torch

from . import refmodel


def build(num_classes: int = 9, level: str = "dropin", device="cuda", seed: int = 42):
    """level: "dropin" (reference modules, scans through libss2d_b200.so) | "fused" (GroupMambaLayer := this repo's fused
    module) | "cpu_ref" (reference PyTorch scan on the host: the CPU baseline)."""
    if level == "cpu_ref":
        m = refmodel.load_reference(scan="cpu_ref")
    else:
        m = refmodel.load_reference(scan="dropin", fused=(level == "fused"))
    import contextlib
    import io
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):           # the reference prints a banner and a checkpoint path
        net = m.build_model(in_channels=3, num_classes=num_classes)
    if level == "fused":
        refmodel.reclass_layernorms(net)
    return net.to(device)


def synthetic_batch(batch: int, size: int, num_classes: int, seed: int = 42, pin: bool = True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, size, size, generator=g)
    y = torch.randint(0, num_classes, (batch, 1, size, size), generator=g).float()
    if pin and torch.cuda.is_available():
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


class TrainStep:
    """One optimisation step of the reference model on one rank's batch shard."""

    def __init__(self, net, num_classes=9, lr=5e-4, weight_decay=1e-3, amp_dtype=torch.bfloat16, reducer=None):
        self.net = net.train()
        self.crit = refmodel.load_losses().DiceCELoss(ce_weight=0.4, dc_weight=0.6)
        self.opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=weight_decay, eps=1e-8, betas=(0.9, 0.999))
        self.amp_dtype = amp_dtype
        self.reducer = reducer
        self.device = next(net.parameters()).device

    def __call__(self, x_host, y_host) -> float:
        x = x_host.to(self.device, non_blocking=True)
        y = y_host.to(self.device, non_blocking=True)
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        with torch.autocast(self.device.type, dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            pred = self.net(x)
            loss = self.crit(pred.float(), y)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return float(loss.item())            # D2H read of the step's result (train_synapse.py:146)


def _sync_all(world, device):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)


def train_bench(device, rank=0, world=1, *, per_gpu_batch=24, size=224, num_classes=9, steps=5, warmup=3, level="dropin",
                amp_dtype=torch.bfloat16, weight_decay=1e-3, graphs=False):
    """-> dict with whole-job slices/s (max-over-ranks device time), loss trace, all-reduce share.
    graphs=True: the step is replayed from CUDA graphs (harness/graph_step.py)."""
    import ceigm_unet_b200 as pkg
    from ceigm_unet_b200 import dist as D
    net = build(num_classes, level, device)
    reducer = D.GradReducer(net) if world > 1 and not graphs else None
    if world > 1:      # same initial weights on every rank, as DDP's constructor broadcast does
        for p in list(net.parameters()) + list(net.buffers()):
            torch.distributed.broadcast(p.data, src=0)
    if graphs:
        from . import graph_step
        graph_step.make_capturable()
        step = graph_step.GraphedTrainStep(net, per_gpu_batch, size, num_classes, amp_dtype=amp_dtype, world=world,
                                           weight_decay=weight_decay)
    else:
        step = TrainStep(net, num_classes, amp_dtype=amp_dtype, reducer=reducer, weight_decay=weight_decay)
    x, y = synthetic_batch(per_gpu_batch, size, num_classes, seed=42 + rank)
    losses = []
    for _ in range(max(warmup, 3)):
        losses.append(step(x, y))
    _sync_all(world, device)
    pkg.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        losses.append(step(x, y))
    e1.record()
    _sync_all(world, device)
    wall = time.perf_counter() - t0
    ms = D.max_over_ranks(e0.elapsed_time(e1), device) / steps
    launches = pkg.launch_count()
    # all-reduce alone (same buckets, nothing to overlap with) for its share of the step
    ar_ms = None
    if world > 1:
        one_allreduce = reducer.finish if reducer is not None else step.allreduce
        _sync_all(world, device)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            one_allreduce()
        a1.record()
        _sync_all(world, device)
        ar_ms = D.max_over_ranks(a0.elapsed_time(a1), device) / 3
        if reducer is not None:
            reducer.remove()
    out = {"slices_per_s": round(world * per_gpu_batch / (ms * 1e-3), 2), "ms_per_step": round(ms, 2),
           "wall_ms_per_step": round(wall / steps * 1e3, 2), "per_gpu_batch": per_gpu_batch, "size": size,
           "num_classes": num_classes, "level": level, "cuda_graphs": bool(graphs), "amp": str(amp_dtype).replace("torch.", "") if amp_dtype else "fp32",
           "steps": steps, "loss_first": round(losses[0], 5), "loss_last": round(losses[-1], 5),
           "ss2d_launches_per_step": launches // steps,
           "h2d_bytes_per_step": x.numel() * 4 + y.numel() * 4, "d2h_bytes_per_step": 4,
           "optimizer": "AdamW lr 5e-4", "loss": "DiceCELoss(0.4, 0.6)"}
    if world > 1:
        out["allreduce"] = {"bytes": reducer.bytes if reducer is not None else step.allreduce_bytes,
                            "buckets": len(reducer.buckets) if reducer is not None else 1, "alone_ms": round(ar_ms, 3),
                            "share_of_step_if_exposed": round(ar_ms / ms, 4), "backend": torch.distributed.get_backend(),
                            "overlapped_with_backward": not graphs}
    del step, net
    torch.cuda.empty_cache()
    return out


@torch.no_grad()
def infer_bench(device, rank=0, world=1, *, per_gpu_batch=64, size=512, num_classes=9, steps=3, warmup=2, level="dropin",
                amp_dtype=torch.bfloat16, graphs=False):
    from ceigm_unet_b200 import dist as D
    net = build(num_classes, level, device).eval()
    x, _ = synthetic_batch(per_gpu_batch, size, num_classes, seed=7 + rank)
    out_host = torch.empty((per_gpu_batch, size, size), dtype=torch.uint8).pin_memory()
    ginf = None
    if graphs:
        from . import graph_step
        graph_step.make_capturable()
        ginf = graph_step.GraphedInference(net, per_gpu_batch, size, amp_dtype=amp_dtype)

    def one():
        if ginf is not None:
            return ginf(x, out_host)
        xd = x.to(device, non_blocking=True)
        with torch.autocast(device.type if hasattr(device, "type") else "cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            logits = net(xd)
        lab = torch.argmax(torch.softmax(logits.float(), dim=1), dim=1).to(torch.uint8)      # eval.py:76
        out_host.copy_(lab, non_blocking=True)
        return lab
    for _ in range(max(warmup, 1)):
        one()
    _sync_all(world, device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    _sync_all(world, device)
    ms = D.max_over_ranks(e0.elapsed_time(e1), device) / steps
    res = {"slices_per_s": round(world * per_gpu_batch / (ms * 1e-3), 2), "ms_per_step": round(ms, 2),
           "per_gpu_batch": per_gpu_batch, "size": size, "level": level, "cuda_graphs": bool(graphs),
           "amp": str(amp_dtype).replace("torch.", "") if amp_dtype else "fp32", "steps": steps,
           "h2d_bytes_per_step": x.numel() * 4, "d2h_bytes_per_step": out_host.numel(),
           "peak_mem_GB": round(torch.cuda.max_memory_allocated(device) / 1e9, 2)}
    del net
    torch.cuda.empty_cache()
    return res


def cpu_reference_step(num_classes=9, size=224, threads=None):
    """BASELINE config 1: the reference model, batch 1, fwd + bwd + DiceCE on the host through the reference's PyTorch
    `selective_scan_ref` path (restated in oracle/selective_scan_ref.py) — the CPU column of the model-level numbers."""
    import os
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = build(num_classes, "cpu_ref", "cpu").train()
    crit = refmodel.load_losses().DiceCELoss(ce_weight=0.4, dc_weight=0.6)
    x, y = synthetic_batch(1, size, num_classes, pin=False)
    t0 = time.perf_counter()
    out = net(x)
    loss = crit(out, y)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    return {"slices_per_s": round(1.0 / (t2 - t0), 4), "fwd_s": round(t1 - t0, 2), "bwd_s": round(t2 - t1, 2), "cores": threads,
            "batch": 1, "size": size, "loss": round(float(loss), 5), "kind": "reference python path (selective_scan_ref + autograd)"}
