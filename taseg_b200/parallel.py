"""Data parallelism for the hot path: one process per GPU, scans sharded across ranks, and ONE
collective — the weight-gradient all-reduce of the training step (the reference's DistributedDataParallel,
R/train.py:247-251).  Inference has no data-path collective at all.

`torch.distributed` is only plumbing here: NCCL over NVLink on the GPUs, gloo in the CPU tests (tests/test_parallel.py).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from .utils.lazy import LazyScalar


def shard_indices(n_items: int, rank: int, world: int, pad: bool = True) -> List[int]:
    """Items of rank `rank`: torch DistributedSampler semantics without shuffling (the reference's eval sampler,
    R/pcseg/data/__init__.py:136-141): round-robin, padded by wrapping so every rank gets ceil(n/world) items."""
    if world <= 1:
        return list(range(n_items))
    idx = list(range(n_items))
    if pad and n_items:
        total = math.ceil(n_items / world) * world
        idx = (idx * math.ceil(total / n_items))[:total]
    return idx[rank::world]


def max_over_ranks(value_ms: float, device) -> float:
    """Time of a step = max over ranks (bench.py contract)."""
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class GradientReducer:
    """Bucketed, overlapped gradient all-reduce (SUM then / world) — the one exchange step of the training path.

    Parameters are packed into buckets of ~`bucket_mb` in REVERSE registration order (gradients become ready roughly
    in that order during backward).  A post-accumulate-grad hook on every parameter counts its bucket down; when a
    bucket is complete its gradients are flattened into one contiguous buffer and `all_reduce` is issued asynchronously,
    so the transfer of bucket b overlaps the wgrad kernels of the layers below it.  `finish()` waits for the
    outstanding work, divides by the world size and scatters the averaged gradients back into `.grad`.
    With world size 1 (or no process group) it is a no-op."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[List[torch.nn.Parameter]] = []
        cap = int(bucket_mb * 1024 * 1024)
        cur, size = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > cap:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): b for b, ps in enumerate(self.buckets) for p in ps}
        self._pending = [len(b) for b in self.buckets]
        self._inflight: List[tuple] = []
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        # bucket membership is FIXED (every rank exchanges the same number of elements whatever it used this step): a
        # parameter without a gradient on this rank contributes zeros, as DistributedDataParallel does for unused ones
        ps = self.buckets[b]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in ps])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((work, flat, ps))

    def finish(self) -> None:
        """Call after backward(): completes the exchange; gradients are then identical on every rank."""
        if self.world <= 1:
            return
        for b, left in enumerate(self._pending):     # buckets with unused parameters never fired
            if left > 0:
                self._launch(b)
        for work, flat, ps in self._inflight:
            work.wait()
            flat /= self.world
            off = 0
            for p in ps:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]

    def bytes_per_step(self) -> int:
        return sum(p.numel() * 4 for p in self.params)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def train_step(model: torch.nn.Module, batch_dict: dict, optimizer: torch.optim.Optimizer,
               reducer: Optional[GradientReducer] = None, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
               clip_grad_norm: Optional[float] = 10.0, comm_events: Optional[list] = None, sync: bool = True):
    """One training step of the reference loop (R/train.py:399-417) with bf16 autocast instead of fp16 + GradScaler:
    forward -> loss -> backward (gradient buckets all-reduced while backward runs) -> clip -> optimizer step.
    comm_events: optional list that receives one (start, end) CUDA event pair per step around `reducer.finish()`.
    sync=True returns the loss as a float (the host waits for the step); sync=False returns a LazyScalar, so a loop that
    logs the loss of step i after it has enqueued step i + 1 never leaves the GPU without queued work."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    dev_type = next(model.parameters()).device.type
    with torch.autocast(device_type=dev_type, dtype=amp_dtype, enabled=amp_dtype is not None):
        ret_dict, _, _ = model(batch_dict)
        loss = ret_dict['loss']
    loss.backward()
    if reducer is not None:
        if comm_events is not None:     # device time between "backward has been issued" and "averaged gradients are in place":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # the EXPOSED part of the exchange
            e0.record()
        reducer.finish()
        if comm_events is not None:
            e1.record()
            comm_events.append((e0, e1))
    if clip_grad_norm:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_grad_norm)
    optimizer.step()
    return float(loss.detach()) if sync else LazyScalar(loss)
