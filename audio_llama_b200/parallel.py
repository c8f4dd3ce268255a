"""Data parallelism for the path: one process per GPU, clips sharded contiguously, frozen weights
replicated. The forward / throughput path needs NO collective (clips are independent; the only reduction, the
mel max, is per clip). The training step has exactly one exchange: a sum-allreduce of the small trainable set
(projector + LoRA, 95.7 M parameters at the README config) over one flat bucket — NCCL over NVLink on the box,
gloo in the CPU tests. The reference has no distributed code at all (grep finds none), so this is new.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of [0, n_items): rank r gets [lo, hi); the first n_items % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Slice every tensor (and the metadata list) of a reference-layout batch dict for this rank."""
    n = batch["input_ids"].shape[0]
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in batch.items():
        out[k] = v[lo:hi]
    return out


class FlatGradBucket:
    """All trainable gradients viewed through ONE contiguous buffer, so the step's exchange is a single
    allreduce (size it for launch latency, not link count: NVSwitch gives every peer full bandwidth)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype=torch.float32):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dtype, device=dev)
        off = 0
        for p in self.params:           # gradients become views into the bucket: no pack / unpack copies
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        """Use this (or optimizer.zero_grad(set_to_none=False)) between steps: the gradients stay views of the bucket."""
        self.rebind()
        self.flat.zero_()

    def rebind(self) -> int:
        """Re-establish p.grad as a view of the bucket for every parameter whose gradient was detached from it.
        The reference loop calls optimizer.zero_grad() (train.py:300), whose default set_to_none=True drops the
        views; autograd then allocates fresh gradient tensors OUTSIDE the bucket and an all-reduce of the bucket
        would average stale zeros. A detached gradient is copied into its slot (a None gradient zeroes the slot);
        returns how many parameters had to be re-bound."""
        off, n = 0, 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            g = p.grad
            if g is None or g.data_ptr() != view.data_ptr() or g.dtype != self.flat.dtype:
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
                p.grad = view
                n += 1
            off += p.numel()
        return n

    def allreduce_mean(self, group=None):
        """sum over ranks then / world (the gradient of the mean loss over the global batch). Gradients that were
        detached from the bucket since the last call (zero_grad(set_to_none=True)) are gathered back first."""
        self.rebind()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat
