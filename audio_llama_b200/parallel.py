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


def init_nccl(device) -> None:
    """One process per GPU: NCCL process group whose collectives run on a HIGH-PRIORITY stream, so that the chunk
    all-reduces launched during the backward pass (FlatGradBucket.arm_overlap) get SMs between the compute kernels
    instead of queueing behind them. Rendezvous comes from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:
        opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=device)


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

    # ------------------------------------------------------------------ overlapped form
    def arm_overlap(self, n_chunks: int = 4, group=None):
        """Overlap the exchange with the backward pass: the bucket is cut into `n_chunks` contiguous chunks (parameter
        boundaries), and a chunk's all-reduce is launched (async, on the communicator's own stream) from the
        post-accumulate-grad hook of the LAST of its parameters to receive its gradient. Parameters are laid out in the
        order given to the constructor; backward produces gradients roughly in reverse order (last LLaMA layers first,
        the projector last), so chunks complete from the back and the early ones ride under the remaining backward.
        Call once; then every step: backward(), finish_overlap() (waits, divides by the world size), clip, step,
        zero() or zero_grad(set_to_none=False). Requires the gradients to stay views of the bucket."""
        if getattr(self, "_ov", None) is not None:
            return self
        n = len(self.params)
        n_chunks = max(1, min(n_chunks, n))
        # chunk boundaries: equal element counts, snapped to parameter boundaries
        sizes = [p.numel() for p in self.params]
        target = self.numel / n_chunks
        bounds, acc, start = [], 0, 0
        for i, sz in enumerate(sizes):
            acc += sz
            if acc >= target * (len(bounds) + 1) and len(bounds) < n_chunks - 1:
                bounds.append((start, i + 1))
                start = i + 1
        bounds.append((start, n))
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + sz)
        chunks = []
        for (a, b) in bounds:
            chunks.append(dict(lo=offs[a], hi=offs[b], n_params=b - a, pending=b - a, work=None))
        owner = {}
        for ci, (a, b) in enumerate(bounds):
            for i in range(a, b):
                owner[i] = ci
        self._ov = dict(chunks=chunks, group=group, handles=[])

        def make_hook(ci):
            def hook(param):
                ch = self._ov["chunks"][ci]
                ch["pending"] -= 1
                if ch["pending"] == 0 and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                    ch["work"] = dist.all_reduce(self.flat[ch["lo"]:ch["hi"]], op=dist.ReduceOp.SUM, group=group, async_op=True)
            return hook

        for i, p in enumerate(self.params):
            self._ov["handles"].append(p.register_post_accumulate_grad_hook(make_hook(owner[i])))
        return self

    def disarm_overlap(self):
        """Remove the gradient hooks (back to the explicit allreduce_mean() form)."""
        ov = getattr(self, "_ov", None)
        if ov is not None:
            for h in ov["handles"]:
                h.remove()
            self._ov = None

    def finish_overlap(self):
        """Wait for the chunk all-reduces of this step (launching those whose parameters received no gradient), divide by
        the world size, re-arm. Returns the flat gradient."""
        ov = self._ov
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(ov["group"]) > 1
        if self.rebind() != 0 and multi:
            raise RuntimeError("FlatGradBucket: gradients were detached from the bucket during an overlapped step "
                               "(use bucket.zero() or zero_grad(set_to_none=False))")
        ov["launched_in_backward"] = sum(1 for ch in ov["chunks"] if ch["work"] is not None)
        for ch in ov["chunks"]:
            if multi:
                if ch["work"] is None:             # some parameter of the chunk got no gradient this step
                    ch["work"] = dist.all_reduce(self.flat[ch["lo"]:ch["hi"]], op=dist.ReduceOp.SUM, group=ov["group"], async_op=True)
                ch["work"].wait()
            ch["work"] = None
            ch["pending"] = ch["n_params"]
        if multi:
            self.flat.div_(dist.get_world_size(ov["group"]))
        return self.flat

    def allreduce_mean(self, group=None):
        """sum over ranks then / world (the gradient of the mean loss over the global batch). Gradients that were
        detached from the bucket since the last call (zero_grad(set_to_none=True)) are gathered back first."""
        self.rebind()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat
