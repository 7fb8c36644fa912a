"""Data-parallel gradient all-reduce — the one collective on the path.

The reference trains with Lightning's DDPStrategy (train.py:630-639: bucketed NCCL all-reduce with
find_unused_parameters=True) and exports NCCL_P2P_DISABLE=1 (train.py:3).  HyenaDNA has 0.4-6.6 M
parameters (1.7-26 MB of fp32 gradients), so on NVLink5/NVSwitch the step is latency- not
bandwidth-bound: all gradients live in ONE flat fp32 buffer (each `p.grad` is a view into it) and a
single `all_reduce` over NCCL (P2P/NVLS enabled) averages them — no bucketing, no graph walk, no
copies.  Sequences shard over ranks along batch with no communication inside the operator.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        seen = set()
        self.params: List[torch.nn.Parameter] = []
        for p in params:
            if p.requires_grad and id(p) not in seen:       # tied weights appear once
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.group = process_group
        off = 0
        for p in self.params:
            n = p.numel()
            if p.dtype != torch.float32:
                raise TypeError("FlatGradAllReduce expects fp32 master parameters")
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach the views if something replaced p.grad (e.g. optimizer.zero_grad(set_to_none=True))."""
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += n

    def allreduce(self, average: bool = True):
        """Sum (or average) the gradients of all ranks in place; no-op without a process group."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(self.group)
        if world == 1:
            return None
        self.rebind()
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        if average:
            self.flat.div_(world)
        return work


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `n_items` sequences for `rank` (weak scaling: batch axis)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def channel_slab(d_model: int, rank: int, world: int):
    """Channel slab [lo, hi) of a single long sequence for `rank` (north_star: channel partition of the
    long convolution at B = 1; conv, short filter, gates and the filter's output columns are all
    per-channel)."""
    if d_model % world:
        raise ValueError("d_model must divide evenly over ranks")
    w = d_model // world
    return rank * w, (rank + 1) * w
