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


# ------------------------------------------------------------------------------------------------
# Channel partition of ONE long sequence (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
class _Exchange(torch.autograd.Function):
    """The one exchange step of the channel partition, as an all-to-all (and its own inverse as backward):

        to_channels=True :  x [B, n*D, L/G]  (this rank's sequence chunk, every channel)
                         -> y [B, n*w, L]    (this rank's channel slab {c, D+c, .., (n-1)D+c : c in slab}, whole sequence)
        to_channels=False:  the inverse.

    Everything outside the operator core (embedding, add + LayerNorm, in_proj, out_proj, MLP, head, loss) is
    per-position and runs on the rank's L/G chunk; the core (short filter, gates, long convolution, filter columns) is
    per-channel and runs on the rank's D/G slab over the whole sequence.  An all-to-all moves every element once
    (each rank sends (G-1)/G of its [n*D, L/G] block), where the all-gather + reduce-scatter form of SURVEY 8(e)
    moves G times as much: at D = 256, L = 1 M, G = 8 that is 168 + 56 MB sent per rank and layer instead of 2 x 448 MB.
    Reference sites: in_proj hyena.py:441, out_proj :504 (both stay whole, applied to a sequence chunk)."""

    @staticmethod
    def forward(ctx, x, part, n, to_channels):
        ctx.part, ctx.n, ctx.to_channels = part, n, to_channels
        return part._exchange(x, n, to_channels)

    @staticmethod
    def backward(ctx, dy):
        return ctx.part._exchange(dy.contiguous(), ctx.n, not ctx.to_channels), None, None, None


class ChannelPartition:
    """Handle of the channel partition: rank / world of `process_group` (default group when None)."""

    def __init__(self, process_group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ChannelPartition needs an initialised torch.distributed process group")
        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        self.bytes_sent = 0          # payload this rank handed to all_to_all_single so far (bench accounting)
        self._timing = False
        self._events = []

    def enable_timing(self, on: bool):
        """bench accounting: bracket every all-to-all with CUDA events on the current stream"""
        self._timing = bool(on)
        if not on:
            self._events = []

    def drain_timing(self) -> float:
        """milliseconds spent between the event pairs recorded since the last drain (synchronises)"""
        if not self._events:
            return 0.0
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        self._events = []
        return ms

    def _a2a(self, recv, send):
        if self._timing and send.is_cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_to_all_single(recv, send, group=self.group)
            b.record()
            self._events.append((a, b))
        else:
            dist.all_to_all_single(recv, send, group=self.group)
        self.bytes_sent += send.numel() * send.element_size() * (self.world - 1) // self.world

    def slab(self, d_model):
        return channel_slab(d_model, self.rank, self.world)

    def chunk(self, L):
        """[lo, hi) of the sequence positions this rank owns outside the operator core."""
        if L % self.world:
            raise ValueError(f"sequence length {L} must divide evenly over {self.world} ranks")
        c = L // self.world
        return self.rank * c, (self.rank + 1) * c

    def slab_rows(self, d_model, n, device):
        """row indices {g*D + c : g < n, c in slab} of a [n*D, ...] parameter (in_proj rows / short_filter channels)."""
        lo, hi = self.slab(d_model)
        return torch.cat([torch.arange(g * d_model + lo, g * d_model + hi, device=device) for g in range(n)])

    def _exchange(self, x, n, to_channels):
        G = self.world
        B = x.shape[0]
        if to_channels:
            C, Lc = x.shape[1], x.shape[2]
            D = C // n
            w = D // G
            if D * n != C or w * G != D:
                raise ValueError(f"channel axis {C} is not {n} groups of a multiple of {G}")
            send = x.reshape(B, n, G, w, Lc).permute(2, 0, 1, 3, 4).contiguous()          # [G(dst slab), B, n, w, Lc]
            recv = torch.empty_like(send)
            self._a2a(recv, send)                                                           # [G(src chunk), B, n, w, Lc]
            return recv.permute(1, 2, 3, 0, 4).reshape(B, n * w, G * Lc)
        nw, L = x.shape[1], x.shape[2]
        w = nw // n
        Lc = L // G
        if w * n != nw or Lc * G != L:
            raise ValueError(f"[{nw}, {L}] is not {n} groups x {G} sequence chunks")
        send = x.reshape(B, n, w, G, Lc).permute(3, 0, 1, 2, 4).contiguous()                # [G(dst chunk), B, n, w, Lc]
        recv = torch.empty_like(send)
        self._a2a(recv, send)                                                               # [G(src slab), B, n, w, Lc]
        return recv.permute(1, 2, 0, 3, 4).reshape(B, n * G * w, Lc)

    def to_channels(self, x, n=1):
        return _Exchange.apply(x, self, n, True)

    def to_sequence(self, x, n=1):
        return _Exchange.apply(x, self, n, False)


def set_channel_partition(module, part):
    """Switch every HyenaOperator under `module` to the channel partition `part` (None = off).  The caller then feeds
    each rank its own sequence chunk (`part.chunk(L)`) and sums gradients over ranks (FlatGradAllReduce.allreduce(
    average=False) with a loss normalised by the GLOBAL token count): parameters stay replicated, every rank produces
    the gradient contribution of its chunk (per-position layers) or its slab (per-channel parameters)."""
    from .hyena import HyenaOperator
    n = 0
    for m in module.modules():
        if isinstance(m, HyenaOperator):
            m.channel_partition = part
            n += 1
    return n
