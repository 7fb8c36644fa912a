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
                view.zero_()            # zero_grad(set_to_none=True): the slot still holds the previous step's gradient
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
        from .hyena import flush_filter_grads
        flush_filter_grads()              # deferred filter backward of operators with filter_reuse: before the exchange
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
    def forward(ctx, x, part, n, to_channels, lane=0):
        ctx.part, ctx.n, ctx.to_channels, ctx.lane = part, n, to_channels, lane
        return part._exchange(x, n, to_channels, lane=lane)

    @staticmethod
    def backward(ctx, dy):
        return ctx.part._exchange(dy.contiguous(), ctx.n, not ctx.to_channels, lane=ctx.lane), None, None, None, None


class _ShapeOnly:
    """shape / dtype / device of a tensor the producer has yet to write (see ChannelPartition.produced_to_channels)"""

    def __init__(self, shape, dtype, device):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device

    def element_size(self):
        return torch.empty((), dtype=self.dtype).element_size()

    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n


class PeerExchange:
    """The exchange over NVLink peer memory (csrc/hy_exchange.cu): every rank owns ONE peer-mappable allocation
    [flag block | two payload buffers], maps the others' through CUDA IPC, and pulls the rows it owns straight into
    their final layout with one kernel — no pack / unpack passes, no NCCL call on the data path.  The payload buffers
    alternate between consecutive exchanges; a buffer is overwritten only after every peer has reported (in this
    rank's flag block) that it finished reading the exchange two epochs back."""

    def __init__(self, group, rank, world, device, max_ctas=0):
        import ctypes as C
        from . import _lib
        self.C, self._lib = C, _lib
        self.group, self.rank, self.world, self.device = group, rank, world, device
        self.max_ctas = int(max_ctas)
        if world > 8:
            raise ValueError("PeerExchange: at most 8 ranks (one NVSwitch box)")
        self.base = None
        self.peer_base = []
        self.cap = 0
        self.epoch = 0
        self.used = [None, None]          # epoch at which payload buffer p was last exposed

    def _ensure(self, nbytes):
        """(re)allocate when a payload does not fit — collective: every rank grows at the same exchange"""
        if nbytes <= self.cap:
            return
        C, lib = self.C, self._lib.lib()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.close()
        cap = (int(nbytes * 1.0) + 4095) // 4096 * 4096
        flag = int(lib.hy_peer_flag_bytes())
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            self._lib.check(lib.hy_peer_alloc(C.c_size_t(flag + 2 * cap), C.byref(ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), cap), group=self.group)
        if any(h[1] != cap for h in handles):
            raise RuntimeError("PeerExchange: ranks disagree on the payload size")
        self.base = ptr.value
        self.peer_base = []
        for j, (h, _) in enumerate(handles):
            if j == self.rank:
                self.peer_base.append(self.base)
            else:
                q = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                with torch.cuda.device(self.device):
                    self._lib.check(lib.hy_peer_open(buf, C.byref(q)))
                self.peer_base.append(q.value)
        self.flag_bytes, self.cap = flag, cap
        self.epoch = 0
        self.used = [None, None]
        dist.barrier(group=self.group)

    def close(self):
        if self.base is None:
            return
        lib = self._lib.lib()
        for j, q in enumerate(self.peer_base):
            if j != self.rank:
                lib.hy_peer_close(self.C.c_void_p(q))
        lib.hy_peer_free(self.C.c_void_p(self.base))
        self.base, self.peer_base, self.cap = None, [], 0

    def _payload(self, p, shape, dtype):
        """torch view of this rank's payload buffer p"""
        n = 1
        for d in shape:
            n *= d
        item = torch.empty((), dtype=dtype).element_size()
        typestr = {torch.bfloat16: "<u2", torch.float16: "<f2", torch.float32: "<f4"}[dtype]

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "version": 2,
                                        "data": (self.base + self.flag_bytes + p * self.cap, False)}
        t = torch.as_tensor(_Arr(), device=self.device)
        if dtype == torch.bfloat16:
            t = t.view(torch.bfloat16)
        assert n * item <= self.cap
        return t.view(shape)

    def begin(self, shape, dtype):
        """The payload view the producer must fill for the next exchange (stream-ordered after the peers' reads of the
        exchange that last used this buffer)."""
        n = 1
        for d in shape:
            n *= d
        self._ensure(n * torch.empty((), dtype=dtype).element_size())
        p = self.epoch & 1
        if self.used[p] is not None:
            self._lib.check(self._lib.lib().hy_peer_wait_done(self.C.c_void_p(self.base), self.world, self.rank,
                                                              self.C.c_uint(self.used[p] & 0xffffffff), self._lib.current_stream_ptr()))
        return self._payload(p, shape, dtype)

    def pull(self, dst, *, n_outer, n_inner, row_bytes, src_base, src_outer, src_inner, dst_outer, dst_inner, dst_peer,
             reduce=False):
        """Run the exchange whose payload was filled after begin(); `dst` receives this rank's rows."""
        C = self.C
        a = self._lib.PeerPullArgs()
        p = self.epoch & 1
        self.epoch += 1
        for j in range(self.world):
            a.src[j] = self.peer_base[j] + self.flag_bytes + p * self.cap
            a.flags[j] = self.peer_base[j]
        a.G, a.self, a.epoch, a.reduce = self.world, self.rank, self.epoch & 0xffffffff, int(reduce)
        a.n_outer, a.n_inner, a.row_bytes = n_outer, n_inner, row_bytes
        a.src_base, a.src_outer, a.src_inner = src_base, src_outer, src_inner
        a.dst_outer, a.dst_inner, a.dst_peer = dst_outer, dst_inner, dst_peer
        a.dst = dst.data_ptr()
        a.max_ctas = self.max_ctas
        self.used[p] = self.epoch
        self._lib.check(self._lib.lib().hy_peer_pull(C.byref(a), self._lib.current_stream_ptr()))
        return dst

    def check(self):
        """raise if one of this rank's exchange kernels gave up waiting for a peer (synchronises)"""
        if self.base is not None and self._lib.lib().hy_peer_error(self.C.c_void_p(self.base)) != 0:
            raise RuntimeError("PeerExchange: a peer never arrived at an exchange (spin timed out)")


class ChannelPartition:
    """Handle of the channel partition: rank / world of `process_group` (default group when None).

    backend: "peer" = pull kernels over NVLink peer memory (PeerExchange; CUDA tensors, one box), "nccl" = packed
    torch.distributed all_to_all_single (any backend, incl. gloo on CPU); None = "peer" for CUDA tensors unless
    HYENA_B200_PEER_EXCHANGE=0."""

    def __init__(self, process_group=None, backend=None):
        import os
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ChannelPartition needs an initialised torch.distributed process group")
        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        if backend is None:
            backend = "nccl" if os.environ.get("HYENA_B200_PEER_EXCHANGE", "1") == "0" else "peer"
        self.backend = backend
        self._peers = {}
        self._side = {}
        # lane 1 = exchanges issued on a side stream beside compute kernels (the implicit-filter path): own buffers,
        # flags and epoch counter, because two streams do not order their kernels alike on every rank
        self.overlap = backend == "peer" and os.environ.get("HYENA_B200_EXCHANGE_OVERLAP", "1") != "0"
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

    def peer(self, device, lane=0):
        if lane not in self._peers:
            self._peers[lane] = PeerExchange(self.group, self.rank, self.world, device, max_ctas=0 if lane == 0 else 32)
        return self._peers[lane]

    def check(self):
        """raise if an exchange kernel of this rank gave up waiting for a peer (synchronises)"""
        for px in self._peers.values():
            px.check()

    def side_stream(self, device):
        key = (device.type, device.index)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=device)
        return self._side[key]

    def _exchange_peer(self, x, n, to_channels, produce=None, lane=0):
        """the same exchange as below with ONE pull kernel reading the peers' buffers in place.  `produce(view)`: the
        producer writes the payload straight into the exposed buffer (x is then only a shape / dtype / device carrier)"""
        G, r = self.world, self.rank
        B = x.shape[0]
        es = x.element_size()
        px = self.peer(x.device, lane)
        src = px.begin(tuple(x.shape), x.dtype)
        if produce is not None:
            produce(src)
        else:
            src.copy_(x)                              # producer -> exposed buffer (local pass)
        ev = self._mark()
        if to_channels:
            C, Lc = x.shape[1], x.shape[2]
            D = C // n
            w = D // G
            if D * n != C or w * G != D:
                raise ValueError(f"channel axis {C} is not {n} groups of a multiple of {G}")
            out = torch.empty((B, n * w, G * Lc), dtype=x.dtype, device=x.device)
            px.pull(out, n_outer=B * n, n_inner=w, row_bytes=Lc * es, src_base=r * w * Lc * es, src_outer=D * Lc * es,
                    src_inner=Lc * es, dst_outer=w * G * Lc * es, dst_inner=G * Lc * es, dst_peer=Lc * es)
        else:
            nw, L = x.shape[1], x.shape[2]
            w = nw // n
            Lc = L // G
            if w * n != nw or Lc * G != L:
                raise ValueError(f"[{nw}, {L}] is not {n} groups x {G} sequence chunks")
            out = torch.empty((B, n * G * w, Lc), dtype=x.dtype, device=x.device)
            px.pull(out, n_outer=B * n, n_inner=w, row_bytes=Lc * es, src_base=r * Lc * es, src_outer=w * L * es,
                    src_inner=L * es, dst_outer=G * w * Lc * es, dst_inner=Lc * es, dst_peer=w * Lc * es)
        self._mark(ev)
        self.bytes_sent += x.numel() * es * (G - 1) // G
        return out

    def _mark(self, ev=None):
        if not self._timing:
            return None
        if ev is None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev
        b = torch.cuda.Event(enable_timing=True)
        b.record()
        self._events.append((ev, b))
        return None

    def _probe_peer(self, device):
        """once, collectively, at the first exchange: can every rank map every peer's memory (CUDA IPC over NVLink)?  If
        any rank cannot (containers without a shared IPC namespace, GPUs without peer access), ALL ranks take the packed
        NCCL all-to-all instead — said on stderr, never silently."""
        self._probed = True
        ok = 1
        try:
            px = PeerExchange(self.group, self.rank, self.world, device)
            px._ensure(1 << 20)
            px.close()
        except Exception as e:  # noqa: BLE001
            ok = 0
            err = str(e)
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            import sys
            if self.rank == 0:
                print("[dna_b200.dp] peer-memory exchange unavailable on this box (%s): using NCCL all_to_all_single"
                      % (err if not ok else "a peer rank failed"), file=sys.stderr, flush=True)
            self.backend = "nccl"
            self.overlap = False

    def prepare(self, device):
        """settle the exchange back end before the first exchange (collective)"""
        if self.backend == "peer" and device.type == "cuda" and not getattr(self, "_probed", False):
            self._probe_peer(device)

    def _exchange(self, x, n, to_channels, lane=0):
        if self.backend == "peer" and x.is_cuda and not getattr(self, "_probed", False):
            self._probe_peer(x.device)
        if self.backend == "peer" and x.is_cuda:
            return self._exchange_peer(x, n, to_channels, lane=lane)
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

    def to_channels(self, x, n=1, lane=0):
        return _Exchange.apply(x, self, n, True, lane)

    def produced_to_channels(self, shape, dtype, device, n, produce):
        """to_channels of a tensor that does not exist yet: `produce(out)` must write it (no autograd; callers wrap this
        in their own autograd.Function).  Peer backend: written directly into the exposed buffer, no local copy."""
        if self.backend == "peer" and device.type == "cuda" and not getattr(self, "_probed", False):
            self._probe_peer(device)
        if self.backend == "peer" and device.type == "cuda":
            return self._exchange_peer(_ShapeOnly(shape, dtype, device), n, True, produce=produce)
        x = torch.empty(shape, dtype=dtype, device=device)
        produce(x)
        return self._exchange(x, n, True)

    def to_sequence(self, x, n=1, lane=0):
        return _Exchange.apply(x, self, n, False, lane)


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
