"""Drop-in for the reference's `src/models/sequence/hyena.py` (and the operator half of
`standalone_hyenadna.py`): same class names, constructor arguments, parameter / buffer names
(state_dict compatible, incl. the three aliased `implicit_filter.{1,3,5}.freq` keys and the `_optim`
hyper-parameter tags) and forward semantics — with the hot path executed by hand-written sm_100a
kernels through the C-ABI of libhyena_b200.so.

Forward of `HyenaOperator` (reference: hyena.py:436-508), order == 2:
    uT  = W_in @ u^T                         cuBLAS, written channel-major [B, 3D, L] (no transpose pass)
    z   = fused( short_filter(uT + b_in) -> x0,x1,v ; g = v*x1 ; y = k (*) g + bias*g ; z = y*x0 )
                                             ONE kernel family (hy_conv_fwd, SHORTCONV mode)
    out = z^T @ W_out^T + b_out              cuBLAS
The implicit filter k comes from the fused filter kernel (hy_filter_fwd) in channel-major layout and
goes through hy_filter_spectrum once per forward.  CUDA tensors only: no CPU fallback.
"""
from __future__ import annotations

import math
import os
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K
from ._lib import IN_PREGATE, IN_SHORTCONV, OUT_PLAIN, OUT_POSTGATE, OUT_SHORTCONV, IN_PLAIN
from .fftconv import circular_conv, fftconv_func, fftconv_ref  # noqa: F401  (re-exported like the reference module)


# ------------------------------------------------------------------------------------------------
# small pieces of the reference's support code the operator depends on
# ------------------------------------------------------------------------------------------------
class OptimModule(nn.Module):
    """Reference: src/utils/train.py:142-156 — tensors registered with lr == 0 become buffers,
    otherwise Parameters tagged with `_optim = {"lr": ..., "weight_decay": ...}` (read by
    train.py:468-487 when building optimizer groups)."""

    def register(self, name, tensor, lr=None, wd=0.0):
        if lr == 0.0:
            self.register_buffer(name, tensor)
        else:
            self.register_parameter(name, nn.Parameter(tensor))
            optim = {}
            if lr is not None:
                optim["lr"] = lr
            if wd is not None:
                optim["weight_decay"] = wd
            setattr(getattr(self, name), "_optim", optim)


def Activation(activation=None, size=None, dim=-1):
    """Subset of src/models/nn/components.py:95 sufficient for HyenaOperator(activation=...)."""
    table = {None: nn.Identity, "id": nn.Identity, "identity": nn.Identity, "linear": nn.Identity,
             "tanh": nn.Tanh, "relu": nn.ReLU, "gelu": nn.GELU, "swish": nn.SiLU, "silu": nn.SiLU,
             "sigmoid": nn.Sigmoid}
    if activation == "glu":
        return nn.GLU(dim=dim)
    if activation not in table:
        raise NotImplementedError(f"hidden activation '{activation}' is not implemented")
    return table[activation]()


class Sin(nn.Module):
    """Reference: hyena.py:100-110."""

    def __init__(self, dim, w=10, train_freq=True):
        super().__init__()
        self.freq = nn.Parameter(w * torch.ones(1, dim)) if train_freq else w * torch.ones(1, dim)

    def forward(self, x):
        return torch.sin(self.freq * x)


class PositionalEmbedding(OptimModule):
    """Reference: hyena.py:113-135 (tables built once for seq_len = l_max, sliced at run time)."""

    def __init__(self, emb_dim: int, seq_len: int, lr_pos_emb: float = 1e-5, **kwargs):
        super().__init__()
        self.seq_len = seq_len
        t = torch.linspace(0, 1, self.seq_len)[None, :, None]
        bands = (emb_dim - 1) // 2
        t_rescaled = torch.linspace(0, seq_len - 1, seq_len)[None, :, None]
        w = 2 * math.pi * t_rescaled / seq_len
        f = torch.linspace(1e-4, bands - 1, bands)[None, None]
        z = torch.exp(-1j * f * w)
        z = torch.cat([t, z.real, z.imag], dim=-1)
        self.register("z", z, lr=lr_pos_emb)
        self.register("t", t, lr=0.0)

    def forward(self, L):
        return self.z[:, :L], self.t[:, :L]


class ExponentialModulation(OptimModule):
    """Reference: hyena.py:138-159 (shift default 0.0) / standalone_hyenadna.py:119-144 (0.05)."""

    def __init__(self, d_model, fast_decay_pct=0.3, slow_decay_pct=1.5, target=1e-2, modulation_lr=0.0,
                 modulate: bool = True, shift: float = 0.0, **kwargs):
        super().__init__()
        self.modulate = modulate
        self.shift = shift
        max_decay = math.log(target) / fast_decay_pct
        min_decay = math.log(target) / slow_decay_pct
        deltas = torch.linspace(min_decay, max_decay, d_model)[None, None]
        self.register("deltas", deltas, lr=modulation_lr)

    def forward(self, t, x):
        if self.modulate:
            x = x * (torch.exp(-t * self.deltas.abs()) + self.shift)
        return x


# ------------------------------------------------------------------------------------------------
# autograd glue around the kernels
# ------------------------------------------------------------------------------------------------
def _mlp_layers(seq: nn.Sequential):
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    return lin


class _FilterFn(torch.autograd.Function):
    """k[D, L] = HyenaFilter.filter(L) in channel-major layout (fused kernel hy_filter_fwd).

    Backward: the MLP is ~0.03 M parameters; its gradient is obtained by re-running the same
    expression with device-side torch ops (cuBLAS GEMMs, fp32, autocast off) under autograd —
    nothing is saved between forward and backward except the inputs."""

    @staticmethod
    def forward(ctx, L, shift, modulate, normalized, z, t, deltas, freq, *wb):
        n_lin = (len(wb) + 1) // 2
        w_in, b_in = wb[0], wb[1]
        w_out = wb[-1]
        hidden = wb[2:-1]
        if hidden:
            w_h = torch.stack([h.detach().float() for h in hidden[0::2]])
            b_h = torch.stack([h.detach().float() for h in hidden[1::2]])
        else:
            w_h = b_h = None
        # when the backward will run the fused trunk kernel, keep the last hidden activation ([L, order] fp32) so that
        # nothing of the MLP is recomputed by torch there
        needs = ctx.needs_input_grad[4:]
        order, emb = w_in.shape
        save_h = (any(needs) and not (needs[0] or needs[1] or needs[2] or normalized)
                  and K.filter_trunk_bwd_supported(order, emb, n_lin - 2))
        # ... and, when it fits the budget (HYENA_B200_TRUNK_SAVE_MAX_MB, default 1024), the trunk's pre-activations:
        # the backward then skips the recompute of the trunk's Linear layers
        save_trunk = save_h and (K.filter_trunk_save_bytes(order, emb, n_lin - 2, L)
                                 <= int(os.environ.get("HYENA_B200_TRUNK_SAVE_MAX_MB", "1024")) << 20)
        k = K.filter_fwd(z.detach()[0], t.detach()[0], w_in.detach().float(), b_in.detach().float(), w_h, b_h,
                         w_out.detach().float(), freq.detach().float().reshape(-1), deltas.detach().float().reshape(-1),
                         shift, modulate, L, save_h=save_h, save_trunk=save_trunk)
        h_last = a_save = None
        if save_trunk:
            k, h_last, a_save = k
        elif save_h:
            k, h_last = k
        if normalized:
            k = k / k.abs().sum(dim=0, keepdim=True)
        ctx.cfg = (L, shift, modulate, normalized, n_lin)
        ctx.has_h = h_last is not None
        ctx.has_a = a_save is not None
        ctx.save_for_backward(z, t, deltas, freq, *wb, *([h_last] if h_last is not None else []),
                              *([a_save] if a_save is not None else []))
        return k

    @staticmethod
    def backward(ctx, dk):
        L, shift, modulate, normalized, n_lin = ctx.cfg
        saved = ctx.saved_tensors
        h_last = a_save = None
        if ctx.has_a:
            a_save, saved = saved[-1], saved[:-1]
        if ctx.has_h:
            h_last, saved = saved[-1], saved[:-1]
        needs = ctx.needs_input_grad[4:]
        if dk.stride(-1) != 1:
            dk = dk.contiguous()
        # z (index 0), t (1) and deltas (2) are buffers in every HyenaDNA config (lr_pos_emb = modulation_lr = 0)
        simple = not (needs[0] or needs[1] or needs[2] or normalized)
        with torch.enable_grad(), torch.autocast(device_type=dk.device.type, enabled=False):
            leaves = [s.detach().float().requires_grad_(bool(n)) for s, n in zip(saved, needs)]
            z, t, deltas, freq = leaves[:4]
            wb = leaves[4:]
            n_hidden = n_lin - 2
            order, emb = wb[0].shape
            fused_trunk = simple and K.filter_trunk_bwd_supported(order, emb, n_hidden)
            if fused_trunk and h_last is not None:
                h = h_last[None]                                # saved by the forward kernel: no recompute
            else:
                with torch.set_grad_enabled(not fused_trunk):  # fused trunk backward needs no autograd graph
                    h = z[:, :L]
                    for i in range(n_lin - 1):
                        h = torch.sin(freq * F.linear(h, wb[2 * i], wb[2 * i + 1]))
            if simple:
                # kernel: dh = dk^T * (decay + shift) in [L, D]; cuBLAS: the last Linear's two GEMMs; autograd: the
                # [L, order] trunk only — the [L, D]-sized elementwise passes of the first cut are gone.
                h2 = h[0]
                if fused_trunk and h_last is not None and K.filter_out_bwd_supported(wb[-1].shape[0], order):
                    # one tensor-core kernel (hy_filter_out_bwd): modulation backward + both GEMMs of the last Linear,
                    # dk streamed once, dh never materialised
                    dh2, g_wout = K.filter_out_bwd(dk.float(), saved[1], saved[2], shift, modulate, wb[-1].detach(),
                                                   h_last, L)
                    if not needs[-1]:
                        g_wout = None
                else:
                    dh = K.filter_modulate_bwd(dk.float(), saved[1], saved[2], shift, modulate, L)
                    g_wout = torch.matmul(dh.t(), h2.detach()) if needs[-1] else None
                    dh2 = torch.matmul(dh, wb[-1].detach())
                trunk_needs = needs[3:-1]                       # freq + every trunk weight / bias
                out = [None, None, None]                        # z, t, deltas: buffers on this path
                if fused_trunk:
                    # fused trunk backward (hy_filter_trunk_bwd): recompute + all parameter gradients in one kernel
                    w_h = torch.stack([wb[2 + 2 * i].detach() for i in range(n_hidden)]) if n_hidden else None
                    b_h = torch.stack([wb[3 + 2 * i].detach() for i in range(n_hidden)]) if n_hidden else None
                    dW_in, db_in, dW_h, db_h, dfreq = K.filter_trunk_bwd(dh2.contiguous(), saved[0], saved[1], wb[0].detach(),
                                                                        wb[1].detach(), w_h, b_h, wb[-1].detach(), freq.detach(), L,
                                                                        a_save=a_save)
                    gl = [dfreq.reshape(saved[3].shape), dW_in, db_in]
                    for i in range(n_hidden):
                        gl += [dW_h[i], db_h[i]]
                    for s_, n, g in zip(saved[3:-1], trunk_needs, gl):
                        out.append(g.to(s_.dtype) if n else None)
                else:
                    req = [x for x, n in zip(leaves[3:-1], trunk_needs) if n]
                    gr = iter(torch.autograd.grad(h2, req, dh2) if req else [])
                    for s_, n in zip(saved[3:-1], trunk_needs):
                        out.append(next(gr).to(s_.dtype) if n else None)
                out.append(g_wout.to(saved[-1].dtype) if g_wout is not None else None)
                return (None, None, None, None, *out)
            h = F.linear(h, wb[-1])
            if modulate:
                h = h * (torch.exp(-t[:, :L] * deltas.abs()) + shift)
            if normalized:
                h = h / torch.norm(h, dim=-1, p=1, keepdim=True)
            req = [x for x, n in zip(leaves, needs) if n]
            grads = torch.autograd.grad(h, req, dk.t().unsqueeze(0).float()) if req else []
        it = iter(grads)
        out = [next(it).to(s.dtype) if n else None for s, n in zip(saved, needs)]
        return (None, None, None, None, *out)


# ---- reuse of the implicit filter across gradient-accumulation micro-batches ---------------
# The filter depends on parameters only (hyena.py:456 regenerates it every forward).  With `filter_reuse` on, an operator
# generates k (and its spectrum) once per parameter version, feeds every micro-batch a detached leaf of it — autograd sums
# dk over the micro-batches in that leaf — and runs the filter's backward ONCE, when the gradients are flushed: by
# flush_filter_grads() (FlatGradAllReduce.allreduce() calls it) or, at the latest, by a global optimizer pre-step hook.
import weakref

_PENDING_FILTER_GRADS = weakref.WeakSet()
_STEP_HOOK = []


def flush_filter_grads(module: nn.Module = None):
    """Run the deferred filter backward of every HyenaOperator with `filter_reuse` (under `module`, or all of them)."""
    ops = list(_PENDING_FILTER_GRADS)
    if module is not None:
        under = set(id(m) for m in module.modules())
        ops = [o for o in ops if id(o) in under]
    for op in ops:
        op._flush_filter_grad()


def _ensure_step_hook():
    if not _STEP_HOOK:
        from torch.optim.optimizer import register_optimizer_step_pre_hook
        _STEP_HOOK.append(register_optimizer_step_pre_hook(lambda opt, args, kwargs: flush_filter_grads()))


# ---- the filter path on a side stream ------------------------------------------------------------------------------------
# The implicit filter depends on parameters only, and in the backward everything downstream of the spectrum gradient
# (hy_conv_dk -> modulation / last Linear / trunk backward, 4.5 ms per layer at 1 M) is independent of the activation-gradient
# chain (short-filter backward -> in_proj backward -> add + LN backward -> the previous block's MLP backward).  The forward
# generates the filter on a side stream; autograd then runs the backward of those nodes on that stream too, and
# _HyenaCoreFn.backward enqueues hy_conv_dk there — so the filter chain overlaps the cuBLAS / elementwise chain instead
# of extending it.  Opt-in (HYENA_B200_FILTER_STREAM=1): measured on B200 at the 1 M configuration the step does not get
# faster (222.7 vs 223.0 ms) — the chip runs under its power cap and every kernel of both chains fills the machine — while
# the per-kernel CUDA-event times of the bench line inflate under the overlap.
_FILTER_STREAMS = {}


def _filter_side_stream(device):
    if os.environ.get("HYENA_B200_FILTER_STREAM", "0") != "1" or device.type != "cuda":
        return None
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _FILTER_STREAMS:
        _FILTER_STREAMS[key] = torch.cuda.Stream(device=device)
    return _FILTER_STREAMS[key]


def _compute_dtype(u: torch.Tensor) -> torch.dtype:
    """dtype the activations of the fused path are stored in: the autocast dtype when autocast is on
    (the reference relies on nn.Linear/Conv1d autocasting), else the input dtype."""
    if u.is_cuda and torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda")
    else:
        dt = u.dtype
    return torch.bfloat16 if dt == torch.bfloat16 else torch.float32


class _InProjT(torch.autograd.Function):
    """uT[b] = W @ u[b]^T  ([3D, D] x [D, L], cuBLAS) — the input projection of hyena.py:441 written
    directly in the channel-major layout the fused kernel reads, so the reference's
    `rearrange(u, 'b l d -> b d l')` (hyena.py:442) never materialises.  The bias is added inside
    the fused kernel.  The backward emits du as a contiguous [B, L, D] tensor."""

    @staticmethod
    def forward(ctx, u, W, cdt):
        uc = u.to(cdt)
        Wc = W.to(cdt)
        uT = torch.matmul(Wc, uc.transpose(-1, -2))
        ctx.save_for_backward(uc, Wc)
        ctx.dtypes = (u.dtype, W.dtype)
        return uT

    @staticmethod
    def backward(ctx, duT):
        uc, Wc = ctx.saved_tensors
        u_dtype, w_dtype = ctx.dtypes
        du = torch.matmul(duT.transpose(-1, -2), Wc).to(u_dtype) if ctx.needs_input_grad[0] else None
        dW = None
        if ctx.needs_input_grad[1]:
            dW = torch.matmul(duT, uc)
            if dW.dim() == 3:
                dW = dW.sum(0)
            dW = dW.to(w_dtype)
        return du, dW, None


class _InProjToSlab(torch.autograd.Function):
    """_InProjT on this rank's sequence chunk followed by the exchange to its channel slab (channel partition of one
    long sequence): the GEMM writes straight into the buffer the peers pull from.  u [B, L/G, D] -> [B, 3 D/G, L]."""

    @staticmethod
    def forward(ctx, u, W, cdt, part):
        uc = u.to(cdt)
        Wc = W.to(cdt)
        B, Lc = uc.shape[0], uc.shape[1]
        ctx.save_for_backward(uc, Wc)
        ctx.dtypes = (u.dtype, W.dtype)
        ctx.part = part
        n = W.shape[0] // W.shape[1]
        return part.produced_to_channels((B, W.shape[0], Lc), cdt, u.device, n,
                                         lambda out: torch.matmul(Wc, uc.transpose(-1, -2), out=out))

    @staticmethod
    def backward(ctx, d_slab):
        uc, Wc = ctx.saved_tensors
        u_dtype, w_dtype = ctx.dtypes
        n = Wc.shape[0] // Wc.shape[1]
        duT = ctx.part._exchange(d_slab.contiguous(), n, False)               # [B, 3D, L/G]
        du = torch.matmul(duT.transpose(-1, -2), Wc).to(u_dtype) if ctx.needs_input_grad[0] else None
        dW = None
        if ctx.needs_input_grad[1]:
            dW = torch.matmul(duT, uc)
            if dW.dim() == 3:
                dW = dW.sum(0)
            dW = dW.to(w_dtype)
        return du, dW, None, None


class _OutProjT(torch.autograd.Function):
    """y[b] = z[b]^T @ W^T + bias  (hyena.py:496-504: the 'b d l -> b l d' rearrange folded into the
    GEMM's operand layout).  The backward emits dz channel-major [B, D, L] directly."""

    @staticmethod
    def forward(ctx, z, W, bias):
        Wc = W.to(z.dtype)
        B, D, L = z.shape
        y = torch.empty((B, L, W.shape[0]), dtype=z.dtype, device=z.device)
        bc = bias.to(z.dtype) if bias is not None else None
        for b in range(B):
            if bc is not None:
                torch.addmm(bc, z[b].t(), Wc.t(), out=y[b])
            else:
                torch.mm(z[b].t(), Wc.t(), out=y[b])
        ctx.save_for_backward(z, Wc)
        ctx.dtypes = (W.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, Wc = ctx.saved_tensors
        w_dtype, b_dtype = ctx.dtypes
        dy = dy.to(z.dtype)
        dz = torch.matmul(Wc.t(), dy.transpose(1, 2)) if ctx.needs_input_grad[0] else None     # [B, D, L]
        dW = torch.matmul(dy.transpose(1, 2), z.transpose(1, 2)).sum(0).to(w_dtype) if ctx.needs_input_grad[1] else None
        db = dy.sum(dim=(0, 1)).to(b_dtype) if (b_dtype is not None and ctx.needs_input_grad[2]) else None
        return dz, dW, db


class _HyenaCoreFn(torch.autograd.Function):
    """uT [B, 3D, L] -> z [B, D, L]: short filter + both gates + long convolution, fused
    (reference: hyena.py:444-503 for order == 2)."""

    @staticmethod
    def forward(ctx, uT, in_bias, sw, sb, k, D, L, kf_cache=None, side=None):
        Dm = uT.shape[1] // 3
        sw32 = sw.detach().float().reshape(3 * Dm, -1).contiguous()
        sb32 = sb.detach().float().contiguous()
        pb32 = in_bias.detach().float().contiguous() if in_bias is not None else None
        if kf_cache is not None and kf_cache.get("Kf") is not None:
            Kf = kf_cache["Kf"]          # same k and D as the micro-batch that computed it (filter_reuse)
        else:
            k32 = k.detach()
            if k32.stride(-1) != 1:
                k32 = k32.contiguous()
            Kf = K.filter_spectrum(k32, D.detach().float(), L)
            if kf_cache is not None:
                kf_cache["Kf"] = Kf
        # long sequences: keep the spectrum of g = v * x1 for the backward (8 B per frequency and channel) instead of
        # transforming g a second time there
        need_bwd = any(ctx.needs_input_grad)
        gs = K.conv_gsave_alloc(uT.shape[0], Dm, L, uT.device) if need_bwd else None
        z, ys = K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw32, sb=sb32, pb=pb32,
                           H=Dm, save_y=True, gsave=gs)
        ctx.L = L
        ctx.side = side              # stream the producers of k and D ran on (their backward runs there): hy_conv_dk joins it
        ctx.meta = (sw.shape, sw.dtype, sb.dtype, None if in_bias is None else in_bias.dtype, k.shape, D.shape, D.dtype)
        ctx.has_gs = gs is not None
        ctx.save_for_backward(uT, Kf, ys, sw32, sb32, pb32, *([gs] if gs is not None else []))
        return z

    @staticmethod
    def backward(ctx, dz):
        uT, Kf, ys, sw32, sb32, pb32 = ctx.saved_tensors[:6]
        gs = ctx.saved_tensors[6] if ctx.has_gs else None
        sw_shape, sw_dtype, sb_dtype, pb_dtype, k_shape, D_shape, D_dtype = ctx.meta
        L = ctx.L
        Dm = uT.shape[1] // 3
        dz = dz.to(uT.dtype)
        if dz.stride(-1) != 1:
            dz = dz.contiguous()
        # dx0 = dz * y is formed by the short-filter backward (it streams those rows anyway) when the layouts allow
        defer = K.shortconv_gate_supported(uT, dz, ys)
        dX, _, _, dKacc, dD = K.conv_bwd(dz, uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw32, sb=sb32,
                                         pb=pb32, ysave=ys, H=Dm, gsave=gs, defer_dx0=defer)
        need_dk = ctx.needs_input_grad[4] or dD is None
        side = ctx.side if (dz.is_cuda and need_dk) else None
        if side is not None:
            # the filter-gradient chain leaves the caller's stream here: its consumers (the backward of the nodes that
            # produced k and D) run on `side` as well, so stream order alone makes dk / dD ready for them
            cur = torch.cuda.current_stream(dz.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dk = K.conv_dk(dKacc, L)
                if dD is None:
                    dD = dk[:, 0]          # y = k * g + D g: the skip weight is one more tap at lag 0
                dD = dD.reshape(D_shape).to(D_dtype)
            dKacc.record_stream(side)
        duT, dsw, dsb, dpb = K.shortconv_bwd(uT, dX, sw32, pb32, L, dout=dz if defer else None, ysave=ys if defer else None)
        if side is None:
            dk = K.conv_dk(dKacc, L) if need_dk else None
            if dD is None:
                dD = dk[:, 0]
        if not ctx.needs_input_grad[4]:
            dk = None
        return (duT,
                dpb.to(pb_dtype) if pb32 is not None else None,
                dsw.reshape(sw_shape).to(sw_dtype),
                dsb.to(sb_dtype),
                dk.reshape(k_shape) if dk is not None else None,
                dD.reshape(D_shape).to(D_dtype),
                None, None, None)


# ------------------------------------------------------------------------------------------------
# HyenaFilter / HyenaOperator
# ------------------------------------------------------------------------------------------------
class HyenaFilter(OptimModule):
    """Implicit long filter with modulation — constructor and attributes of hyena.py:162-271."""

    def __init__(self, d_model, emb_dim=3, order=16, fused_fft_conv=False, seq_len=1024, lr=1e-3, lr_pos_emb=1e-5,
                 dropout=0.0, w=1, wd=0, bias=True, num_inner_mlps=2, linear_mixer=False, modulate: bool = True,
                 normalized=False, bidirectional=False, **kwargs):
        super().__init__()
        self.d_model = d_model
        self.emb_dim = emb_dim
        self.seq_len = seq_len
        self.modulate = modulate
        self.use_bias = bias
        self.fused_fft_conv = fused_fft_conv      # accepted for config compatibility: always fused here
        self.bias = nn.Parameter(torch.randn(self.d_model))
        self.dropout = nn.Dropout(dropout)        # never applied by the reference either (hyena.py:190-191)
        self.bidirectional = bidirectional
        if linear_mixer:
            raise NotImplementedError("HyenaFilter(linear_mixer=True) is not implemented")
        if order > 64 or emb_dim > 64:
            raise NotImplementedError("hyena-b200 filter kernel supports MLP width / emb_dim up to 64")
        act = Sin(dim=order, w=w)
        assert emb_dim % 2 != 0 and emb_dim >= 3, "emb_dim must be odd and greater or equal to 3 (time, sine and cosine)"
        self.pos_emb = PositionalEmbedding(emb_dim, seq_len, lr_pos_emb)
        self.implicit_filter = nn.Sequential(nn.Linear(emb_dim, order), act)
        for _ in range(num_inner_mlps):
            self.implicit_filter.append(nn.Linear(order, order))
            self.implicit_filter.append(act)
        self.implicit_filter.append(nn.Linear(order, d_model, bias=False))
        self.modulation = ExponentialModulation(d_model, **kwargs)
        self.normalized = normalized
        for c in self.implicit_filter.children():
            for name, _ in c.state_dict().items():
                setattr(getattr(c, name), "_optim", {"weight_decay": wd, "lr": lr})

    # channel-major filter [D, L] (the layout the kernels consume)
    def filter_cm(self, L, positions=None):
        """positions = (lo, hi): only the filter taps of sequence positions [lo, hi) — [D, hi - lo].  The MLP, the sine
        and the modulation are per-position, so under the channel partition of one long sequence
        every rank generates its chunk of positions for all channels and the exchange step hands each rank its channel
        slab over the whole length; the backward runs the same way in reverse."""
        z, t = self.pos_emb.z, self.pos_emb.t
        if positions is not None:
            lo, hi = positions
            z, t, L = z[:, lo:hi], t[:, lo:hi], hi - lo
        lins = _mlp_layers(self.implicit_filter)
        wb = []
        for lin in lins[:-1]:
            wb += [lin.weight, lin.bias]
        wb.append(lins[-1].weight)
        freq = self.implicit_filter[1].freq
        mod = self.modulation
        modulate = bool(self.modulate) and bool(getattr(mod, "modulate", True))
        return _FilterFn.apply(L, float(mod.shift), modulate, bool(self.normalized), z, t, mod.deltas, freq, *wb)

    def filter(self, L, *args, **kwargs):
        """[1, L, D] like the reference (hyena.py:233-242); a transposed view of the kernel output."""
        return self.filter_cm(L).t().unsqueeze(0)

    def forward(self, x, L, k=None, bias=None, *args, **kwargs):
        """Reference: hyena.py:244-271 — y = fftconv(x, k, bias). x: [B, D, L] or the 5-D 'b h v z l' view of
        HyenaOperator (z = num_blocks sequence blocks of length L / z each, convolved separately, hyena.py:447-453)."""
        if k is None:
            k = self.filter_cm(L)
        k = k[0] if type(k) is tuple else k
        if bias is None:
            bias = self.bias
        bias = bias if self.use_bias else 0 * bias
        shape = x.shape
        H = k.shape[-2] if k.dim() >= 2 else self.d_model
        k2 = k.reshape(H, -1)
        if x.dim() == 5 and shape[3] > 1:
            x3 = x.permute(0, 1, 3, 2, 4).reshape(-1, H, shape[-1])           # (b h z) v l
        else:
            x3 = x.reshape(-1, H, shape[-1])
        l = shape[-1]
        b1 = bias.reshape(-1).float()
        if self.bidirectional:
            y = fftconv_func(x3, k2, b1, dropout_mask=None, gelu=False, bidirectional=True)
        elif k2.shape[-1] > l:
            # num_blocks > 1: the reference's rfft(k, n=2l) crops the full-length filter to 2l taps and the product of
            # the spectra makes the convolution CIRCULAR with period 2l (hyena.py:61-63,84)
            N = 2 * l
            kk = k2[..., :N]
            if kk.shape[-1] < N:
                kk = F.pad(kk, (0, N - kk.shape[-1]))
            y = (circular_conv(x3, kk, l) + x3.float() * b1[:, None]).to(x3.dtype)
        else:
            y = fftconv_func(x3, k2, b1, dropout_mask=None, gelu=False)
        if x.dim() == 5 and shape[3] > 1:
            y = y.reshape(shape[0], shape[1], shape[3], shape[2], shape[4]).permute(0, 1, 3, 2, 4)
        return y.reshape(shape).to(dtype=x.dtype)


_FILTER_REGISTRY = {"hyena-filter": HyenaFilter}


class HyenaOperator(nn.Module):
    """Constructor signature of the reference (hyena.py:312-333); `standalone_hyenadna.HyenaOperator`
    (standalone:227-271) is the same class with fewer keywords."""

    def __init__(self, d_model, l_max, order=2, filter_order=64, num_heads=1, inner_factor=1, num_blocks=1,
                 fused_bias_fc=False, outer_mixing=False, dropout=0.0, filter_dropout=0.0, filter_cls="hyena-filter",
                 post_order_ffn=False, jit_filter=False, short_filter_order=3, activation="id", return_state=False,
                 bidirectional=False, **filter_args):
        super().__init__()
        assert d_model % num_heads == 0, f"Model dimension {d_model} must be divisible by num heads {num_heads}"
        assert l_max % num_blocks == 0, f"Maximum signal length {l_max} must be divisible by block dimension {num_blocks}"
        assert order >= 2, f"Order must be at least 2, (got {order})"
        # num_heads > 1 and inner_factor != 1 do not run in the reference either (its forward raises: the split by
        # d_model at hyena.py:455 / the conv1d channel count at :407-413 do not match — checked against the reference);
        # fused_bias_fc needs flash-attn's FusedDense, jit_filter references an attribute that does not exist (:426)
        unsupported = dict(num_heads=(num_heads, 1), inner_factor=(inner_factor, 1),
                           fused_bias_fc=(fused_bias_fc, False), jit_filter=(jit_filter, False))
        for name, (val, ok) in unsupported.items():
            if val != ok:
                raise NotImplementedError(f"hyena-b200 HyenaOperator: {name}={val} is not implemented (only {ok})")
        self.d_model, self.l_max, self.order = d_model, l_max, order
        self.num_heads, self.inner_factor, self.num_blocks = num_heads, inner_factor, num_blocks
        self.block_dim, self.head_dim = l_max // num_blocks, d_model // num_heads
        self.filter_order, self.short_filter_order = filter_order, short_filter_order
        self.post_order_ffn, self.outer_mixing, self.jit_filter = post_order_ffn, outer_mixing, jit_filter
        self.filter_dropout, self.return_state, self.bidirectional = filter_dropout, return_state, bidirectional
        self.activation = Activation(activation)
        self.dropout = nn.Dropout(dropout)
        # projections stay nn.Linear-typed attributes: the backbone re-initialises every nn.Linear and
        # `out_proj.weight` by name (long_conv_lm.py:270-318, standalone:612-641)
        self.out_proj = nn.Linear(d_model * inner_factor, d_model)
        self.in_proj = nn.Linear(d_model, (order + 1) * d_model)
        if post_order_ffn:
            self.ord_proj_w = nn.Parameter(torch.randn(order, num_heads, num_heads) / math.sqrt(self.head_dim))
        total_width = d_model * inner_factor * (order + 1)
        self.short_filter = nn.Conv1d(total_width, total_width, short_filter_order, groups=total_width,
                                      padding=short_filter_order - 1)
        # drop keys the reference swallows through **kwargs (layer_idx/device/dtype from create_mixer_cls)
        filter_args = {k: v for k, v in filter_args.items() if k not in ("layer_idx", "device", "dtype")}
        fcls = _FILTER_REGISTRY[filter_cls] if isinstance(filter_cls, str) else filter_cls
        self.filter_fn = fcls(self.head_dim * inner_factor * (order - 1), order=filter_order, seq_len=l_max, channels=1,
                              dropout=filter_dropout, bidirectional=bidirectional, **filter_args)
        # channel order of the (order-1) filters inside filter_fn's d_model axis: "src" = '(v o)'
        # (hyena.py:460), "standalone" = '(o v)' (standalone:283). Identical for order == 2.
        self.filter_channel_order = "src"
        self.cache_filter_spectrum = True      # under torch.no_grad(): reuse the filter spectrum until a parameter changes
        self._kf_cache = None
        self.channel_partition = None          # dna_b200.dp.ChannelPartition: one long sequence split over the ranks
        # training with gradient accumulation: generate the filter once per parameter version and run its backward once
        # per optimizer step (see flush_filter_grads); off by default — the reference regenerates it every forward
        self.filter_reuse = False
        self._filter_train_cache = None

    def recurrence(self, u, state):
        raise NotImplementedError("Working on it!")

    def _split_filter(self, k_cm, bias):
        D, o = self.d_model, self.order - 1
        if o == 1:
            return [k_cm], [bias]
        if self.filter_channel_order == "src":
            return [k_cm[i::o] for i in range(o)], [bias[i::o] for i in range(o)]
        return [k_cm[i * D:(i + 1) * D] for i in range(o)], [bias[i * D:(i + 1) * D] for i in range(o)]

    def _needs_reference_structure(self):
        """options outside the HyenaDNA configurations: run the reference's own sequence of steps (hyena.py:444-503)
        around the long-convolution kernels instead of the single fused core"""
        return (self.num_blocks != 1 or self.outer_mixing or self.post_order_ffn or self.bidirectional
                or self.short_filter_order != 3 or (self.training and self.dropout.p > 0))

    def forward(self, u, *args, **kwargs):
        K._check_dev(u)
        l = u.size(-2)
        L = min(l, self.l_max)
        if L < l:
            u = u[..., :L, :]          # causal + truncated output (hyena.py:439,444): later inputs never matter
        D = self.d_model
        squeeze = u.dim() == 2
        if squeeze:
            u = u.unsqueeze(0)
        cdt = _compute_dtype(u)
        # the reference's out_proj returns the autocast dtype (fp16 too; the kernels then compute in fp32 and the result
        # is cast once), else the input dtype
        out_dtype = torch.get_autocast_dtype("cuda") if (u.is_cuda and torch.is_autocast_enabled()) else u.dtype
        part = self.channel_partition
        if self._needs_reference_structure():
            if part is not None and part.world > 1:
                raise NotImplementedError("channel partition: only the fused order-2 configuration is implemented")
            return self._finish(self._forward_reference_structure(u, L, cdt), out_dtype, squeeze)
        if part is not None and part.world > 1:
            return self._finish(self._forward_channel_partition(u, part, cdt), out_dtype, squeeze)
        # in_proj written channel-major: uT[b] = W_in @ u[b]^T (bias is added inside the fused kernel)
        if self.order == 2 and torch.is_grad_enabled() and not self.filter_reuse and u.is_cuda:
            side = _filter_side_stream(u.device)
            if side is not None:
                return self._finish(self._forward_side_stream(u, L, cdt, side), out_dtype, squeeze)
        uT = _InProjT.apply(u, self.in_proj.weight, cdt)
        if self.order == 2 and not torch.is_grad_enabled() and self.cache_filter_spectrum:
            # inference: the filter is input-independent, so its spectrum is generated once and reused until a
            # parameter changes (the reference regenerates it every call, hyena.py:456)
            z = self._forward_cached(uT, L)
            return self._finish(z, out_dtype, squeeze)
        kf_cache = None
        if self.filter_reuse and self.order == 2 and torch.is_grad_enabled():
            k_cm, kf_cache = self._reused_filter(L, u.device)
        else:
            k_cm = self.filter_fn.filter_cm(L)                                  # [D*(order-1), L] fp32
        fbias = self.filter_fn.bias if self.filter_fn.use_bias else 0 * self.filter_fn.bias
        ks, bs = self._split_filter(k_cm, fbias)
        if self.order == 2:
            z = _HyenaCoreFn.apply(uT, self.in_proj.bias, self.short_filter.weight, self.short_filter.bias, ks[0], bs[0], L,
                                   kf_cache)
        else:
            z = self._forward_general(uT, ks, bs, L)
        return self._finish(z, out_dtype, squeeze)

    def _filter_state_key(self, L, device):
        """Identity of everything the filter spectrum depends on: tensor storage + version counters of filter_fn's
        parameters / buffers and the python-level switches.  Returns None (= do not cache) for inference tensors,
        which have no version counter.  Writes through `.data` bypass the counter: call invalidate_filter_cache()."""
        ts = list(self.filter_fn.parameters()) + list(self.filter_fn.buffers())
        if any(t.is_inference() for t in ts):
            return None
        mod = self.filter_fn.modulation
        flags = (bool(self.filter_fn.modulate), bool(getattr(mod, "modulate", True)), float(getattr(mod, "shift", 0.0)),
                 bool(self.filter_fn.use_bias), bool(self.filter_fn.normalized))
        return (L, str(device), flags, tuple((t.data_ptr(), t._version) for t in ts))

    def _forward_side_stream(self, u, L, cdt, side):
        """order-2 training forward with the filter path on `side` (see _filter_side_stream)"""
        cur = torch.cuda.current_stream(u.device)
        side.wait_stream(cur)                  # the parameters' last update (optimizer step) happened on `cur`
        with torch.cuda.stream(side):
            k_cm = self.filter_fn.filter_cm(L)
            fbias = self.filter_fn.bias if self.filter_fn.use_bias else 0 * self.filter_fn.bias
            fbias = fbias.clone()              # a node of the side stream: the skip weight's gradient arrives there too
        uT = _InProjT.apply(u, self.in_proj.weight, cdt)
        cur.wait_stream(side)
        k_cm.record_stream(cur)
        fbias.record_stream(cur)
        return _HyenaCoreFn.apply(uT, self.in_proj.bias, self.short_filter.weight, self.short_filter.bias, k_cm, fbias, L,
                                  None, side)

    def _reused_filter(self, L, device):
        """the detached leaf of the filter shared by the micro-batches of one parameter version (+ its spectrum cache)"""
        key = self._filter_state_key(L, device)
        c = self._filter_train_cache
        if key is None:
            return self.filter_fn.filter_cm(L), None
        if c is None or c["key"] != key:
            self._flush_filter_grad()
            k_graph = self.filter_fn.filter_cm(L)
            c = dict(key=key, k_graph=k_graph, k_leaf=k_graph.detach().requires_grad_(True), spec={})
            self._filter_train_cache = c
            _PENDING_FILTER_GRADS.add(self)
            _ensure_step_hook()
        return c["k_leaf"], c["spec"]

    def _flush_filter_grad(self):
        c, self._filter_train_cache = self._filter_train_cache, None
        _PENDING_FILTER_GRADS.discard(self)
        if c is not None and c["k_leaf"].grad is not None:
            c["k_graph"].backward(c["k_leaf"].grad)       # ONE filter backward for all micro-batches

    def invalidate_filter_cache(self):
        """Drop the cached filter spectrum (needed after in-place updates through `.data`, which autograd's version
        counters do not see; train() and load_state_dict() call it themselves)."""
        self._kf_cache = None

    def train(self, mode: bool = True):
        self._kf_cache = None
        return super().train(mode)

    def _load_from_state_dict(self, *args, **kwargs):
        self._kf_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _forward_cached(self, uT, L):
        key = self._filter_state_key(L, uT.device)
        if key is None or self._kf_cache is None or self._kf_cache[0] != key:
            k_cm = self.filter_fn.filter_cm(L)
            fbias = self.filter_fn.bias if self.filter_fn.use_bias else 0 * self.filter_fn.bias
            k32 = k_cm if k_cm.stride(-1) == 1 else k_cm.contiguous()
            Kf = K.filter_spectrum(k32, fbias.detach().float(), L)
            self._kf_cache = (key, Kf) if key is not None else None
        else:
            Kf = self._kf_cache[1]
        Dm = self.d_model
        sw32 = self.short_filter.weight.detach().float().reshape(3 * Dm, -1).contiguous()
        sb32 = self.short_filter.bias.detach().float().contiguous()
        pb32 = self.in_proj.bias.detach().float().contiguous() if self.in_proj.bias is not None else None
        z, _ = K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw32, sb=sb32, pb=pb32, H=Dm)
        return z

    def _forward_channel_partition(self, u, part, cdt):
        """ONE sequence split over the ranks (BASELINE.json configs[3]): `u` is this rank's chunk
        [B, L/G, D] of the sequence.  in_proj (hyena.py:441) runs on the chunk for all 3D channels; one all-to-all hands
        every rank its slab {c, D+c, 2D+c} over the whole sequence; short filter, gates and the long convolution
        (hyena.py:444-503) — all per-channel — run there with the slab's short_filter taps, filters and skip weights; a
        second all-to-all returns z to sequence chunks for out_proj (hyena.py:504).  The implicit filter (hyena.py:211-242)
        is generated per chunk of positions and exchanged the same way.  No collective inside the core; parameters stay replicated and their gradients are
        summed over ranks by the caller."""
        if self.order != 2:
            raise NotImplementedError("channel partition: order > 2 is not implemented")
        D = self.d_model
        L = u.shape[-2] * part.world
        if L > self.l_max:
            raise ValueError(f"channel partition: global length {L} exceeds l_max {self.l_max}")
        lo, hi = part.slab(D)
        rows = part.slab_rows(D, 3, u.device)
        # the implicit filter is per-position too: generate this rank's chunk of taps for every channel, then the same
        # exchange hands over the slab's filters for the whole length (fp32, 4 D L / G bytes per rank).  It does not
        # depend on the activations: it runs on a side stream (its own exchange lane) beside in_proj and the exchange of
        # uT, and autograd runs its backward (dk exchange + filter backward) there too, beside the du path.
        if hasattr(part, "prepare"):
            part.prepare(u.device)
        overlap = u.is_cuda and getattr(part, "overlap", False)
        if overlap:
            cur = torch.cuda.current_stream(u.device)
            side = part.side_stream(u.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                k_chunk = self.filter_fn.filter_cm(L, positions=part.chunk(L))     # [D, L/G] fp32
                k_cm = part.to_channels(k_chunk.unsqueeze(0), 1, lane=1)[0]        # [w, L]
        uT = _InProjToSlab.apply(u, self.in_proj.weight, cdt, part)           # [B, 3D, L/G] -> exchange -> [B, 3w, L]
        if overlap:
            cur.wait_stream(side)
            k_cm.record_stream(cur)
        else:
            k_chunk = self.filter_fn.filter_cm(L, positions=part.chunk(L))
            k_cm = part.to_channels(k_chunk.unsqueeze(0), 1)[0]
        fbias = self.filter_fn.bias if self.filter_fn.use_bias else 0 * self.filter_fn.bias
        in_bias = self.in_proj.bias[rows] if self.in_proj.bias is not None else None
        z = _HyenaCoreFn.apply(uT, in_bias, self.short_filter.weight[rows], self.short_filter.bias[rows], k_cm,
                               fbias[lo:hi], L)                                # [B, w, L]
        return part.to_sequence(z, 1)                                          # [B, D, L/G]

    def _finish(self, z, out_dtype, squeeze):
        if isinstance(self.activation, nn.Identity):
            y = _OutProjT.apply(z, self.out_proj.weight, self.out_proj.bias)
        else:
            y = F.linear(self.activation(z.transpose(1, 2)), self.out_proj.weight.to(z.dtype),
                         None if self.out_proj.bias is None else self.out_proj.bias.to(z.dtype))
        if y.dtype != out_dtype:
            y = y.to(out_dtype)
        if squeeze:
            y = y[0]
        if self.return_state:
            return y, None
        return y

    def _forward_reference_structure(self, u, L, cdt):
        """The reference's forward step by step (hyena.py:444-503) for num_blocks > 1, outer_mixing, post_order_ffn,
        bidirectional, short_filter_order != 3 and dropout > 0: in_proj and the long convolutions run on the kernels
        (HyenaFilter.forward -> fftconv_func), the rearranges / gates / mixing between them are elementwise torch ops on
        the reference's 'b h v z l' view (h = 1)."""
        D, B = self.d_model, u.shape[0]
        uT = _InProjT.apply(u, self.in_proj.weight, cdt)                        # [B, (order+1) D, l], bias added below
        if self.short_filter_order == 3:
            uc = _ShortConvFn.apply(uT, self.in_proj.bias, self.short_filter.weight, self.short_filter.bias, L)
        else:
            xb = uT if self.in_proj.bias is None else uT + self.in_proj.bias.to(uT.dtype)[None, :, None]
            uc = F.conv1d(xb, self.short_filter.weight.to(uT.dtype), self.short_filter.bias.to(uT.dtype),
                          padding=self.short_filter_order - 1, groups=uT.shape[1])[..., :L]
        z = self.num_blocks
        if L % z:
            raise ValueError(f"sequence length {L} is not divisible by num_blocks {z}")
        uc = uc.reshape(B, 1, uc.shape[1], z, L // z)                            # b ho v z l
        *x, v = uc.split(D, dim=2)
        k_cm = self.filter_fn.filter_cm(L)
        fbias = self.filter_fn.bias if self.filter_fn.use_bias else 0 * self.filter_fn.bias
        ks, bs = self._split_filter(k_cm, fbias)
        for o, x_i in enumerate(reversed(x[1:])):
            if self.outer_mixing:
                v = self.dropout(v.unsqueeze(2) * x_i.unsqueeze(3)).sum(dim=2)  # 'b h 1 v z l' * 'b h v 1 z l'
            else:
                v = self.dropout(v * x_i)
            v = self.filter_fn(v, L, k=ks[o], bias=bs[o][None, :, None])
            if self.post_order_ffn:
                w = self.ord_proj_w[o]                                          # mul_sum over h1 (hyena.py:487-492)
                v = (w[None, :, :, None, None, None].to(v.dtype) * v.unsqueeze(2)).sum(dim=1)
        y = v * x[0]                                                            # b h v z l
        return y.reshape(B, D, L)                                               # h = 1: channel-major [B, D, (z l)]

    def _forward_general(self, uT, ks, bs, L):
        """order > 2 (hyena.py:475-484): v <- fftconv(v * x_i, k[o], bias[o]) for x_i = x[order-1] .. x[1],
        each step one fused PREGATE kernel; the last step also applies the x[0] gate (POSTGATE)."""
        D = self.d_model
        ucx = _ShortConvFn.apply(uT, self.in_proj.bias, self.short_filter.weight, self.short_filter.bias, L)
        xs = list(ucx.split(D, dim=1))
        v = xs.pop()
        gates = list(reversed(xs[1:]))
        for o, x_i in enumerate(gates):
            last = o == len(gates) - 1
            v = fftconv_func(v, ks[o], bs[o].float(), gelu=False, v=x_i, q=xs[0] if last else None)
        return v

    @property
    def d_output(self):
        return self.d_model


class _ShortConvFn(torch.autograd.Function):
    """Standalone short filter (used by the order > 2 path)."""

    @staticmethod
    def forward(ctx, uT, in_bias, sw, sb, L):
        C3 = uT.shape[1]
        sw32 = sw.detach().float().reshape(C3, -1).contiguous()
        sb32 = sb.detach().float().contiguous()
        pb32 = in_bias.detach().float().contiguous() if in_bias is not None else None
        ctx.L = L
        ctx.meta = (sw.shape, sw.dtype, sb.dtype, None if in_bias is None else in_bias.dtype)
        ctx.save_for_backward(uT, sw32, pb32)
        return K.shortconv_fwd(uT, sw32, sb32, pb32, L)

    @staticmethod
    def backward(ctx, dxc):
        uT, sw32, pb32 = ctx.saved_tensors
        sw_shape, sw_dtype, sb_dtype, pb_dtype = ctx.meta
        dX = torch.empty_strided(uT.shape, uT.stride(), dtype=uT.dtype, device=uT.device)
        dX[:, :, :ctx.L].copy_(dxc)
        duT, dsw, dsb, dpb = K.shortconv_bwd(uT, dX, sw32, pb32, ctx.L)
        return duT, (dpb.to(pb_dtype) if pb32 is not None else None), dsw.reshape(sw_shape).to(sw_dtype), dsb.to(sb_dtype), None


def standalone_hyena_operator(d_model, l_max, order=2, filter_order=64, dropout=0.0, filter_dropout=0.0, **filter_args):
    """`standalone_hyenadna.HyenaOperator(...)` (standalone:227-271): modulation shift defaults to 0.05
    there (standalone:129) and the (order-1) filters are laid out '(o d)' (standalone:283-284)."""
    filter_args.setdefault("shift", 0.05)
    op = HyenaOperator(d_model, l_max, order=order, filter_order=filter_order, dropout=dropout,
                       filter_dropout=filter_dropout, **filter_args)
    op.filter_channel_order = "standalone"
    return op
