"""Block glue around the operator: residual add + LayerNorm in one kernel.

The reference's prenorm Block does `residual = dropped + residual; hidden = norm(residual.to(norm.weight.dtype));
if residual_in_fp32: residual = residual.to(float32)` twice per layer (standalone_hyenadna.py:520-541; the src
tree has the same fusion behind `fused_dropout_add_ln`, src/models/sequence/long_conv_lm.py:560-575).  That is
`add_layer_norm` below: one pass, each element read once; an active dropout (embed_dropout = 0.1 on the first block in
the training configs, hg38_hyena.yaml:13) is a keep mask applied to x inside the same kernel.

Autocast note: under autocast the reference's LayerNorm returns fp32 and the next Linear casts it to the autocast
dtype; `add_layer_norm` writes that dtype directly (same values: the cast is the only thing between them).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import kernels as K


class _AddLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, y_dtype, res_dtype, keep=None, keep_scale=1.0):
        # the residual stream r = x + residual is returned (and saved) unless it IS x (first block, same dtype, no dropout)
        alias_x = residual is None and x.dtype == res_dtype and keep is None
        xc = x.contiguous()
        rc = residual.contiguous() if residual is not None else None
        y, res_out, mean, rstd = K.add_ln_fwd(xc, rc, gamma, beta, eps, y_dtype, res_dtype, write_res=not alias_x, keep=keep,
                                              keep_scale=keep_scale)
        ctx.save_for_backward(xc if alias_x else res_out, mean, rstd, gamma, *([keep] if keep is not None else []))
        ctx.x_dtype = x.dtype
        ctx.has_res = residual is not None
        ctx.keep_scale = keep_scale
        return y, res_out       # res_out is None when the stream is x itself (the caller keeps x)

    @staticmethod
    def backward(ctx, dy, dr):
        r, mean, rstd, gamma = ctx.saved_tensors[:4]
        keep = ctx.saved_tensors[4] if len(ctx.saved_tensors) > 4 else None
        dy = dy.contiguous()
        if dr is not None:
            dr = dr.contiguous()
            if dr.dtype != r.dtype:
                dr = dr.to(r.dtype)
        want_dx = ctx.needs_input_grad[0]
        want_dres = ctx.has_res and ctx.needs_input_grad[1]
        if not (want_dx or want_dres):
            want_dx = True
        dx, dres, dgamma, dbeta = K.add_ln_bwd(dy, dr, r, mean, rstd, gamma, ctx.x_dtype, want_dx, want_dres, keep=keep,
                                               keep_scale=ctx.keep_scale)
        return dx, dres, dgamma, dbeta, None, None, None, None, None


def _stream_dtype(x, residual):
    return x.dtype if residual is None else torch.promote_types(x.dtype, residual.dtype)


def add_layer_norm_supported(norm: torch.nn.Module, x: torch.Tensor, residual: Optional[torch.Tensor] = None,
                             residual_in_fp32: bool = False) -> bool:
    """True when `add_layer_norm` reproduces the reference's three statements for these operands."""
    ok = (isinstance(norm, torch.nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
          and len(norm.normalized_shape) == 1 and norm.weight.dtype == torch.float32
          and x.dtype in (torch.float32, torch.bfloat16)
          and (residual is None or residual.dtype in (torch.float32, torch.bfloat16))
          and x.shape[-1] == norm.normalized_shape[0] and K.add_ln_supported(norm.normalized_shape[0]))
    if not ok:
        return False
    if residual_in_fp32 and _stream_dtype(x, residual) != torch.float32:
        return False      # norm would see a bf16-rounded sum while the stored stream is fp32: keep the unfused path
    if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") != torch.bfloat16:
        return False
    return True


def add_layer_norm(x: torch.Tensor, residual: Optional[torch.Tensor], norm: torch.nn.LayerNorm,
                   residual_in_fp32: bool = False, keep_norm_dtype: bool = False, dropout_p: float = 0.0,
                   keep_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(hidden, residual) of one prenorm step: residual' = x + residual, hidden = norm(residual').

    Dtypes follow the reference's expressions: the sum takes torch's promoted dtype (fp32 once either side is fp32,
    or when residual_in_fp32), LayerNorm computes in fp32, and `hidden` comes out in the autocast dtype when
    autocast is on (what the consuming Linear would cast it to), else in the norm weight's dtype."""
    if not add_layer_norm_supported(norm, x, residual, residual_in_fp32):
        raise NotImplementedError("add_layer_norm: unsupported operands (see add_layer_norm_supported)")
    res_dtype = _stream_dtype(x, residual)
    if residual is not None and residual.dtype != res_dtype:
        residual = residual.to(res_dtype)
    # keep_norm_dtype: no Linear consumes `hidden` (the backbone's final ln_f when the model returns embeddings): the
    # reference's LayerNorm returns the weight's dtype (fp32) under autocast too
    y_dtype = torch.bfloat16 if (torch.is_autocast_enabled() and not keep_norm_dtype) else norm.weight.dtype
    keep, scale = None, 1.0
    if dropout_p > 0.0:
        # `dropped = dropout(hidden)` of the Block (standalone_hyenadna.py:521,534) folded in: ONE extra kernel draws the
        # keep mask (1 B per element; keep_mask lets a test hand in the reference's own mask), the fused kernel applies it
        keep = keep_mask if keep_mask is not None else torch.empty(x.shape, dtype=torch.bool, device=x.device).bernoulli_(1.0 - dropout_p)
        keep = keep.contiguous().view(torch.uint8) if keep.dtype == torch.bool else keep.contiguous()
        scale = 1.0 / (1.0 - dropout_p)
    y, res_out = _AddLayerNormFn.apply(x, residual, norm.weight, norm.bias, norm.eps, y_dtype, res_dtype, keep, scale)
    return y, (x if res_out is None else res_out)
