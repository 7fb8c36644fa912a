"""ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/hyena_oracle.py for the rules).

CPU restatement of the reference's HyenaDNA backbone forward around the oracle operator:
standalone_hyenadna.py:723-734 (LMBackbone.forward), :510-541 (prenorm Block.forward),
:420-451 (Mlp, tanh-GELU from create_mlp_cls :583-590) and the next-token head + cross-entropy of the
training path (src/models/sequence/long_conv_lm.py:771-786, src/tasks/metrics.py:182).  It consumes a
state_dict with the reference model's key names.  Used by tests and by bench.py's cpu_baseline /
`--impl reference` legs (the reference source tree does not exist on the GPU box).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import hyena_oracle as O


def block_add_norm(hidden: torch.Tensor, residual, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
                   residual_in_fp32: bool = False):
    """The dropout(p=0) -> add -> LayerNorm step of the prenorm Block, standalone_hyenadna.py:521-525 (= :534-538):
    residual = hidden + residual; hidden = norm(residual.to(weight.dtype)); optionally residual.to(float32).
    Returns (hidden, residual).  This is what dna_b200.block_ops.add_layer_norm is checked against."""
    residual = hidden + residual if residual is not None else hidden
    out = F.layer_norm(residual.to(weight.dtype), (hidden.shape[-1],), weight, bias, eps)
    if residual_in_fp32:
        residual = residual.to(torch.float32)
    return out, residual


def backbone_forward(ids: torch.Tensor, sd: Dict[str, torch.Tensor], *, n_layer: int, l_max: int, shift: float = 0.05,
                     eps: float = 1e-5) -> torch.Tensor:
    h = F.embedding(ids, sd["backbone.embeddings.word_embeddings.weight"])
    residual = None
    d = h.shape[-1]
    for i in range(n_layer):
        pre = f"backbone.layers.{i}."
        x, residual = block_add_norm(h, residual, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
        mixer = {k[len(pre + "mixer."):]: v for k, v in sd.items() if k.startswith(pre + "mixer.")}
        x = O.hyena_operator(x, mixer, l_max=l_max, shift=shift)
        x, residual = block_add_norm(x, residual, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
        x = F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
        x = F.gelu(x, approximate="tanh")
        h = F.linear(x, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return block_add_norm(h, residual, sd["backbone.ln_f.weight"], sd["backbone.ln_f.bias"], eps)[0]


def lm_loss(ids: torch.Tensor, targets: torch.Tensor, sd: Dict[str, torch.Tensor], **kw) -> torch.Tensor:
    h = backbone_forward(ids, sd, **kw)
    logits = F.linear(h, sd["backbone.embeddings.word_embeddings.weight"])   # tied head
    return F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), targets.reshape(-1))
