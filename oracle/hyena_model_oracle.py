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
                   residual_in_fp32: bool = False, keep_mask=None, dropout_p: float = 0.0):
    """The dropout(p=0) -> add -> LayerNorm step of the prenorm Block, standalone_hyenadna.py:521-525 (= :534-538):
    residual = hidden + residual; hidden = norm(residual.to(weight.dtype)); optionally residual.to(float32).
    Returns (hidden, residual).  This is what dna_b200.block_ops.add_layer_norm is checked against."""
    if keep_mask is not None:
        # dropped = self.dropout1(hidden_states) (standalone_hyenadna.py:521) for a GIVEN keep mask: nn.Dropout scales the
        # kept elements by 1 / (1 - p) in the input's dtype
        hidden = (hidden.float() * keep_mask.to(torch.float32) * (1.0 / (1.0 - dropout_p))).to(hidden.dtype)
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


def init_state_dict(d_model: int, n_layer: int, d_inner: int, vocab_size: int, l_max: int, *, emb_dim: int = 5,
                    filter_order: int = 64, w: float = 10.0, num_inner_mlps: int = 2, seed: int = 2222,
                    initializer_range: float = 0.02) -> Dict[str, torch.Tensor]:
    """A random-init state_dict with the reference model's key names and shapes, drawn the way the reference draws it:
    GPT-2 init of standalone_hyenadna.py:612-641 (Linear / Embedding ~ N(0, 0.02), zero Linear biases, out_proj.weight
    and fc2.weight ~ N(0, 0.02 / sqrt(2 n_layer))), LayerNorm (1, 0), Conv1d default init (short_filter, :262-268),
    filter bias ~ N(0, 1) (:165), Sin.freq = w (:89), positional tables (:99-113) and decay rates (:119-136).
    Lets bench.py's CPU / GPU baseline legs run the oracle without importing the product package."""
    import math
    g = torch.Generator().manual_seed(seed)
    n = lambda *shape, std=initializer_range: torch.randn(*shape, generator=g) * std
    sd = {"backbone.embeddings.word_embeddings.weight": n(vocab_size, d_model)}
    z, t = O.positional_tables(emb_dim, l_max)
    for i in range(n_layer):
        pre = f"backbone.layers.{i}."
        for nm in ("norm1", "norm2"):
            sd[pre + nm + ".weight"] = torch.ones(d_model)
            sd[pre + nm + ".bias"] = torch.zeros(d_model)
        m = pre + "mixer."
        sd[m + "in_proj.weight"] = n(3 * d_model, d_model)
        sd[m + "in_proj.bias"] = torch.zeros(3 * d_model)
        sd[m + "out_proj.weight"] = n(d_model, d_model, std=initializer_range / math.sqrt(2 * n_layer))
        sd[m + "out_proj.bias"] = torch.zeros(d_model)
        bound = 1.0 / math.sqrt(3.0)                       # Conv1d(groups=C, kernel 3): fan_in = 3
        sd[m + "short_filter.weight"] = (torch.rand(3 * d_model, 1, 3, generator=g) * 2 - 1) * bound
        sd[m + "short_filter.bias"] = (torch.rand(3 * d_model, generator=g) * 2 - 1) * bound
        f = m + "filter_fn."
        sd[f + "bias"] = torch.randn(d_model, generator=g)
        widths = [emb_dim] + [filter_order] * (num_inner_mlps + 1)
        for j in range(num_inner_mlps + 1):
            sd[f + f"implicit_filter.{2 * j}.weight"] = n(widths[j + 1], widths[j])
            sd[f + f"implicit_filter.{2 * j}.bias"] = torch.zeros(widths[j + 1])
            sd[f + f"implicit_filter.{2 * j + 1}.freq"] = w * torch.ones(1, filter_order)
        sd[f + f"implicit_filter.{2 * (num_inner_mlps + 1)}.weight"] = n(d_model, filter_order)
        sd[f + "pos_emb.z"], sd[f + "pos_emb.t"] = z.clone(), t.clone()
        sd[f + "modulation.deltas"] = O.modulation_deltas(d_model)
        sd[pre + "mlp.fc1.weight"] = n(d_inner, d_model)
        sd[pre + "mlp.fc1.bias"] = torch.zeros(d_inner)
        sd[pre + "mlp.fc2.weight"] = n(d_model, d_inner, std=initializer_range / math.sqrt(2 * n_layer))
        sd[pre + "mlp.fc2.bias"] = torch.zeros(d_model)
    sd["backbone.ln_f.weight"] = torch.ones(d_model)
    sd["backbone.ln_f.bias"] = torch.zeros(d_model)
    return sd
