"""Host harness around the operator: a HyenaDNA backbone with the construction surface of
`standalone_hyenadna.HyenaDNAModel` (standalone_hyenadna.py:869-919) so the reference's tiny / small /
large configurations (BASELINE.json) can be stepped end-to-end on the GPU box, where /root/reference
does not exist.  The mixer is ours (dna_b200.hyena.HyenaOperator) and so is the add -> LayerNorm glue either side of it
(dna_b200.block_ops); embeddings, MLP and the head are ordinary PyTorch modules
(cuBLAS / ATen), i.e. the reference's own callers restated: prenorm Block of standalone_hyenadna.py:467-541, LMBackbone :692-734, init :612-641.
state_dict keys match the reference model so checkpoints (huggingface.py:54-65) load unchanged.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, block_ops
from .hyena import standalone_hyena_operator


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, activation=F.gelu):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.activation = activation
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.activation(self.fc1(x)))


class Block(nn.Module):
    """Prenorm block: (dropout -> add -> LN -> mixer) then (dropout -> add -> LN -> MLP); returns
    (hidden_states, residual) like the reference (standalone_hyenadna.py:510-541)."""

    def __init__(self, dim, mixer_cls, mlp_cls, norm_cls=nn.LayerNorm, resid_dropout1=0.0, resid_dropout2=0.0,
                 residual_in_fp32=False, fused_add_norm=True):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.mixer = mixer_cls()
        self.dropout1 = nn.Dropout(resid_dropout1)
        self.norm1 = norm_cls(dim)
        self.mlp = mlp_cls(dim)
        self.dropout2 = nn.Dropout(resid_dropout2)
        self.norm2 = norm_cls(dim)

    def forward(self, hidden_states, residual=None):
        hidden_states, residual = add_norm(self.dropout1, self.norm1, hidden_states, residual, self.residual_in_fp32,
                                           self.fused_add_norm)
        hidden_states = self.mixer(hidden_states)
        hidden_states, residual = add_norm(self.dropout2, self.norm2, hidden_states, residual, self.residual_in_fp32,
                                           self.fused_add_norm)
        hidden_states = self.mlp(hidden_states)
        return hidden_states, residual


def add_norm(dropout, norm, hidden_states, residual, residual_in_fp32, fused, keep_norm_dtype=False):
    """dropout -> add -> norm of the prenorm block (standalone_hyenadna.py:521-525 / :534-538). With a LayerNorm our
    kernel supports this is one launch (dna_b200.block_ops.add_layer_norm; an active dropout adds the launch that draws
    the keep mask); otherwise the reference's own three statements."""
    active_dropout = dropout.p > 0.0 and dropout.training
    if (fused and (hidden_states.is_cuda or _lib.is_emulation())
            and block_ops.add_layer_norm_supported(norm, hidden_states, residual, residual_in_fp32)):
        return block_ops.add_layer_norm(hidden_states, residual, norm, residual_in_fp32, keep_norm_dtype=keep_norm_dtype,
                                        dropout_p=dropout.p if active_dropout else 0.0)
    dropped = dropout(hidden_states)
    residual = dropped + residual if residual is not None else dropped
    hidden_states = norm(residual.to(dtype=norm.weight.dtype))
    if residual_in_fp32:
        residual = residual.to(torch.float32)
    return hidden_states, residual


class GPT2Embeddings(nn.Module):
    def __init__(self, embed_dim, vocab_size, max_position_embeddings=0):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, embed_dim)
        self.max_position_embeddings = max_position_embeddings
        if max_position_embeddings > 0:
            self.position_embeddings = nn.Embedding(max_position_embeddings, embed_dim)

    def forward(self, input_ids, position_ids=None):
        x = self.word_embeddings(input_ids)
        if self.max_position_embeddings > 0:
            if position_ids is None:
                position_ids = torch.arange(input_ids.shape[1], dtype=torch.long, device=input_ids.device)
            x = x + self.position_embeddings(position_ids)
        return x


def _init_weights(module, n_layer, initializer_range=0.02, rescale_prenorm_residual=True):
    """GPT-2 style init of the reference (standalone_hyenadna.py:612-641): every nn.Linear / Embedding
    ~ N(0, 0.02); parameters literally named out_proj.weight / fc2.weight ~ N(0, 0.02/sqrt(2 n_layer))."""
    if isinstance(module, nn.Linear):
        nn.init.normal_(module.weight, std=initializer_range)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ("out_proj.weight", "fc2.weight"):
                nn.init.normal_(p, mean=0.0, std=initializer_range / math.sqrt(2 * n_layer))


class LMBackbone(nn.Module):
    def __init__(self, d_model, n_layer, d_inner, vocab_size, layer=None, max_position_embeddings=0,
                 resid_dropout=0.0, embed_dropout=0.1, layer_norm_epsilon=1e-5, initializer_cfg=None,
                 residual_in_fp32=False, checkpoint_blocks=False, fused_add_norm=True, **kwargs):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.checkpoint_blocks = checkpoint_blocks
        self.fused_add_norm = fused_add_norm       # the src tree's `fused_dropout_add_ln` (long_conv_lm.py:560-575)
        self.final_norm_fp32 = True                # cleared by HyenaDNAModel when an lm_head (a Linear) consumes ln_f
        self.embeddings = GPT2Embeddings(d_model, vocab_size, max_position_embeddings)
        norm_cls = partial(nn.LayerNorm, eps=layer_norm_epsilon)
        mlp_cls = partial(Mlp, hidden_features=d_inner if d_inner is not None else 4 * d_model,
                          activation=partial(F.gelu, approximate="tanh"))
        self.layers = nn.ModuleList()
        for i in range(n_layer):
            blk = Block(d_model, partial(standalone_hyena_operator, **layer), mlp_cls, norm_cls=norm_cls,
                        resid_dropout1=embed_dropout if i == 0 else resid_dropout, resid_dropout2=resid_dropout,
                        residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm)
            blk.layer_idx = i
            self.layers.append(blk)
        self.drop_f = nn.Dropout(resid_dropout)
        self.ln_f = nn.LayerNorm(d_model, eps=layer_norm_epsilon)
        self.apply(partial(_init_weights, n_layer=n_layer, **(initializer_cfg or {})))

    def forward(self, input_ids, position_ids=None):
        hidden_states = self.embeddings(input_ids, position_ids=position_ids)
        residual = None
        for layer in self.layers:
            if self.checkpoint_blocks and self.training:
                from torch.utils.checkpoint import checkpoint
                if residual is None:
                    hidden_states, residual = checkpoint(lambda h, l=layer: l(h, None), hidden_states, use_reentrant=False)
                else:
                    hidden_states, residual = checkpoint(layer, hidden_states, residual, use_reentrant=False)
            else:
                hidden_states, residual = layer(hidden_states, residual)
        # final_norm_fp32: the hidden states leave the model (no Linear follows) -> fp32 like the reference's LayerNorm
        hidden_states, _ = add_norm(self.drop_f, self.ln_f, hidden_states, residual, False, self.fused_add_norm,
                                    keep_norm_dtype=self.final_norm_fp32)
        return hidden_states


class HyenaDNAModel(nn.Module):
    """`HyenaDNAModel(d_model, n_layer, d_inner, vocab_size, layer=dict(...), ...)` -> hidden states
    [B, L, d_model] (standalone_hyenadna.py:869-919).  `lm_head=True` adds the next-token head of the
    training path (ConvLMHeadModel, src/models/sequence/long_conv_lm.py:684-786: Linear(d_model, vocab),
    weight tied to the embedding)."""

    def __init__(self, d_model, n_layer, d_inner, vocab_size, layer=None, max_position_embeddings=0,
                 resid_dropout=0.0, embed_dropout=0.1, layer_norm_epsilon=1e-5, initializer_cfg=None,
                 residual_in_fp32=False, pad_vocab_size_multiple=1, lm_head=False, checkpoint_blocks=False, **kwargs):
        super().__init__()
        if vocab_size % pad_vocab_size_multiple != 0:
            vocab_size += pad_vocab_size_multiple - (vocab_size % pad_vocab_size_multiple)
        layer = dict(layer or {})
        layer.setdefault("d_model", d_model)
        self.backbone = LMBackbone(d_model=d_model, n_layer=n_layer, d_inner=d_inner, vocab_size=vocab_size, layer=layer,
                                   max_position_embeddings=max_position_embeddings, resid_dropout=resid_dropout,
                                   embed_dropout=embed_dropout, layer_norm_epsilon=layer_norm_epsilon,
                                   initializer_cfg=initializer_cfg, residual_in_fp32=residual_in_fp32,
                                   checkpoint_blocks=checkpoint_blocks)
        self.lm_head = None
        if lm_head:
            self.lm_head = nn.Linear(d_model, vocab_size, bias=False)
        self.apply(partial(_init_weights, n_layer=n_layer, **(initializer_cfg or {})))
        if self.lm_head is not None:
            self.lm_head.weight = self.backbone.embeddings.word_embeddings.weight
            self.backbone.final_norm_fp32 = False  # the head casts to the autocast dtype anyway: same single rounding

    def forward(self, input_ids, position_ids=None, state=None):
        h = self.backbone(input_ids, position_ids=position_ids)
        if self.lm_head is not None:
            return self.lm_head(h)
        return h
