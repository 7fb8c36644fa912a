"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's HyenaDNA hot path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this file, and only as the checker or as the reported CPU baseline — never as
part of the product path (dna_b200/ never imports oracle/).

What it restates (all paths relative to /root/reference):
  * fftconv_ref            src/ops/fftconv.py:15-34, src/models/sequence/hyena.py:60-92,
                           standalone_hyenadna.py:45-60
  * fftconv_h3_ref         src/ops/fftconv.py:38-55   (head_dim == 1 gating semantics)
  * positional tables      src/models/sequence/hyena.py:113-135
  * implicit filter MLP    src/models/sequence/hyena.py:203-242 (+ Sin :100-110, modulation :138-159)
  * short filter + gates   src/models/sequence/hyena.py:436-508 (standalone_hyenadna.py:273-293)
  * reverse complement     src/dataloaders/datasets/hg38_dataset.py:28-38
  * operator options       src/models/sequence/hyena.py:447-453,476-492 (num_blocks, outer_mixing, post_order_ffn, ...)
  * LongConv               src/models/sequence/long_conv.py:107-165, long_conv_kernel.py:68-81

The arithmetic itself lives in a third-party dependency of the reference, PyTorch (pinned
torch==2.0.0+cu118 in environment.yml:192 / torch==2.1.0 in README.md:16; 2.11.0 is installed here):
torch.fft.rfft/irfft (pocketfft on CPU), nn.Conv1d, nn.Linear, sin, exp.  The restatement therefore
calls the same torch primitives in the same order and dtype as the reference does, written as plain
functions over explicit tensors instead of nn.Modules.

Pinning: the reference has no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference) and committed as tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# long convolution (hyena.py:60-92)
# ------------------------------------------------------------------------------------------------
def fftconv_ref(u: torch.Tensor, k: torch.Tensor, D: torch.Tensor, dropout_mask=None, gelu: bool = True,
                k_rev: Optional[torch.Tensor] = None, bidirectional: bool = False) -> torch.Tensor:
    """y[t] = sum_{s<=t} k[s] u[t-s] + D u[t], FFT size exactly 2L, FFT in k.dtype, result in u.dtype."""
    L = u.shape[-1]
    n = 2 * L
    k_f = torch.fft.rfft(k, n=n) / n
    if k_rev is not None:
        k_f = k_f + (torch.fft.rfft(k_rev, n=n) / n).conj()
    if bidirectional:
        half = L // 2
        total = L + 2 * half
        before = total // 2 - half
        after = total - L - before
        u_f = torch.fft.rfft(F.pad(u, (before, after)).to(k.dtype), n=n)
    else:
        u_f = torch.fft.rfft(u.to(k.dtype), n=n)
    if u.dim() > 3 and k_f.dim() != 2:
        k_f = k_f.reshape(u_f.shape).contiguous()
    elif u.dim() > 3:
        k_f = k_f.unsqueeze(1)
    y = torch.fft.irfft(u_f * k_f, n=n, norm="forward")[..., :L]
    out = y + u * D.unsqueeze(-1)
    if gelu:
        out = F.gelu(out)
    if dropout_mask is not None:
        out = out * dropout_mask.unsqueeze(-1)
    return out.to(u.dtype)


def fftconv_h3_ref(k, ssm_kernel, D, q, v, head_dim: int = 1, ssm_kernel_rev=None):
    """H3 gating (src/ops/fftconv.py:38-55): out = (conv(ssm_kernel, k*v) + D*(k*v)) * q, summed over d1."""
    L = k.shape[-1]
    n = 2 * L
    B = k.shape[0]
    kk = k.reshape(B, -1, head_dim, L).permute(0, 2, 1, 3).unsqueeze(2)    # b d1 1 h l
    vv = v.reshape(B, -1, head_dim, L).permute(0, 2, 1, 3).unsqueeze(1)    # b 1 d2 h l
    kv = kk * vv
    kv_f = torch.fft.rfft(kv.to(ssm_kernel.dtype), n=n) / n
    s_f = torch.fft.rfft(ssm_kernel, n=n)
    if ssm_kernel_rev is not None:
        s_f = s_f + torch.fft.rfft(ssm_kernel_rev, n=n).conj()
    y = torch.fft.irfft(kv_f * s_f, n=n, norm="forward")[..., :L]
    out = y + kv * D.unsqueeze(-1)
    qq = q.reshape(B, -1, head_dim, L).permute(0, 2, 1, 3).unsqueeze(2)    # b d1 1 h l
    if head_dim > 1:
        out = (out * qq).sum(dim=1)                                        # b d2 h l
        return out.permute(0, 2, 1, 3).reshape(B, -1, L).to(k.dtype)
    return (out * qq)[:, 0, 0].to(k.dtype)


# ------------------------------------------------------------------------------------------------
# implicit filter (hyena.py:113-159, 203-242)
# ------------------------------------------------------------------------------------------------
def positional_tables(emb_dim: int, seq_len: int):
    """z [1, seq_len, emb_dim], t [1, seq_len, 1] (hyena.py:113-135)."""
    t = torch.linspace(0, 1, seq_len)[None, :, None]
    bands = (emb_dim - 1) // 2
    t_rescaled = torch.linspace(0, seq_len - 1, seq_len)[None, :, None]
    w = 2 * math.pi * t_rescaled / seq_len
    f = torch.linspace(1e-4, bands - 1, bands)[None, None]
    zc = torch.exp(-1j * f * w)
    z = torch.cat([t, zc.real, zc.imag], dim=-1)
    return z, t


def modulation_deltas(d_model: int, fast_decay_pct: float = 0.3, slow_decay_pct: float = 1.5, target: float = 1e-2):
    """deltas [1, 1, d_model] (hyena.py:148-153)."""
    max_decay = math.log(target) / fast_decay_pct
    min_decay = math.log(target) / slow_decay_pct
    return torch.linspace(min_decay, max_decay, d_model)[None, None]


def hyena_filter(p: Dict[str, torch.Tensor], L: int, *, shift: float, modulate: bool = True,
                 normalized: bool = False) -> torch.Tensor:
    """h [1, L, D] (hyena.py:233-242).  `p` uses the reference's state_dict names relative to filter_fn:
    pos_emb.z, pos_emb.t, implicit_filter.{0,2,4,...}.{weight,bias}, implicit_filter.<last>.weight,
    implicit_filter.1.freq, modulation.deltas."""
    z = p["pos_emb.z"][:, :L]
    t = p["pos_emb.t"][:, :L]
    freq = p["implicit_filter.1.freq"]
    idx = sorted({int(key.split(".")[1]) for key in p if key.startswith("implicit_filter.") and key.endswith(".weight")})
    h = z
    for i in idx[:-1]:
        h = F.linear(h, p[f"implicit_filter.{i}.weight"], p[f"implicit_filter.{i}.bias"])
        h = torch.sin(freq * h)
    h = F.linear(h, p[f"implicit_filter.{idx[-1]}.weight"])
    if modulate:
        decay = torch.exp(-t * p["modulation.deltas"].abs())
        h = h * (decay + shift)
    if normalized:
        h = h / torch.norm(h, dim=-1, p=1, keepdim=True)
    return h


# ------------------------------------------------------------------------------------------------
# operator (hyena.py:436-508 with order=2, num_heads=num_blocks=inner_factor=1, dropout=0, activation=id)
# ------------------------------------------------------------------------------------------------
def short_filter(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, L: int) -> torch.Tensor:
    """Depthwise Conv1d(k, padding=k-1, groups=C)[..., :L] (hyena.py:407-413, 444). x: [B, C, l]."""
    ksz = weight.shape[-1]
    return F.conv1d(x, weight, bias, padding=ksz - 1, groups=x.shape[1])[..., :L]


def hyena_operator(u: torch.Tensor, p: Dict[str, torch.Tensor], *, l_max: int, shift: float, modulate: bool = True,
                   use_bias: bool = True, return_parts: bool = False, order: int = 2, channel_order: str = "src",
                   bidirectional: bool = False):
    """u [B, l, D] -> y [B, min(l, l_max), D].  `p`: the reference HyenaOperator state_dict.

    order > 2 follows the recurrence of hyena.py:475-484 (standalone_hyenadna.py:286-288): the in_proj output splits
    into x[0..order-1] and v; for x_i = x[order-1] .. x[1]: v <- fftconv(v * x_i, k[o], bias[o]); the result is gated by
    x[0].  The (order-1) filters live side by side on filter_fn's channel axis: '(v o)' in src (hyena.py:460,463-465),
    '(o v)' in standalone (standalone_hyenadna.py:283-284) — `channel_order` "src" / "standalone"."""
    D = u.shape[-1]
    l = u.shape[-2]
    L = min(l, l_max)
    x = F.linear(u, p["in_proj.weight"], p["in_proj.bias"]).transpose(1, 2)            # b (order+1)d l
    uc = short_filter(x, p["short_filter.weight"], p["short_filter.bias"], L)
    *xs, v = uc.split(D, dim=1)
    fp = {key[len("filter_fn."):]: val for key, val in p.items() if key.startswith("filter_fn.")}
    kf = hyena_filter(fp, L, shift=shift, modulate=modulate)[0]                          # l (order-1)d
    bias = fp["bias"] if use_bias else 0 * fp["bias"]
    o_n = order - 1
    if channel_order == "src":
        k = kf.reshape(L, D, o_n).permute(2, 1, 0)                                       # o v l
        bias = bias.reshape(D, o_n).t()
    else:
        k = kf.reshape(L, o_n, D).permute(1, 2, 0)                                       # o d l
        bias = bias.reshape(o_n, D)
    g = y = None
    for o, x_i in enumerate(reversed(xs[1:])):
        g = v * x_i
        y = fftconv_ref(g, k[o], bias[o], None, gelu=False, bidirectional=bidirectional).to(g.dtype)
        v = y
    z = (v * xs[0]).transpose(1, 2)                                                       # b l d
    out = F.linear(z, p["out_proj.weight"], p["out_proj.bias"])
    if return_parts:
        return out, dict(x0=xs[0], x1=xs[-1], v=uc.split(D, dim=1)[-1], k=k[0], g=g, y=y, z=z)
    return out


def hyena_operator_options(u: torch.Tensor, p: Dict[str, torch.Tensor], *, l_max: int, shift: float, order: int = 2,
                           num_blocks: int = 1, outer_mixing: bool = False, post_order_ffn: bool = False,
                           short_filter_order: int = 3, bidirectional: bool = False, dropout_p: float = 0.0,
                           training: bool = False, activation=None, modulate: bool = True) -> torch.Tensor:
    """HyenaOperator.forward (hyena.py:436-508) with the options outside the HyenaDNA configs, num_heads = 1, src filter
    layout '(v o)': num_blocks (:447-453 — the sequence is cut into z blocks of l / z positions convolved separately
    with the FULL-length filter cropped by rfft(k, n = 2 l / z), i.e. circularly), outer_mixing (:476-479),
    post_order_ffn (:487-492, parameter ord_proj_w), short_filter_order (:407-413), dropout (:481; F.dropout draws from
    the global CPU stream exactly like nn.Dropout), activation before out_proj (:496)."""
    D = u.shape[-1]
    l = u.shape[-2]
    L = min(l, l_max)
    B = u.shape[0]
    x = F.linear(u, p["in_proj.weight"], p["in_proj.bias"]).transpose(1, 2)
    uc = F.conv1d(x, p["short_filter.weight"], p["short_filter.bias"], padding=short_filter_order - 1,
                  groups=x.shape[1])[..., :L]
    z = num_blocks
    uc = uc.reshape(B, 1, uc.shape[1], z, L // z)                                         # b ho v z l
    *xs, v = uc.split(D, dim=2)
    fp = {key[len("filter_fn."):]: val for key, val in p.items() if key.startswith("filter_fn.")}
    kf = hyena_filter(fp, L, shift=shift, modulate=modulate)[0]                           # l (v o)
    o_n = order - 1
    k = kf.reshape(L, D, o_n).permute(2, 1, 0)                                            # o v l
    bias = fp["bias"].reshape(D, o_n).t()                                                 # o v
    for o, x_i in enumerate(reversed(xs[1:])):
        if outer_mixing:
            v = F.dropout(v.unsqueeze(2) * x_i.unsqueeze(3), dropout_p, training).sum(dim=2)
        else:
            v = F.dropout(v * x_i, dropout_p, training)
        v = fftconv_ref(v, k[o], bias[o][None, :, None], None, gelu=False, bidirectional=bidirectional).to(v.dtype)
        if post_order_ffn:
            w = p["ord_proj_w"][o]
            v = (w[None, :, :, None, None, None] * v.unsqueeze(2)).sum(dim=1)
    y = (v * xs[0]).permute(0, 3, 4, 1, 2).reshape(B, L, D)                               # b (z l) (h v)
    if activation is not None:
        y = activation(y)
    return F.linear(y, p["out_proj.weight"], p["out_proj.bias"])


def long_conv_ref(u: torch.Tensor, p: Dict[str, torch.Tensor], *, channels: int = 1, bidirectional: bool = False,
                  lam: float = 0.1, transposed: bool = True, activation=F.gelu, postact: Optional[str] = "glu") -> torch.Tensor:
    """LongConv.forward (src/models/sequence/long_conv.py:107-165) with LongConvKernel.forward (long_conv_kernel.py:68-81),
    eval mode (no dropout): soft-thresholded explicit kernel, rfft/irfft of length L_kernel + L, skip, activation,
    position-wise output Linear (+ GLU)."""
    if not transposed:
        u = u.transpose(-1, -2)
    L = u.size(-1)
    k = p["kernel.kernel"]
    k = F.relu(torch.abs(k) - lam) * torch.sign(k)
    Lk = k.shape[-1]
    if bidirectional:
        k0, k1 = k[:channels], k[channels:]
        k = F.pad(k0, (0, L)) + F.pad(k1.flip(-1), (L, 0))
    # L_kernel = min(L, l_max) (long_conv.py:123) although the kernel module returns all l_max taps: for L < l_max the
    # transform length n = L_kernel + L crops the kernel and the product of spectra wraps around (reference behaviour)
    n = min(L, Lk) + L
    k_f = torch.fft.rfft(k, n=n)
    u_f = torch.fft.rfft(u, n=n)
    y = torch.fft.irfft(torch.einsum("bhl,chl->bchl", u_f, k_f), n=n)[..., :L]
    y = y + torch.einsum("bhl,ch->bchl", u, p["D"])
    y = y.reshape(y.shape[0], -1, L)                                                      # ... (c h) l
    if not transposed:
        y = y.transpose(-1, -2)
    y = activation(y) if activation is not None else y
    if postact is None:
        return y
    W, b = p["output_linear.0.weight"], p["output_linear.0.bias"]
    if transposed:
        y = torch.einsum("bul,vu->bvl", y, W) + b[:, None]
        return F.glu(y, dim=1) if postact == "glu" else y
    y = F.linear(y, W, b)
    return F.glu(y, dim=-1) if postact == "glu" else y


# ------------------------------------------------------------------------------------------------
# tokenizer (hg38_char_tokenizer.py:58-94, standalone_hyenadna.py:1003-1018, hg38_dataset.py:194-223,383-386)
# ------------------------------------------------------------------------------------------------
VOCAB = {"[CLS]": 0, "[SEP]": 1, "[BOS]": 2, "[MASK]": 3, "[PAD]": 4, "[RESERVED]": 5, "[UNK]": 6,
         "A": 7, "C": 8, "G": 9, "T": 10, "N": 11}


_COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a"}


def reverse_complement_ref(seq: str) -> str:
    """string_reverse_complement, src/dataloaders/datasets/hg38_dataset.py:28-38: walk the string backwards,
    complement the eight listed letters, keep every other character (N, n, '.', IUPAC codes)."""
    return "".join(_COMPLEMENT.get(ch, ch) for ch in reversed(seq))


def tokenize_ref(text: str, max_length: int, *, add_special_tokens: bool = True, cls_token: bool = False):
    """transformers==4.28 call `tokenizer(seq, add_special_tokens=..., padding="max_length",
    max_length=..., truncation=True)["input_ids"]` with padding_side='left' (hg38_dataset.py:194-199).
    `cls_token=True` is the standalone_hyenadna.py variant ([CLS] ... [SEP])."""
    ids = [VOCAB.get(ch, VOCAB["[UNK]"]) for ch in text]
    n_special = (1 + int(cls_token)) if add_special_tokens else 0
    ids = ids[: max(max_length - n_special, 0)]          # truncation=True (longest_first, right side)
    if add_special_tokens:
        ids = ([VOCAB["[CLS]"]] if cls_token else []) + ids + [VOCAB["[SEP]"]]
    pad = max_length - len(ids)
    return [VOCAB["[PAD]"]] * pad + ids                  # padding_side='left'


def dataset_item_ref(text: str, max_length: int, *, add_eos: bool = True, replace_N_token: bool = False,
                     nucleotide_encode: bool = False):
    """HG38Dataset.__getitem__ post-processing (hg38_dataset.py:216-223; encode branch :383-386)."""
    seq = torch.LongTensor(tokenize_ref(text, max_length, add_special_tokens=add_eos))
    if nucleotide_encode:
        seq = seq - 7
        seq[(seq >= 4) | (seq < 0)] = 4
    if replace_N_token:
        seq[seq == VOCAB["N"]] = VOCAB["[PAD]"]
    return seq[:-1].clone(), seq[1:].clone()


# ------------------------------------------------------------------------------------------------
# data ingest ahead of the tokenizer (hg38_dataset.py:72-124, 238-286)
# ------------------------------------------------------------------------------------------------
def fetch_interval_ref(chromosome: str, start: int, end: int, max_length: int, *, pad_interval: bool = False,
                       reverse_complement: bool = False) -> str:
    """FastaInterval.__call__ (hg38_dataset.py:72-124) on a chromosome held as a Python string, without the random
    draws (shift_augs: the caller shifts start / end; rc_aug: the caller passes the coin flip)."""
    interval_length = end - start
    chromosome_length = len(chromosome)
    left_padding = right_padding = 0
    if interval_length < max_length:
        extra_seq = max_length - interval_length
        extra_left_seq = extra_seq // 2
        extra_right_seq = extra_seq - extra_left_seq
        start -= extra_left_seq
        end += extra_right_seq
    if start < 0:
        left_padding = -start
        start = 0
    if end > chromosome_length:
        right_padding = end - chromosome_length
        end = chromosome_length
    if interval_length > max_length:
        end = start + max_length
    seq = chromosome[start:end]
    if reverse_complement:
        seq = reverse_complement_ref(seq)
    if pad_interval:
        seq = ("." * left_padding) + seq + ("." * right_padding)
    return seq


def bert_mask_ref(seq: torch.Tensor, mask_token_id: int, pad_token_id: int, vocab_size: int, mask_prob: float = 0.15,
                  random_token_prob: float = 0.1, unchanged_token_prob: float = 0.1, special_token_ids=None):
    """bert_mask (hg38_dataset.py:238-286): same statements, same order of draws from torch's global CPU generator.
    Returns (seq, mask, labels) like the reference plus the draws (r_mask, r_kind, random_tokens) so that a kernel can be
    fed the very same randomness."""
    seq = seq.clone()
    r_mask = torch.rand(seq.shape)
    mask = (seq != pad_token_id) & (r_mask < mask_prob)
    labels = seq.clone()
    labels[~mask] = -100
    rand = torch.rand(seq.shape)
    indices_masked = mask & (rand < (1 - random_token_prob - unchanged_token_prob))
    seq[indices_masked] = mask_token_id
    indices_random = mask & (rand >= (1 - random_token_prob - unchanged_token_prob)) & (rand < (1 - unchanged_token_prob))
    random_tokens = torch.randint(0, vocab_size, seq.shape, dtype=torch.long)
    special = torch.tensor(special_token_ids)
    while torch.isin(random_tokens, special).any():
        bad = torch.isin(random_tokens, special)
        random_tokens[bad] = torch.randint(0, vocab_size, (random_tokens[bad].shape[0],), dtype=torch.long)
    seq[indices_random] = random_tokens[indices_random]
    return (seq, mask, labels), (r_mask, rand, random_tokens)
