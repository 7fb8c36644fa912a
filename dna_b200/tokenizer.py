"""Character tokenizer / nucleotide encoder on the GPU — call surface of the reference's
`CharacterTokenizer` (src/dataloaders/datasets/hg38_char_tokenizer.py:15-149; standalone twin
standalone_hyenadna.py:939-1073 which also prepends [CLS]) as used by HG38Dataset
(hg38_dataset.py:194-223, 383-386), executed by the `hy_tokenize` kernel.

The reference tokenises one string at a time with per-character dict lookups inside DataLoader
workers; here a whole batch of raw ASCII bytes (1 B/nt over PCIe instead of 8 B/nt int64) is turned
into left-padded int64 ids by one kernel launch in the training process.  No CPU implementation
lives in this package: every entry point runs the kernel.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import kernels as K
from ._lib import TOK_ADD_CLS, TOK_ADD_SEP, TOK_N_TO_PAD, TOK_NUC_ENCODE

_SPECIALS = ["[CLS]", "[SEP]", "[BOS]", "[MASK]", "[PAD]", "[RESERVED]", "[UNK]"]


class CharacterTokenizer:
    def __init__(self, characters: Sequence[str], model_max_length: int, padding_side: str = "left",
                 cls_token: bool = False, **kwargs):
        """`cls_token=True` selects the standalone_hyenadna.py layout ([CLS] ... [SEP], :1010-1018);
        the default is the training-pipeline layout (... [SEP], hg38_char_tokenizer.py:86-94)."""
        if list(characters) != ["A", "C", "G", "T", "N"]:
            raise NotImplementedError("hy_tokenize implements the HyenaDNA alphabet ['A','C','G','T','N'] only")
        if padding_side != "left":
            raise NotImplementedError("only padding_side='left' (the reference's setting, :16) is implemented")
        self.characters = list(characters)
        self.model_max_length = model_max_length
        self.padding_side = padding_side
        self.use_cls_token = cls_token
        self._vocab_str_to_int = {**{s: i for i, s in enumerate(_SPECIALS)},
                                  **{ch: i + 7 for i, ch in enumerate(self.characters)}}
        self._vocab_int_to_str = {v: k for k, v in self._vocab_str_to_int.items()}
        self.cls_token, self.sep_token, self.bos_token, self.mask_token = "[CLS]", "[SEP]", "[BOS]", "[MASK]"
        self.pad_token, self.unk_token, self.eos_token = "[PAD]", "[UNK]", "[SEP]"
        self.cls_token_id, self.sep_token_id, self.bos_token_id, self.mask_token_id = 0, 1, 2, 3
        self.pad_token_id, self.unk_token_id, self.eos_token_id = 4, 6, 1

    # ---- vocabulary surface (hg38_char_tokenizer.py:70-84) ---------------------------------------
    @property
    def vocab_size(self) -> int:
        return len(self._vocab_str_to_int)

    def __len__(self) -> int:
        return self.vocab_size

    def get_vocab(self) -> Dict[str, int]:
        return dict(self._vocab_str_to_int)

    def _tokenize(self, text: str) -> List[str]:
        return list(text)

    def _convert_token_to_id(self, token: str) -> int:
        return self._vocab_str_to_int.get(token, self.unk_token_id)

    def _convert_id_to_token(self, index: int) -> str:
        return self._vocab_int_to_str[index]

    def convert_tokens_to_string(self, tokens):
        return "".join(tokens)

    def decode(self, ids, skip_special_tokens: bool = False) -> str:
        toks = [self._vocab_int_to_str[int(i)] for i in ids]
        if skip_special_tokens:
            toks = [t for t in toks if t not in _SPECIALS]
        return "".join(toks)

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        head = [self.cls_token_id] if self.use_cls_token else []
        out = head + list(token_ids_0) + [self.sep_token_id]
        if token_ids_1 is not None:
            out += list(token_ids_1) + [self.sep_token_id]
        return out

    def get_special_tokens_mask(self, token_ids_0, token_ids_1=None, already_has_special_tokens=False):
        if already_has_special_tokens:
            return [1 if t in (0, 1, 2, 3, 4, 5, 6) else 0 for t in token_ids_0]
        out = ([1] if self.use_cls_token else []) + [0] * len(token_ids_0) + [1]
        if token_ids_1 is not None:
            out += [0] * len(token_ids_1) + [1]
        return out

    # ---- config round trip (hg38_char_tokenizer.py:125-149) --------------------------------------
    def get_config(self) -> Dict:
        return {"char_ords": [ord(ch) for ch in self.characters], "model_max_length": self.model_max_length}

    @classmethod
    def from_config(cls, config: Dict) -> "CharacterTokenizer":
        return cls(characters=[chr(i) for i in config["char_ords"]], model_max_length=config["model_max_length"])

    def save_pretrained(self, save_directory: Union[str, os.PathLike], **kwargs):
        with open(Path(save_directory) / "tokenizer_config.json", "w") as f:
            json.dump(self.get_config(), f, indent=4)

    @classmethod
    def from_pretrained(cls, save_directory: Union[str, os.PathLike], **kwargs):
        with open(Path(save_directory) / "tokenizer_config.json") as f:
            return cls.from_config(json.load(f))

    # ---- the kernel path ------------------------------------------------------------------------
    @staticmethod
    def pack_bytes(texts: Sequence[Union[str, bytes]], pin: bool = True):
        """Host staging: ragged strings -> (uint8 [B, max_chars] pinned, int32 [B] lengths)."""
        raw = [t.encode("latin-1", "replace") if isinstance(t, str) else bytes(t) for t in texts]
        lens = np.array([len(r) for r in raw], dtype=np.int32)
        width = max(int(lens.max()) if len(raw) else 0, 1)
        buf = np.zeros((len(raw), width), dtype=np.uint8)
        for i, r in enumerate(raw):
            buf[i, :len(r)] = np.frombuffer(r, dtype=np.uint8)
        tb, tl = torch.from_numpy(buf), torch.from_numpy(lens)
        if pin and torch.cuda.is_available():
            tb, tl = tb.pin_memory(), tl.pin_memory()
        return tb, tl

    @staticmethod
    def reverse_complement_cuda(seqs: torch.Tensor, lens: Optional[torch.Tensor] = None,
                                apply: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The rc_aug augmentation of the FASTA reader (`string_reverse_complement`, hg38_dataset.py:28-38, :118-119)
        on a device byte batch: rows with apply[b] != 0 (all rows when None) are reversed and complemented over
        their first lens[b] bytes; everything else is copied. Feed the result to encode_bytes_cuda."""
        return K.reverse_complement(seqs, lens, apply)

    def encode_bytes_cuda(self, seqs: torch.Tensor, lens: Optional[torch.Tensor], max_length: int,
                          add_special_tokens: bool = True, replace_N_token: bool = False,
                          nucleotide_encode: bool = False) -> torch.Tensor:
        """uint8 [B, max_chars] (device) -> int64 ids [B, max_length] (device): truncation=True,
        padding='max_length', left padding — the exact call of hg38_dataset.py:194-199 (+ :218-220 /
        :383-386 when the flags are set)."""
        flags = 0
        if add_special_tokens:
            flags |= TOK_ADD_SEP | (TOK_ADD_CLS if self.use_cls_token else 0)
        if replace_N_token:
            flags |= TOK_N_TO_PAD
        if nucleotide_encode:
            flags |= TOK_NUC_ENCODE
        return K.tokenize(seqs, lens, max_length, flags)

    def encode_batch_cuda(self, texts: Sequence[Union[str, bytes]], max_length: Optional[int] = None,
                          add_special_tokens: bool = True, device=None, **kw) -> torch.Tensor:
        max_length = max_length or self.model_max_length
        tb, tl = self.pack_bytes(texts)
        device = device or torch.device("cuda")
        return self.encode_bytes_cuda(tb.to(device, non_blocking=True), tl.to(device, non_blocking=True), max_length,
                                      add_special_tokens=add_special_tokens, **kw)

    def __call__(self, text, add_special_tokens: bool = True, padding=False, max_length: Optional[int] = None,
                 truncation: bool = False, return_tensors: Optional[str] = None, **kwargs):
        """HF-style call used by the reference dataset. Runs the GPU kernel; returns {"input_ids": ...}."""
        single = isinstance(text, (str, bytes))
        texts = [text] if single else list(text)
        n_special = (1 + int(self.use_cls_token)) if add_special_tokens else 0
        longest = max((len(t) for t in texts), default=0) + n_special
        if padding == "max_length":
            width = max_length or self.model_max_length
            if not truncation and longest > width:
                raise ValueError("sequence longer than max_length and truncation=False")
        else:
            if len({len(t) for t in texts}) > 1 and padding not in (True, "longest"):
                raise ValueError("ragged batch needs padding")
            width = min(longest, max_length) if (truncation and max_length) else longest
        ids = self.encode_batch_cuda(texts, max_length=max(width, 1), add_special_tokens=add_special_tokens)
        if return_tensors == "pt":
            return {"input_ids": ids[0] if single else ids}
        out = ids.cpu().tolist()
        return {"input_ids": out[0] if single else out}



# ------------------------------------------------------------------------------------------------
# data ingest ahead of the tokenizer: interval fetch + BERT masking on the device
# ------------------------------------------------------------------------------------------------
class DeviceFastaInterval:
    """`FastaInterval` (src/dataloaders/datasets/hg38_dataset.py:40-124) with the chromosomes resident in HBM as raw
    bytes (1 B / nt: all of hg38 is 3.1 GB): __call__ takes BATCHES of intervals of one chromosome and returns the padded
    byte rows + lengths the tokenizer kernel consumes — fetch, symmetric widening, clipping, '.' padding and the
    reverse-complement augmentation in ONE launch, no Python per-character work.  The random draws (shift, coin flip)
    use torch's generator on the host exactly where the reference draws them (:82-90, :118)."""

    def __init__(self, chromosomes, *, shift_augs=None, rc_aug=False, pad_interval=False, device=None):
        device = device or torch.device("cuda")
        self.seqs = {}
        for name, seq in chromosomes.items():
            if isinstance(seq, str):
                seq = seq.encode()
            if isinstance(seq, (bytes, bytearray)):
                seq = torch.frombuffer(bytearray(seq), dtype=torch.uint8)
            self.seqs[name] = seq.to(device)
        self.chr_lens = {k: int(v.numel()) for k, v in self.seqs.items()}
        self.shift_augs, self.rc_aug, self.pad_interval = shift_augs, rc_aug, pad_interval

    def __call__(self, chr_name, starts, ends, max_length, generator=None):
        chrom = self.seqs[chr_name]
        starts = torch.as_tensor(starts, dtype=torch.int64)
        ends = torch.as_tensor(ends, dtype=torch.int64)
        if self.shift_augs is not None:
            lo, hi = self.shift_augs
            n = self.chr_lens[chr_name]
            min_shift = torch.clamp(starts + lo, min=0) - starts
            max_shift = torch.clamp(ends + hi + 1, max=n) - ends
            span = torch.clamp(max_shift - min_shift, min=1)
            r = (torch.rand(starts.shape, generator=generator) * span).floor().to(torch.int64) + min_shift
            starts, ends = starts + r, ends + r
        rc = None
        if self.rc_aug:
            rc = torch.rand(starts.shape, generator=generator) > 0.5
        dev = chrom.device
        return K.fetch_intervals(chrom, starts.to(dev), ends.to(dev), max_length, rc=None if rc is None else rc.to(dev),
                                 pad_interval=self.pad_interval)


def bert_mask_cuda(seq, mask_token_id, pad_token_id, vocab_size, mask_prob=0.15, random_token_prob=0.1,
                   unchanged_token_prob=0.1, special_token_ids=None, generator=None):
    """`bert_mask` (hg38_dataset.py:238-286) on the device: draws the two uniform fields and the replacement tokens
    (uniform over the non-special ids — what the reference's re-draw loop at :270-273 converges to) with torch's CUDA
    generator, then ONE kernel forms the masked sequence, the mask and the labels.  Returns (seq, mask, labels) like the
    reference (which also overwrites `seq` in place; here the input is left untouched)."""
    dev = seq.device
    r_mask = torch.rand(seq.shape, device=dev, generator=generator)
    r_kind = torch.rand(seq.shape, device=dev, generator=generator)
    special = set(int(i) for i in (special_token_ids or []))
    allowed = torch.tensor([i for i in range(vocab_size) if i not in special], dtype=torch.int64, device=dev)
    rand_tok = allowed[torch.randint(0, allowed.numel(), seq.shape, device=dev, generator=generator)]
    return K.bert_mask(seq, r_mask, r_kind, rand_tok, mask_token_id, pad_token_id, mask_prob, random_token_prob,
                       unchanged_token_prob)
