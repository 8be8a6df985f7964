"""Real-text data path of the three roles (reference neurons/miner.py:54-106, validator.py:45-98, averager.py:58-94).

The reference loads WikiText-103 through ``datasets`` and tokenises with the GPT-2 BPE + an added ``[PAD]``; neither the
dataset nor the BPE files are reachable offline, so the source here is any local TEXT FILE (one text per line, blank and
``= Heading =`` lines kept, exactly like the raw WikiText lines the reference feeds) and the tokenizer is

* an HF tokenizer DIRECTORY (``--data.tokenizer /path/with/tokenizer.json``): loaded with ``AutoTokenizer`` and given the
  ``[PAD]`` token the way every reference role does (``add_special_tokens({'pad_token': '[PAD]'})``), or
* ``byte`` (default): UTF-8 bytes with ``[PAD] = vocab - 1`` -- always available.

Everything downstream is the reference's pipeline: per-item tokenisation to exactly ``max_length`` ids with right padding
(:class:`WikitextDataset`), ``labels = input_ids`` (PAD not masked), ``attention_mask`` passed to the model, a plain
``DataLoader`` without shuffle; batches are pinned so the per-step host->device copy is asynchronous.
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, List, Optional

import torch

from .synthetic import ByteTokenizer, WikitextDataset, custom_collate_fn


def read_lines(path: str, limit: Optional[int] = None) -> List[str]:
    out: List[str] = []
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        for line in f:
            out.append(line.rstrip("\n"))
            if limit is not None and len(out) >= limit:
                break
    return out


def build_tokenizer(spec: Optional[str], vocab_size: int):
    """``byte`` | HF tokenizer directory.  The returned tokenizer pads with id ``vocab_size - 1`` (the model's [PAD] row)."""
    if spec and spec != "byte" and os.path.isdir(spec):
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(spec)
        if tok.pad_token is None:
            tok.add_special_tokens({"pad_token": "[PAD]"})  # reference neurons/miner.py:61-62
        if len(tok) > vocab_size:
            raise ValueError(f"tokenizer has {len(tok)} tokens but the model's vocabulary is {vocab_size}")
        return tok
    if vocab_size < 258:
        raise ValueError("the byte tokenizer needs a vocabulary of at least 258 ids")
    return ByteTokenizer(pad_id=vocab_size - 1)


class PinnedLoader:
    """Iterable over ``DataLoader`` batches with (a) the ``kv_len`` reduction of the attention mask attached, (b) pinned
    memory on GPU boxes, (c) optional endless repetition (the reference trains for 3e16 "epochs")."""

    def __init__(self, dataset, batch_size: int, drop_last: bool, repeat: bool, max_batches: Optional[int] = None):
        self.dl = torch.utils.data.DataLoader(dataset, batch_size=batch_size, collate_fn=custom_collate_fn, shuffle=False,
                                              drop_last=drop_last, pin_memory=torch.cuda.is_available())
        self.repeat, self.max_batches = repeat, max_batches
        self.bytes_per_batch = batch_size * dataset.max_length * 4 + batch_size * 4

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        n = 0
        while True:
            for b in self.dl:
                b["kv_len"] = b["attention_mask"].sum(dim=1).clamp_(min=1).to(torch.int32)
                yield b
                n += 1
                if self.max_batches is not None and n >= self.max_batches:
                    return
            if not self.repeat:
                return

    def __len__(self) -> int:
        return len(self.dl) if self.max_batches is None else min(len(self.dl), self.max_batches)


def build_text_loader(path: str, tokenizer_spec: Optional[str], vocab_size: int, batch_size: int, max_length: int,
                      limit: Optional[int] = None, drop_last: bool = False, repeat: bool = False,
                      max_batches: Optional[int] = None) -> PinnedLoader:
    texts = read_lines(path, limit)
    if not texts:
        raise ValueError(f"{path} holds no text lines")
    tok = build_tokenizer(tokenizer_spec, vocab_size)
    ds = WikitextDataset(texts, tok, max_length=max_length)
    return PinnedLoader(ds, batch_size, drop_last=drop_last, repeat=repeat, max_batches=max_batches)
