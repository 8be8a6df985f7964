"""Data pipeline stand-ins (there is no network: WikiText-103 and the GPT-2 BPE files are unreachable).

Reference pipeline: ``WikitextDataset`` tokenises each raw line to exactly ``max_length`` ids,
right-padded with ``[PAD]`` (id 50257), ``labels = input_ids.clone()`` with PAD *not* masked, plain ``DataLoader`` without
shuffle (reference neurons/miner.py:69-106).  :class:`SyntheticTokens` yields batches of the same dict shape from a
pre-generated pool in **pinned host memory** (so the per-step H2D copy is a real, asynchronous PCIe transfer);
:class:`WikitextDataset` + :func:`custom_collate_fn` keep the reference's interface for real text with any tokenizer.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence

import torch

PAD_ID = 50257


class SyntheticTokens:
    """Infinite (or ``steps``-long) iterable of ``{"input_ids", "attention_mask", "labels"}`` batches [B, T] int32."""

    def __init__(self, batch: int, seq: int, vocab: int = 50258, pad_id: int = PAD_ID, pad_fraction: float = 0.25,
                 seed: int = 0, pool: int = 8, steps: Optional[int] = None, pin: Optional[bool] = None, device: str = "cpu",
                 zipf: bool = True):
        g = torch.Generator().manual_seed(seed)
        pin = torch.cuda.is_available() if pin is None else pin
        self.steps = steps
        self.pool: List[Dict[str, torch.Tensor]] = []
        real_vocab = min(vocab, pad_id) if pad_id < vocab else vocab
        for _ in range(pool):
            if zipf:  # Zipf-like marginal, closer to text than uniform ids
                u = torch.rand(batch, seq, generator=g)
                ids = (real_vocab ** u - 1).long().clamp_(0, real_vocab - 1)
            else:
                ids = torch.randint(0, real_vocab, (batch, seq), generator=g)
            lens = ((1.0 - pad_fraction * torch.rand(batch, generator=g)) * seq).long().clamp_(1, seq)
            mask = torch.arange(seq)[None, :] < lens[:, None]
            if pad_id < vocab:
                ids = torch.where(mask, ids, torch.full_like(ids, pad_id))  # right padding, like tokenizer(padding="max_length")
            ids = ids.to(torch.int32)
            # "kv_len" = attention_mask.sum(-1): the padding mask in the form the attention kernels consume (4 B per row)
            b = {"input_ids": ids, "attention_mask": mask.to(torch.int32), "labels": ids.clone(),
                 "kv_len": lens.to(torch.int32) if pad_id < vocab else torch.full((batch,), seq, dtype=torch.int32)}
            if device != "cpu":
                b = {k: v.to(device) for k, v in b.items()}
            elif pin:
                b = {k: v.pin_memory() for k, v in b.items()}
            self.pool.append(b)
        self.bytes_per_batch = batch * seq * 4 + batch * 4  # input_ids + kv_len: labels / mask are derived on the device

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        i = 0
        while self.steps is None or i < self.steps:
            yield self.pool[i % len(self.pool)]
            i += 1

    def __len__(self) -> int:
        return self.steps if self.steps is not None else 1 << 62


class ByteTokenizer:
    """Offline fallback tokenizer: UTF-8 bytes (ids 0..255) + ``[PAD]``; same call signature subset as HF tokenizers."""

    def __init__(self, pad_id: int = 256):
        self.pad_token_id = pad_id
        self.vocab_size = pad_id + 1

    def __len__(self) -> int:
        return self.vocab_size

    def add_special_tokens(self, d: Dict[str, str]) -> int:
        return 0

    def __call__(self, text: str, max_length: int = 64, padding: str = "max_length", truncation: bool = True,
                 return_tensors: Optional[str] = None) -> Dict[str, torch.Tensor]:
        ids = list(text.encode("utf-8"))[:max_length]
        mask = [1] * len(ids) + [0] * (max_length - len(ids))
        ids = ids + [self.pad_token_id] * (max_length - len(ids))
        return {"input_ids": torch.tensor([ids], dtype=torch.long), "attention_mask": torch.tensor([mask], dtype=torch.long)}


class WikitextDataset(torch.utils.data.Dataset):
    """Per-item tokenisation to exactly ``max_length`` ids (reference neurons/miner.py:69-92)."""

    def __init__(self, texts: Sequence[str], tokenizer, max_length: int = 64):
        self.texts, self.tokenizer, self.max_length = list(texts), tokenizer, max_length

    def __len__(self) -> int:
        return len(self.texts)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        enc = self.tokenizer(self.texts[idx], max_length=self.max_length, padding="max_length", truncation=True,
                             return_tensors="pt")
        ids = enc["input_ids"].squeeze(0)
        return {"input_ids": ids, "attention_mask": enc["attention_mask"].squeeze(0), "labels": ids.clone()}


def custom_collate_fn(batch: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Stack dict items (reference neurons/miner.py:95-99); ids become int32 for the kernels."""
    ids = torch.stack([b["input_ids"] for b in batch]).to(torch.int32)
    return {"input_ids": ids, "attention_mask": torch.stack([b["attention_mask"] for b in batch]).to(torch.int32),
            "labels": ids.clone()}


def make_loader(dataset, batch_size: int, drop_last: bool = False):
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, collate_fn=custom_collate_fn, shuffle=False,
                                       drop_last=drop_last)


class SyntheticMNIST:
    """Learnable 10-class toy images (class-dependent blobs + noise) standing in for MNIST in the simulations."""

    def __init__(self, n: int = 512, batch: int = 64, seed: int = 0, shape=(1, 28, 28)):
        g = torch.Generator().manual_seed(seed)
        protos = torch.randn(10, *shape, generator=torch.Generator().manual_seed(1234))
        self.y = torch.randint(0, 10, (n,), generator=g)
        self.x = protos[self.y] + 0.5 * torch.randn(n, *shape, generator=g)
        self.batch = batch

    def __iter__(self):
        for i in range(0, len(self.y), self.batch):
            yield self.x[i:i + self.batch], self.y[i:i + self.batch]

    def __len__(self):
        return (len(self.y) + self.batch - 1) // self.batch
