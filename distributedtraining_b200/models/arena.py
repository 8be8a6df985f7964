"""Flat parameter arenas + manifest.

The reference moves model state around as ``dict[str, Tensor]`` with 148 separate tensors and loops over them in
Python for every operation (delta emit: reference hivetrain/training_manager.py:417-421; weighted average:
hivetrain/averaging_logic.py:431-448; meta-gradient: :513-522).  Here every model's parameters live in ONE contiguous
buffer; sibling arenas (grad, Adam m/v, base snapshot, delta) share the same :class:`Manifest`, so each of those loops
becomes one kernel launch and a delta is a single peer-addressable window.

The per-tensor mixing matrix ``w[N, P]`` of the learned averager indexes manifest entries.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass
from typing import Dict, Iterator, List, Sequence, Tuple

import torch

ALIGN = 256  # elements; keeps every tensor 512 B (bf16) / 1 KB (fp32) aligned -> TMA- and uint4-friendly
CHUNK = 4096  # elements per work chunk of the segmented kernels (weighted average / multi-dot)


@dataclass(frozen=True)
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int
    init: str = "normal"  # normal | zeros | ones | normal_resid
    decay: bool = True

    @property
    def padded(self) -> int:
        return (self.numel + ALIGN - 1) // ALIGN * ALIGN


class Manifest:
    """Ordered list of parameter tensors and their offsets in the flat arena."""

    def __init__(self, entries: Sequence[Tuple[str, Sequence[int], str, bool]]):
        specs: List[ParamSpec] = []
        off = 0
        for name, shape, init, decay in entries:
            numel = int(math.prod(shape))
            spec = ParamSpec(name, tuple(int(s) for s in shape), off, numel, init, decay)
            specs.append(spec)
            off += spec.padded
        self.specs = specs
        self.total = off
        self._index = {s.name: i for i, s in enumerate(specs)}
        self._chunk_cache: Dict[Tuple[str, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    # -- basic queries -------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.specs)

    def __iter__(self) -> Iterator[ParamSpec]:
        return iter(self.specs)

    def __getitem__(self, key) -> ParamSpec:
        if isinstance(key, str):
            return self.specs[self._index[key]]
        return self.specs[key]

    def index(self, name: str) -> int:
        return self._index[name]

    @property
    def names(self) -> List[str]:
        return [s.name for s in self.specs]

    @property
    def num_params(self) -> int:
        return sum(s.numel for s in self.specs)

    def fingerprint(self) -> str:
        h = hashlib.sha256()
        for s in self.specs:
            h.update(f"{s.name}:{s.shape}:{s.offset};".encode())
        return h.hexdigest()[:16]

    def same_layout(self, shapes: Dict[str, Sequence[int]]) -> bool:
        """Shape screen of the averager (reference averaging_logic.py:406-410) done once on metadata."""
        if set(shapes) != set(self._index):
            return False
        return all(tuple(shapes[s.name]) == s.shape for s in self.specs)

    # -- views ---------------------------------------------------------------------------------------------------
    def view(self, flat: torch.Tensor, name_or_idx) -> torch.Tensor:
        s = self[name_or_idx]
        return flat[s.offset:s.offset + s.numel].view(s.shape)

    def views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {s.name: flat[s.offset:s.offset + s.numel].view(s.shape) for s in self.specs}

    def pack(self, tensors: Dict[str, torch.Tensor], out: torch.Tensor) -> torch.Tensor:
        """Flatten a name->tensor dict into ``out``.  Shapes must match the manifest exactly (the reference's shape
        screen, averaging_logic.py:406-410): a transposed / foreign-layout tensor of the right numel is an error, never a
        silent scramble -- convert HF-layout dicts with ``models.transformer.from_hf_state_dict`` first."""
        for s in self.specs:
            t = tensors[s.name]
            if tuple(t.shape) != s.shape:
                raise ValueError(f"{s.name}: shape {tuple(t.shape)} does not match the manifest's {s.shape}")
            out[s.offset:s.offset + s.numel].copy_(t.reshape(-1))
        return out

    # -- segment tables for the segmented kernels ------------------------------------------------------------------
    def seg_table(self, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(chunk_start[int64], chunk_len[int32], chunk_tensor[int32]): fixed-size chunks that never straddle tensors."""
        key = (str(device), CHUNK)
        if key not in self._chunk_cache:
            starts, lens, tids = [], [], []
            for j, s in enumerate(self.specs):
                for c0 in range(0, s.padded, CHUNK):
                    starts.append(s.offset + c0)
                    lens.append(min(CHUNK, s.padded - c0))
                    tids.append(j)
            self._chunk_cache[key] = (
                torch.tensor(starts, dtype=torch.int64, device=device),
                torch.tensor(lens, dtype=torch.int32, device=device),
                torch.tensor(tids, dtype=torch.int32, device=device),
            )
        return self._chunk_cache[key]

    def tensor_ids(self, device) -> torch.Tensor:
        """int32[total]: tensor index of every arena element (reference implementation of the segmented ops)."""
        key = (str(device), -1)
        if key not in self._chunk_cache:
            ids = torch.empty(self.total, dtype=torch.int64, device=device)
            for j, s in enumerate(self.specs):
                ids[s.offset:s.offset + s.padded] = j
            self._chunk_cache[key] = (ids, ids, ids)
        return self._chunk_cache[key][0]

    def decay_mask(self, device) -> torch.Tensor:
        m = torch.zeros(self.total, dtype=torch.float32, device=device)
        for s in self.specs:
            if s.decay:
                m[s.offset:s.offset + s.numel] = 1.0
        return m


class Arena:
    """One flat buffer laid out by a :class:`Manifest`."""

    def __init__(self, manifest: Manifest, dtype=torch.float32, device="cpu", flat: torch.Tensor | None = None):
        self.manifest = manifest
        if flat is None:
            flat = torch.zeros(manifest.total, dtype=dtype, device=device)
        assert flat.numel() == manifest.total and flat.dim() == 1
        self.flat = flat

    @property
    def device(self):
        return self.flat.device

    @property
    def dtype(self):
        return self.flat.dtype

    def view(self, name) -> torch.Tensor:
        return self.manifest.view(self.flat, name)

    def views(self) -> Dict[str, torch.Tensor]:
        return self.manifest.views(self.flat)

    def state_dict(self, clone: bool = True) -> Dict[str, torch.Tensor]:
        v = self.views()
        return {k: (t.detach().clone() if clone else t.detach()) for k, t in v.items()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        for s in self.manifest:
            if s.name not in sd:
                if strict:
                    raise KeyError(s.name)
                continue
            self.view(s.name).copy_(sd[s.name].to(self.flat.dtype).reshape(s.shape))

    def clone(self) -> "Arena":
        return Arena(self.manifest, flat=self.flat.clone())

    def zero_(self) -> "Arena":
        self.flat.zero_()
        return self

    def sha256(self) -> str:
        """Model hash (reference validation_logic.py:198-203 hashes every parameter on the CPU)."""
        h = hashlib.sha256()
        h.update(self.flat.detach().to("cpu", torch.float32).contiguous().numpy().tobytes())
        return h.hexdigest()


def init_arena_(arena: Arena, std: float = 0.02, n_layer: int = 1, seed: int = 0) -> Arena:
    """GPT-2 style random init (no pretrained weights are reachable offline)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    for s in arena.manifest:
        v = arena.view(s.name)
        if s.init == "zeros":
            v.zero_()
        elif s.init == "ones":
            v.fill_(1.0)
        else:
            sd = std / math.sqrt(2.0 * n_layer) if s.init == "normal_resid" else std
            v.copy_((torch.randn(s.shape, generator=g, dtype=torch.float32) * sd).to(v.dtype))
    return arena
