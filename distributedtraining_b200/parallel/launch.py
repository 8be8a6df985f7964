"""Process bootstrap: one process per GPU, ``torch.distributed`` for the plumbing (NCCL on GPUs, gloo on CPU).

The reference parses ``--rank/--world-size/--store-address/--store-port`` and never reads them
(reference hivetrain/config/hivetrain_config.py:21-32); here they (or torchrun's environment) drive the rendezvous.
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def env_rank_world(config=None) -> Tuple[int, int, int]:
    rank = int(os.environ.get("RANK", getattr(config, "rank", None) or 0))
    world = int(os.environ.get("WORLD_SIZE", getattr(config, "world_size", None) or 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    return rank, world, local


def init_distributed(backend: Optional[str] = None, config=None, timeout_s: int = 600) -> Tuple[int, int, torch.device]:
    """Initialise the default process group (no-op for a single process) and select this rank's device."""
    rank, world, local = env_rank_world(config)
    use_cuda = torch.cuda.is_available() and (backend in (None, "nccl", "peer"))
    if use_cuda:
        torch.cuda.set_device(local % torch.cuda.device_count())
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", getattr(config, "store_address", None) or "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(getattr(config, "store_port", None) or 4999))
        os.environ.setdefault("RANK", str(rank))
        os.environ.setdefault("WORLD_SIZE", str(world))
        be = "nccl" if use_cuda else "gloo"
        kw = {"device_id": device} if use_cuda else {}
        dist.init_process_group(backend=be, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=timeout_s), **kw)
    return rank, world, device


def parse_roles(spec: str, world: int) -> Dict[str, List[int]]:
    """``'miner:0-6,validator:7,averager:0'`` -> {'miner': [0..6], 'validator': [7], 'averager': [0]}.
    Empty spec: every rank mines and rank 0 also averages (BASELINE.json config 2)."""
    roles: Dict[str, List[int]] = {"miner": [], "validator": [], "averager": []}
    if not spec:
        roles["miner"] = list(range(world))
        roles["averager"] = [0]
        return roles
    for part in spec.split(","):
        name, rng = part.split(":")
        ranks: List[int] = []
        for piece in rng.split("+"):
            if "-" in piece:
                a, b = piece.split("-")
                ranks += list(range(int(a), int(b) + 1))
            else:
                ranks.append(int(piece))
        roles.setdefault(name.strip(), []).extend(r for r in ranks if r < world)
    return roles


def barrier_sync(device: torch.device) -> None:
    if dist.is_initialized():
        if device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: torch.device) -> float:
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
