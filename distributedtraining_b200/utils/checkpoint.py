"""Checkpoint / resume (the reference parses ``--save_every`` and never reads it; its only durable state is the
averaged model in the hub: SURVEY.md section 5.4).  Saved per rank: master, base, Adam moments, step, hyper-parameters,
round counter; plus role extras (mixing weights ``w``, validator score EMA)."""
from __future__ import annotations

import glob
import os
import re
from typing import Any, Dict, Optional

import torch

from .logging import logger


def _path(cfg, rank: int, round: int) -> str:
    return os.path.join(cfg.checkpoint_dir, f"rank{rank}_round{round:08d}.pt")


def save_checkpoint(cfg, trainer, rank: int, round: int, extra: Optional[Dict[str, Any]] = None) -> str:
    os.makedirs(cfg.checkpoint_dir, exist_ok=True)
    blob = {"round": round, "fingerprint": trainer.man.fingerprint(),
            "trainer": {k: v.detach().cpu() for k, v in trainer.state_dict().items()}, "extra": extra or {}}
    path = _path(cfg, rank, round)
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(blob, tmp)
    os.replace(tmp, path)
    return path


def latest_checkpoint(cfg, rank: int) -> Optional[str]:
    files = glob.glob(os.path.join(cfg.checkpoint_dir, f"rank{rank}_round*.pt"))
    if not files:
        return None
    return max(files, key=lambda f: int(re.search(r"round(\d+)", f).group(1)))


def load_checkpoint(path: str, trainer) -> Dict[str, Any]:
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob["fingerprint"] != trainer.man.fingerprint():
        raise ValueError("checkpoint was written for a different model layout")
    trainer.load_state_dict({k: v.to(trainer.master.device) for k, v in blob["trainer"].items()})
    return blob


def maybe_resume(cfg, trainer, rank: int) -> int:
    if not getattr(cfg, "resume", False):
        return 0
    p = latest_checkpoint(cfg, rank)
    if p is None:
        logger.info("--resume: no checkpoint found, starting fresh")
        return 0
    blob = load_checkpoint(p, trainer)
    logger.info(f"resumed from {p} (round {blob['round']})")
    return int(blob["round"])
