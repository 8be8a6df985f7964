"""Checkpoint / resume of ALL durable state, periodically (``--save_every N`` rounds) and at exit.

The reference parses ``--save_every`` and never reads it (hivetrain/config/hivetrain_config.py:43-50); its only durable state is
the averaged model in the hub; the optimizer, the mixing weights ``w`` (reset each round) and the validator's score EMA
(``base_scores``, btt_connector.py:305-307) live in memory only.  Saved here, per role and rank:

* miner / co-located rank: master, base, Adam moments, step, hyper-parameters, round counter, global step;
* averager: the model arenas + ``w[N, P]``, the per-miner consumed-round table, the published base round;
* validator: the model arenas + raw / normalised / loss scores, base loss, the network's score EMA;
* co-located coordinator: ``w``, round and meta-step counters.

Files are written atomically (tmp + rename), so a process killed mid-write leaves the previous checkpoint intact; only the
newest ``keep`` files per (role, rank) are retained.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Any, Dict, Optional

import torch

from .logging import logger


def _path(cfg, rank: int, round: int, role: str = "") -> str:
    prefix = f"{role}_" if role else ""
    return os.path.join(cfg.checkpoint_dir, f"{prefix}rank{rank}_round{round:08d}.pt")


def _cpu(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu()
    if isinstance(x, dict):
        return {k: _cpu(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_cpu(v) for v in x)
    return x


def save_checkpoint(cfg, trainer, rank: int, round: int, extra: Optional[Dict[str, Any]] = None, role: str = "", keep: int = 2) -> str:
    os.makedirs(cfg.checkpoint_dir, exist_ok=True)
    blob = {"round": int(round), "role": role, "fingerprint": trainer.man.fingerprint(),
            "trainer": {k: v.detach().cpu() for k, v in trainer.state_dict().items()}, "extra": _cpu(extra or {})}
    path = _path(cfg, rank, round, role)
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(blob, tmp)
    os.replace(tmp, path)
    if keep > 0:  # prune: only the newest ``keep`` checkpoints of this (role, rank) stay on disk
        for old in _all(cfg, rank, role)[:-keep]:
            try:
                os.remove(old)
            except OSError:
                pass
    return path


def _all(cfg, rank: int, role: str = ""):
    prefix = f"{role}_" if role else ""
    files = [f for f in glob.glob(os.path.join(cfg.checkpoint_dir, f"{prefix}rank{rank}_round*.pt")) if ".tmp." not in f]
    return sorted(files, key=lambda f: int(re.search(r"round(\d+)", f).group(1)))


def latest_checkpoint(cfg, rank: int, role: str = "") -> Optional[str]:
    files = _all(cfg, rank, role)
    return files[-1] if files else None


def load_checkpoint(path: str, trainer) -> Dict[str, Any]:
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob["fingerprint"] != trainer.man.fingerprint():
        raise ValueError("checkpoint was written for a different model layout")
    trainer.load_state_dict({k: v.to(trainer.master.device) for k, v in blob["trainer"].items()})
    return blob


def maybe_resume(cfg, trainer, rank: int, role: str = "") -> int:
    """``--resume``: load the newest checkpoint of (role, rank) into ``trainer``; returns its round (0 = fresh start).
    The blob (with the role's ``extra``) is kept on ``trainer._resume_blob`` for the role object to pick up."""
    trainer._resume_blob = None
    if not getattr(cfg, "resume", False):
        return 0
    p = latest_checkpoint(cfg, rank, role)
    if p is None:
        logger.info("--resume: no checkpoint found, starting fresh")
        return 0
    blob = load_checkpoint(p, trainer)
    trainer._resume_blob = blob
    logger.info(f"resumed from {p} (round {blob['round']})")
    return int(blob["round"])


class PeriodicCheckpointer:
    """``hook(obj, round)`` for the role loops: every ``save_every`` rounds write trainer state + ``obj.state_dict()``."""

    def __init__(self, cfg, trainer, rank: int, role: str):
        self.cfg, self.trainer, self.rank, self.role = cfg, trainer, rank, role
        self.every = int(getattr(cfg, "save_every", 0) or 0)
        self.saved = 0

    def __call__(self, obj, round: int, force: bool = False) -> Optional[str]:
        if self.every <= 0 or (not force and round % self.every != 0):
            return None
        extra = obj.state_dict() if hasattr(obj, "state_dict") else {}
        self.saved += 1
        return save_checkpoint(self.cfg, self.trainer, self.rank, round, extra=extra, role=self.role)
