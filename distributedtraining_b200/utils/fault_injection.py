"""Fault injection (absent from the reference): ``--inject kind:rank[,kind:rank...]`` with
kind in {nan, shape, stall, drop}.  Applied at delta-publish time on the named ranks so the averager's NaN screen, shape
screen, stale-flag handling and missing-miner path can be exercised end to end."""
from __future__ import annotations

import time
from typing import Dict, List

import torch

KINDS = ("nan", "shape", "stall", "drop")


def parse_inject(spec: str) -> Dict[int, List[str]]:
    out: Dict[int, List[str]] = {}
    for part in filter(None, (spec or "").split(",")):
        kind, rank = part.split(":")
        if kind not in KINDS:
            raise ValueError(f"unknown fault {kind!r}; choose from {KINDS}")
        out.setdefault(int(rank), []).append(kind)
    return out


class FaultyExchange:
    """Wraps an exchange; corrupts / delays / drops this rank's publishes according to the plan."""

    def __init__(self, exchange, rank: int, plan: Dict[int, List[str]], stall_s: float = 2.0):
        self._ex, self._rank, self._faults, self._stall = exchange, rank, plan.get(rank, []), stall_s

    def __getattr__(self, k):
        return getattr(self._ex, k)

    def publish_delta(self, trainer, round: int, *a, **kw):
        if "drop" in self._faults:
            return  # never publishes: the flag stays stale, receivers see None
        if "stall" in self._faults:
            time.sleep(self._stall)
        if "shape" in self._faults and hasattr(self._ex, "_delta_path"):
            torch.save({"round": round, "fingerprint": "wrong-shape", "delta": torch.zeros(7)}, self._ex._delta_path(self._rank))
            return
        if "nan" in self._faults:
            class _NaN:
                master = trainer.master
                def emit_delta(_, out, scales=None, bad=None):
                    # poison one MASTER element for the duration of the emit so that the delta kernel's own NaN screen
                    # (the verdict that travels with the publish flag) sees it, exactly like a diverged miner would
                    k = trainer.master.numel() // 3
                    keep = trainer.master[k].clone()
                    trainer.master[k] = float("nan")
                    try:
                        trainer.emit_delta(out, scales, bad) if bad is not None else trainer.emit_delta(out, scales)
                    finally:
                        trainer.master[k] = keep
                    return out
            return self._ex.publish_delta(_NaN(), round, *a, **kw)
        return self._ex.publish_delta(trainer, round, *a, **kw)
