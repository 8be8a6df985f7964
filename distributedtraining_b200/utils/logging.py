"""Logging + structured JSONL metrics (replaces ``bittensor.logging`` and stray prints of the reference)."""
from __future__ import annotations

import json
import logging as _logging
import os
import sys
import time
from typing import Any, Dict, Optional

logger = _logging.getLogger("dtb200")
if not logger.handlers:
    _h = _logging.StreamHandler(sys.stderr)
    _h.setFormatter(_logging.Formatter("%(asctime)s | %(levelname)-7s | r" + os.environ.get("RANK", "0") + " | %(message)s",
                                       "%H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(_logging.DEBUG if os.environ.get("DTB200_DEBUG") else _logging.INFO)
    logger.propagate = False


def enable_debug() -> None:
    logger.setLevel(_logging.DEBUG)


class MetricsLogger:
    """One JSON object per line: ``{"ts":…, "rank":…, "role":…, "round":…, <metrics>}``."""

    def __init__(self, path: Optional[str] = None, role: str = "", rank: int = 0):
        self.path, self.role, self.rank = path, role, rank
        self._f = None
        if path:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self._f = open(path, "a", buffering=1)
        self.last: Dict[str, Any] = {}

    def log(self, **metrics) -> Dict[str, Any]:
        rec = {"ts": time.time(), "rank": self.rank, "role": self.role, **metrics}
        self.last = rec
        if self._f:
            self._f.write(json.dumps(rec, default=float) + "\n")
        return rec

    def close(self) -> None:
        if self._f:
            self._f.close()
            self._f = None
