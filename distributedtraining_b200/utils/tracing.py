"""Tracing / profiling hooks (absent from the reference): NVTX ranges per round phase, CUDA-event
phase timers (device time, no host sync until read), optional torch.profiler export."""
from __future__ import annotations

import contextlib
from collections import defaultdict
from typing import Dict, List, Tuple

import torch

PHASES = ("local_steps", "delta_emit", "gather_avg", "meta_learning", "broadcast_gemm", "validate")


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class PhaseTimer:
    """``with timer.phase("gather_avg"): ...`` records a CUDA-event pair; ``summary()`` syncs once and returns ms."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled and torch.cuda.is_available()
        self._events: Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = defaultdict(list)

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            with nvtx_range(name):
                yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with nvtx_range(name):
            yield
        e1.record()
        self._events[name].append((e0, e1))

    def summary(self, reset: bool = True) -> Dict[str, float]:
        if not self.enabled:
            return {}
        torch.cuda.synchronize()
        out = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self._events.items()}
        if reset:
            self._events.clear()
        return out


@contextlib.contextmanager
def torch_profile(path: str, enabled: bool = True):
    if not enabled:
        yield None
        return
    from torch.profiler import ProfilerActivity, profile
    acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
    with profile(activities=acts) as prof:
        yield prof
    prof.export_chrome_trace(path)
