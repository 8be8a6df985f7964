"""Fake peer for exercising a validator/averager without training (reference hivetrain/utils/dummy_miner.py:25-82 posts
random "loss" metrics with a hotkey-signed timestamp to an HTTP endpoint that no longer exists).

The in-box equivalent publishes a *synthetic delta* (random, zero, NaN or wrong-shape) through any exchange every
``interval`` seconds, signed with an HMAC of (hotkey, round) so a receiver can authenticate the sender."""
from __future__ import annotations

import hashlib
import hmac
import time
from typing import Optional

import torch


class _FlatDelta:
    def __init__(self, flat):
        self.flat, self.master = flat, flat

    def emit_delta(self, out, scales=None, bad=None):
        out[: self.flat.numel()].copy_(self.flat.to(out.dtype))
        return out


class ValidationCommunicator:
    def __init__(self, exchange, manifest, hotkey: str = "dummy", secret: bytes = b"dtb200", kind: str = "random",
                 scale: float = 1e-3, interval: float = 60.0, device="cpu", seed: int = 0):
        self.exchange, self.man, self.hotkey, self.secret = exchange, manifest, hotkey, secret
        self.kind, self.scale, self.interval, self.device = kind, scale, interval, device
        self.gen = torch.Generator().manual_seed(seed)
        self.round = 0

    def create_signed_message(self, round: int) -> dict:
        msg = f"{self.hotkey}:{round}:{int(time.time())}"
        return {"message": msg, "signature": hmac.new(self.secret, msg.encode(), hashlib.sha256).hexdigest()}

    @staticmethod
    def verify(message: dict, secret: bytes = b"dtb200") -> bool:
        want = hmac.new(secret, message["message"].encode(), hashlib.sha256).hexdigest()
        return hmac.compare_digest(want, message["signature"])

    def make_delta(self) -> torch.Tensor:
        n = self.man.total
        if self.kind == "zero":
            d = torch.zeros(n)
        elif self.kind == "nan":
            d = torch.zeros(n)
            d[n // 2] = float("nan")
        else:
            d = torch.randn(n, generator=self.gen) * self.scale
        return d.to(self.device)

    def send(self) -> dict:
        self.round += 1
        self.exchange.publish_delta(_FlatDelta(self.make_delta()), self.round)
        return self.create_signed_message(self.round)

    def start(self, rounds: Optional[int] = None) -> None:
        i = 0
        while rounds is None or i < rounds:
            self.send()
            i += 1
            if rounds is None or i < rounds:
                time.sleep(self.interval)
