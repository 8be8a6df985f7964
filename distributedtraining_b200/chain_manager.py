"""Address book: which exchange endpoint ("repo") belongs to which hotkey.

Reference: hivetrain/chain_manager.py -- ``run_in_subprocess`` hang guard (:14-54), ``ChainMultiAddressStore`` writing a
string commitment to the Bittensor chain (:57-115), ``LocalAddressStore`` JSON fake (:124-168).  Here the "chain" is
the ledger of :mod:`.btt_connector` (TCPStore / JSON file / memory); the API is unchanged.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
from typing import Any, Callable, Optional

from .utils.logging import logger


def _wrapped_func(func: Callable, queue, args, kwargs) -> None:
    try:
        queue.put(("ok", func(*args, **kwargs)))
    except BaseException as e:  # marshal the exception to the parent
        try:
            queue.put(("err", e))
        except Exception:
            queue.put(("err", RuntimeError(repr(e))))


def run_in_subprocess(func: Callable, ttl: float, *args, **kwargs) -> Any:
    """Run ``func`` in a forked child with a time-to-live; terminate and raise ``TimeoutError`` if it hangs."""
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_wrapped_func, args=(func, q, args, kwargs))
    p.start()
    p.join(timeout=ttl)
    if p.is_alive():
        p.terminate()
        p.join()
        raise TimeoutError(f"Failed to {getattr(func, '__name__', 'run')} after {ttl} seconds")
    try:
        status, payload = q.get(timeout=1.0)
    except Exception:
        raise RuntimeError(f"{getattr(func, '__name__', 'func')} died without a result (exit code {p.exitcode})")
    if status == "err":
        raise payload
    return payload


class ChainMultiAddressStore:
    """``store_hf_repo`` / ``retrieve_hf_repo`` over a ledger.  ``subtensor`` may be any object with put/get."""

    def __init__(self, subtensor, netuid: int = 1, wallet=None, ttl: float = 60.0, guard: bool = False):
        self.subtensor, self.netuid, self.wallet, self.ttl, self.guard = subtensor, netuid, wallet, ttl, guard

    def _key(self, hotkey: str) -> str:
        return f"commit/{self.netuid}/{hotkey}"

    def store_hf_repo(self, hf_repo: str) -> None:
        if self.wallet is None:
            raise ValueError("No wallet available to write to the chain.")
        hk = getattr(self.wallet, "hotkey_str", None) or self.wallet.hotkey.ss58_address
        if self.guard:
            run_in_subprocess(self.subtensor.put, self.ttl, self._key(hk), hf_repo)
        else:
            self.subtensor.put(self._key(hk), hf_repo)

    def retrieve_hf_repo(self, hotkey: str) -> Optional[str]:
        try:
            if self.guard:
                return run_in_subprocess(self.subtensor.get, self.ttl, self._key(hotkey))
            return self.subtensor.get(self._key(hotkey))
        except TimeoutError as e:
            logger.warning(str(e))
            return None


class LocalAddressStore:
    """``storage.json`` dict hotkey -> repo path (reference chain_manager.py:124-168)."""

    def __init__(self, subtensor=None, netuid: int = 1, wallet=None, path: str = "storage.json"):
        self.wallet, self.netuid, self.path = wallet, netuid, path
        if not os.path.exists(path):
            with open(path, "w") as f:
                json.dump({}, f)

    def _load(self):
        with open(self.path) as f:
            return json.load(f)

    def store_hf_repo(self, hf_repo: str) -> None:
        hk = getattr(self.wallet, "hotkey_str", None) or self.wallet.hotkey.ss58_address
        d = self._load()
        d[hk] = hf_repo
        tmp = f"{self.path}.tmp.{os.getpid()}"
        with open(tmp, "w") as f:
            json.dump(d, f)
        os.replace(tmp, self.path)

    def retrieve_hf_repo(self, hotkey: str) -> Optional[str]:
        return self._load().get(hotkey)
