"""Identity / membership / score sink -- the ``BittensorNetwork`` role of the reference without a blockchain.

Reference: hivetrain/btt_connector.py:264-506 (``BittensorNetwork`` class-level singleton: ``initialize``, ``sync``,
``set_weights`` with EMA 0.333333 + uint16 normalisation, ``should_set_weights``, ``get_validator_uids``,
``detect_metric_anomaly``, ``rate_limiter``) and :514-671 (``LocalBittensorNetwork`` JSON-file fake: 100 simulated
hotkeys, uids 91-99 are validators).

Here "the chain" is a pluggable *ledger*: in-memory (``--mock``), a JSON file shared through the filesystem
(``LocalBittensorNetwork``), or the ``torch.distributed`` TCPStore of the in-box job (hotkeys == ranks).
"""
from __future__ import annotations

import json
import os
import threading
import time
from types import SimpleNamespace
from typing import Dict, Iterable, List, Optional, Sequence

import torch

from . import __spec_version__
from .utils.logging import logger

BLOCK_SECONDS = 12.0
U16_MAX = 65535


# ---------------------------------------------------------------------------------------------------------------------
# ledgers
# ---------------------------------------------------------------------------------------------------------------------
class MemoryLedger:
    def __init__(self):
        self._d: Dict[str, str] = {}
        self._lock = threading.Lock()

    def put(self, key: str, value: str) -> None:
        with self._lock:
            self._d[key] = value

    def get(self, key: str) -> Optional[str]:
        with self._lock:
            return self._d.get(key)

    def keys(self, prefix: str = "") -> List[str]:
        with self._lock:
            return [k for k in self._d if k.startswith(prefix)]


class JsonFileLedger:
    """Single JSON file shared by all local processes; atomic replace on write, lock-file serialised."""

    def __init__(self, path: str):
        self.path = path
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if not os.path.exists(path):
            self._write({})

    def _read(self) -> Dict[str, str]:
        for _ in range(50):
            try:
                with open(self.path) as f:
                    return json.load(f)
            except (json.JSONDecodeError, FileNotFoundError):
                time.sleep(0.01)
        return {}

    def _write(self, d: Dict[str, str]) -> None:
        tmp = f"{self.path}.tmp.{os.getpid()}.{threading.get_ident()}"
        with open(tmp, "w") as f:
            json.dump(d, f)
        os.replace(tmp, self.path)

    def _locked(self):
        lock = self.path + ".lock"
        t0 = time.time()
        while True:
            try:
                fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.close(fd)
                return lock
            except FileExistsError:
                if time.time() - t0 > 5.0:  # stale lock
                    try:
                        os.remove(lock)
                    except FileNotFoundError:
                        pass
                time.sleep(0.005)

    def put(self, key: str, value: str) -> None:
        lock = self._locked()
        try:
            d = self._read()
            d[key] = value
            self._write(d)
        finally:
            os.remove(lock)

    def get(self, key: str) -> Optional[str]:
        return self._read().get(key)

    def keys(self, prefix: str = "") -> List[str]:
        return [k for k in self._read() if k.startswith(prefix)]


class StoreLedger:
    """``torch.distributed`` TCPStore / FileStore backed ledger (the in-box control plane)."""

    def __init__(self, store):
        self.store = store

    def put(self, key: str, value: str) -> None:
        self.store.set(key, value)
        idx = set(filter(None, self._index()))
        if key not in idx:
            idx.add(key)
            self.store.set("__index__", "\n".join(sorted(idx)))

    def _index(self) -> List[str]:
        try:
            if hasattr(self.store, "check") and not self.store.check(["__index__"]):
                return []
            return self.store.get("__index__").decode().split("\n")
        except Exception:
            return []

    def get(self, key: str) -> Optional[str]:
        try:
            if hasattr(self.store, "check") and not self.store.check([key]):
                return None
            return self.store.get(key).decode()
        except Exception:
            return None

    def keys(self, prefix: str = "") -> List[str]:
        return [k for k in self._index() if k and k.startswith(prefix)]


# ---------------------------------------------------------------------------------------------------------------------
# metagraph / wallet
# ---------------------------------------------------------------------------------------------------------------------
class Metagraph:
    """hotkeys, stakes and the validator->miner weight matrix ``W`` read back from the ledger."""

    def __init__(self, hotkeys: Sequence[str], stakes: Optional[Sequence[float]] = None):
        self.hotkeys: List[str] = list(hotkeys)
        self.n = len(self.hotkeys)
        self.uids = list(range(self.n))
        self.S = torch.tensor(list(stakes) if stakes is not None else [0.0] * self.n, dtype=torch.float32)
        self.W = torch.zeros(self.n, self.n)
        self.last_update = [0] * self.n
        self.block = 0

    def uid_of(self, hotkey: str) -> int:
        return self.hotkeys.index(hotkey)


class Wallet:
    def __init__(self, name: str = "default", hotkey: str = "default"):
        self.name = name
        self.hotkey = SimpleNamespace(ss58_address=hotkey)
        self.hotkey_str = hotkey

    def __repr__(self):
        return f"Wallet({self.name}, {self.hotkey_str})"


class LocalHotkey(SimpleNamespace):
    """``wallet.hotkey`` stand-in with only ``ss58_address`` (reference btt_connector.py:514-520)."""

    def __init__(self, ss58_address: str = "simulated_hotkey_0"):
        super().__init__(ss58_address=ss58_address)


# names the reference's simulation layer uses (btt_connector.py:514-585); same objects here
LocalMetagraph = Metagraph
LocalWallet = Wallet


def current_block() -> int:
    return int(time.time() // BLOCK_SECONDS)


# ---------------------------------------------------------------------------------------------------------------------
# the network singleton
# ---------------------------------------------------------------------------------------------------------------------
class BittensorNetwork:
    _instance = None
    _lock = threading.Lock()
    _weights_lock = threading.Lock()
    _anomaly_lock = threading.Lock()
    _config_lock = threading.Lock()
    _rate_limit_lock = threading.Lock()
    metrics_data: Dict[str, Dict[str, float]] = {}
    model_checksums: Dict[str, str] = {}
    request_counts: Dict[str, List[float]] = {}
    blacklisted_addresses: Dict[str, float] = {}
    last_sync_time = 0.0
    sync_interval = 600.0
    moving_average_alpha = 0.333333

    wallet: Optional[Wallet] = None
    metagraph: Optional[Metagraph] = None
    ledger = None
    config = None
    uid: int = 0
    base_scores: Optional[torch.Tensor] = None
    block_fn = staticmethod(current_block)
    last_set_block = 0

    def __new__(cls):
        with cls._lock:
            if cls._instance is None:
                cls._instance = super().__new__(cls)
        return cls._instance

    # -- bootstrap -----------------------------------------------------------------------------------------------
    @classmethod
    def initialize(cls, config, ignore_regs: bool = False, ledger=None, hotkeys: Optional[Sequence[str]] = None,
                   stakes: Optional[Sequence[float]] = None) -> None:
        with cls._lock:
            cls.config = config
            hk = getattr(getattr(config, "wallet", None), "hotkey", "default")
            cls.wallet = Wallet(getattr(getattr(config, "wallet", None), "name", "default"), hk)
            cls.ledger = ledger if ledger is not None else MemoryLedger()
            known = list(hotkeys) if hotkeys is not None else cls._ledger_hotkeys()
            if hk not in known:
                if hotkeys is not None and not ignore_regs:
                    raise SystemExit(f"hotkey {hk!r} is not registered (reference btt_connector.py:297-303 exits)")
                known.append(hk)
            cls.ledger.put(f"hotkey/{hk}", "1")
            cls.metagraph = Metagraph(known, stakes)
            cls.uid = cls.metagraph.uid_of(hk)
            cls.base_scores = torch.zeros(cls.metagraph.n)
            alpha = getattr(getattr(config, "neuron", None), "moving_average_alpha", None)
            if alpha:
                cls.moving_average_alpha = float(alpha)
            cls.last_set_block = cls.block_fn()
            cls.last_sync_time = time.time()

    @classmethod
    def _ledger_hotkeys(cls) -> List[str]:
        return sorted(k.split("/", 1)[1] for k in cls.ledger.keys("hotkey/"))

    # -- membership ------------------------------------------------------------------------------------------------
    @classmethod
    def resync_metagraph(cls, lite: bool = True) -> None:
        known = cls._ledger_hotkeys()
        mg = cls.metagraph
        for hk in known:
            if hk not in mg.hotkeys:
                mg.hotkeys.append(hk)
        if len(mg.hotkeys) != mg.n:
            n_old, mg.n = mg.n, len(mg.hotkeys)
            mg.uids = list(range(mg.n))
            mg.S = torch.cat([mg.S, torch.zeros(mg.n - n_old)])
            W = torch.zeros(mg.n, mg.n)
            W[:n_old, :n_old] = mg.W
            mg.W = W
            mg.last_update += [0] * (mg.n - n_old)
            cls.base_scores = torch.cat([cls.base_scores, torch.zeros(mg.n - n_old)])
        if not lite:
            for uid, hk in enumerate(mg.hotkeys):
                row = cls.ledger.get(f"weights/{hk}")
                if row:
                    blob = json.loads(row)
                    for u, wv in zip(blob["uids"], blob["weights"]):
                        if u < mg.n:
                            mg.W[uid, u] = wv / U16_MAX
                    mg.last_update[uid] = blob.get("block", 0)
        mg.block = cls.block_fn()

    @classmethod
    def sync(cls, lite: bool = True) -> None:
        if time.time() - cls.last_sync_time > cls.sync_interval or not lite:
            try:
                cls.resync_metagraph(lite)
                cls.last_sync_time = time.time()
            except Exception as e:  # tolerated, as in the reference (btt_connector.py:500-504)
                logger.warning(f"Failed to resync metagraph: {e}")
        else:
            cls.metagraph.block = cls.block_fn()

    @classmethod
    def get_validator_uids(cls, vpermit_tao_limit: float = 1024) -> List[int]:
        return [uid for uid in cls.metagraph.uids if float(cls.metagraph.S[uid]) >= vpermit_tao_limit]

    # -- scores ------------------------------------------------------------------------------------------------------
    @classmethod
    def should_set_weights(cls) -> bool:
        with cls._lock:
            disabled = getattr(getattr(cls.config, "neuron", None), "disable_set_weights", False)
            epoch = getattr(getattr(cls.config, "neuron", None), "epoch_length", 100)
            return (not disabled) and (cls.block_fn() - cls.last_set_block) > epoch

    @classmethod
    def set_weights(cls, scores: Dict[str, float]) -> bool:
        """EMA the normalised scores into ``base_scores`` and commit the row (reference btt_connector.py:311-356)."""
        with cls._weights_lock:
            mg = cls.metagraph
            a = cls.moving_average_alpha
            for hk, sc in scores.items():
                if hk not in mg.hotkeys:
                    continue
                uid = mg.uid_of(hk)
                sc = float(sc)
                if sc != sc:  # NaN guard
                    sc = 0.0
                cls.base_scores[uid] = a * sc + (1 - a) * cls.base_scores[uid]
            raw = cls.base_scores.clamp_min(0)
            tot = float(raw.sum())
            norm = raw / tot if tot > 0 else torch.zeros_like(raw)
            mx = float(norm.max())
            u16 = [int(round(float(x) / mx * U16_MAX)) if mx > 0 else 0 for x in norm]
            uids = [u for u, wv in enumerate(u16) if wv > 0]
            blob = {"uids": uids, "weights": [u16[u] for u in uids], "version_key": __spec_version__,
                    "block": cls.block_fn()}
            try:
                cls.ledger.put(f"weights/{cls.wallet.hotkey_str}", json.dumps(blob))
                mg.W[cls.uid].zero_()
                for u in uids:
                    mg.W[cls.uid, u] = u16[u] / U16_MAX
                cls.last_set_block = cls.block_fn()
                logger.info(f"set_weights: {len(uids)} non-zero weights committed")
                return True
            except Exception as e:
                logger.warning(f"set_weights failed: {e}")
                return False

    # -- Byzantine screens (present-but-unused in the reference; wired into the validator here) ---------------------
    @classmethod
    def detect_metric_anomaly(cls, metric: str = "loss", OUTLIER_THRESHOLD: float = 2.0,
                              MEDIAN_ABSOLUTE_DEVIATION: bool = True) -> Dict[str, bool]:
        """hotkey -> is_outlier, by MAD (default) or sigma distance (reference btt_connector.py:388-426)."""
        with cls._anomaly_lock:
            keys = [k for k, v in cls.metrics_data.items() if metric in v]
            if not keys:
                return {}
            vals = torch.tensor([cls.metrics_data[k][metric] for k in keys], dtype=torch.float64)
            if MEDIAN_ABSOLUTE_DEVIATION:
                med = vals.median()
                mad = (vals - med).abs().median() * 1.4826
                dev = (vals - med).abs() / mad if mad > 0 else torch.zeros_like(vals)
            else:
                sd = vals.std(unbiased=False)
                dev = (vals - vals.mean()).abs() / sd if sd > 0 else torch.zeros_like(vals)
            return {k: bool(d > OUTLIER_THRESHOLD) for k, d in zip(keys, dev)}

    @classmethod
    def run_evaluation(cls) -> Dict[str, float]:
        """Scores of non-anomalous hotkeys = 1, anomalous = 0 (reference btt_connector.py:430-452)."""
        flags = cls.detect_metric_anomaly()
        scores = {hk: 0.0 if bad else 1.0 for hk, bad in flags.items()}
        if scores and cls.should_set_weights():
            cls.set_weights(scores)
        return scores

    @classmethod
    def rate_limiter(cls, public_address: str, n: int = 10, t: float = 60.0) -> bool:
        """True if the request is allowed; > n requests within t seconds blacklists the address for t seconds."""
        with cls._rate_limit_lock:
            now = time.time()
            until = cls.blacklisted_addresses.get(public_address)
            if until is not None:
                if now < until:
                    return False
                del cls.blacklisted_addresses[public_address]
            q = [x for x in cls.request_counts.get(public_address, []) if now - x < t]
            q.append(now)
            cls.request_counts[public_address] = q
            if len(q) > n:
                cls.blacklisted_addresses[public_address] = now + t
                return False
            return True


class LocalBittensorNetwork(BittensorNetwork):
    """JSON-file ledger + ``n`` simulated hotkeys; the top tenth (by uid) hold validator stake
    (reference btt_connector.py:587-594: 100 hotkeys, stake 10000 for uid > 90 else 10)."""

    @classmethod
    def initialize(cls, config, ignore_regs: bool = True, n: int = 100, path: str = "bittensor_network/metagraph.json",
                   **kw) -> None:
        hotkeys = [f"simulated_hotkey_{i}" for i in range(n)]
        cut = n - max(n // 10, 1)
        stakes = [10000.0 if i > cut else 10.0 for i in range(n)]
        BittensorNetwork.initialize.__func__(cls, config, ignore_regs=True, ledger=JsonFileLedger(path), hotkeys=hotkeys,
                                             stakes=stakes)
        for hk in hotkeys:
            cls.ledger.put(f"hotkey/{hk}", "1")


def rank_network(config, rank: int, world: int, store=None, validator_ranks: Iterable[int] = ()) -> type:
    """In-box network: hotkey == ``rank{r}``, ledger == the job's TCPStore (or memory for a single process)."""
    hotkeys = [f"rank{r}" for r in range(world)]
    vr = set(validator_ranks)
    stakes = [10000.0 if r in vr else 10.0 for r in range(world)]
    config.wallet.hotkey = f"rank{rank}"
    BittensorNetwork.initialize(config, ignore_regs=True, ledger=StoreLedger(store) if store is not None else MemoryLedger(),
                                hotkeys=hotkeys, stakes=stakes)
    return BittensorNetwork


# ---------------------------------------------------------------------------------------------------------------------
# module-level helpers kept for API parity (reference btt_connector.py:19-63, 99-260)
# ---------------------------------------------------------------------------------------------------------------------
def initialize_bittensor_objects(config, ignore_regs: bool = True, ledger=None):
    """(wallet, ledger, metagraph) -- the reference's version refers to undefined Mock* classes under ``--mock``; here
    ``--mock`` simply selects the in-memory ledger."""
    BittensorNetwork.initialize(config, ignore_regs=ignore_regs, ledger=ledger or MemoryLedger())
    return BittensorNetwork.wallet, BittensorNetwork.ledger, BittensorNetwork.metagraph


def resync_metagraph(lite: bool = True) -> None:
    BittensorNetwork.resync_metagraph(lite)


def sync(lite: bool = True) -> None:
    BittensorNetwork.sync(lite)


def serve_extrinsic(ledger, wallet, ip: str, port: int, netuid: int = 1, protocol: int = 4) -> bool:
    """Advertise this neuron's endpoint (the reference copies bittensor's axon-serve extrinsic; nothing calls it)."""
    try:
        ledger.put(f"axon/{netuid}/{wallet.hotkey_str}", json.dumps({"ip": ip, "port": port, "protocol": protocol,
                                                                      "block": current_block()}))
        return True
    except Exception as e:
        logger.warning(f"serve_extrinsic failed: {e}")
        return False


def serve_axon(netuid: int, host_address: str, external_address: str, host_port: int, external_port: int) -> bool:
    return serve_extrinsic(BittensorNetwork.ledger, BittensorNetwork.wallet, external_address or host_address, external_port or host_port,
                           netuid)
