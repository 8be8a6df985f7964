"""Role bootstrap shared by the three neurons: config -> rendezvous -> network -> exchange -> hf_manager -> address book.

Mirrors the top-level wiring of the reference entry scripts (reference neurons/miner.py:26-129, validator.py:26-115,
averager.py:39-106) with every networked dependency replaced by its in-box equivalent.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from .btt_connector import BittensorNetwork, JsonFileLedger, MemoryLedger, StoreLedger
from .chain_manager import ChainMultiAddressStore
from .config import Config, Configurator
from .hf_manager import HFManager
from .models.transformer import build_manifest, get_config
from .parallel.exchange import DiskExchange, PeerExchange
from .parallel.launch import init_distributed, parse_roles
from .utils.logging import MetricsLogger, logger


@dataclass
class Context:
    config: Config
    rank: int
    world: int
    device: torch.device
    roles: dict
    network: type
    chain: ChainMultiAddressStore
    hf_manager: HFManager
    exchange: object
    manifest: object
    metrics: MetricsLogger
    hotkey: str


def build_context(role: str, argv=None, config: Optional[Config] = None) -> Context:
    cfg = config or Configurator.combine_configs(argv)
    backend = cfg.backend
    if cfg.device == "cpu" and backend in ("peer", "nccl"):
        backend = "disk"
    rank, world, device = init_distributed("gloo" if cfg.device == "cpu" else "nccl", cfg)
    roles = parse_roles(cfg.roles, world)
    model_cfg = get_config(cfg.model)
    man = build_manifest(model_cfg)
    hotkey = f"rank{rank}"
    cfg.wallet.hotkey = hotkey
    # ---- ledger: the job's store when distributed, a JSON file for independent local processes, memory otherwise ----
    if dist.is_initialized():
        from torch.distributed.distributed_c10d import _get_default_store
        ledger = StoreLedger(_get_default_store())
    elif backend == "disk":
        ledger = JsonFileLedger(os.path.join(cfg.storage.model_dir, "ledger.json"))
    else:
        ledger = MemoryLedger()
    hotkeys = [f"rank{r}" for r in range(world)]
    stakes = [10000.0 if r in roles.get("validator", []) else 10.0 for r in range(world)]
    BittensorNetwork.initialize(cfg, ignore_regs=True, ledger=ledger, hotkeys=hotkeys, stakes=stakes)
    # ---- exchange plane ----
    if backend == "peer":
        exchange = PeerExchange(man, delta_dtype=cfg.delta_dtype)
        scheme = "peer"
    else:
        exchange = DiskExchange(cfg.storage.model_dir, rank, man, cfg.delta_dtype if cfg.delta_dtype != "fp8" else "bf16")
        scheme = "disk"
    if cfg.inject:  # fault injection on the named ranks (utils/fault_injection.py)
        from .utils.fault_injection import FaultyExchange, parse_inject
        exchange = FaultyExchange(exchange, rank, parse_inject(cfg.inject))
    my_repo = cfg.storage.my_repo_id or f"{scheme}://{rank}"
    hf = HFManager(local_dir=cfg.storage.gradient_dir, my_repo_id=my_repo if role == "miner" else None,
                   averaged_model_repo_id=cfg.storage.averaged_model_repo_id, model_dir=cfg.storage.model_dir,
                   device=str(device), exchange=exchange, manifest=man, model_config=model_cfg)
    chain = ChainMultiAddressStore(BittensorNetwork.ledger, cfg.netuid, BittensorNetwork.wallet)
    if role == "miner":  # register my endpoint once (the reference commits its HF repo id to the chain)
        if chain.retrieve_hf_repo(hotkey) != my_repo:
            chain.store_hf_repo(my_repo)
    metrics = MetricsLogger(cfg.metrics_jsonl or None, role, rank)
    logger.info(f"{role} rank {rank}/{world} on {device}, exchange={scheme}, model={cfg.model}")
    return Context(cfg, rank, world, device, roles, BittensorNetwork, chain, hf, exchange, man, metrics, hotkey)
