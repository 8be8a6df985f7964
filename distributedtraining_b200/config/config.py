"""Configuration: one argparse parser with dotted flags -> nested namespace.

Mirrors the reference's ``Configurator.combine_configs()`` (reference hivetrain/config/config.py:44-60), which merges
bittensor's wallet/subtensor/logging/axon groups with the local groups and returns a nested ``bt.config``.  Flag names
are kept where they exist in the reference (``--batch_size``, ``--device``, ``--storage.*``, ``--rank``,
``--world-size``, ``--store-address``, ``--store-port``, ``--save_every``, ``--netuid``, ``--neuron.*``,
``--blacklist.*``); the vestigial rendezvous flags are *actually used* here for the in-box launcher.
"""
from __future__ import annotations

import argparse
from typing import Any, Dict, Optional, Sequence

from .base_subnet_config import add_miner_args, add_neuron_args, add_validator_args
from .hivetrain_config import add_b200_args, add_meta_miner_args, add_orchestrator_args, add_torch_miner_args


class Config(dict):
    """dict with attribute access and dotted-key nesting (``cfg.storage.my_repo_id``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @classmethod
    def from_flat(cls, flat: Dict[str, Any]) -> "Config":
        root = cls()
        for key, val in flat.items():
            node = root
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node or not isinstance(node[p], Config):
                    node[p] = cls()
                node = node[p]
            node[parts[-1]] = val
        return root

    def flat(self, prefix: str = "") -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for k, v in self.items():
            if isinstance(v, Config):
                out.update(v.flat(prefix + k + "."))
            else:
                out[prefix + k] = v
        return out


def add_identity_args(parser: argparse.ArgumentParser) -> None:
    """Stand-ins for bittensor's wallet / subtensor / logging / axon flag groups (identity = a string hotkey)."""
    parser.add_argument("--wallet.name", type=str, default="default")
    parser.add_argument("--wallet.hotkey", type=str, default="default")
    parser.add_argument("--wallet.path", type=str, default="~/.dtb200/wallets")
    parser.add_argument("--subtensor.network", type=str, default="local")
    parser.add_argument("--subtensor.chain_endpoint", type=str, default="")
    parser.add_argument("--logging.debug", action="store_true")
    parser.add_argument("--logging.trace", action="store_true")
    parser.add_argument("--logging.logging_dir", type=str, default="~/.dtb200/logs")
    parser.add_argument("--axon.port", type=int, default=8091)
    parser.add_argument("--axon.ip", type=str, default="127.0.0.1")


class Configurator:
    @staticmethod
    def build_parser() -> argparse.ArgumentParser:
        parser = argparse.ArgumentParser(description="distributedtraining_b200 configuration", allow_abbrev=False)
        add_identity_args(parser)
        add_torch_miner_args(parser)
        add_meta_miner_args(parser)
        add_orchestrator_args(parser)
        add_neuron_args(parser)
        add_miner_args(parser)
        add_validator_args(parser)
        add_b200_args(parser)
        return parser

    @staticmethod
    def combine_configs(argv: Optional[Sequence[str]] = None, strict: bool = False) -> Config:
        parser = Configurator.build_parser()
        if strict:
            ns = parser.parse_args(argv)
        else:
            ns, _ = parser.parse_known_args(argv)
        return Config.from_flat(vars(ns))
