"""Subnet-template flag groups (reference hivetrain/config/base_subnet_config.py:25-183)."""
from __future__ import annotations

import argparse
import os


def check_config(config) -> None:
    """Create the logging directory for this identity (the reference's version refers to an undefined ``logger``)."""
    path = os.path.expanduser(os.path.join(config.logging.logging_dir, config.wallet.name, config.wallet.hotkey,
                                           f"netuid{config.netuid}", config.neuron.name))
    config.neuron.full_path = path
    os.makedirs(path, exist_ok=True)


def add_neuron_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--netuid", type=int, default=1)
    parser.add_argument("--neuron.device", type=str, default="cuda")
    parser.add_argument("--neuron.epoch_length", type=int, default=100, help="blocks between set_weights calls")
    parser.add_argument("--mock", action="store_true", help="use the in-memory ledger (no files)")
    parser.add_argument("--neuron.events_retention_size", type=str, default="2 GB")
    parser.add_argument("--neuron.dont_save_events", action="store_true")
    parser.add_argument("--neuron.initial_peers", type=str, nargs="*", default=[])
    parser.add_argument("--neuron.name", type=str, default="neuron")


def add_miner_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--blacklist.force_validator_permit", action="store_true")
    parser.add_argument("--blacklist.allow_non_registered", action="store_true")


def add_validator_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--neuron.timeout", type=float, default=10.0)
    parser.add_argument("--neuron.num_concurrent_forwards", type=int, default=1)
    parser.add_argument("--neuron.sample_size", type=int, default=50)
    parser.add_argument("--neuron.disable_set_weights", action="store_true")
    parser.add_argument("--neuron.moving_average_alpha", type=float, default=0.333333)
    parser.add_argument("--neuron.axon_off", action="store_true")
    parser.add_argument("--neuron.vpermit_tao_limit", type=int, default=1024)
