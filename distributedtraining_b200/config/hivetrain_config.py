"""Local flag groups (reference hivetrain/config/hivetrain_config.py:6-56) + the B200-specific group."""
from __future__ import annotations

import argparse


def add_meta_miner_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--miner.batch-size", type=int, default=64)
    parser.add_argument("--miner.epochs", type=int, default=100)
    parser.add_argument("--miner.send_interval", type=int, default=800, help="seconds between delta pushes (wall-clock mode)")
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--storage.gradient_dir", type=str, default="./dtb200_store/gradients")
    parser.add_argument("--storage.model_dir", type=str, default="./dtb200_store/model")
    parser.add_argument("--storage.my_repo_id", type=str, default=None)
    parser.add_argument("--storage.averaged_model_repo_id", type=str, default="averaged_model")


def add_torch_miner_args(parser: argparse.ArgumentParser) -> None:
    # parsed-but-unused in the reference (hivetrain_config.py:21-32); here they drive the real rendezvous
    parser.add_argument("--rank", type=int, default=None, help="rank of this process (default: $RANK)")
    parser.add_argument("--world-size", type=int, default=None, help="number of ranks (default: $WORLD_SIZE)")
    parser.add_argument("--store-address", type=str, default="127.0.0.1")
    parser.add_argument("--store-port", type=int, default=4999)
    parser.add_argument("--initial_peers", type=str, nargs="*", default=[])
    parser.add_argument("--batch_size", type=int, default=1)
    parser.add_argument("--save_every", type=int, default=0, help="checkpoint every N rounds (0 = off)")


def add_orchestrator_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--port", type=int, default=5000)
    parser.add_argument("--host-address", type=str, default="127.0.0.1")


def add_b200_args(parser: argparse.ArgumentParser) -> None:
    g = parser.add_argument_group("b200")
    g.add_argument("--model", type=str, default="gpt2", help="gpt2 | gpt2-medium | llama-3.2-1b | gpt2-tiny | llama-tiny")
    g.add_argument("--seq_len", type=int, default=64)
    g.add_argument("--local_steps", type=int, default=100, help="optimizer steps per round (replaces send_interval=800 s)")
    g.add_argument("--lr", type=float, default=5e-4)
    g.add_argument("--dropout", type=float, default=None, help="train-mode embd/attn/resid dropout (default: the model preset; GPT-2 = 0.1)")
    g.add_argument("--post_pull_lr", type=float, default=5e-5)
    g.add_argument("--no_reset_optimizer", action="store_true")
    g.add_argument("--delta_dtype", type=str, default="fp32", choices=["fp32", "bf16", "fp8"])
    g.add_argument("--mixer", type=str, default="learned", choices=["learned", "uniform", "score", "genetic"])
    g.add_argument("--meta_epochs", type=int, default=7)
    g.add_argument("--meta_lr", type=float, default=0.01)
    g.add_argument("--val_batch", type=int, default=8, help="co-located job: validation batch size of the learned mixer")
    g.add_argument("--meta_dropout", action="store_true", help="keep dropout on in the averager's meta-gradient passes (reference behaviour; default: deterministic)")
    g.add_argument("--roles", type=str, default="", help="e.g. 'miner:0-6,validator:7,averager:0'")
    g.add_argument("--backend", type=str, default="peer", choices=["peer", "nccl", "gloo", "disk"])
    g.add_argument("--resume", action="store_true")
    g.add_argument("--checkpoint_dir", type=str, default="./dtb200_store/checkpoints")
    g.add_argument("--rounds", type=int, default=0, help="stop after N rounds (0 = run forever)")
    g.add_argument("--wall_clock", action="store_true", help="reference cadence: time-based send/poll intervals")
    g.add_argument("--inject", type=str, default="", help="fault injection 'nan|shape|stall|drop:rank[,..]'")
    g.add_argument("--metrics_jsonl", type=str, default="")
    # real-text data path (reference: WikiText-103 train / test[:100]); unset = synthetic tokens of the same shape
    g.add_argument("--data.train_file", type=str, default="", help="text file, one text per line (miner)")
    g.add_argument("--data.val_file", type=str, default="", help="text file, one text per line (validator / averager: first 100 lines)")
    g.add_argument("--data.tokenizer", type=str, default="byte", help="'byte' or an HF tokenizer directory")
    g.add_argument("--validate_every", type=int, default=0, help="co-located job: all ranks score the deltas every N rounds (0 = off)")
    g.add_argument("--eval_rows", type=int, default=52, help="rows per eval batch of the co-located validator")
    g.add_argument("--flag_timeout", type=float, default=13.0, help="seconds before a device-side peer flag wait gives up")
