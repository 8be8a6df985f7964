#!/usr/bin/env python
"""Miner neuron (reference neurons/miner.py): train a private copy, publish ``delta = theta - theta_base`` every round,
adopt every new averaged base.

    python neurons/miner.py --device cuda --model gpt2 --batch_size 256 --local_steps 100 --backend peer
    torchrun --nproc-per-node 8 neurons/miner.py ...          # one miner per GPU
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedtraining_b200.data import SyntheticTokens, build_text_loader  # noqa: E402
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.runtime import build_context  # noqa: E402
from distributedtraining_b200.training_manager import DeltaLoop  # noqa: E402
from distributedtraining_b200.utils.checkpoint import maybe_resume, save_checkpoint  # noqa: E402


def main(argv=None):
    ctx = build_context("miner", argv)
    cfg = ctx.config
    batch_size = cfg.batch_size
    trainer = Trainer(cfg.model, device=ctx.device, batch=batch_size, seq=cfg.seq_len, lr=cfg.lr, seed=0,
                      dropout=getattr(cfg, "dropout", None), dropout_seed=ctx.rank)
    resumed_round = maybe_resume(cfg, trainer, ctx.rank)
    # reference: WikiText-103 train split @ max_length 64, no shuffle (neurons/miner.py:54-106); offline: synthetic tokens
    if cfg.data.train_file:  # real text: per-item tokenisation to max_length ids, right-padded, no shuffle (reference :69-106)
        data_loader = build_text_loader(cfg.data.train_file, cfg.data.tokenizer, trainer.cfg.vocab_size, batch_size, cfg.seq_len,
                                        drop_last=True, repeat=True)
    else:
        data_loader = SyntheticTokens(batch_size, cfg.seq_len, trainer.cfg.vocab_size, pad_id=trainer.cfg.vocab_size - 1,
                                      seed=ctx.rank)
    max_steps = cfg.rounds * cfg.local_steps if cfg.rounds else None
    loop = DeltaLoop(ctx.device, cfg.model, data_loader, send_interval=cfg.miner.send_interval, learning_rate=cfg.lr,
                     hf_manager=ctx.hf_manager, trainer=trainer, local_steps=None if cfg.wall_clock else cfg.local_steps,
                     post_pull_lr=cfg.post_pull_lr, reset_optimizer=not cfg.no_reset_optimizer, max_steps=max_steps,
                     metrics=ctx.metrics, my_hotkey=ctx.hotkey)
    if resumed_round:  # continue the round numbering: a fresh_only averager ignores rounds it has already consumed
        loop.rounds_sent = resumed_round
        loop.global_step = resumed_round * (cfg.local_steps or 0)
        if ctx.hf_manager is not None:
            ctx.hf_manager.round = resumed_round
        if max_steps is not None:
            loop.max_steps = max_steps + loop.global_step
    if cfg.save_every:  # periodic: every ``save_every`` rounds (the reference parses the flag and never reads it)
        loop.checkpoint_hook = lambda lp: (lp.rounds_sent % cfg.save_every == 0) and save_checkpoint(
            cfg, trainer, ctx.rank, lp.rounds_sent, extra={"global_step": lp.global_step})
    loop.train(epochs=int(3e16) if max_steps is None else 1)
    if cfg.save_every:
        save_checkpoint(cfg, trainer, ctx.rank, loop.rounds_sent, extra={"global_step": loop.global_step})
    return loop


if __name__ == "__main__":
    main()
