#!/usr/bin/env python
"""Validator neuron (reference neurons/validator.py): score every miner's delta by the perplexity drop on held-out data,
normalise, EMA, commit the weights."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedtraining_b200.data import SyntheticTokens  # noqa: E402
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.runtime import build_context  # noqa: E402
from distributedtraining_b200.validation_logic import DeltaValidator  # noqa: E402

EVAL_BATCH = 8      # reference neurons/validator.py:98
EVAL_SEQ = 512      # :63
EVAL_TEXTS = 100    # :49
INTERVAL = 1800     # :40


def main(argv=None):
    ctx = build_context("validator", argv)
    cfg = ctx.config
    seq = min(EVAL_SEQ, cfg.seq_len * 8) if cfg.model.endswith("tiny") else EVAL_SEQ
    trainer = Trainer(cfg.model, device=ctx.device, batch=EVAL_BATCH, seq=seq, lr=cfg.lr, seed=0, use_graph=False)
    n_batches = (EVAL_TEXTS + EVAL_BATCH - 1) // EVAL_BATCH
    test_loader = list(SyntheticTokens(EVAL_BATCH, seq, trainer.cfg.vocab_size, pad_id=trainer.cfg.vocab_size - 1, seed=4242,
                                       steps=n_batches, pool=n_batches))
    validator = DeltaValidator(ctx.device, trainer, None, test_loader, ctx.network, ctx.hf_manager,
                               interval=0 if cfg.rounds else INTERVAL, chain_manager=ctx.chain, metrics=ctx.metrics,
                               max_rounds=cfg.rounds or None, check_update_interval=0 if cfg.rounds else 300)
    validator.start_periodic_validation()
    return validator


if __name__ == "__main__":
    main()
