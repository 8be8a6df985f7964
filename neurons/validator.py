#!/usr/bin/env python
"""Validator neuron (reference neurons/validator.py): score every miner's delta by the perplexity drop on held-out data,
normalise, EMA, commit the weights."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedtraining_b200.data import SyntheticTokens, build_text_loader  # noqa: E402
from distributedtraining_b200.utils.checkpoint import PeriodicCheckpointer, maybe_resume  # noqa: E402
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
    if cfg.data.val_file:  # reference: first 100 test texts @512, batch 8 (neurons/validator.py:49,63,98)
        test_loader = list(build_text_loader(cfg.data.val_file, cfg.data.tokenizer, trainer.cfg.vocab_size, EVAL_BATCH, seq,
                                             limit=EVAL_TEXTS))
    else:
        test_loader = list(SyntheticTokens(EVAL_BATCH, seq, trainer.cfg.vocab_size, pad_id=trainer.cfg.vocab_size - 1, seed=4242,
                                           steps=n_batches, pool=n_batches))
    maybe_resume(cfg, trainer, ctx.rank, role="validator")
    validator = DeltaValidator(ctx.device, trainer, None, test_loader, ctx.network, ctx.hf_manager,
                               interval=0 if cfg.rounds else INTERVAL, chain_manager=ctx.chain, metrics=ctx.metrics,
                               max_rounds=cfg.rounds or None, check_update_interval=0 if cfg.rounds else 300)
    if getattr(trainer, "_resume_blob", None):  # scores + the score EMA survive a restart (the reference loses them)
        validator.load_state_dict(trainer._resume_blob["extra"])
    ckpt = PeriodicCheckpointer(cfg, trainer, ctx.rank, "validator")
    validator.checkpoint_hook = ckpt
    validator.start_periodic_validation()
    ckpt(validator, validator.rounds_done, force=True)
    return validator


if __name__ == "__main__":
    main()
