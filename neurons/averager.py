#!/usr/bin/env python
"""Averager neuron (reference neurons/averager.py): gather every miner's delta, learn the per-(miner, tensor) mixing
weights on validation data, publish the weighted average as the next base model."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedtraining_b200.averaging_logic import GeneticAverager, ParameterizedAverager  # noqa: E402
from distributedtraining_b200.data import SyntheticTokens, build_text_loader  # noqa: E402
from distributedtraining_b200.utils.checkpoint import PeriodicCheckpointer, maybe_resume  # noqa: E402
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.runtime import build_context  # noqa: E402

EVAL_SEQ = 512     # reference neurons/averager.py:72
EVAL_TEXTS = 100   # :61
PERIOD = 1200      # :106


def main(argv=None):
    ctx = build_context("averager", argv)
    cfg = ctx.config
    bs = cfg.batch_size
    seq = min(EVAL_SEQ, cfg.seq_len * 8) if cfg.model.endswith("tiny") else EVAL_SEQ
    trainer = Trainer(cfg.model, device=ctx.device, batch=bs, seq=seq, lr=cfg.lr, seed=0, use_graph=False,
                      meta_dropout=bool(getattr(cfg, "meta_dropout", False)))
    n_batches = (EVAL_TEXTS + bs - 1) // bs if not cfg.rounds else max(1, min(4, EVAL_TEXTS // bs))
    if cfg.data.val_file:  # reference: first 100 test texts @512, batch --batch_size (neurons/averager.py:58-94)
        val_loader = list(build_text_loader(cfg.data.val_file, cfg.data.tokenizer, trainer.cfg.vocab_size, bs, seq, limit=EVAL_TEXTS,
                                            max_batches=n_batches))
    else:
        val_loader = list(SyntheticTokens(bs, seq, trainer.cfg.vocab_size, pad_id=trainer.cfg.vocab_size - 1, seed=4242,
                                          steps=n_batches, pool=n_batches))
    maybe_resume(cfg, trainer, ctx.rank, role="averager")
    ckpt = PeriodicCheckpointer(cfg, trainer, ctx.rank, "averager")

    def _arm(avg):  # durable w / consumed-round table / published round; periodic --save_every
        if getattr(trainer, "_resume_blob", None):
            avg.load_state_dict(trainer._resume_blob["extra"])
        avg.checkpoint_hook = ckpt
        return avg
    kw = dict(hf_manager=ctx.hf_manager, local_dir=cfg.storage.model_dir, gradients_dir=cfg.storage.gradient_dir,
              chain_manager=ctx.chain, bittensor_network=ctx.network, metrics=ctx.metrics, fresh_only=bool(cfg.rounds))
    period = 0 if cfg.rounds else PERIOD
    if cfg.mixer == "genetic":
        avg = _arm(GeneticAverager(trainer, ctx.device, **kw))
        for _ in range(cfg.rounds or 1 << 62):
            if avg.cache_params_locally():
                avg.run_evolution(val_loader)
                avg.save_model(); avg._adopt_as_base(); avg.push_to_hf_hub()
    elif cfg.mixer in ("uniform", "score"):
        avg = _arm(ParameterizedAverager(trainer, ctx.device, **kw))
        avg.run_periodic_averaging(val_loader, 0, cfg.meta_lr, period, max_rounds=cfg.rounds or None)
    else:
        avg = _arm(ParameterizedAverager(trainer, ctx.device, **kw))
        # reference: run_periodic_averaging(test_loader, 7, 0.01, 1200)  (neurons/averager.py:106)
        avg.run_periodic_averaging(val_loader, cfg.meta_epochs, cfg.meta_lr, period, max_rounds=cfg.rounds or None)
    ckpt(avg, avg.round, force=True)
    return avg


if __name__ == "__main__":
    main()
