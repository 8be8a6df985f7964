#!/usr/bin/env python
"""Co-located job (BASELINE.json config 2): every rank of the box MINES, and all ranks together are the AVERAGER.

    torchrun --nproc-per-node 8 neurons/colocated.py --model gpt2 --batch_size 512 --local_steps 100 --rounds 10 \\
        --meta_epochs 7 --save_every 1 [--resume] [--data.train_file train.txt --data.val_file val.txt]

One round = ``local_steps`` optimizer steps on every rank (reference miner loop, hivetrain/training_manager.py:345-433), then
the exchange on the device: delta emit into the symmetric windows -> the learned mixer run by ALL ranks (reference
hivetrain/averaging_logic.py:490-541 + neurons/averager.py:106; parallel/meta.py) -> sharded weighted average -> pull
all-gather fused with the optimizer reset (parallel/local_sgd.py).  Durable state (arenas, optimizer, ``w``, counters) is
checkpointed every ``--save_every`` rounds; ``--resume`` continues from the newest checkpoint.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedtraining_b200 import ops  # noqa: E402
from distributedtraining_b200.data import SyntheticTokens, build_text_loader  # noqa: E402
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.parallel.exchange import CollectiveExchange, PeerExchange  # noqa: E402
from distributedtraining_b200.parallel.local_sgd import LocalSGDCoordinator  # noqa: E402
from distributedtraining_b200.runtime import build_context  # noqa: E402
from distributedtraining_b200.training_manager import DeltaLoop  # noqa: E402
from distributedtraining_b200.utils.checkpoint import PeriodicCheckpointer, maybe_resume  # noqa: E402
from distributedtraining_b200.utils.logging import logger  # noqa: E402

EVAL_SEQ, EVAL_TEXTS = 512, 100  # reference neurons/averager.py:61,72


def main(argv=None):
    ctx = build_context("miner", argv)
    cfg = ctx.config
    B = cfg.batch_size
    # the trainer's theta_base / bf16 copy live in the symmetric window (peers land the new base there directly)
    buffers = ctx.exchange.trainer_buffers() if isinstance(ctx.exchange, PeerExchange) else None
    trainer = Trainer(cfg.model, device=ctx.device, batch=B, seq=cfg.seq_len, lr=cfg.lr, seed=0, dropout_seed=ctx.rank,
                      dropout=getattr(cfg, "dropout", None), buffers=buffers)
    V = trainer.cfg.vocab_size
    done = maybe_resume(cfg, trainer, ctx.rank, role="colocated")
    # ---- exchange plane: the peer windows built by the runtime, or the collective (NCCL / gloo) baseline plane ----
    ex = ctx.exchange if isinstance(ctx.exchange, PeerExchange) else (CollectiveExchange(trainer.man) if ctx.world > 1 else None)
    if isinstance(ex, PeerExchange):
        ops.set_flag_timeout(cfg.flag_timeout)
    # ---- data ----
    if cfg.data.train_file:
        train = build_text_loader(cfg.data.train_file, cfg.data.tokenizer, V, B, cfg.seq_len, drop_last=True, repeat=True)
    else:
        train = SyntheticTokens(B, cfg.seq_len, V, pad_id=V - 1, seed=1000 + ctx.rank)
    vseq = min(EVAL_SEQ, cfg.seq_len * 8, trainer.cfg.n_positions) if cfg.model.endswith("tiny") else min(EVAL_SEQ, trainer.cfg.n_positions)
    vb = max(1, min(getattr(cfg, "val_batch", 8) or 8, EVAL_TEXTS))
    if cfg.data.val_file:
        val = list(build_text_loader(cfg.data.val_file, cfg.data.tokenizer, V, vb, vseq, limit=EVAL_TEXTS))
    else:
        n_val = EVAL_TEXTS if not cfg.model.endswith("tiny") else 2 * vb
        vs = SyntheticTokens(n_val, vseq, V, pad_id=V - 1, seed=7, pool=1).pool[0]
        val = [{k: v[i:i + vb] for k, v in vs.items()} for i in range(0, n_val, vb)]
    validator = None
    if cfg.validate_every > 0 and isinstance(ex, PeerExchange):
        from distributedtraining_b200.validation_logic import CollectiveDeltaValidator
        validator = CollectiveDeltaValidator(ctx.device, trainer, val, ctx.network, ex, list(range(ctx.world)),
                                             validator_rank=(ctx.roles.get("validator") or [0])[0], eval_rows=cfg.eval_rows or None,
                                             metrics=ctx.metrics)
    learned = cfg.mixer == "learned" and cfg.meta_epochs > 0
    coord = LocalSGDCoordinator(trainer, ex, mixer="learned" if learned else "uniform", meta_epochs=cfg.meta_epochs if learned else 0,
                                meta_lr=cfg.meta_lr, val_batches=val, post_pull_lr=cfg.post_pull_lr,
                                reset_optimizer=not cfg.no_reset_optimizer, meta_dropout=bool(cfg.meta_dropout),
                                meta_log=(lambda k, loss, wm: ctx.metrics.log(meta_pass=k, loss_averaged=loss, w_mean=wm))
                                if ctx.rank == 0 else None, validator=validator, validate_every=cfg.validate_every)
    if getattr(trainer, "_resume_blob", None):
        coord.load_state_dict(trainer._resume_blob["extra"].get("coordinator", {}))
    ckpt = PeriodicCheckpointer(cfg, trainer, ctx.rank, "colocated")

    class _State:  # what a checkpoint holds besides the trainer arenas
        def state_dict(self_inner):
            return {"coordinator": coord.state_dict(), "global_step": loop.global_step, "rounds_sent": loop.rounds_sent}

    def round_hook(lp):
        coord.finish_round(lp)
        if ckpt.every > 0 and coord.rounds_total % ckpt.every == 0:
            coord.sync_base()  # a pushed base must have landed completely before the arenas are read for the checkpoint
            ckpt(_State(), coord.rounds_total)

    max_steps = cfg.rounds * cfg.local_steps if cfg.rounds else None
    loop = DeltaLoop(ctx.device, cfg.model, train, learning_rate=cfg.lr, hf_manager=None, trainer=trainer,
                     local_steps=cfg.local_steps, round_hook=round_hook, max_steps=max_steps, metrics=ctx.metrics,
                     post_pull_lr=cfg.post_pull_lr, my_hotkey=ctx.hotkey)
    if done:
        loop.rounds_sent = done
        loop.global_step = done * cfg.local_steps
        if max_steps is not None:
            loop.max_steps = loop.global_step + max_steps
        logger.info(f"rank {ctx.rank}: continuing after round {done}")
    loop.train(epochs=int(3e16) if max_steps is None else 1)
    coord.sync_base()
    ckpt(_State(), coord.rounds_total, force=True)
    if isinstance(ex, PeerExchange):
        ex.win.check_errors()
    return loop, coord


if __name__ == "__main__":
    main()
