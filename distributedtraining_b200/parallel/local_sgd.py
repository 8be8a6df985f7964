"""Co-located synchronous local-SGD rounds: every rank mines, the averaging is sharded over all ranks.

One round (BASELINE.json config 2: "8 miner-GPUs GPT-2-small bf16, 100 local steps/round, fused delta
all-gather -> weighted-avg") after the miners' ``local_steps`` optimizer steps, entirely stream-ordered on the device:

  1. every rank emits ``delta = theta - theta_base`` straight into its symmetric window and release-stores its round flag;
  2. (learned mixer) ALL ranks run the SGD steps on the mixing matrix ``w[N, P]`` together (parallel/meta.py): delta
     all-to-all by pull once, then per step sharded ``theta_bar`` rebuild -> bf16 all-gather by pull -> fwd/bwd on a
     validation batch -> sharded segmented multi-dot -> ``w -= lr G`` (identical on every rank, no broadcast);
  3. every rank reduces ITS shard of the arena, ``s_j*base + sum_i w_ij delta_i``, into its base window;
  4. every rank pulls the other shards, fused with the base / optimizer reset (optimizer re-created, lr -> post_pull_lr).

The same object degrades to a single rank (N = 1) and to the NCCL/gloo collective baseline.
Reference semantics: miners hivetrain/training_manager.py:345-433, averager hivetrain/averaging_logic.py:335-583.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..utils.tracing import PhaseTimer
from .exchange import CollectiveExchange, PeerExchange
from .meta import DistributedMetaLearner


class LocalSGDCoordinator:
    def __init__(self, trainer, exchange=None, miners: Optional[List[int]] = None, averager_rank: int = 0,
                 mixer: str = "learned", meta_steps: int = 0, meta_lr: float = 0.01, val_batches: Optional[list] = None,
                 post_pull_lr: Optional[float] = 5e-5, reset_optimizer: bool = True, meta_epochs: int = 0,
                 meta_mode: str = "auto", meta_dropout: bool = False, reset_w: bool = True, meta_log=None,
                 validator=None, validate_every: int = 0, fused_first_forward: bool = True):
        self.trainer = trainer
        self.ex = exchange
        self.rank = getattr(exchange, "rank", 0)
        self.world = getattr(exchange, "world", 1)
        self.miners = list(range(self.world)) if miners is None else miners
        self.averager_rank = averager_rank  # kept for API parity: the learned mixer now runs on ALL ranks (parallel/meta.py)
        self.mixer, self.meta_steps, self.meta_lr, self.meta_epochs = mixer, meta_steps, meta_lr, meta_epochs
        self.val_batches = val_batches or []
        self.post_pull_lr, self.reset_optimizer = post_pull_lr, reset_optimizer
        self.reset_w = reset_w  # reference: w is rebuilt (1/N) at every averaging round (averaging_logic.py:492)
        self.meta_log = meta_log
        # optional co-located validator (validation_logic.CollectiveDeltaValidator): scores the round's deltas against the
        # OLD base, i.e. after the publish and before the averaging replaces theta_base
        self.validator, self.validate_every = validator, int(validate_every)
        # path (b): after a pushed base the first forward GEMMs acquire the owners' base flags in-kernel (no wait kernel)
        self.fused_first_forward = False
        if fused_first_forward and isinstance(exchange, PeerExchange) and exchange.can_push(trainer) and hasattr(trainer, "engine") \
                and exchange.world > 1:
            nch = trainer.man.seg_table("cpu")[0].numel()
            per = (nch + exchange.world - 1) // exchange.world
            trainer.enable_fused_first_forward(exchange.win.flag_ptr(exchange.F_BASE), exchange.base_target, per, exchange.world)
            self.fused_first_forward = trainer.ready_capable
        self.timer = PhaseTimer(enabled=True)  # CUDA-event phase timers (read once, at the end of a bench)
        self.round = 0
        self.round_base = 0  # rounds completed before a resume
        self.meta_steps_done = 0
        N, P = len(self.miners), len(trainer.man)
        dev = trainer.master.device
        self._new_base = torch.empty_like(trainer.master) if not isinstance(exchange, PeerExchange) else None
        self._local_delta = torch.empty_like(trainer.master) if exchange is None else None
        self.meta = None
        if self.learning and getattr(exchange, "name", "") != "nvls":
            self.meta = DistributedMetaLearner(trainer, exchange, self.miners, self.val_batches, meta_lr=meta_lr, mode=meta_mode,
                                               meta_dropout=meta_dropout)
            self.w = self.meta.w
        else:
            self.w = torch.full((N, P), 1.0 / N, dtype=torch.float32, device=dev)  # softmax(ones) == 1/N

    # -- durable state --------------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        """``round_base + round`` rounds completed so far (the flag words of a restarted job begin at zero again, so the
        live counter restarts while the total keeps counting), the mixing matrix and the meta-step counter."""
        return {"rounds_total": self.round_base + self.round, "w": self.w.detach().clone(), "meta_steps_done": self.meta_steps_done}

    def load_state_dict(self, sd: dict) -> None:
        self.round_base = int(sd.get("rounds_total", 0))
        self.meta_steps_done = int(sd.get("meta_steps_done", 0))
        if sd.get("w") is not None and tuple(sd["w"].shape) == tuple(self.w.shape):
            self.w.copy_(sd["w"].to(self.w.device))

    @property
    def rounds_total(self) -> int:
        return self.round_base + self.round

    @property
    def learning(self) -> bool:
        return self.mixer == "learned" and (self.meta_steps > 0 or self.meta_epochs > 0) and bool(self.val_batches)

    def _meta_learn(self) -> None:
        if self.meta_epochs > 0:
            self.meta_steps_done += self.meta.run(self.meta_epochs, self.meta_log)
        else:
            self.meta_steps_done += self.meta.run_steps(self.meta_steps)

    # -- the round ------------------------------------------------------------------------------------------------------
    def finish_round(self, loop=None) -> None:
        t = self.trainer
        self.round += 1
        r = self.round
        if isinstance(self.ex, PeerExchange) or isinstance(getattr(self.ex, "_ex", None), PeerExchange):
            ex = self.ex
            with self.timer.phase("delta_emit"):
                ex.publish_delta(t, r)
            if self.validator is not None and self.validate_every > 0 and self.rounds_total % self.validate_every == 0:
                with self.timer.phase("validate"):
                    ex.win.wait(ex.F_DELTA, r, self.miners)  # every miner's delta of this round has landed
                    torch.cuda.current_stream().synchronize()  # the scorer reads the flag page on the host (once per round)
                    self.validator.validate_and_score(round=r)
            # push mode: the trainer's arenas are multicast-bound window regions and the optimizer is re-created anyway -> the
            # broadcast of the new base is the multimem.st of the averaging kernel; nothing is pulled, nothing is copied
            push = self.reset_optimizer and ex.can_push(t)
            self.last_round_mode = "push (NVLS multicast stores from the averaging kernel)" if push else "pull"
            if self.learning:
                # learned mixer on ALL ranks: delta all-to-all by pull once, then meta_steps sharded SGD steps on w
                with self.timer.phase("meta_prepare"):
                    self.meta.begin_round(r, reset_w=self.reset_w)
                with self.timer.phase("meta_learning"):
                    self._meta_learn()
                with self.timer.phase("gather_avg"):
                    if push:
                        self.meta.final_average_push(r)
                    else:
                        per = self.meta.final_average_shard(r)
            else:
                # uniform mixer: w = 1/N_active (a miner whose emit kernel flagged NaN/Inf is skipped by every rank), then
                # reduce-scatter by pull straight from the miners' windows
                with self.timer.phase("gather_avg"):
                    active = ex.prepare_round(r, self.miners, self.w, init_w=True)
                    if push:
                        d, s = ex._delta_ptrs(r, self.miners)
                        ex.push_average(t.base, d, s, self.w, r, {"fp32": 0, "bf16": 1, "fp8": 2}[ex.delta_dtype_name], active=active)
                    else:
                        per = ex.reduce_scatter_average(t.base, self.w, r, self.miners, active=active)
            # the Adam moments are NOT rewritten: opt.reset() puts the step counter at 0 and the first step kernel treats them as
            # zero and reads theta from theta_base (the "fresh" flag) -- the optimizer re-creation of the reference without stores
            with self.timer.phase("broadcast_reset"):
                if push:
                    if not self.fused_first_forward:
                        ex.wait_base()      # every owner's shard has landed in my theta_base / bf16 copy
                    # else: the first forward GEMMs of the next step acquire the owners' flags themselves (TransformerEngine.
                    # configure_ready): the broadcast overlaps the step instead of preceding it
                    t.master_stale = True   # theta == theta_base until the first step writes the master arena
                else:
                    # all-gather by pull fused with the base / master / bf16 reset (P2P stores are the slow direction)
                    ex.all_gather_reset(t, per, reset_moments=False)
            if self.reset_optimizer:
                t.opt.reset()
            if self.post_pull_lr is not None:
                t.opt.set_lr(self.post_pull_lr)
            ex.win.poll_errors()  # non-blocking: a flag wait that timed out raises here one round later
            return
        if getattr(self.ex, "name", "") == "nvls":
            # uniform mixer on the NVLS plane: in-switch sum of the deltas + multicast of the new base (1 kernel per rank)
            with self.timer.phase("delta_emit"):
                self.ex.publish_delta(t, r)
            with self.timer.phase("gather_avg"):
                new_base = self.ex.average_broadcast(t.base)
        elif isinstance(self.ex, CollectiveExchange):
            # baseline plane: NCCL/gloo all_gather of the deltas; the learned mixer runs the same sharded step sequence with
            # all_reduce in place of the peer kernels; otherwise every rank computes the full weighted sum
            with self.timer.phase("delta_emit"):
                g = self.ex.allgather_deltas(t)
            if self.learning:
                with self.timer.phase("meta_prepare"):
                    self.meta.begin_round(r, deltas=[g[i] for i in range(g.shape[0])], reset_w=self.reset_w)
                with self.timer.phase("meta_learning"):
                    self._meta_learn()
                with self.timer.phase("gather_avg"):
                    new_base = self.meta.final_average_full(self._new_base)
            else:
                with self.timer.phase("gather_avg"):
                    self.ex.allgather_average(t, self.w, self._new_base, gathered=g)
                new_base = self._new_base
        else:  # single process, no exchange object: N = 1
            t.emit_delta(self._local_delta)
            if self.learning:
                self.meta.begin_round(r, deltas=[self._local_delta], reset_w=self.reset_w)
                self._meta_learn()
                new_base = self.meta.final_average_full(self._new_base)
            else:
                ops.weighted_avg(t.base, [self._local_delta], self.w, t.man, [self._new_base])
                new_base = self._new_base
        with self.timer.phase("broadcast_reset"):
            t.load_base(new_base, lr=self.post_pull_lr, reset_optimizer=self.reset_optimizer)

    def sync_base(self) -> None:
        """Stream-ordered wait for a pushed base to have landed completely (needed before anything but ``Trainer.step`` reads
        the arenas right after a round in fused-first-forward mode: evaluation, checkpoints, checksums)."""
        if isinstance(self.ex, PeerExchange) and self.ex._base_round > 0 and self.ex.world > 1:
            self.ex.wait_base()

    def __call__(self, loop=None) -> None:
        self.finish_round(loop)
