"""Co-located synchronous local-SGD rounds: every rank mines, the averaging is sharded over all ranks.

One round (BASELINE.json config 2: "8 miner-GPUs GPT-2-small bf16, 100 local steps/round, fused delta
all-gather -> weighted-avg") after the miners' ``local_steps`` optimizer steps, entirely stream-ordered on the device:

  1. every rank emits ``delta = theta - theta_base`` straight into its symmetric window and release-stores its round flag;
  2. (learned mixer) the averager rank runs ``meta_steps`` SGD steps on the mixing matrix ``w[N, P]``: fused pull-average of
     all peer deltas -> fwd/bwd on a validation batch -> segmented multi-dot over the peer deltas -> ``w -= lr G``; ``w`` is
     then pushed into every rank's window;
  3. every rank runs ONE fused kernel on its shard of the arena: P2P-load the N deltas, ``s_j*base + sum_i w_ij delta_i``,
     P2P-store the result into every rank's base window (reduce-scatter + all-gather in a single launch, no NCCL);
  4. flag barrier, then each miner adopts the new base (optimizer re-created, lr -> post_pull_lr).

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


class LocalSGDCoordinator:
    def __init__(self, trainer, exchange=None, miners: Optional[List[int]] = None, averager_rank: int = 0,
                 mixer: str = "learned", meta_steps: int = 0, meta_lr: float = 0.01, val_batches: Optional[list] = None,
                 post_pull_lr: Optional[float] = 5e-5, reset_optimizer: bool = True):
        self.trainer = trainer
        self.ex = exchange
        self.rank = getattr(exchange, "rank", 0)
        self.world = getattr(exchange, "world", 1)
        self.miners = list(range(self.world)) if miners is None else miners
        self.averager_rank = averager_rank
        self.mixer, self.meta_steps, self.meta_lr = mixer, meta_steps, meta_lr
        self.val_batches = val_batches or []
        self.post_pull_lr, self.reset_optimizer = post_pull_lr, reset_optimizer
        self.timer = PhaseTimer(enabled=True)  # CUDA-event phase timers (read once, at the end of a bench)
        self.pull_only = True  # measured: P2P loads reach ~780 GB/s, P2P stores ~210 GB/s -> both halves of the round pull
        self.round = 0
        N, P = len(self.miners), len(trainer.man)
        dev = trainer.master.device
        self.w = torch.full((N, P), 1.0 / N, dtype=torch.float32, device=dev)  # softmax(ones) == 1/N
        self._G = torch.empty(N, P, dtype=torch.float32, device=dev)
        self._new_base = torch.empty_like(trainer.master) if not isinstance(exchange, PeerExchange) else None
        self._local_delta = torch.empty_like(trainer.master) if exchange is None else None
        self._w_epoch = 0

    # -- learned mixer on the averager rank (peer plane) ------------------------------------------------------------
    def _meta_learn_peer(self, r: int) -> None:
        ex, t = self.ex, self.trainer
        for k in range(self.meta_steps):
            batch = self.val_batches[(self.round * self.meta_steps + k) % len(self.val_batches)]
            ex.gather_average(t.base, self.w, r, self.miners, t.master, t.p16 if t.is_cuda else None, wait=(k == 0))
            t.loss_and_grad(batch)
            d, s = ex._delta_ptrs(r, self.miners)
            ops.multi_dot(t.grad, d, t.base, t.master, t.man, self._G, dscales=s,
                          mode={"fp32": 0, "bf16": 1, "fp8": 2}[ex.delta_dtype_name])
            self.w.add_(self._G, alpha=-self.meta_lr)

    def _meta_learn_collective(self, g: torch.Tensor) -> None:
        """Same learned-mixer steps on the collective (NCCL/gloo) plane: ``g`` = this round's all-gathered deltas, w is
        broadcast afterwards."""
        t = self.trainer
        if self.rank == self.averager_rank:
            deltas = [g[i] for i in range(g.shape[0])]
            for k in range(self.meta_steps):
                batch = self.val_batches[(self.round * self.meta_steps + k) % len(self.val_batches)]
                ops.weighted_avg(t.base, deltas, self.w, t.man, [t.master], [t.p16] if t.is_cuda else None)
                t.loss_and_grad(batch)
                ops.multi_dot(t.grad, deltas, t.base, t.master, t.man, self._G)
                self.w.add_(self._G, alpha=-self.meta_lr)
        if dist.is_initialized():
            dist.broadcast(self.w, src=self.averager_rank)

    def _share_w_peer(self) -> None:
        """Averager -> all: the mixing matrix travels through the windows (a few KB), flag-synchronised."""
        from .symm import F_HEART
        ex = self.ex
        self._w_epoch += 1
        if ex.world > 1:
            if self.rank == self.averager_rank:
                for rk in range(ex.world):
                    ex.win.peer("w", rk, torch.float32)[:self.w.numel()].copy_(self.w.view(-1), non_blocking=True)
                ex.win.publish(F_HEART, self._w_epoch)
            ex.win.wait(F_HEART, self._w_epoch, [self.averager_rank])
            if self.rank != self.averager_rank:
                self.w.view(-1).copy_(ex.win.local("w", torch.float32)[:self.w.numel()])

    # -- the round ------------------------------------------------------------------------------------------------------
    def finish_round(self, loop=None) -> None:
        t = self.trainer
        self.round += 1
        r = self.round
        if isinstance(self.ex, PeerExchange):
            with self.timer.phase("delta_emit"):
                self.ex.publish_delta(t, r)
            if self.mixer == "learned" and self.meta_steps > 0 and self.val_batches:
                with self.timer.phase("meta_learning"):
                    if self.rank == self.averager_rank:
                        self._meta_learn_peer(r)
                    self._share_w_peer()
            if self.pull_only and getattr(t, "is_cuda", False) and hasattr(t, "engine"):
                # reduce-scatter by pull, then all-gather by pull fused with the base/optimizer reset (2 kernels, no pushes)
                with self.timer.phase("gather_avg"):
                    per = self.ex.reduce_scatter_average(t.base, self.w, r, self.miners)
                with self.timer.phase("broadcast_reset"):
                    self.ex.all_gather_reset(t, per, reset_moments=self.reset_optimizer)
                if self.reset_optimizer:
                    t.opt.reset()
                if self.post_pull_lr is not None:
                    t.opt.set_lr(self.post_pull_lr)
                return
            new_base = self.ex.sharded_average_broadcast(t.base, self.w, r, self.miners)
        elif getattr(self.ex, "name", "") == "nvls":
            # uniform mixer on the NVLS plane: in-switch sum of the deltas + multicast of the new base (1 kernel per rank)
            with self.timer.phase("delta_emit"):
                self.ex.publish_delta(t, r)
            with self.timer.phase("gather_avg"):
                new_base = self.ex.average_broadcast(t.base)
        elif isinstance(self.ex, CollectiveExchange):
            g = self.ex.allgather_deltas(t)  # once per round, BEFORE the mixer touches the averager's master copy
            if self.mixer == "learned" and self.meta_steps > 0 and self.val_batches:
                self._meta_learn_collective(g)
            # baseline plane: NCCL/gloo all_gather + torch weighted sum (every rank computes the full average)
            self.ex.allgather_average(t, self.w, self._new_base, gathered=g)
            new_base = self._new_base
        else:  # single process, no exchange object: N = 1
            t.emit_delta(self._local_delta)
            ops.weighted_avg(t.base, [self._local_delta], self.w, t.man, [self._new_base])
            new_base = self._new_base
        t.load_base(new_base, lr=self.post_pull_lr, reset_optimizer=self.reset_optimizer)

    def __call__(self, loop=None) -> None:
        self.finish_round(loop)
