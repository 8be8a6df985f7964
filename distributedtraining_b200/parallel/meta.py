"""The learned mixer (the averager's meta-learning) executed by ALL ranks of the box.

Reference: ``ParameterizedAverager.meta_learning`` (hivetrain/averaging_logic.py:490-541) runs ``meta_epochs**2`` passes over
the validation loader on ONE machine, each batch = rebuild ``theta_bar(w)`` (2N disk loads + N x P axpys), fwd/bwd,
``G_ij = <g_j, theta_ij - theta_bar_j>`` (another N disk loads), ``w -= lr G`` -- ``neurons/averager.py:106`` calls it with
``(test_loader, 7, 0.01)``, i.e. 49 x ceil(100 / B) strictly sequential SGD steps per averaging round.

The sequence of steps is kept (it is the algorithm); each step is spread over the R ranks that hold the miners:

* once per round every rank pulls ITS shard of every miner's delta over NVLink (delta all-to-all, ``shard_transpose``), so all
  later passes over the N deltas touch local HBM only and each rank touches 1/R of them;
* per step: ``theta_bar`` shard (fused weighted sum, local) -> bf16 all-gather by pull (the only bulk NVLink traffic) ->
  fwd/bwd (``replicate``: the same batch on every rank, no gradient traffic at all; ``dp``: the validation rows are split
  over the ranks and the gradient reduce-scatter is FUSED into the dot kernel, which reads the R gradient arenas through
  peer pointers) -> segmented multi-dot on the local shard -> the [N, P] partials are stored into every rank's slot table
  -> ``w -= lr * sum_r partial_r`` in a fixed order on every rank (``w`` stays bit-identical everywhere, no broadcast).

Cost per step at N = R = 8, GPT-2-small, fp32 deltas: 0.56 GB of local HBM reads + 0.22 GB NVLink ingress per rank, against
4.5 GB of NVLink ingress twice for the one-rank formulation of round 1.

Two back-ends with the same step sequence: ``peer`` (symmetric windows + the kernels of csrc/meta_avg.cu, flags instead of
barriers, everything stream-ordered) and ``collective`` (torch.distributed all_gather / all_reduce, or a single process) --
the latter is the CPU/gloo test oracle and the NCCL baseline.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .. import ops
from ..models.transformer import TransformerEngine, kv_len_of


def _rows_of(total: int, world: int, rank: int):
    base, rem = divmod(total, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


class DistributedMetaLearner:
    def __init__(self, trainer, exchange, miners: Optional[Sequence[int]] = None, val_batches: Optional[List[Dict]] = None,
                 meta_lr: float = 0.01, mode: str = "auto", meta_dropout: bool = False, use_graph: Optional[bool] = None):
        from .exchange import PeerExchange
        self.t = trainer
        self.ex = exchange
        self.peer = isinstance(exchange, PeerExchange)
        self.rank = getattr(exchange, "rank", 0)
        self.world = getattr(exchange, "world", 1)
        self.miners = list(range(self.world)) if miners is None else list(miners)
        self.N, self.P = len(self.miners), len(trainer.man)
        self.meta_lr, self.meta_dropout = float(meta_lr), bool(meta_dropout)
        self.dev = trainer.master.device
        self.val_batches = [self._norm_batch(b) for b in (val_batches or [])]
        assert self.val_batches, "the learned mixer needs validation batches"
        self.Bv = max(int(b["input_ids"].shape[0]) for b in self.val_batches)
        self.Tv = int(self.val_batches[0]["input_ids"].shape[1])
        if mode == "auto":
            # data-parallel as soon as every rank gets at least one row: measured faster than replicating the batch at N = 2 and
            # N = 8 even with ONE 512-token row per rank (5.2 vs 6.7 ms per step, profiles/meta_check_n8_gpt2_b8.json), and 6 x
            # faster at 12 rows per rank (9.2 vs 55.9 ms, ..._b96.json); the fused reduce-scatter of g costs ~0.6 ms
            mode = "dp" if (self.world > 1 and self.Bv >= self.world) else "replicate"
        if self.world == 1:
            mode = "replicate"
        self.mode = mode
        self.r0, self.r1 = _rows_of(self.Bv, self.world, self.rank) if mode == "dp" else (0, self.Bv)
        rows_static = -(-self.Bv // self.world) if mode == "dp" else self.Bv
        man = trainer.man
        # ---- this rank's shard of the chunk table / element range ----
        cs, cl, _ = man.seg_table("cpu")
        self.nchunks = cs.numel()
        self.per = -(-self.nchunks // self.world)
        self.c0, self.c1 = min(self.nchunks, self.rank * self.per), min(self.nchunks, (self.rank + 1) * self.per)
        self.e0 = int(cs[self.c0]) if self.c0 < self.nchunks else man.total
        self.e1 = int(cs[self.c1 - 1]) + int(cl[self.c1 - 1]) if self.c1 > self.c0 else self.e0
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.w = torch.full((self.N, self.P), 1.0 / max(self.N, 1), **f32)
        self.dT = torch.zeros(self.N, max(self.e1 - self.e0, 8), **f32)  # local fp32 shards of all miners' deltas
        self.loss_acc = torch.zeros(2, **f32)
        self.tick = 0
        self.steps_done = 0
        # ---- validation engine: theta_bar lives in the trainer's (idle during the exchange) master / p16 arenas ----
        if self.peer and getattr(exchange, "with_meta", False):
            self.g = exchange.win.local("g", torch.float32)[:man.total]
            self.g.zero_()
        else:
            self.g = torch.zeros(man.total, **f32)
        self.engine = TransformerEngine(trainer.cfg, man, trainer.p16, self.g, rows_static, self.Tv, seed=1234 + self.rank)
        # push mode: the trainer's arenas are window regions bound to the NVSwitch multicast object (see PeerExchange.can_push)
        self.push = bool(self.peer and exchange.can_push(trainer))
        self.use_graph = (self.dev.type == "cuda") if use_graph is None else (use_graph and self.dev.type == "cuda")
        self._graphs = {}
        if self.peer:
            self._dT_ptrs = [self.dT[i].data_ptr() - 4 * self.e0 for i in range(self.N)]  # virtual bases: element e at ptr + 4e
            self._partial = torch.empty(max(self.c1 - self.c0, 1) * (self.N + 1), **f32)
        else:
            self._tid = man.tensor_ids(self.dev)
            self._active_host = None

    # ------------------------------------------------------------------------------------------------------------------
    def _norm_batch(self, b) -> Dict[str, torch.Tensor]:
        if not isinstance(b, dict):
            b = {"input_ids": b}
        ids = b["input_ids"]
        out = {"input_ids": ids.to(self.dev)}
        out["labels"] = b["labels"].to(self.dev) if b.get("labels") is not None else out["input_ids"]
        kv = b.get("kv_len")
        if kv is None and b.get("attention_mask") is not None:
            kv = kv_len_of(b["attention_mask"], ids.shape[-1])
        if kv is not None:
            out["kv_len"] = kv.to(self.dev)
        return out

    def describe(self) -> dict:
        return {"mode": self.mode, "val_batch": [int(self.Bv), int(self.Tv)], "val_batches": len(self.val_batches),
                "rows_per_rank": int(self.r1 - self.r0), "backend": "peer" if self.peer else "collective",
                "theta_bar_all_gather": "multimem.st push (NVLS)" if self.push else ("pull" if self.peer else "all_reduce"),
                "shard_elems": int(self.e1 - self.e0)}

    # -- round start -----------------------------------------------------------------------------------------------------
    def begin_round(self, round: int, deltas: Optional[Sequence[torch.Tensor]] = None, reset_w: bool = True) -> None:
        """Collect this rank's shard of every miner's delta and (reference: ``self.weights = None`` every round) reset
        ``w = 1/N_active``.  Peer plane: flag waits + NaN verdicts + the NVLink pull happen in two kernels, no host sync.
        Collective plane: ``deltas`` = the list of full flat deltas (already all-gathered by the caller)."""
        if self.peer:
            ex = self.ex
            ex.prepare_round(round, self.miners, self.w, init_w=reset_w)
            d, s = ex._delta_ptrs(round, self.miners)
            mode = {"fp32": 0, "bf16": 1, "fp8": 2}[ex.delta_dtype_name]
            ops.shard_transpose(d, s, [self.dT[i] for i in range(self.N)], ex.active, self.e0, self.e1, mode)
            self.active = ex.active
        else:
            assert deltas is not None and len(deltas) == self.N
            fin = torch.stack([torch.isfinite(dl.float()).all() for dl in deltas]).to(self.dev)
            self.active = fin.to(torch.int32)
            for i, dl in enumerate(deltas):
                sh = dl[self.e0:self.e1].to(self.dev, torch.float32)
                self.dT[i, :self.e1 - self.e0].copy_(torch.where(fin[i], sh, torch.zeros_like(sh)))
            if reset_w:
                na = fin.sum().clamp(min=1).float()
                self.w.copy_((fin.float() / na)[:, None].expand(self.N, self.P))
        self.loss_acc.zero_()

    # -- one SGD step on w ---------------------------------------------------------------------------------------------
    def _fwd_bwd(self) -> None:
        e = self.engine
        if not self.use_graph:
            e.forward_backward(True, dropout=self.meta_dropout)
            return
        # the CE normaliser (1 / valid tokens) is a kernel ARGUMENT, so a captured graph is only valid for one
        # (row count, denominator) pair: one graph per distinct pair (the last validation batch is usually smaller)
        key = (e.n_rows, e.loss_denominator)
        g = self._graphs.get(key)
        if g is None:  # eager warm-up on a side stream, then capture (same protocol as Trainer.step)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                e.forward_backward(True, dropout=self.meta_dropout)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e.forward_backward(True, dropout=self.meta_dropout)
            self._graphs[key] = g
        g.replay()

    def _set_batch(self, k: int) -> None:
        b = self.val_batches[k % len(self.val_batches)]
        bsz = b["input_ids"].shape[0]
        if self.mode == "dp":  # rows of THIS batch owned by this rank; gradients are normalised by the batch's global token count
            r0, r1 = _rows_of(bsz, self.world, self.rank)
            self.engine.loss_denominator = bsz * (self.Tv - 1)
        else:
            r0, r1 = 0, bsz
        kv = b.get("kv_len")
        self.engine.set_batch(b["input_ids"][r0:r1], b["labels"][r0:r1], None, kv[r0:r1] if kv is not None else None)

    def step(self, k: int) -> None:
        """One reference meta-step (averaging_logic.py:499-528) on validation batch ``k``."""
        t, man = self.t, self.t.man
        # the tick sequence belongs to the EXCHANGE (its flag words are monotonic): a second learner on the same windows continues it
        tick = self.tick = (getattr(self.ex, "_meta_tick", 0) + 1) if self.ex is not None else self.tick + 1
        if self.ex is not None:
            self.ex._meta_tick = tick
        if self.peer:
            ex, win = self.ex, self.ex.win
            from .symm import F_G, F_GP, F_TB
            R = self.world
            if self.push and R > 1:
                # 1.+2. theta_bar shard from the LOCAL delta shards: fp32 -> master (shard only, scratch), bf16 -> multimem.st into
                # EVERY rank's compute copy (the all-gather is the store of the rebuild kernel); then wait for the other shards
                ops.weighted_avg(t.base, self._dT_ptrs, self.w, man, [t.master], None, chunk_range=(self.c0, self.c1), mode=0,
                                 active=self.active, mc_bf16=win.mc("base16"))
                win.publish(F_TB, tick, multicast=True)  # the flag follows the multicast data through the switch (release)
                win.wait(F_TB, tick)
            else:
                # 1. theta_bar shard from the LOCAL delta shards: fp32 -> master (shard only), bf16 -> my window + my p16
                own16 = win.ptr("base16", self.rank)
                outs16 = ([own16] if t.p16.data_ptr() == own16 else [own16, t.p16]) if ex.with_base16 and t.is_cuda else None
                ops.weighted_avg(t.base, self._dT_ptrs, self.w, man, [t.master], outs16, chunk_range=(self.c0, self.c1), mode=0,
                                 active=self.active)
                if R > 1:
                    win.publish(F_TB, tick)
                    # 2. bf16 all-gather by pull
                    ops.shard_pull16([win.ptr("base16", r) for r in range(R)], man, self.per, self.rank, t.p16,
                                     wait_flags=[win.flag_ptr(F_TB + r) for r in range(R)], wait_value=tick,
                                     error_flag=win.error_flag)
            # 3. fwd/bwd on the (local rows of the) validation batch
            self._set_batch(k)
            self._fwd_bwd()
            # 4./5. meta-gradient partial of my shard -> every rank's slot table
            if self.mode == "dp":
                win.publish(F_G, tick)
                gs = [win.ptr("g", r) for r in range(R)]
                gflags = [win.flag_ptr(F_G + r) for r in range(R)]
                loss_scale = 1.0
            else:
                gs, gflags, loss_scale = [self.g], None, 1.0 / R
            sw = ex.slot_words * 4
            dsts = [win.ptr("gslot", r, self.rank * sw) for r in range(R)]
            ops.seg_dot(gs, self._dT_ptrs, t.base, t.master, man, dsts, N=self.N, wait_flags=gflags, wait_value=tick, mode=0,
                        chunk_range=(self.c0, self.c1), active=self.active, loss=self.engine.loss, loss_scale=loss_scale,
                        error_flag=win.error_flag, partial=self._partial)
            win.publish(F_GP, tick)
            # 6. w -= lr * sum_r partial_r  (identical arithmetic on every rank)
            ops.w_update([win.ptr("gslot", self.rank, r * sw) for r in range(R)], self.w, self.meta_lr,
                         wait_flags=[win.flag_ptr(F_GP + r) for r in range(R)], wait_value=tick, loss_acc=self.loss_acc,
                         error_flag=win.error_flag)
        else:
            self._step_collective(k)
        self.steps_done += 1

    # -- collective / single-process back-end (CPU oracle, NCCL baseline) ---------------------------------------------------
    def _allreduce(self, x: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(x, group=getattr(self.ex, "group", None))

    def _step_collective_cuda(self, k: int) -> None:
        """The FAIR NCCL formulation on a GPU: the same sharded step sequence and the same local kernels (fused weighted
        average, segmented multi-dot) -- only the communication differs: ncclAllReduce of the zero-padded bf16 theta_bar, of the
        validation gradient (dp mode) and of the [N, P] partials, instead of peer-memory loads / multicast stores + flags."""
        t, man = self.t, self.t.man
        dptr = [self.dT[i].data_ptr() - 4 * self.e0 for i in range(self.N)]
        if self.world > 1:
            t.p16.zero_()
        ops.weighted_avg(t.base, dptr, self.w, man, [t.master], [t.p16], chunk_range=(self.c0, self.c1), mode=0, active=self.active)
        self._allreduce(t.p16)  # exact: every element has exactly one non-zero contributor
        self._set_batch(k)
        self._fwd_bwd()
        if self.mode == "dp":
            self._allreduce(self.g)
        if getattr(self, "_table", None) is None:
            self._table = torch.empty(self.N * self.P + 1, dtype=torch.float32, device=self.dev)
            self._partial = torch.empty(max(self.c1 - self.c0, 1) * (self.N + 1), dtype=torch.float32, device=self.dev)
        ops.seg_dot([self.g], dptr, t.base, t.master, man, [self._table], N=self.N, mode=0, chunk_range=(self.c0, self.c1),
                    active=self.active, loss=self.engine.loss, loss_scale=1.0 if self.mode == "dp" else 1.0 / self.world,
                    partial=self._partial)
        self._allreduce(self._table)
        self.w.add_(self._table[:self.N * self.P].view(self.N, self.P), alpha=-self.meta_lr)
        self.loss_acc[0] += self._table[self.N * self.P]
        self.loss_acc[1] = self._table[self.N * self.P]

    def _step_collective(self, k: int) -> None:
        t, man = self.t, self.t.man
        if t.is_cuda and ops.have_kernels():
            return self._step_collective_cuda(k)
        e0, e1 = self.e0, self.e1
        # 1. theta_bar shard, 2. "all-gather" = all-reduce of the zero-padded arena (exact: every element has one owner)
        keep = self.active.bool()
        wa = torch.where(keep[:, None], self.w, torch.zeros_like(self.w))
        tid = self._tid[e0:e1]
        sh = t.base[e0:e1] * wa.sum(0)[tid]
        for i in range(self.N):
            sh = sh + self.dT[i, :e1 - e0] * wa[i][tid]
        if self.world > 1:
            t.master.zero_()
        t.master[e0:e1] = sh
        self._allreduce(t.master)
        if t.is_cuda:
            ops.cast_copy(t.master, t.p16)
        # 3. fwd/bwd
        self._set_batch(k)
        self._fwd_bwd()
        if self.mode == "dp":
            self._allreduce(self.g)  # local gradients are normalised by the GLOBAL token count
        # 4./5. partial meta-gradient of my shard, summed over ranks
        gsh = self.g[e0:e1]
        G = torch.zeros(self.N + 1, self.P, dtype=torch.float32, device=self.dev)
        G[self.N].index_add_(0, tid, gsh * (t.base[e0:e1] - sh))
        for i in range(self.N):
            G[i].index_add_(0, tid, gsh * self.dT[i, :e1 - e0])
        Gf = torch.where(keep[:, None], G[:self.N] + G[self.N][None], torch.zeros_like(G[:self.N]))
        loss = self.engine.loss.detach().clone().reshape(1)
        if self.mode != "dp":
            loss = loss / self.world
        self._allreduce(Gf)
        self._allreduce(loss)
        # 6.
        self.w.add_(Gf, alpha=-self.meta_lr)
        self.loss_acc[0] += loss[0]
        self.loss_acc[1] = loss[0]

    # -- the nested loop of the reference --------------------------------------------------------------------------------
    def run(self, meta_epochs: int, log=None) -> int:
        """``for epoch in range(meta_epochs): for epoch in range(meta_epochs): for batch in val_loader`` (reference :493-495).
        ``log(pass_index, avg_loss, w_mean)`` is called once per pass (one host read each) when given."""
        nb = len(self.val_batches)
        done = 0
        for outer in range(meta_epochs):
            for inner in range(meta_epochs):
                if log is not None:
                    self.loss_acc.zero_()
                for k in range(nb):
                    self.step(k)
                    done += 1
                if log is not None:
                    avg = float(self.loss_acc[0]) / max(nb, 1)
                    log(outer * meta_epochs + inner, avg, [float(x) for x in self.w.mean(dim=1)])
        return done

    def run_steps(self, n: int) -> int:
        for k in range(n):
            self.step(self.steps_done)
        return n

    # -- the round's result ----------------------------------------------------------------------------------------------
    def final_average_shard(self, round: int) -> int:
        """theta_bar(w_final) for my shard, straight into my ``base`` window, then the base flag (peer plane).  The
        all-gather half (``PeerExchange.all_gather_reset``) pulls the shards and resets the optimizer in the same pass."""
        assert self.peer
        ex, win, t = self.ex, self.ex.win, self.t
        ops.weighted_avg(t.base, self._dT_ptrs, self.w, t.man, [win.ptr("base", self.rank)], None, chunk_range=(self.c0, self.c1),
                         mode=0, active=self.active)
        ex._base_round = round + 1
        win.publish(ex.F_BASE, ex._base_round)
        return self.per

    def final_average_push(self, round: int) -> None:
        """Push mode: theta_bar(w_final) of my shard goes by ``multimem.st`` into EVERY rank's theta_base and bf16 compute copy
        (in place on my own shard) -- the averaged-base broadcast is the store of the averaging kernel."""
        assert self.peer and self.push
        self.ex.push_average(self.t.base, self._dT_ptrs, None, self.w, round, 0, active=self.active)

    def final_average_full(self, out: torch.Tensor) -> torch.Tensor:
        """Collective plane: the complete new base on every rank."""
        t = self.t
        e0, e1 = self.e0, self.e1
        keep = self.active.bool()
        wa = torch.where(keep[:, None], self.w, torch.zeros_like(self.w))
        tid = self._tid[e0:e1]
        sh = t.base[e0:e1] * wa.sum(0)[tid]
        for i in range(self.N):
            sh = sh + self.dT[i, :e1 - e0] * wa[i][tid]
        if self.world > 1:
            out.zero_()
        out[e0:e1] = sh
        self._allreduce(out)
        return out
