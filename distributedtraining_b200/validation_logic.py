"""Validator role: score every miner's delta by the loss / perplexity drop it produces on held-out data.

Reference: hivetrain/validation_logic.py -- ``ModelValidator`` (:29-203), ``DeltaValidator`` (:251-259),
``LocalValidator`` / ``LocalDeltaValidator`` (:206-262), ``MNISTValidator`` / ``MNISTDeltaValidator`` (:265-318).
Algorithm spec: SURVEY.md section 2.6-B.

B200-first differences (scores are computed by the same formulas):
* ``theta_base`` is never mutated: ``theta_base + delta_i`` is materialised by ONE fused kernel straight into the
  engine's (fp32 master, bf16 compute) arenas -- reading the delta from the miner's peer window over NVLink -- so the
  reference's ``deepcopy(state_dict)`` / ``load_state_dict`` per miner (:123,139) and its sha256-on-CPU round trip
  (:133) disappear;
* the base loss is re-evaluated whenever the base changes (the reference computes it once in the constructor and never
  refreshes it, :48);
* ``sum(ppl_score) == 0`` no longer divides by zero (:186-187).
"""
from __future__ import annotations

import math
import time
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import ops
from .training_manager import calculate_model_hash
from .utils.logging import MetricsLogger, logger


def _batch_ids_labels(batch):
    if isinstance(batch, dict):
        return batch["input_ids"], batch.get("labels")
    return batch, None


class ModelValidator:
    def __init__(self, device, model, optimizer=None, data_loader: Optional[Iterable] = None, bittensor_network=None,
                 hf_manager=None, interval: float = 3600, chain_manager=None, check_update_interval: float = 300,
                 metrics: Optional[MetricsLogger] = None, compute_hash: bool = False, max_rounds: Optional[int] = None):
        self.device = device
        self.model = model  # Trainer / ModuleTrainer; model.base is theta_base
        self.optimizer = optimizer  # accepted for signature parity; the validator never steps it
        self.data_loader = data_loader
        self.bittensor_network = bittensor_network
        self.hf_manager = hf_manager
        self.chain_manager = chain_manager
        self.interval = interval
        self.check_update_interval = check_update_interval
        self.last_pull_time = 0.0
        self.metrics = metrics or MetricsLogger(None, "validator")
        self.compute_hash = compute_hash
        self.max_rounds = max_rounds
        self.base_loss, self.base_perplexity = self.evaluate_model()
        hotkeys = list(bittensor_network.metagraph.hotkeys) if bittensor_network is not None else []
        self.scores = {hk: 0.0 for hk in hotkeys}
        self.normalized_scores = {hk: 0.0 for hk in hotkeys}
        self.loss_scores = {hk: 0.0 for hk in hotkeys}
        self.losses: Dict[str, float] = {}
        self.rounds_done = 0
        self.checkpoint_hook = None  # callable(validator, round): periodic --save_every

    # -- durable state (the reference keeps scores and the EMA in memory only: btt_connector.py:305-307) --------------------
    def state_dict(self) -> Dict:
        net = self.bittensor_network
        ema = getattr(net, "base_scores", None)
        return {"scores": dict(self.scores), "normalized_scores": dict(self.normalized_scores), "loss_scores": dict(self.loss_scores),
                "losses": dict(self.losses), "base_loss": self.base_loss, "base_perplexity": self.base_perplexity,
                "rounds_done": self.rounds_done, "score_ema": ema.clone() if isinstance(ema, torch.Tensor) else None}

    def load_state_dict(self, sd: Dict) -> None:
        self.scores.update(sd.get("scores", {}))
        self.normalized_scores.update(sd.get("normalized_scores", {}))
        self.loss_scores.update(sd.get("loss_scores", {}))
        self.losses.update(sd.get("losses", {}))
        self.rounds_done = int(sd.get("rounds_done", 0))
        ema, net = sd.get("score_ema"), self.bittensor_network
        if ema is not None and net is not None and getattr(net, "base_scores", None) is not None \
                and net.base_scores.numel() == ema.numel():
            net.base_scores.copy_(ema)

    # -- delta application ------------------------------------------------------------------------------------------
    def update_model_weights(self, gradients, alpha: float = 5e-4) -> None:
        """Gradient-style update theta -= alpha * g (reference :72-76).  ``gradients``: flat tensor or name->tensor."""
        g = self._as_flat(gradients)
        self.model.master.add_(g.to(self.model.master.device, torch.float32), alpha=-alpha)
        self._sync_compute_copy()

    def _as_flat(self, t) -> torch.Tensor:
        if isinstance(t, dict):  # name->tensor dict in the engine's or in the reference's (HF) layout
            if hasattr(self.model, "flat_from"):
                return self.model.flat_from(t)
            return self.model.man.pack(t, torch.zeros(self.model.man.total, dtype=torch.float32))
        return t

    def _sync_compute_copy(self) -> None:
        m = self.model
        if getattr(m, "is_cuda", False) and m.p16.data_ptr() != m.master.data_ptr():
            ops.cast_copy(m.master, m.p16)

    def restore_base(self) -> None:
        m = self.model
        ops.round_reset(m.base, m.master, m.p16 if getattr(m, "is_cuda", False) else None, m.m, m.v, reset_moments=False)

    # -- evaluation -------------------------------------------------------------------------------------------------------
    def evaluate_model(self, metric: str = "loss") -> Tuple[float, float]:
        """Mean CE over the eval set (batch-size weighted mean of per-batch means, reference :78-97) and exp() of it."""
        total = torch.zeros((), dtype=torch.float64)
        n = 0
        acc = None
        for batch in self.data_loader or []:
            ids, labels = _batch_ids_labels(batch)
            bs = ids.shape[0] if hasattr(ids, "shape") else len(ids[0])
            # dict batches go through whole: the engine applies their attention_mask as the padding mask (reference :85-90)
            loss = self.model.eval_loss(batch) if isinstance(batch, dict) and hasattr(self.model, "engine") else \
                self.model.eval_loss(ids, labels)
            acc = loss.detach().double() * bs if acc is None else acc + loss.detach().double() * bs  # stays on device
            n += bs
        if acc is None:
            return float("nan"), float("nan")
        avg = float(acc) / max(n, 1)  # ONE host read per evaluation (the reference syncs every batch)
        return avg, math.exp(min(avg, 50.0))

    def calculate_model_hash(self) -> str:
        return calculate_model_hash(self.model)

    # -- the scoring round ------------------------------------------------------------------------------------------------
    def _maybe_pull(self) -> bool:
        if self.hf_manager is None or time.time() - self.last_pull_time < self.check_update_interval:
            return False
        self.last_pull_time = time.time()
        if self.hf_manager.check_for_new_submissions(self.hf_manager.model_repo_id):
            logger.info("Averaged model updated. Pulling latest model...")
            self.hf_manager.pull_latest_model()
            self.model = self.hf_manager.update_model(self.model, reset_optimizer=False)
            self.base_loss, self.base_perplexity = self.evaluate_model()  # refresh (reference never does)
            return True
        return False

    def _receive(self, hotkey: str):
        repo = self.chain_manager.retrieve_hf_repo(hotkey) if self.chain_manager is not None else hotkey
        if repo is None:
            return None
        return self.hf_manager.receive_flat(repo) if hasattr(self.hf_manager, "receive_flat") else \
            self.hf_manager.receive_gradients(repo)

    def validate_and_score(self) -> Dict[str, float]:
        net = self.bittensor_network
        net.sync(lite=True)
        self._maybe_pull()
        logger.info(f"Base loss {self.base_loss:.4f} ppl {self.base_perplexity:.3f}")
        for uid, hotkey in enumerate(net.metagraph.hotkeys):
            delta = self._receive(hotkey)
            if delta is not None and self._finite(delta):
                try:
                    self.update_model_weights(delta)
                    if self.compute_hash:
                        logger.info(f"Model hash is: {self.calculate_model_hash()}")
                    loss, perplexity = self.evaluate_model()
                    loss_score = max(0.0, self.base_loss - loss)
                    perplexity_score = max(0.0, self.base_perplexity - perplexity)
                finally:
                    self.restore_base()
            else:
                loss, perplexity, loss_score, perplexity_score = float("nan"), float("nan"), 0.0, 0.0
            self.losses[hotkey] = loss
            self.loss_scores[hotkey] = loss_score
            self.scores[hotkey] = perplexity_score
            net.metrics_data[hotkey] = {"loss": loss if loss == loss else 1e9}
            self.metrics.log(hotkey=hotkey, loss=loss, perplexity=perplexity, loss_score=loss_score,
                             perplexity_score=perplexity_score)
            logger.info(f"uid {uid} {hotkey}: loss {loss:.4f} ppl {perplexity:.3f} loss_score {loss_score:.4f} "
                        f"ppl_score {perplexity_score:.4f}")
        total = sum(self.scores.values())
        self.normalized_scores = {hk: (max(0.0, sc / total) if total > 0 else 0.0) for hk, sc in self.scores.items()}
        if net.should_set_weights():
            net.set_weights(self.normalized_scores)
        return self.normalized_scores

    @staticmethod
    def _finite(delta) -> bool:
        if isinstance(delta, dict):
            return all(bool(torch.isfinite(v).all()) for v in delta.values())
        return bool(torch.isfinite(delta.float()).all()) if delta.dtype != torch.uint8 else True

    def start_periodic_validation(self) -> None:
        rounds = 0
        while True:
            t0 = time.time()
            self.validate_and_score()
            if self.hf_manager is not None:
                self.hf_manager.clear_hf_cache()
            rounds += 1
            self.rounds_done += 1
            if self.checkpoint_hook is not None:
                self.checkpoint_hook(self, self.rounds_done)
            if self.max_rounds is not None and rounds >= self.max_rounds:
                return
            time.sleep(max(0.0, self.interval - (time.time() - t0)))


class DeltaValidator(ModelValidator):
    """theta <- delta + theta (reference :251-259).

    Fused form (``fused_eval=True``, delta stored in the compute dtype -- bf16 windows on a B200): ``theta_base + delta_i``
    is NEVER materialised for the weight matrices.  Every eval GEMM computes ``x (W + dW_i)^T`` as two accumulating
    tcgen05 passes with ``dW_i`` read straight from miner i's peer window over NVLink (csrc/sm100_gemm.cu, dual-B), the
    embedding adds the delta rows on the fly, and only the non-matrix tensors (< 1 % of the bytes) are updated by the
    fused apply kernel restricted to their chunks.  Otherwise: one fused kernel writes base+delta into master AND the
    bf16 copy (N = 1, w = 1)."""

    def __init__(self, *a, fused_eval: bool = False, **kw):
        # fused_eval=False (default): ONE fused apply kernel (peer read) materialises base+delta_i, then plain eval GEMMs --
        #   the faster choice when a miner is scored on many tokens (the reference evaluates 51 200 tokens per miner):
        #   the dual-B form spends two tensor-core passes per GEMM and is only ahead when the eval is weight-bandwidth
        #   bound (few tokens per miner).  Measured: profiles/validator_bench_*.json.
        self.fused_eval = fused_eval
        self._fused_delta = None
        self._dcache = None
        super().__init__(*a, **kw)

    def _ones(self):
        m = self.model
        return torch.ones(1, len(m.man), dtype=torch.float32, device=m.master.device)

    def update_model_weights(self, weight_deltas, alpha: float = 5e-4) -> None:
        m = self.model
        d = self._as_flat(weight_deltas)
        eng = getattr(m, "engine", None)
        same_dev = isinstance(d, torch.Tensor) and d.device == m.master.device
        if self.fused_eval and eng is not None and same_dev and d.dtype == m.p16.dtype:
            if m.is_cuda:  # one NVLink read: cache the peer delta locally, later batches hit HBM/L2 instead of the link
                if self._dcache is None:
                    self._dcache = torch.empty_like(m.p16)
                self._dcache.copy_(d, non_blocking=True)
                d = self._dcache
            ops.weighted_avg(m.base, [d], self._ones(), m.man, [m.master], [m.p16] if m.is_cuda else None,
                             chunk_ids=eng.small_chunk_ids(), unit_base=True)
            eng.set_delta(d)
            self._fused_delta = d
        elif getattr(m, "is_cuda", False) and same_dev:
            ops.weighted_avg(m.base, [d], self._ones(), m.man, [m.master], [m.p16], unit_base=True)
        else:
            m.master.copy_(m.base + d.to(m.master.device, torch.float32))
            self._sync_compute_copy()

    def restore_base(self) -> None:
        m = self.model
        if self._fused_delta is not None:
            m.engine.set_delta(None)
            ops.weighted_avg(m.base, [self._fused_delta], torch.zeros_like(self._ones()), m.man, [m.master],
                             [m.p16] if m.is_cuda else None, chunk_ids=m.engine.small_chunk_ids(), unit_base=True)
            self._fused_delta = None
        else:
            super().restore_base()


class EvalModel:
    """Forward-only scoring model: one bf16 parameter arena + an eval-only engine (activation buffers shared by all layers)
    + one CUDA graph per batch shape.  ``load()`` materialises ``theta_base (+ delta_i)`` with ONE fused kernel that reads
    the delta straight from the miner's peer window; nothing else of the model state is touched (theta_base is never mutated:
    the reference's deepcopy / load_state_dict pair per miner, validation_logic.py:123,139, has no counterpart)."""

    def __init__(self, cfg, man, device, rows: int, seq: int, lm_chunk: int = 8192):
        from .models.transformer import TransformerEngine
        self.cfg, self.man, self.device = cfg, man, torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.p16 = torch.empty(man.total, dtype=torch.bfloat16 if self.is_cuda else torch.float32, device=self.device)
        self.engine = TransformerEngine(cfg, man, self.p16, None, rows, seq, lm_chunk=lm_chunk, eval_only=True)
        self._graphs = {}
        self._acc = torch.zeros((), dtype=torch.float64, device=self.device)
        self._ones = torch.ones(1, len(man), dtype=torch.float32, device=self.device)

    def load(self, base: torch.Tensor, delta=None, dscale=None, mode: int = 0) -> None:
        """p16 = compute-dtype(base + delta); ``delta``: tensor or raw peer address (typed by ``mode``: 0 fp32, 1 bf16, 2 fp8)."""
        if delta is None:
            if self.is_cuda:
                ops.cast_copy(base, self.p16)
            else:
                self.p16.copy_(base)
            return
        if self.is_cuda:
            ops.weighted_avg(base, [delta], self._ones, self.man, [None], [self.p16], dscales=[dscale] if dscale is not None else None,
                             mode=mode, unit_base=True)
        else:
            self.p16.copy_(base + delta.to(base.device, torch.float32))

    def _forward(self) -> None:
        e = self.engine
        if not self.is_cuda:
            e.forward_loss()
            return
        key = e.n_rows  # the CE normaliser is a kernel argument: one graph per row count
        g = self._graphs.get(key)
        if g is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                e.forward_loss()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e.forward_loss()
            self._graphs[key] = g
        g.replay()

    @torch.no_grad()
    def mean_loss(self, batches) -> torch.Tensor:
        """Row-weighted mean of the per-batch mean CE (reference evaluate_model, :78-97) as a DEVICE scalar (no host sync)."""
        self._acc.zero_()
        n = 0
        for b in batches:
            self.engine.set_batch(b)
            self._forward()
            self._acc += self.engine.loss.double() * self.engine.n_rows
            n += self.engine.n_rows
        return self._acc / max(n, 1)


def rebatch(batches, rows: int):
    """Concatenate dict batches into batches of ``rows`` sequences.  Every row carries the same number of targets (PAD is
    not masked in the labels), so the row-weighted mean of per-batch means is the same number for any batch size -- the
    validator may evaluate 100 texts in one or two big GEMM-friendly batches instead of 13 small ones."""
    keys = [k for k in batches[0] if isinstance(batches[0][k], torch.Tensor)]
    cat = {k: torch.cat([b[k] for b in batches], dim=0) for k in keys}
    n = cat["input_ids"].shape[0]
    return [{k: v[i:i + rows] for k, v in cat.items()} for i in range(0, n, rows)]


class CollectiveDeltaValidator(DeltaValidator):
    """Co-located validator: ALL ranks of the box score the miners' deltas together (the reference scores them serially on
    one machine, validation_logic.py:126-183).

    Jobs = the N miners + the base model; job k runs on rank ``k % world``: fused delta-apply straight from the miner's peer
    window (one NVLink read of the delta) -> graph-captured eval forward over the validation set in large batches -> one
    device scalar.  The NaN / missing-miner verdicts come from the publish flags (no pass over the delta, no host sync per
    miner); the loss table is combined with one tiny all-reduce, after which the score maths of the reference (:136-187)
    runs unchanged on every rank and the validator rank commits the weights.  Every rank calls :meth:`validate_and_score`.
    """

    def __init__(self, device, model, data_loader, bittensor_network, exchange, miner_ranks, validator_rank: int = 0,
                 eval_rows: Optional[int] = None, group=None, **kw):
        self.exchange = exchange
        self.miner_ranks = list(miner_ranks)
        self.validator_rank = validator_rank
        self.group = group
        import torch.distributed as dist
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        batches = list(data_loader)
        if eval_rows:
            batches = rebatch(batches, eval_rows)
        rows = max(int(b["input_ids"].shape[0]) for b in batches)
        seq = int(batches[0]["input_ids"].shape[1])
        self.batches = [{k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in b.items()} for b in batches]
        self.evalm = EvalModel(model.cfg, model.man, device, rows, seq)
        self.round = 0
        super().__init__(device, model, None, self.batches, bittensor_network, kw.pop("hf_manager", None), **kw)

    def evaluate_model(self, metric: str = "loss") -> Tuple[float, float]:
        """Loss / perplexity of the CURRENT base (constructor, base refresh)."""
        self.evalm.load(self.model.base)
        avg = float(self.evalm.mean_loss(self.batches))
        return avg, math.exp(min(avg, 50.0))

    def _miner_ok(self, flags, src: int, round: int) -> bool:
        ex = self.exchange
        return int(flags[ex.F_DELTA + src]) >= round and int(flags[ex.F_BAD + src]) != round

    def validate_and_score(self, round: Optional[int] = None, base_changed: bool = True) -> Dict[str, float]:
        """``base_changed=False``: theta_base is the one scored last time -- its loss is reused instead of re-evaluated (the
        reference evaluates the base once, in the constructor, :48)."""
        import torch.distributed as dist
        ex, net = self.exchange, self.bittensor_network
        self.round = round if round is not None else self.round + 1
        r = self.round
        hot = list(net.metagraph.hotkeys)
        reuse_base = (not base_changed) and self.base_loss == self.base_loss
        jobs = list(self.miner_ranks) + ([] if reuse_base else [-1])   # -1 = the base model
        table = torch.zeros(len(jobs), dtype=torch.float64, device=self.evalm.device)
        flags = ex.win.flags().tolist() if hasattr(ex, "win") else None  # ONE host read of the local flag page per round
        mode = {"fp32": 0, "bf16": 1, "fp8": 2}.get(getattr(ex, "delta_dtype_name", "fp32"), 0)
        valid = []
        for k, src in enumerate(jobs):
            ok = src < 0 or flags is None or self._miner_ok(flags, src, r)
            valid.append(ok)
            if not ok or k % self.world != self.rank:
                continue
            if src < 0:
                self.evalm.load(self.model.base)
            else:
                d, s = ex._delta_ptrs(r, [src])
                self.evalm.load(self.model.base, d[0], s[0] if s else None, mode)
            table[k] = self.evalm.mean_loss(self.batches)
        if self.world > 1:
            dist.all_reduce(table, group=self.group)
        losses = table.tolist()                         # the round's single device->host read of results
        if not reuse_base:
            self.base_loss = losses[-1]
            self.base_perplexity = math.exp(min(self.base_loss, 50.0))
        for k, src in enumerate(self.miner_ranks):
            hk = f"rank{src}" if f"rank{src}" in hot else (hot[src] if src < len(hot) else str(src))
            if valid[k]:
                loss = losses[k]
                ppl = math.exp(min(loss, 50.0))
                ls, ps = max(0.0, self.base_loss - loss), max(0.0, self.base_perplexity - ppl)
            else:
                loss, ppl, ls, ps = float("nan"), float("nan"), 0.0, 0.0
            self.losses[hk], self.loss_scores[hk], self.scores[hk] = loss, ls, ps
            net.metrics_data[hk] = {"loss": loss if loss == loss else 1e9}
            if self.rank == self.validator_rank:
                self.metrics.log(hotkey=hk, loss=loss, perplexity=ppl, loss_score=ls, perplexity_score=ps)
        total = sum(self.scores.values())
        self.normalized_scores = {hk: (max(0.0, sc / total) if total > 0 else 0.0) for hk, sc in self.scores.items()}
        if self.rank == self.validator_rank and net.should_set_weights():
            net.set_weights(self.normalized_scores)
        return self.normalized_scores


class LocalValidator(ModelValidator):
    """Deltas come from a local directory tree ``<repo>/gradients.pt`` (reference :206-248)."""

    def _receive(self, hotkey: str):
        import os
        repo = self.chain_manager.retrieve_hf_repo(hotkey) if self.chain_manager is not None else hotkey
        if repo is None:
            return None
        path = os.path.join(str(repo), "gradients.pt")
        if os.path.exists(path):
            try:
                return torch.load(path, map_location="cpu", weights_only=False)
            except Exception as e:
                logger.warning(f"Error loading gradients from {path}: {e}")
                return None
        return self.hf_manager.receive_flat(repo) if self.hf_manager is not None else None


class LocalDeltaValidator(DeltaValidator, LocalValidator):
    pass


class MNISTValidator(LocalValidator):
    """Accuracy-aware evaluation for (images, labels) batches (reference :265-310)."""

    def evaluate_model(self, metric: str = "loss") -> Tuple[float, float]:
        mod = self.model.module
        was = mod.training
        mod.eval()
        tot, correct, n = 0.0, 0, 0
        with torch.no_grad():
            for x, y in self.data_loader or []:
                x, y = x.to(self.device), y.to(self.device)
                out = mod(x)
                tot += float(torch.nn.functional.cross_entropy(out, y, reduction="sum"))
                correct += int((out.argmax(1) == y).sum())
                n += y.numel()
        mod.train(was)
        self.accuracy = correct / max(n, 1)
        avg = tot / max(n, 1)
        return avg, math.exp(min(avg, 50.0))


class MNISTDeltaValidator(DeltaValidator, MNISTValidator):
    pass
