"""Averager role: merge the miners' deltas into the next base model.

Reference: hivetrain/averaging_logic.py -- ``Averager`` (v0, score-weighted *gradient* averaging, :27-197),
``DeltaAverager`` (v1, score-weighted *weight* averaging, :200-269), ``LocalAverager`` (:272-332),
``ParameterizedAverager`` (v2, learned per-(miner, tensor) mixing matrix -- production, :335-583), its ``Local*``
twins (:586-827) and ``GeneticAverager`` (:830-970).  Algorithm spec: SURVEY.md section 2.6-C.

Maths preserved, execution re-designed:

* ``theta_bar_j = sum_i w_ij (theta_base_j + delta_ij)``  ==  ``s_j theta_base_j + sum_i w_ij delta_ij``  (``s_j = sum_i
  w_ij``; rows of ``w`` drift from 1 and may go negative -- kept): ONE fused kernel over the flat arena that reads the
  deltas (resident in HBM or straight from the miners' NVLink peer windows), never ``N x 148`` axpys + 2N disk loads;
* meta-gradient ``G_ij = <dL/dtheta_bar_j, theta_ij - theta_bar_j>`` = ``<g_j, delta_ij> + <g_j, theta_base_j - theta_bar_j>``:
  one segmented multi-dot pass producing the whole ``[N, P]`` matrix;
* the nested ``meta_epochs x meta_epochs`` loop, ``w <- w - lr G`` with no re-normalisation, and ``w = 1/N`` init are
  exactly the reference's (:423-430, 493-494, 528).
"""
from __future__ import annotations

import math
import os
import time
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from . import ops
from .utils.logging import MetricsLogger, logger
from .validation_logic import _batch_ids_labels


class Averager:
    """v0: score-weighted gradient averaging applied as theta -= alpha * g_bar, alpha = 1e-5 (reference :27-197)."""

    def __init__(self, model, local_dir: str, bittensor_network=None, chain_manager=None, hf_manager=None, hf_token=None,
                 device=None, gradients_dir: Optional[str] = None, metrics: Optional[MetricsLogger] = None):
        self.model = model
        self.local_dir = local_dir
        self.gradients_dir = gradients_dir or os.path.join(local_dir, "gradients")
        self.bittensor_network = bittensor_network
        self.chain_manager = chain_manager
        self.hf_manager = hf_manager
        self.hf_token = hf_token
        self.device = device if device is not None else model.master.device
        self.metrics = metrics or MetricsLogger(None, "averager")
        self.last_pull_time = 0.0
        self.check_update_interval = 300
        self.round = 0
        self.checkpoint_hook = None  # callable(averager, round): periodic --save_every

    # -- durable state (the reference re-creates w every round and never saves it: averaging_logic.py:492) ------------------
    def state_dict(self) -> Dict:
        hf = self.hf_manager
        return {"round": self.round, "weights": getattr(self, "weights", None), "consumed": dict(getattr(self, "_consumed", {})),
                "miner_hotkeys": list(getattr(self, "miner_hotkeys", [])), "published_round": getattr(hf, "_published_round", 0),
                "avg_round": getattr(self, "_avg_round", 0)}

    def load_state_dict(self, sd: Dict) -> None:
        self.round = int(sd.get("round", 0))
        if sd.get("weights") is not None and hasattr(self, "weights"):
            self.weights = sd["weights"].to(self.model.master.device)
        if hasattr(self, "_consumed"):
            self._consumed.update(sd.get("consumed", {}))
        self._avg_round = int(sd.get("avg_round", 0))
        if self.hf_manager is not None:
            self.hf_manager._published_round = max(getattr(self.hf_manager, "_published_round", 0), int(sd.get("published_round", 0)))

    # -- inputs -----------------------------------------------------------------------------------------------------
    def _repo_of(self, hotkey: str):
        return self.chain_manager.retrieve_hf_repo(hotkey) if self.chain_manager is not None else hotkey

    @staticmethod
    def have_nans(t) -> bool:
        if isinstance(t, dict):
            return any(bool(torch.isnan(v).any()) for v in t.values())
        return bool(torch.isnan(t.float()).any())

    def receive_gradients(self, repo_id, gradient_file_name: str = "gradients.pt") -> Optional[torch.Tensor]:
        """Flat delta/gradient of one miner, ``None`` if missing, malformed or NaN (reference :60-78)."""
        if repo_id is None:
            return None
        try:
            flat = self.hf_manager.receive_flat(repo_id)
            if flat is None or flat.numel() != self.model.man.total:
                return None
            if flat.dtype != torch.uint8 and self.have_nans(flat):
                return None
            return flat
        except Exception as e:
            logger.warning(f"Error receiving gradients from {repo_id}: {e}")
            return None

    def receive_and_score_gradients(self) -> Tuple[List[Optional[torch.Tensor]], torch.Tensor]:
        """Per-miner score = mean over validator rows of the metagraph weight matrix (reference :80-119)."""
        net = self.bittensor_network
        net.sync(lite=False)
        vuids = net.get_validator_uids(vpermit_tao_limit=1024)
        W = net.metagraph.W
        scores = W[vuids].mean(0) if len(vuids) else torch.zeros(net.metagraph.n)
        grads = [self.receive_gradients(self._repo_of(hk)) for hk in net.metagraph.hotkeys]
        return grads, scores

    # -- maths ------------------------------------------------------------------------------------------------------
    def average_gradients(self, gradients: Sequence[Optional[torch.Tensor]], scores: torch.Tensor, beta: float = 1.0) -> torch.Tensor:
        """g_bar = sum_i g_i * score_i * beta over the miners that delivered (reference :129-147)."""
        m = self.model
        out = torch.zeros_like(m.master)
        for g, sc in zip(gradients, scores):
            if g is None:
                continue
            out.add_(g.to(out.device, torch.float32), alpha=float(sc) * beta)
        return out

    def apply_averaged_gradients(self, averaged: torch.Tensor, alpha: float = 1e-5) -> None:
        m = self.model
        m.master.add_(averaged, alpha=-alpha)
        if getattr(m, "is_cuda", False):
            ops.cast_copy(m.master, m.p16)

    # -- outputs ----------------------------------------------------------------------------------------------------
    def save_model(self) -> str:
        os.makedirs(self.local_dir, exist_ok=True)
        path = os.path.join(self.local_dir, "averaged_model.pt")
        tmp = f"{path}.tmp.{os.getpid()}"
        # the reference's checkpoint format: ``model.state_dict()`` of the HF model (:481-488) -- Conv1D weights [in, out],
        # tied lm_head present -- so ``averaged_model.pt`` loads into GPT2LMHeadModel and into every loader of this package
        m = self.model
        sd = m.hf_state_dict("master") if hasattr(m, "hf_state_dict") else m.man.views(m.master.detach().cpu())
        torch.save(sd, tmp)
        os.replace(tmp, path)
        return path

    def push_to_hf_hub(self, commit_message: str = "Pushing model to Hub") -> None:
        if self.hf_manager is not None:
            self.round += 1
            self.hf_manager.push_to_hf_hub(base=self.model.master, commit_message=commit_message)

    def _adopt_as_base(self) -> None:
        """The published average becomes everybody's next theta_base -- including the averager's own."""
        self.model.base.copy_(self.model.master)

    def run_periodic_averaging(self, t: float, max_rounds: Optional[int] = None) -> None:
        rounds = 0
        while True:
            t0 = time.time()
            logger.info("Averaging gradients...")
            grads, scores = self.receive_and_score_gradients()
            self.apply_averaged_gradients(self.average_gradients(grads, scores))
            self.save_model()
            self._adopt_as_base()
            self.push_to_hf_hub(commit_message="Updated model with new gradients")
            if self.checkpoint_hook is not None:
                self.checkpoint_hook(self, self.round)
            rounds += 1
            if max_rounds is not None and rounds >= max_rounds:
                return
            time.sleep(max(0.0, t - (time.time() - t0)))


class DeltaAverager(Averager):
    """v1: theta_bar = sum_i (theta_base + delta_i) * score_i * beta / N_present (reference :221-242).  The reference's
    ``apply_averaged_gradients`` is a no-op bug (:244-248); here it really installs the average."""

    def average_gradients(self, gradients, scores, beta: float = 1.0) -> torch.Tensor:
        m = self.model
        present = [(g, float(sc)) for g, sc in zip(gradients, scores) if g is not None]
        if not present:
            return m.base.clone()
        n = len(present)
        P = len(m.man)
        w = torch.tensor([[sc * beta / n] * P for _, sc in present], dtype=torch.float32, device=m.master.device)
        out = torch.empty_like(m.master)
        ops.weighted_avg(m.base, [g.to(m.master.device) for g, _ in present], w, m.man, [out])
        return out

    def apply_averaged_gradients(self, averaged: torch.Tensor, alpha: float = 1.0) -> None:
        m = self.model
        m.master.copy_(averaged)
        if getattr(m, "is_cuda", False):
            ops.cast_copy(m.master, m.p16)


class LocalAverager(DeltaAverager):
    """Simulation twin: ``push_to_hf_hub`` just saves ``averaged_model.pt`` (reference :272-332)."""

    def push_to_hf_hub(self, commit_message: str = "Pushing model to Hub") -> None:
        self.save_model()
        if self.hf_manager is not None:
            super().push_to_hf_hub(commit_message)


class ParameterizedAverager(DeltaAverager):
    """v2 (production): learned mixing matrix ``w[N, P]`` trained by SGD on validation loss (reference :335-583)."""

    def __init__(self, model, device=None, hf_manager=None, local_dir: str = ".", gradients_dir: Optional[str] = None,
                 chain_manager=None, bittensor_network=None, hf_token=None, metrics: Optional[MetricsLogger] = None,
                 cache_to_disk: bool = False, exchange=None, fresh_only: bool = False, idle_timeout: float = 600.0):
        super().__init__(model, local_dir, bittensor_network, chain_manager, hf_manager, hf_token, device, gradients_dir, metrics)
        self.cache_to_disk = cache_to_disk
        self.fresh_only = fresh_only  # only average deltas published since the last averaging (reference: re-averages stale ones)
        self.idle_timeout = idle_timeout
        self._consumed: Dict[str, int] = {}
        self.exchange = exchange if exchange is not None else getattr(hf_manager, "exchange", None)
        self.weights: Optional[torch.Tensor] = None  # w[N, P]
        self.deltas: List[torch.Tensor] = []         # resident flat deltas of the miners that delivered
        self.dscales: Optional[List[torch.Tensor]] = None
        self.miner_hotkeys: List[str] = []
        self._G: Optional[torch.Tensor] = None

    # -- gathering ----------------------------------------------------------------------------------------------------
    def get_model_paths(self) -> List[Tuple[str, Optional[str]]]:
        """(hotkey, repo) for every registered hotkey (reference :365-376)."""
        return [(hk, self._repo_of(hk)) for hk in self.bittensor_network.metagraph.hotkeys]

    def _delta_file(self, hotkey: str) -> str:
        return os.path.join(self.gradients_dir, f"weight_delta_{hotkey}.pt")

    def store_weight_delta(self, flat: torch.Tensor, hotkey: str) -> None:
        os.makedirs(self.gradients_dir, exist_ok=True)
        tmp = self._delta_file(hotkey) + f".tmp.{os.getpid()}"
        torch.save(flat.detach().cpu(), tmp)
        os.replace(tmp, self._delta_file(hotkey))

    def load_weight_delta(self, hotkey: str) -> Optional[torch.Tensor]:
        p = self._delta_file(hotkey)
        return torch.load(p, map_location=self.device, weights_only=False) if os.path.exists(p) else None

    def cache_params_locally(self) -> int:
        """Collect every valid, shape-matching delta (reference :396-419).  Deltas stay RESIDENT (HBM or peer windows);
        a miner without a repo / with a stale flag / with NaNs is skipped instead of crashing the round."""
        self.deltas, self.miner_hotkeys = [], []
        for hotkey, repo in self.get_model_paths():
            if repo is None:
                continue
            if self.fresh_only and self.hf_manager is not None:
                r = self.hf_manager.delta_round(repo)
                if r <= self._consumed.get(hotkey, 0):
                    continue
            flat = self.receive_gradients(repo)
            if flat is None:
                continue
            if self.fresh_only:
                self._consumed[hotkey] = r
            flat = flat if flat.device == self.model.master.device else flat.to(self.model.master.device)
            self.deltas.append(flat)
            self.miner_hotkeys.append(hotkey)
            if self.cache_to_disk:
                self.store_weight_delta(flat, hotkey)
        ex = getattr(self.hf_manager, "exchange", None)
        if hasattr(ex, "stale_miners"):  # peer plane: name the ranks whose heartbeat lags behind the averaging round
            self._avg_round = getattr(self, "_avg_round", 0) + 1
            stale = ex.stale_miners(self._avg_round)
            if stale:
                logger.warning(f"averaging round {self._avg_round}: ranks {stale} have not published (skipped)")
        return len(self.deltas)

    @property
    def num_models(self) -> int:
        return len(self.deltas)

    # -- mixing -------------------------------------------------------------------------------------------------------
    def _ensure_weights(self) -> torch.Tensor:
        N, P = self.num_models, len(self.model.man)
        if self.weights is None or tuple(self.weights.shape) != (N, P):
            # softmax(ones, dim=0) == 1/N everywhere; plain tensor, never re-normalised (reference :423-430)
            self.weights = torch.softmax(torch.ones(N, P, dtype=torch.float32, device=self.model.master.device), dim=0)
        return self.weights

    def get_averaged_params(self, out: Optional[torch.Tensor] = None, out16: Optional[torch.Tensor] = None) -> torch.Tensor:
        """theta_bar(w) in one fused pass (reference :422-448 does N x P axpys after 2N torch.loads)."""
        m = self.model
        w = self._ensure_weights()
        out = out if out is not None else torch.empty_like(m.master)
        ops.weighted_avg(m.base, self.deltas, w, m.man, [out], [out16] if out16 is not None else None, dscales=self.dscales)
        return out

    def lazy_load_params(self) -> Iterator[torch.Tensor]:
        """theta_i = theta_base + delta_i, one miner at a time (API parity with reference :450-470)."""
        for d in self.deltas:
            yield self.model.base + d.to(self.model.base.dtype)

    def get_averaged_model(self):
        """Install theta_bar(w) into the model's master (+bf16 compute copy) and return the model (reference :472-479)."""
        m = self.model
        self.get_averaged_params(m.master, m.p16 if getattr(m, "is_cuda", False) else None)
        return m

    # -- learning the mixing weights ------------------------------------------------------------------------------------
    def meta_step(self, batch, lr: float) -> torch.Tensor:
        """One SGD step on ``w``: rebuild theta_bar, fwd/bwd on the batch, G = multi-dot, w -= lr * G."""
        m = self.model
        self.get_averaged_model()
        ids, labels = _batch_ids_labels(batch)
        if isinstance(batch, dict) and hasattr(m, "engine"):
            loss = m.loss_and_grad(batch)  # labels + attention_mask travel with the dict (reference :502-506)
        else:
            loss = m.loss_and_grad(batch if not isinstance(batch, dict) else ids, labels)
        N, P = self.weights.shape
        if self._G is None or tuple(self._G.shape) != (N, P):
            self._G = torch.empty(N, P, dtype=torch.float32, device=m.master.device)
        ops.multi_dot(m.grad, self.deltas, m.base, m.master, m.man, self._G, dscales=self.dscales)
        self.weights.add_(self._G, alpha=-lr)
        return loss

    def meta_learning(self, val_loader: Iterable, meta_epochs: int, lr: float):
        """The reference nests ``for epoch in range(meta_epochs)`` twice (:493-494) -> meta_epochs^2 passes; kept."""
        self._ensure_weights()
        for outer in range(meta_epochs):
            for inner in range(meta_epochs):
                tot, n = None, 0
                for batch in val_loader:
                    ids, _ = _batch_ids_labels(batch)
                    bs = ids.shape[0] if hasattr(ids, "shape") else len(batch[0])
                    loss = self.meta_step(batch, lr)
                    tot = loss.detach().double() * bs if tot is None else tot + loss.detach().double() * bs
                    n += bs
                if tot is not None:
                    avg = float(tot) / max(n, 1)  # one host read per pass
                    wm = self.weights.mean(dim=1)
                    self.metrics.log(meta_epoch=outer * meta_epochs + inner, loss_averaged=avg,
                                     perplexity_averaged=math.exp(min(avg, 50.0)), w_mean=[float(x) for x in wm])
                    logger.info(f"Meta-epoch {outer}.{inner}: val loss {avg:.4f} ppl {math.exp(min(avg, 50.0)):.3f} "
                                f"w_mean {[round(float(x), 4) for x in wm]}")
        return self.get_averaged_model()

    def run_periodic_averaging(self, val_loader: Iterable, meta_epochs: int, lr: float, t: float,
                               max_rounds: Optional[int] = None) -> None:
        rounds = 0
        while True:
            t0 = time.time()
            if self.hf_manager is not None and time.time() - self.last_pull_time >= self.check_update_interval:
                self.last_pull_time = time.time()
                if self.hf_manager.check_for_new_submissions(self.hf_manager.model_repo_id):
                    self.hf_manager.pull_latest_model()
                    self.model = self.hf_manager.update_model(self.model, reset_optimizer=False)
            n = self.cache_params_locally()
            if n > 0:
                self.weights = None  # N (and w) are rebuilt from scratch each round (reference :492)
                if meta_epochs > 0:
                    self.meta_learning(val_loader, meta_epochs, lr)
                else:
                    self.get_averaged_model()  # uniform mixer: w = 1/N, no meta-steps
                self.save_model()
                self._adopt_as_base()
                self.push_to_hf_hub(commit_message="Updated model with new gradients")
                if self.checkpoint_hook is not None:
                    self.checkpoint_hook(self, self.round)
                idle_since = time.time()
            else:
                logger.debug("No valid deltas this round")
                if self.fresh_only:  # wait for fresh submissions instead of burning a round
                    idle_since = locals().get("idle_since", t0)
                    if time.time() - idle_since > self.idle_timeout:
                        logger.warning("averager idle timeout")
                        return
                    time.sleep(0.05)
                    continue
            rounds += 1
            if max_rounds is not None and rounds >= max_rounds:
                return
            time.sleep(max(0.0, t - (time.time() - t0)))


class LocalParameterizedAverager(ParameterizedAverager):
    """Simulation twin for toy models: deltas come from ``<miner_dir>/gradients.pt`` or the disk exchange; also reports
    accuracy when the batches are (x, y) pairs (reference :586-760)."""

    def receive_gradients(self, repo_id, gradient_file_name: str = "gradients.pt"):
        if repo_id is not None and isinstance(repo_id, str) and os.path.isdir(repo_id):
            p = os.path.join(repo_id, gradient_file_name)
            if os.path.exists(p):
                blob = torch.load(p, map_location="cpu", weights_only=False)
                flat = self.model.flat_from(blob) if hasattr(self.model, "flat_from") else (
                    blob if isinstance(blob, torch.Tensor) else self.model.man.pack(
                        blob, torch.zeros(self.model.man.total, dtype=torch.float32)))
                return None if self.have_nans(flat) else flat
        return super().receive_gradients(repo_id, gradient_file_name)

    def push_to_hf_hub(self, commit_message: str = "Pushing model to Hub") -> None:
        self.save_model()
        if self.hf_manager is not None:
            super().push_to_hf_hub(commit_message)


class LocalLLMParameterizedAverager(LocalParameterizedAverager):
    """Twin for dict batches ``{"input_ids", "attention_mask", "labels"}`` (reference :763-827) -- the generic
    :meth:`ParameterizedAverager.meta_step` already handles them."""


class GeneticAverager(ParameterizedAverager):
    """Evolutionary search over per-miner scalar weights: population 10 x 10 generations, sigma = 0.1 Gaussian mutation,
    top 50 % survive, fitness = -validation loss, theta_bar = sum_i w_i (theta_base + delta_i) / N (reference :830-970)."""

    def __init__(self, *a, population_size: int = 10, num_generations: int = 10, sigma: float = 0.1, topk_percent: float = 0.5,
                 seed: int = 0, **kw):
        super().__init__(*a, **kw)
        self.population_size, self.num_generations, self.sigma, self.topk_percent = population_size, num_generations, sigma, topk_percent
        self.gen = torch.Generator().manual_seed(seed)

    def get_averaged_model(self, weights: Optional[torch.Tensor] = None):
        m = self.model
        N, P = self.num_models, len(m.man)
        if weights is None:
            return super().get_averaged_model()
        w = (weights.to(m.master.device).float() / N)[:, None].expand(N, P).contiguous()
        ops.weighted_avg(m.base, self.deltas, w, m.man, [m.master], [m.p16] if getattr(m, "is_cuda", False) else None,
                         dscales=self.dscales)
        return m

    def evaluate_population(self, population: torch.Tensor, val_loader: Iterable) -> torch.Tensor:
        fit = []
        for ind in population:
            self.get_averaged_model(ind)
            tot, n = 0.0, 0
            for batch in val_loader:
                ids, labels = _batch_ids_labels(batch)
                bs = ids.shape[0] if hasattr(ids, "shape") else len(batch[0])
                full = isinstance(batch, dict) and hasattr(self.model, "engine")
                tot += float(self.model.eval_loss(batch) if full else self.model.eval_loss(batch if not isinstance(batch, dict) else ids, labels)) * bs
                n += bs
            fit.append(-tot / max(n, 1))
        return torch.tensor(fit)

    def evolve_population(self, population: torch.Tensor, fitness: torch.Tensor) -> torch.Tensor:
        k = max(1, int(len(population) * self.topk_percent))
        top = population[fitness.argsort(descending=True)[:k]]
        children = []
        while len(children) < len(population) - k:
            parent = top[int(torch.randint(0, k, (1,), generator=self.gen))]
            children.append(parent + torch.randn(parent.shape, generator=self.gen) * self.sigma)
        return torch.cat([top, torch.stack(children)]) if children else top

    def run_evolution(self, val_loader: Iterable):
        N = self.num_models
        pop = torch.ones(self.population_size, N) + torch.randn(self.population_size, N, generator=self.gen) * self.sigma
        pop[0] = 1.0
        best, best_fit = pop[0].clone(), -float("inf")
        for g in range(self.num_generations):
            fit = self.evaluate_population(pop, val_loader)
            i = int(fit.argmax())
            if float(fit[i]) > best_fit:
                best, best_fit = pop[i].clone(), float(fit[i])
            self.metrics.log(generation=g, best_loss=-best_fit)
            logger.info(f"Generation {g}: best loss {-best_fit:.4f}")
            pop = self.evolve_population(pop, fit)
        self.best_weights = best
        return self.get_averaged_model(best)
