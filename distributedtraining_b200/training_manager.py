"""Miner role: train a private copy, emit ``delta = theta - theta_base``, adopt new averaged bases.

Reference: hivetrain/training_manager.py -- ``TrainingLoop`` (:28-168, legacy gradient-accumulating miner),
``DeltaLoop`` (:345-433, the production loop), ``LocalTrainingLoop`` / ``LocalDeltaLoop`` (:326-342, 436-437),
``MNISTTrain`` / ``MNISTDeltaTrain`` (:462-803), ``FeedforwardNN`` (:440-459) and hivetrain/new_training_manager.py
``TrainingLoopNew`` (:20-169).  Algorithm spec: SURVEY.md section 2.6-A.

B200-first differences (semantics preserved):
* the model is a :class:`~distributedtraining_b200.models.trainer.Trainer` (flat arenas, hand-written kernels, the whole
  step in one CUDA graph) instead of an HF ``nn.Module``; the loss stays on the device (no ``loss.item()`` per step);
* cadence is *rounds of ``local_steps`` optimizer steps* (default 100) instead of ``send_interval=800 s`` wall-clock
  (wall-clock mode is kept: pass ``local_steps=None``);
* a push is "write delta into my symmetric window + release-store the round flag", not ``torch.save`` + git push.
"""
from __future__ import annotations

import hashlib
import math
import os
import time
from typing import Callable, Dict, Iterable, Optional

import torch

from . import ops  # noqa: F401  (re-exported: callers reach the op layer through this module as upstream scripts did)
from .config.mlflow_config import MLFLOW_ACTIVE
from .models.toys import FeedforwardNN, ModuleTrainer, SimpleCNN  # noqa: F401  (re-exported like the reference)
from .models.trainer import Trainer
from .utils.logging import MetricsLogger, logger
from .utils.mlflow_utils import initialize_mlflow, log_model_metrics


def _ids_of(batch):
    return batch["input_ids"] if isinstance(batch, dict) else batch


def _labels_of(batch):
    # the reference miner passes labels = input_ids (training_manager.py:383): PAD is NOT masked
    return None


def calculate_model_hash(trainer) -> str:
    """sha256 over all parameters (reference training_manager.py:697-702 / validation_logic.py:198-203)."""
    h = hashlib.sha256()
    h.update(trainer.master.detach().float().cpu().numpy().tobytes())
    return h.hexdigest()


def normalize_gradients(flat_grad: torch.Tensor, manifest, threshold: float = 0.1) -> torch.Tensor:
    """Per-tensor norm clipping (reference training_manager.py:493-508): tensors whose norm exceeds ``threshold`` are
    rescaled to it."""
    for s in manifest:
        g = flat_grad[s.offset:s.offset + s.numel]
        n = g.norm()
        if n > threshold:
            g.mul_(threshold / n)
    return flat_grad


class TrainingLoop:
    """Base miner.  ``train`` here is the legacy loop: accumulate raw gradients over ``send_interval`` and publish the
    *accumulated gradient* (what reference :75-79,116-118 set out to do; its :146-147 saves the full state dict)."""

    def __init__(self, device, model_name="gpt2", data_loader: Optional[Iterable] = None, learning_rate: float = 5e-5,
                 check_update_interval: float = 300, send_interval: float = 300, hf_manager=None, *, trainer=None,
                 batch_size: int = 1, seq_len: int = 64, local_steps: Optional[int] = 100, post_pull_lr: float = 5e-5,
                 reset_optimizer: bool = True, max_steps: Optional[int] = None, metrics: Optional[MetricsLogger] = None,
                 round_hook: Optional[Callable] = None, seed: int = 0, my_hotkey: str = "miner",
                 host_loss_every_step: bool = False):
        self.device = device
        self.model_name = model_name
        # the reference loads tokenizer + AutoModelForCausalLM.from_pretrained here (:39-46); offline we build the
        # named architecture with random init (or take an injected trainer, like TrainingLoopNew does)
        # ``model_name`` may also be an HF checkpoint DIRECTORY: it is then loaded and grown by the [PAD] row exactly like
        # the reference constructor does (from_pretrained + add [PAD] + resize_token_embeddings, :39-46)
        self.model = trainer if trainer is not None else Trainer(model_name, device=device, batch=batch_size, seq=seq_len,
                                                                 lr=learning_rate, seed=seed)
        self.hf_manager = hf_manager
        self.learning_rate = learning_rate
        self.data_loader = data_loader
        self.check_update_interval = check_update_interval
        self.send_interval = send_interval
        self.local_steps = local_steps
        self.post_pull_lr = post_pull_lr
        self.reset_optimizer = reset_optimizer
        self.max_steps = max_steps
        self.metrics = metrics or MetricsLogger(None, "miner")
        self.round_hook = round_hook
        self.checkpoint_hook = None  # callable(loop) run after every completed round (periodic --save_every)
        self.last_pull_time = 0.0
        self.last_send_time = time.time()
        self.global_step = 0
        self.rounds_sent = 0
        self.my_hotkey = my_hotkey
        self._loss_acc = torch.zeros((), dtype=torch.float32, device=self.model.master.device)
        self._loss_n = 0
        # optional per-step device->host read-back of the loss into pinned memory (asynchronous; the reference blocks on
        # loss.item() every step, training_manager.py:388)
        self.host_losses = None
        if host_loss_every_step:
            n = max_steps or 1024
            self.host_losses = torch.zeros(n, dtype=torch.float32)
            if torch.cuda.is_available():
                self.host_losses = self.host_losses.pin_memory()
        if MLFLOW_ACTIVE:
            initialize_mlflow(role="miner", device=device, version=None, my_hotkey=my_hotkey, learning_rate=learning_rate,
                              send_interval=send_interval, check_update_interval=check_update_interval)
        else:
            logger.debug("mlflow inactive")

    # -- shared pieces ----------------------------------------------------------------------------------------------
    def _due_poll(self) -> bool:
        if self.local_steps is not None:
            return False  # round mode: bases are adopted at round boundaries (see _end_round)
        return time.time() - self.last_pull_time >= self.check_update_interval

    def _due_send(self) -> bool:
        if self.local_steps is not None:
            return self.global_step > 0 and self.global_step % self.local_steps == 0
        return time.time() - self.last_send_time >= self.send_interval

    def _maybe_pull(self, force_check: bool = False) -> bool:
        """Poll the averaged model; on change: pull, load, re-create the optimizer at ``post_pull_lr``, re-snapshot base
        (reference :361-378)."""
        if self.hf_manager is None:
            return False
        pulled = False
        if self.hf_manager.check_for_new_submissions(self.hf_manager.model_repo_id):
            logger.info("Averaged model updated. Pulling latest model...")
            self.hf_manager.pull_latest_model()
            self.model = self.hf_manager.update_model(self.model, lr=self.post_pull_lr, reset_optimizer=self.reset_optimizer)
            pulled = True
        self.last_pull_time = time.time()
        return pulled

    def get_gradient_staleness(self) -> float:
        """Seconds since the last successful send (logged only, as in the reference :156-168)."""
        return time.time() - self.last_send_time

    def _log_round(self, epoch: int) -> Dict:
        avg = float(self._loss_acc) / max(self._loss_n, 1)  # ONE device->host read per round
        self._loss_acc.zero_()
        self._loss_n = 0
        rec = self.metrics.log(round=self.rounds_sent, epoch=epoch, step=self.global_step, train_loss=avg,
                               perplexity=math.exp(min(avg, 50.0)), gradient_staleness=self.get_gradient_staleness())
        logger.info(f"Epoch: {epoch}, step {self.global_step}, Loss: {avg:.4f}")
        if MLFLOW_ACTIVE:
            log_model_metrics(step=self.global_step, train_loss=avg, gradient_staleness=rec["gradient_staleness"])
        return rec

    # -- legacy loop -------------------------------------------------------------------------------------------------
    def train(self, epochs: int):
        self.last_send_time = time.time()
        agg = torch.zeros_like(self.model.master)
        for epoch in range(int(epochs)):
            for step, batch in enumerate(self.data_loader):
                if self._due_poll():
                    self._maybe_pull()
                loss = self.model.step(batch if isinstance(batch, dict) and hasattr(self.model, "engine") else _ids_of(batch))
                agg.add_(self.model.grad)
                self._loss_acc += loss
                self._loss_n += 1
                self.global_step += 1
                if self._due_send():
                    self._log_round(epoch)
                    self.store_gradients(agg)
                    agg.zero_()
                    self.last_send_time = time.time()
                    self.rounds_sent += 1
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    return

    def store_gradients(self, flat: torch.Tensor) -> None:
        if self.hf_manager is not None:
            self.hf_manager.push_changes(trainer=_FlatAsDelta(flat))


class _FlatAsDelta:
    def __init__(self, flat):
        self.flat, self.master = flat, flat

    def emit_delta(self, out, scales=None, bad=None):
        out.copy_(self.flat.to(out.dtype))
        if bad is not None and not bool(torch.isfinite(self.flat).all()):
            bad.fill_(1)
        return out


class DeltaLoop(TrainingLoop):
    """The production miner loop (reference :345-433)."""

    def train(self, epochs: int):
        self.last_send_time = time.time()
        m = self.model
        # theta_base = snapshot of theta at start / at the last pull; NOT reset on send -> deltas are cumulative
        # w.r.t. the last pulled base (reference :405-431)
        for epoch in range(int(epochs)):
            logger.debug(f"Starting Epoch: {epoch}")
            for step, batch in enumerate(self.data_loader):
                if self._due_poll():
                    self._maybe_pull()
                # dict batches go through whole: the engine reads input_ids (labels = input_ids, PAD not masked) and applies the
                # attention_mask as padding mask, as the reference's model(...) call does (:380-384)
                loss = m.step(batch if isinstance(batch, dict) and hasattr(m, "engine") else _ids_of(batch), _labels_of(batch))
                if self.host_losses is not None:
                    self.host_losses[self.global_step % self.host_losses.numel()].copy_(loss, non_blocking=True)
                self._loss_acc += loss
                self._loss_n += 1
                self.global_step += 1
                if self._due_send():
                    self._end_round(epoch)
                    m = self.model
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    return

    def _end_round(self, epoch: int) -> None:
        try:
            logger.debug("Attempting to send weights")
            if self.round_hook is not None:
                # co-located synchronous round: publish -> fused sharded gather/avg/broadcast -> adopt the new base,
                # all stream-ordered on the device (parallel/local_sgd.py)
                self.round_hook(self)
            elif self.hf_manager is not None:
                self.hf_manager.push_changes("weight_diff.pt", trainer=self.model)
                if self.local_steps is not None:
                    self._maybe_pull()
            self.rounds_sent += 1
            self._log_round(epoch)
            self.last_send_time = time.time()
            if getattr(self, "checkpoint_hook", None) is not None:
                self.checkpoint_hook(self)
        except Exception as e:  # best effort, as in the reference (:428-431)
            logger.warning(f"Sending gradients failed: {e}")
            self.last_send_time = time.time()


class LocalTrainingLoop:
    """Mixin: store deltas/gradients in a local directory instead of pushing (reference :326-342)."""

    @staticmethod
    def store_gradients(aggregated_gradients, local_dir: str, gradient_file_name: str = "gradients.pt") -> str:
        """Reference signature (training_manager.py:327-342): save a name->tensor dict (or a flat tensor) locally."""
        if isinstance(aggregated_gradients, dict):
            return LocalTrainingLoop.store_gradients_to_dir(aggregated_gradients, local_dir, gradient_file_name)
        os.makedirs(local_dir, exist_ok=True)
        path = os.path.join(local_dir, gradient_file_name)
        torch.save(aggregated_gradients.detach().cpu(), path + ".tmp")
        os.replace(path + ".tmp", path)
        return path

    @staticmethod
    def store_gradients_to_dir(tensors: Dict[str, torch.Tensor], local_dir: str, gradient_file_name: str = "gradients.pt") -> str:
        os.makedirs(local_dir, exist_ok=True)
        path = os.path.join(local_dir, gradient_file_name)
        tmp = f"{path}.tmp.{os.getpid()}"
        torch.save({k: v.detach().cpu() for k, v in tensors.items()}, tmp)
        os.replace(tmp, path)
        return path


class LocalDeltaLoop(DeltaLoop, LocalTrainingLoop):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# toy-workload miners (CPU-second simulations of the whole subnet)
# ---------------------------------------------------------------------------------------------------------------------
class MNISTTrain(LocalTrainingLoop):
    """Accumulate norm-clipped gradients, apply them every ``n_steps``, then RETURN (train_loss, test_loss, test_acc)
    -- an experiment harness (reference :462-644)."""

    def __init__(self, model=None, device="cpu", lr: float = 0.1, train_loader=None, test_loader=None, clip: float = 0.1):
        self.device = device
        self.trainer = ModuleTrainer(model or FeedforwardNN(), device=device, lr=lr, optimizer="sgd")
        self.model = self.trainer.module
        self.train_loader, self.test_loader = train_loader, test_loader
        self.clip = clip

    def normalize_gradients(self, threshold: Optional[float] = None) -> None:
        normalize_gradients(self.trainer.grad, self.trainer.man, self.clip if threshold is None else threshold)

    def test(self):
        tot, correct, n = 0.0, 0, 0
        with torch.no_grad():
            for x, y in self.test_loader:
                x, y = x.to(self.device), y.to(self.device)
                out = self.model(x)
                tot += float(torch.nn.functional.cross_entropy(out, y, reduction="sum"))
                correct += int((out.argmax(1) == y).sum())
                n += y.numel()
        return tot / max(n, 1), correct / max(n, 1)

    def save_model(self, path: str) -> None:
        torch.save(self.trainer.man.views(self.trainer.master), path)

    def train(self, epochs: int = 1, hf_manager=None, n_steps: int = 10):
        t = self.trainer
        agg = torch.zeros_like(t.master)
        tot, cnt = 0.0, 0
        for epoch in range(epochs):
            for i, (x, y) in enumerate(self.train_loader):
                if hf_manager is not None and hf_manager.check_for_new_submissions():
                    hf_manager.pull_latest_model()
                    hf_manager.update_model(t)
                loss = t.loss_and_grad((x.to(self.device), y.to(self.device)))
                self.normalize_gradients()
                agg.add_(t.grad)
                tot += float(loss)
                cnt += 1
                if (i + 1) % n_steps == 0:
                    t.master.add_(agg, alpha=-t.opt.host["lr"] / n_steps)
                    if hf_manager is not None:
                        hf_manager.push_changes(trainer=_FlatAsDelta(agg / n_steps))
                    te_loss, te_acc = self.test() if self.test_loader is not None else (float("nan"), float("nan"))
                    return tot / cnt, te_loss, te_acc
        return tot / max(cnt, 1), float("nan"), float("nan")


class MNISTDeltaTrain(LocalTrainingLoop):
    """MNIST simulation of DeltaLoop: SGD lr 0.1 (0.001 after a pull), per-batch base polling, periodic delta store
    (reference :647-803)."""

    def __init__(self, model=None, device="cpu", lr: float = 0.1, post_pull_lr: float = 0.001, train_loader=None,
                 test_loader=None, send_every: int = 50):
        self.device = device
        self.trainer = ModuleTrainer(model or FeedforwardNN(), device=device, lr=lr, optimizer="sgd")
        self.model = self.trainer.module
        self.train_loader, self.test_loader = train_loader, test_loader
        self.post_pull_lr, self.send_every = post_pull_lr, send_every

    def calculate_model_hash(self) -> str:
        return calculate_model_hash(self.trainer)

    test = MNISTTrain.test
    save_model = MNISTTrain.save_model

    def train(self, epochs: int = 1, hf_manager=None, max_steps: Optional[int] = None):
        t = self.trainer
        step = 0
        losses = []
        for epoch in range(epochs):
            for x, y in self.train_loader:
                if hf_manager is not None and hf_manager.check_for_new_submissions():
                    logger.info("Model updated from the hub. Continuing training with new model...")
                    hf_manager.pull_latest_model()
                    hf_manager.update_model(t, lr=self.post_pull_lr)
                losses.append(float(t.step((x.to(self.device), y.to(self.device)))))
                step += 1
                if step % self.send_every == 0 and hf_manager is not None:
                    hf_manager.push_changes(trainer=t)
                if max_steps is not None and step >= max_steps:
                    return losses
        return losses


class MNISTDeltaTrainHugging(MNISTDeltaTrain):
    """Reference :171-323 is a near-duplicate of MNISTDeltaTrain that subclasses TrainingLoop with a broken
    ``super().__init__()``; kept as a working alias."""


class TrainingLoopNew(DeltaLoop):
    """Rewrite with the model *injected* and Adam lr 1e-3 (reference new_training_manager.py:20-169)."""

    def __init__(self, device, model, data_loader, test_loader=None, hf_manager=None, learning_rate: float = 1e-3,
                 send_interval: float = 800, check_update_interval: float = 300, **kw):
        trainer = model if hasattr(model, "master") else ModuleTrainer(model, device=device, lr=learning_rate)
        super().__init__(device, getattr(model, "name", "injected"), data_loader, learning_rate, check_update_interval,
                         send_interval, hf_manager, trainer=trainer, **kw)
        self.test_loader = test_loader

    def calculate_model_hash(self) -> str:
        return calculate_model_hash(self.model)

    def test(self) -> float:
        tot, n = 0.0, 0
        for batch in self.test_loader or []:
            tot += float(self.model.eval_loss(_ids_of(batch) if isinstance(batch, dict) else batch))
            n += 1
        return tot / max(n, 1)
