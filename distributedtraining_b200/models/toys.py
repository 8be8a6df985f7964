"""Toy workloads + a generic ``nn.Module`` adapter onto flat arenas.

The reference uses MNIST-sized models so the whole miner -> averager -> validator loop can run on a CPU in seconds
(``FeedforwardNN``: reference hivetrain/training_manager.py:440-459; ``SimpleCNN``: hivetrain/new_training_manager.py:
173-189).  :class:`ModuleTrainer` re-homes any module's parameters (and their ``.grad``) into flat arenas so that every
flat-arena op of the framework (delta emit, fused weighted average, multi-dot, fused AdamW) applies to it unchanged.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .arena import Manifest


class FeedforwardNN(nn.Module):
    """784-512-512-128-128-10 ReLU MLP (reference hivetrain/training_manager.py:440-459)."""

    def __init__(self, in_dim: int = 784, hidden=(512, 512, 128, 128), out_dim: int = 10):
        super().__init__()
        dims = [in_dim, *hidden, out_dim]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x):
        x = x.reshape(x.shape[0], -1)
        for l in self.layers[:-1]:
            x = F.relu(l(x))
        return self.layers[-1](x)


class SimpleCNN(nn.Module):
    """Two-conv MNIST net (reference hivetrain/new_training_manager.py:173-189)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 32, 3, 1)
        self.conv2 = nn.Conv2d(32, 64, 3, 1)
        self.fc1 = nn.Linear(9216, 128)
        self.fc2 = nn.Linear(128, 10)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = torch.flatten(x, 1)
        return self.fc2(F.relu(self.fc1(x)))


def manifest_of(module: nn.Module) -> Manifest:
    return Manifest([(n, tuple(p.shape), "normal", p.dim() > 1) for n, p in module.named_parameters()])


def classification_loss(module: nn.Module, batch) -> torch.Tensor:
    x, y = batch
    return F.cross_entropy(module(x), y)


class ModuleTrainer:
    """Same interface as :class:`distributedtraining_b200.models.trainer.Trainer` for an arbitrary ``nn.Module``."""

    def __init__(self, module: nn.Module, loss_fn: Callable = classification_loss, device="cpu", lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, optimizer: str = "adamw"):
        self.module = module.to(device)
        self.loss_fn = loss_fn
        self.device = torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.man = manifest_of(self.module)
        n = self.man.total
        f32 = dict(dtype=torch.float32, device=self.device)
        self.master = torch.zeros(n, **f32)
        self.grad = torch.zeros(n, **f32)
        for name, p in self.module.named_parameters():
            v = self.man.view(self.master, name)
            v.copy_(p.data)
            p.data = v  # parameters now alias the arena
            p.grad = self.man.view(self.grad, name)
        self.base = self.master.clone()
        self.m = torch.zeros(n, **f32)
        self.v = torch.zeros(n, **f32)
        self.p16 = self.master
        self.optimizer = optimizer
        self.opt = ops.AdamState(self.device, lr, betas[0], betas[1], eps, weight_decay)
        self.steps_done = 0
        self.cfg = None

    def loss_and_grad(self, batch, labels=None, zero_grad: bool = True) -> torch.Tensor:
        if zero_grad:
            self.grad.zero_()
        loss = self.loss_fn(self.module, batch)
        loss.backward()
        return loss.detach()

    def step(self, batch, labels=None) -> torch.Tensor:
        loss = self.loss_and_grad(batch)
        if self.optimizer == "sgd":
            self.master.add_(self.grad, alpha=-self.opt.host["lr"])
        else:
            ops.adamw_step(self.master, None, self.grad, self.m, self.v, self.opt)
        self.steps_done += 1
        return loss

    @torch.no_grad()
    def eval_loss(self, batch, labels=None) -> torch.Tensor:
        was = self.module.training
        self.module.eval()
        out = self.loss_fn(self.module, batch)
        self.module.train(was)
        return out

    def emit_delta(self, out, scales=None, bad=None):
        return ops.delta_emit(self.master, self.base, out, scales, bad)

    def load_base(self, new_base, lr: Optional[float] = None, reset_optimizer: bool = True):
        if new_base.data_ptr() != self.base.data_ptr():
            self.base.copy_(new_base)
        ops.round_reset(self.base, self.master, None, self.m, self.v, reset_optimizer)
        if reset_optimizer:
            self.opt.reset()
        if lr is not None:
            self.opt.set_lr(lr)
