"""Per-rank training state: flat arenas (fp32 master, bf16 compute copy, fp32 grads, Adam m/v, base snapshot) + the
engine + the fused optimizer, with the whole step captured in ONE CUDA graph on a B200.

This is the miner's inner loop of the reference (reference hivetrain/training_manager.py:380-392: forward, backward,
``optimizer.step()``, ``zero_grad()``, ``loss.item()`` sync every step) re-designed for the hardware: no per-step host
sync (the loss stays on the device), no per-tensor Python loops, launch overhead removed by graph replay.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from .. import ops
from .arena import Manifest
from .transformer import (ModelConfig, TransformerEngine, build_manifest, from_hf_state_dict, get_config, load_hf_checkpoint,
                          new_model, pack_any, to_hf_state_dict)


class Trainer:
    def __init__(self, model="gpt2", device="cpu", batch: int = 1, seq: int = 64, lr: float = 5e-4, seed: int = 0,
                 betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0, use_graph: Optional[bool] = None,
                 lm_chunk: int = 32768, init_flat: Optional[torch.Tensor] = None, fp8_forward: bool = False,
                 dropout: Optional[float] = None, meta_dropout: bool = False, dropout_seed: Optional[int] = None,
                 buffers: Optional[Dict[str, torch.Tensor]] = None, fp8_backward: bool = False):
        if isinstance(model, str) and os.path.isdir(model) and os.path.exists(os.path.join(model, "config.json")):
            # an HF checkpoint directory, as the reference's ``AutoModelForCausalLM.from_pretrained(model_name)`` + [PAD]
            # resize (hivetrain/training_manager.py:39-46)
            model, _, init_flat = load_hf_checkpoint(model, add_pad_token=True)
        self.cfg: ModelConfig = get_config(model) if isinstance(model, str) else model
        if dropout is not None and dropout != self.cfg.dropout:  # override the preset's train-mode dropout (0 disables)
            import dataclasses
            self.cfg = dataclasses.replace(self.cfg, dropout=float(dropout))
        self.man: Manifest = build_manifest(self.cfg)
        self.device = torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.batch, self.seq = batch, seq
        n = self.man.total
        f32 = dict(dtype=torch.float32, device=self.device)
        if init_flat is None:
            _, _, a = new_model(self.cfg, seed=seed)
            init_flat = a.flat
        self.master = init_flat.to(**f32).clone()
        # ``buffers``: externally owned arenas to live in -- PeerExchange.trainer_buffers() hands out the ``base`` (fp32) and
        # ``p16`` (bf16) regions of this rank's symmetric window, so that the averaging kernels of the peers can land the new
        # base / theta_bar DIRECTLY in the trainer's arenas (multimem.st through the NVSwitch), with no copy pass afterwards
        buffers = buffers or {}
        if "base" in buffers:
            self.base = buffers["base"][:n]
            self.base.copy_(self.master)
        else:
            self.base = self.master.clone()  # theta_base: the last pulled averaged model
        self.master_stale = False  # True between a pushed base and the first optimizer step (theta == theta_base, master unwritten)
        self.grad = torch.zeros(n, **f32)
        self.m = torch.zeros(n, **f32)
        self.v = torch.zeros(n, **f32)
        if self.is_cuda:
            self.p16 = buffers["p16"][:n] if "p16" in buffers else torch.empty(n, dtype=torch.bfloat16, device=self.device)
            ops.cast_copy(self.master, self.p16)
        else:
            self.p16 = self.master  # CPU: compute directly on the fp32 master
        self.opt = ops.AdamState(self.device, lr, betas[0], betas[1], eps, weight_decay)
        self.engine = TransformerEngine(self.cfg, self.man, self.p16, self.grad, batch, seq, lm_chunk=lm_chunk,
                                        fp8_forward=fp8_forward and self.is_cuda, fp8_backward=fp8_backward and self.is_cuda,
                                        seed=seed if dropout_seed is None else dropout_seed)
        # ``seed`` fixes the (shared) initial weights; ``dropout_seed`` decorrelates the dropout masks of co-located miners
        # (every rank is built with seed=0 so that all share theta_base -- their masks must still differ: pass the rank)
        # gradients WITHOUT an optimizer step (the averager's meta-learning) are deterministic by default; True reproduces the
        # reference, whose averager leaves dropout on
        self.meta_dropout = bool(meta_dropout)
        self.use_graph = self.is_cuda if use_graph is None else (use_graph and self.is_cuda)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graphs: Dict[bool, torch.cuda.CUDAGraph] = {}  # step graphs: plain / first-step-after-a-pushed-base (in-GEMM flag waits)
        self.ready_capable = False
        self._eval_graph: Optional[torch.cuda.CUDAGraph] = None
        self._lg_graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.steps_done = 0
        self.tokens_per_step = batch * seq

    # ------------------------------------------------------------------------------------------------------------
    def _step_body(self) -> torch.Tensor:
        loss = self.engine.forward_backward()
        ops.adamw_step(self.master, self.p16 if self.is_cuda else None, self.grad, self.m, self.v, self.opt, fresh_src=self.base)
        self.engine.roll_fp8_scales()
        return loss

    def enable_fused_first_forward(self, flags_address: int, target: torch.Tensor, chunks_per_rank: int, world: int) -> None:
        """Path (b): after a pushed base the FIRST step runs a variant of the step graph whose forward GEMMs acquire the shard
        owners' base flags inside the kernel (TransformerEngine.configure_ready) -- no wait kernel between round and step."""
        self.engine.configure_ready(flags_address, target, chunks_per_rank, world)
        self.ready_capable = self.is_cuda and not self.engine.fp8

    def step(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimizer step on a [B,T] batch (host-pinned or device).  Returns the device-resident mean loss."""
        if isinstance(input_ids, dict):  # miner: labels = input_ids (PAD not masked), attention_mask -> padding-masked attention
            self.engine.set_batch(input_ids["input_ids"], None, input_ids.get("attention_mask"), input_ids.get("kv_len"))
        else:
            self.engine.set_batch(input_ids, labels)
        assert self.engine.n_rows == self.batch, "training batches must fill the static batch"
        ready = bool(self.ready_capable and self.master_stale)  # first step after a pushed base: in-GEMM flag acquires
        if self.ready_capable:
            self.engine._ready_on = ready
        self._graph = self._graphs.get(ready)
        if not self.use_graph:
            c0 = ops.launch_count()
            loss = self._step_body()
            self.launches_per_step = ops.launch_count() - c0
        else:
            if self._graph is None:
                c0 = ops.launch_count()
                # snapshot state, run the warm-up eagerly on a side stream (required before capture), restore, capture.
                # The warm-up mutates optimizer state; replaying the graph will redo the step on restored state.
                snap = (self.master.clone(), self.m.clone(), self.v.clone(), self.opt.step.clone(), self.opt.host_step)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._step_body()
                torch.cuda.current_stream().wait_stream(s)
                self.launches_per_step = ops.launch_count() - c0
                self.master.copy_(snap[0]); self.m.copy_(snap[1]); self.v.copy_(snap[2]); self.opt.step.copy_(snap[3])
                self.opt.host_step = snap[4]
                ops.cast_copy(self.base if self.master_stale else self.master, self.p16)  # (stale master: theta == theta_base)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._step_body()
                self._graphs[ready] = self._graph
                self.opt.host_step = snap[4]
            self._graph.replay()
            self.opt.host_step += 1
            loss = self.engine.loss
        self.steps_done += 1
        self.master_stale = False
        return loss

    def sync_master(self) -> None:
        """After a pushed base (``master_stale``) theta lives in ``base`` only until the first optimizer step: materialise
        it for anyone who reads ``master`` in between (checkpoints, hashes, eval)."""
        if self.master_stale:
            self.master.copy_(self.base)
            self.master_stale = False

    def loss_and_grad(self, input_ids, labels: Optional[torch.Tensor] = None, zero_grad: bool = True) -> torch.Tensor:
        """Forward + backward at the CURRENT master/p16 without an optimizer step: grads land in ``self.grad``.
        (The averager's meta-learning needs dL/dtheta_bar: reference hivetrain/averaging_logic.py:502-511.)"""
        self.engine.set_batch(input_ids, labels)  # dict batches carry labels / attention_mask themselves
        md = self.meta_dropout
        if not (self.use_graph and zero_grad and self.engine.n_rows == self.batch):
            return self.engine.forward_backward(zero_grad, dropout=md)
        if self._lg_graph is None:  # same capture protocol as step(): eager warm-up on a side stream, then capture
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.engine.forward_backward(True, dropout=md)
            torch.cuda.current_stream().wait_stream(s)
            self._lg_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._lg_graph):
                self.engine.forward_backward(True, dropout=md)
        self._lg_graph.replay()
        return self.engine.loss

    @torch.no_grad()
    def eval_loss(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.engine.set_batch(input_ids, labels)
        return self.engine.forward_loss()

    # ------------------------------------------------------------------------------------------------------------
    def emit_delta(self, out: torch.Tensor, scales: Optional[torch.Tensor] = None, bad: Optional[torch.Tensor] = None) -> torch.Tensor:
        """delta = theta - theta_base into ``out`` (usually this rank's symmetric window); ``bad`` (int32[1]) is raised when
        the delta holds a NaN/Inf."""
        self.sync_master()
        return ops.delta_emit(self.master, self.base, out, scales, bad)

    def load_base(self, new_base: torch.Tensor, lr: Optional[float] = None, reset_optimizer: bool = True) -> None:
        """Adopt a new averaged base: theta = theta_base = new_base; optimizer re-created (moments dropped) with the
        post-pull learning rate, exactly as reference hivetrain/training_manager.py:365-378."""
        if new_base.data_ptr() != self.base.data_ptr():
            self.base.copy_(new_base)
        ops.round_reset(self.base, self.master, self.p16 if self.is_cuda else None, self.m, self.v, reset_optimizer)
        if reset_optimizer:
            self.opt.reset()
        if lr is not None:
            self.opt.set_lr(lr)

    # -- HF interoperability -------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path_or_state_dict, model=None, add_pad_token: bool = True, **kw) -> "Trainer":
        """Start from pretrained weights like the reference miner / validator / averager do (training_manager.py:39-46,
        neurons/validator.py:52-58): an HF checkpoint directory, or an HF-layout ``state_dict`` plus the ``model`` preset
        it belongs to.  The embedding table is grown by the ``[PAD]`` row when the checkpoint is one row short."""
        if isinstance(path_or_state_dict, str):
            cfg, _, flat = load_hf_checkpoint(path_or_state_dict, add_pad_token=add_pad_token)
            return cls(cfg, init_flat=flat, **kw)
        cfg = get_config(model) if isinstance(model, str) else model
        return cls(cfg, init_flat=from_hf_state_dict(cfg, path_or_state_dict), **kw)

    def hf_state_dict(self, which: str = "master") -> Dict[str, torch.Tensor]:
        """HF / reference-format state dict of ``master`` or ``base`` (loads into ``GPT2LMHeadModel`` as is)."""
        self.sync_master()
        return to_hf_state_dict(self.cfg, getattr(self, which).detach().float().cpu())

    def flat_from(self, blob) -> torch.Tensor:
        """Flat fp32 arena from a flat tensor / engine-layout dict / HF-layout dict (shape-screened)."""
        return pack_any(self.man, blob, self.cfg)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        self.sync_master()
        return {"master": self.master, "base": self.base, "m": self.m, "v": self.v, "step": self.opt.step,
                "hyper": self.opt.hyper}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.master.copy_(sd["master"]); self.base.copy_(sd["base"]); self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        self.opt.step.copy_(sd["step"]); self.opt.hyper[:sd["hyper"].numel()].copy_(sd["hyper"])
        self.opt.host_step = int(sd["step"])
        self.opt.host["lr"] = float(sd["hyper"][0])
        if self.is_cuda:
            ops.cast_copy(self.master, self.p16)
