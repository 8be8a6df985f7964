"""Reference arm of ``bench.py --impl reference``: the UNMODIFIED reference miner loop.

``baseline/_ref/hivetrain`` is a ``pip install --no-deps --target baseline/_ref`` of ``/root/reference``.
This harness runs the reference's own public API -- ``hivetrain.training_manager.DeltaLoop(...).train(epochs)`` with HF
``GPT2LMHeadModel`` in fp32 eager, ``AdamW``, per-step ``loss.item()`` and ``torch.save`` of ``weight_diff.pt`` -- on the
same metric/config as our arm (GPT-2-small + [PAD] = 124 440 576 params, B x 64 tokens per step, synthetic tokens,
random init).  None of this repository's models, kernels or engine are on that path.

What is shimmed (third-party packages that cannot be installed offline, NOT reference code):
  * ``bittensor``  -> baseline/shims/bittensor  (config/wallet/subtensor/logging stand-ins; the reference connects to
                      the chain at import time: hivetrain/training_manager.py:22-24)
  * ``mlflow``     -> baseline/shims/mlflow     (gated off in the reference anyway)
  * ``transformers.AdamW`` (removed upstream)   -> ``torch.optim.AdamW`` with the old defaults (eps 1e-6, wd 0)
  * ``huggingface_hub.Repository/HfFolder`` (removed upstream) -> dummy names so the import succeeds; the hub itself is
    replaced through the reference's own dependency-injection seam: ``DeltaLoop(hf_manager=<local-disk stub>)``
  * pretrained ``openai-community/gpt2`` files  -> a local directory with a random-init GPT2LMHeadModel + a 50 257-entry
    tokenizer, passed as ``model_name`` (no network).
One miner process per GPU (the reference has no multi-GPU mode: miners are independent replicas; its averager is a separate,
asynchronous process and is not on the miner's critical path).  One delta push happens inside the timed region.
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _unavailable(why: str) -> dict:
    return {"impl": "reference", "unavailable": why}


def _make_local_gpt2(path: str) -> None:
    """Random-init GPT-2-small + a tokenizer with the GPT-2 vocabulary size, saved like a HF checkpoint directory."""
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace
    from transformers import GPT2Config, GPT2LMHeadModel, PreTrainedTokenizerFast

    if os.path.exists(os.path.join(path, "config.json")) and os.path.exists(os.path.join(path, "tokenizer.json")):
        return
    os.makedirs(path, exist_ok=True)
    vocab = {f"tok{i}": i for i in range(50256)}
    vocab["<|endoftext|>"] = 50256
    tok = Tokenizer(WordLevel(vocab, unk_token="<|endoftext|>"))
    tok.pre_tokenizer = Whitespace()
    PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|endoftext|>", unk_token="<|endoftext|>").save_pretrained(path)
    import torch
    torch.manual_seed(0)
    GPT2LMHeadModel(GPT2Config()).save_pretrained(path)


class DiskHubStub:
    """Local-disk hub with exactly the methods ``DeltaLoop.train`` calls (reference hivetrain/training_manager.py:361-423)."""

    def __init__(self, root: str):
        self.root = root
        self.model_repo_id = os.path.join(root, "averaged")
        self.grad_dir = os.path.join(root, "gradients")
        os.makedirs(self.model_repo_id, exist_ok=True)
        os.makedirs(self.grad_dir, exist_ok=True)
        self.pushes = 0

    def check_for_new_submissions(self, repo_id=None) -> bool:
        return False  # no averager process in this arm

    def pull_latest_model(self):
        pass

    def update_model(self, model, model_file_name="averaged_model.pt"):
        return model

    def get_local_gradient_directory(self):
        return self.grad_dir

    def push_changes(self, file_to_send):
        assert os.path.exists(os.path.join(self.grad_dir, file_to_send))
        self.pushes += 1


def run_reference(args) -> dict:
    if not os.path.isdir(os.path.join(REF, "hivetrain")):
        return _unavailable("baseline/_ref/hivetrain missing: run pip install --no-deps --target baseline/_ref on a copy of /root/reference")
    try:
        import torch
        import torch.distributed as dist
    except Exception as e:  # pragma: no cover
        return _unavailable(f"torch import failed: {e}")
    if not torch.cuda.is_available():
        return _unavailable("no CUDA device")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    if args.gpus > 1 and world == 1:
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.join(os.path.dirname(HERE), "bench.py")] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B, T, K, W = args.batch_size, args.seq_len, args.steps, max(args.warmup, 3)
    try:
        # ---- third-party shims, then the unmodified reference ----
        sys.path.insert(0, os.path.join(HERE, "shims"))
        sys.path.insert(0, REF)
        import transformers

        class _AdamW(torch.optim.AdamW):
            def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True, **kw):
                super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

        from transformers import AutoModelForCausalLM, AutoTokenizer, Trainer, TrainingArguments  # noqa: F401  settle the lazy module
        sys.modules["transformers"].AdamW = _AdamW  # removed upstream; the reference imports it (training_manager.py:10)
        import huggingface_hub
        for name in ("Repository", "HfFolder"):
            if not hasattr(huggingface_hub, name):
                setattr(huggingface_hub, name, type(name, (), {}))
        try:
            import dotenv  # noqa: F401
        except Exception:
            import types
            sys.modules["dotenv"] = types.SimpleNamespace(load_dotenv=lambda *a, **k: None)
        work = os.environ.get("DTB200_REF_WORKDIR", "/tmp/dtb200_reference_arm")
        model_dir = os.path.join(work, "gpt2-random-init")
        if rank == 0:
            _make_local_gpt2(model_dir)
        if world > 1:
            dist.barrier(device_ids=[local])
        saved_argv = sys.argv
        sys.argv = [sys.argv[0], "--batch_size", str(B)]  # the reference parses sys.argv at import time
        cwd = os.getcwd()
        os.chdir(REF)  # the reference reads ./template/__init__.py for its version at import (utils/mlflow_utils.py:72-82)
        try:
            import hivetrain  # noqa: F401  (unmodified reference; connects to the shimmed "chain" on import)
            from hivetrain.training_manager import DeltaLoop
        finally:
            os.chdir(cwd)
            sys.argv = saved_argv
    except SystemExit:
        raise
    except Exception as e:
        return _unavailable(f"reference import failed: {type(e).__name__}: {e}"[:300])

    # ---- data: same shape as the reference's collate output (dict of [B,T] int64 CPU tensors, right-padded) ----
    g = torch.Generator().manual_seed(2000 + rank)
    pool = []
    for _ in range(8):
        u = torch.rand(B, T, generator=g)
        ids = (50257 ** u - 1).long().clamp_(0, 50256)
        lens = ((1.0 - 0.25 * torch.rand(B, generator=g)) * T).long().clamp_(1, T)
        mask = torch.arange(T)[None, :] < lens[:, None]
        ids = torch.where(mask, ids, torch.full_like(ids, 50257))
        # PINNED host memory: every step's ``.to(device)`` in the reference loop (training_manager.py:381-383) is then a real
        # asynchronous PCIe transfer, as the end-to-end rule of the bench contract asks
        pool.append({k: v.pin_memory() for k, v in {"input_ids": ids, "attention_mask": mask.long(), "labels": ids.clone()}.items()})

    hub = DiskHubStub(os.path.join(work, f"hub_rank{rank}"))
    events = {}

    class TimedLoader:
        """W warm-up batches, then TWO separately timed regions of exactly K batches each (``value`` and ``e2e``: the
        reference has a single mode -- pinned-host batch -> ``.to(device)`` x3 -> step -> ``loss.item()`` -- so both regions
        run the same end-to-end loop, measured independently).  ``send_interval`` is flipped so that ONE delta push
        (torch.save of weight_diff.pt + hub push) happens right after the last step of EACH region, inside it."""

        def __init__(self):
            self.loop = None

        def _mark(self, name):
            if world > 1:
                dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
            events[name] = torch.cuda.Event(enable_timing=True)
            events[name].record()

        def __iter__(self):
            n_regions = 1 if args.no_e2e else 2
            for i in range(W + n_regions * K):
                if i == W:
                    self._mark("e0")
                if i == W + K:  # the push of region 1 ran after step W+K-1: close region 1, open region 2
                    self._mark("e1")
                    self.loop.send_interval = 1e12
                    self._mark("f0")
                if i in (W + K - 1, W + 2 * K - 1):
                    self.loop.send_interval = 0.0
                yield pool[i % len(pool)]

    loader = TimedLoader()
    loop = DeltaLoop(device, model_dir, loader, send_interval=1e12, learning_rate=args.lr, hf_manager=hub)
    loop.check_update_interval = 1e12
    loader.loop = loop
    nparams = sum(p.numel() for p in loop.model.parameters())
    from bench import ClockSampler
    sampler = ClockSampler(local)
    sampler.start()
    loop.train(1)
    loader._mark("f1" if not args.no_e2e else "e1")
    clocks = sampler.stop()

    def mx(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = mx(events["e0"].elapsed_time(events["e1"]))
    ms_e2e = mx(events["f0"].elapsed_time(events["f1"])) if not args.no_e2e else None
    if world > 1:
        dist.destroy_process_group()
    tokens = K * B * T * world
    from bench import GPT2_SMALL_DESC, shared_config
    out = {
        "impl": "reference", "metric": "tokens/sec (GPT-2-small local-SGD training, all miners; per-miner = value / n_gpus)",
        "value": tokens / ms * 1e3, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (the reference's only precision)",
        "data": "synthetic tokens (Zipf ids, right-padded), random-init weights",
        "config": shared_config(GPT2_SMALL_DESC, B, T, world),
        "detail": {"model": f"transformers.GPT2LMHeadModel + [PAD] ({nparams} params), fp32 eager",
                   "parallelism": f"{world} independent reference miners (upstream has no multi-GPU mode; no averager in this arm)",
                   "delta_pushes": hub.pushes, "api": "hivetrain.training_manager.DeltaLoop.train (unmodified, baseline/_ref)"},
        "clocks": clocks, "gpu_launches": 0,
    }
    if ms_e2e is not None:
        # per step: 3 pinned int64 [B,T] tensors host->device (input_ids, attention_mask, labels), loss.item() device->host
        out["e2e"] = {"value": tokens / ms_e2e * 1e3, "unit": "tokens/s", "ms_per_step": ms_e2e / K,
                      "h2d_bytes_per_step": 3 * B * T * 8, "d2h_bytes_per_step": 4,
                      "api": "hivetrain.training_manager.DeltaLoop.train (separately timed second region of K steps)"}
    return out if rank == 0 else {}
