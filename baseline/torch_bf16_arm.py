"""``bench.py --impl torch-bf16``: the strongest STOCK-LIBRARY formulation of the same job, as an honest secondary baseline.

The unmodified reference arm is fp32 eager with no averager (that is what upstream ships); most of the gap to it is precision
and library choice.  This arm removes that excuse: HF ``GPT2LMHeadModel`` (random init, vocab 50258) under bf16 autocast with
SDPA attention, ``torch.optim.AdamW(fused=True)`` (lr 5e-4, eps 1e-6, wd 0 -- the reference's hyper-parameters), the same
batches (input_ids + attention_mask + labels = input_ids), and every ``--local-steps`` steps a delta-averaging round done with
NCCL: all_gather of the fp32 deltas + torch weighted sum + base add, then optimizer re-creation at lr 5e-5 (reference
semantics).  None of this repository's kernels, engine or exchange code is on this path.
"""
from __future__ import annotations

import os
import sys


def run_torch_bf16(args) -> dict:
    import torch
    import torch.distributed as dist
    from transformers import GPT2Config, GPT2LMHeadModel

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import GPT2_SMALL_DESC, ClockSampler, shared_config

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    B, T, K, W = args.batch_size, args.seq_len, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)  # same theta_base on every rank
    cfg = GPT2Config(vocab_size=50258)
    cfg._attn_implementation = "sdpa"
    model = GPT2LMHeadModel(cfg).to(device).train()
    params = [p for p in model.parameters() if p.requires_grad]
    nparams = sum(p.numel() for p in params)

    def make_opt(lr):
        return torch.optim.AdamW(params, lr=lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, fused=True)

    state = {"opt": make_opt(args.lr)}
    base = [p.detach().clone() for p in params]
    flat_n = sum(p.numel() for p in params)
    gathered = torch.empty(world, flat_n, device=device) if world > 1 else None
    mine = torch.empty(flat_n, device=device)
    w = torch.full((world,), 1.0 / world, device=device)

    def round_():
        """delta = theta - base -> NCCL all_gather -> base + sum_i w_i delta_i -> theta = base = new; AdamW re-created."""
        with torch.no_grad():
            off = 0
            for p, b in zip(params, base):
                mine[off:off + p.numel()].copy_((p - b).reshape(-1))
                off += p.numel()
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(-1), mine)
                avg = (gathered * w[:, None]).sum(0)
            else:
                avg = mine
            off = 0
            for p, b in zip(params, base):
                b.add_(avg[off:off + p.numel()].view_as(b))
                p.copy_(b)
                off += p.numel()
        state["opt"] = make_opt(5e-5)

    def step(batch):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["input_ids"]).loss
        loss.backward()
        state["opt"].step()
        state["opt"].zero_grad(set_to_none=True)
        return loss.detach()

    def make_pool(seed, dev, pin):
        g = torch.Generator().manual_seed(seed)
        pool = []
        for _ in range(8):
            u = torch.rand(B, T, generator=g)
            ids = (50257 ** u - 1).long().clamp_(0, 50256)
            lens = ((1.0 - 0.25 * torch.rand(B, generator=g)) * T).long().clamp_(1, T)
            mask = torch.arange(T)[None, :] < lens[:, None]
            ids = torch.where(mask, ids, torch.full_like(ids, 50257))
            b = {"input_ids": ids, "attention_mask": mask.long()}
            pool.append({k: (v.to(dev) if dev is not None else (v.pin_memory() if pin else v)) for k, v in b.items()})
        return pool

    dev_pool = make_pool(1000 + rank, device, False)
    host_pool = make_pool(2000 + rank, None, True)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def run(n, pool, host):
        did = False
        last = None
        for i in range(n):
            b = pool[i % len(pool)]
            if host:
                b = {k: v.to(device, non_blocking=True) for k, v in b.items()}
            last = step(b)
            if host:
                host_loss.copy_(last, non_blocking=True)
            if (i + 1) % args.local_steps == 0:
                round_()
                did = True
        if not did:
            round_()
        return last

    host_loss = torch.zeros((), dtype=torch.float32).pin_memory()
    run(W, dev_pool, False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run(K, dev_pool, False)
    e1.record()
    barrier()
    clocks = sampler.stop()

    def mx(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = mx(e0.elapsed_time(e1))
    tokens = K * B * T * world
    out = {"impl": "torch-bf16", "metric": "tokens/sec (GPT-2-small local-SGD training, all miners; per-miner = value / n_gpus)",
           "value": tokens / ms * 1e3, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (autocast; fp32 master weights)",
           "data": "synthetic tokens (Zipf ids, right-padded), random-init weights",
           "config": shared_config(GPT2_SMALL_DESC, B, T, world),
           "detail": {"model": f"transformers.GPT2LMHeadModel ({nparams} params), sdpa attention, bf16 autocast",
                      "optimizer": "torch.optim.AdamW(fused=True)", "round": "NCCL all_gather of fp32 deltas + torch weighted sum",
                      "local_steps": args.local_steps},
           "clocks": clocks, "gpu_launches": 0}
    if not args.no_e2e:
        barrier()
        e0.record()
        last = run(K, host_pool, True)
        e1.record()
        barrier()
        ms2 = mx(e0.elapsed_time(e1))
        out["e2e"] = {"value": tokens / ms2 * 1e3, "unit": "tokens/s", "ms_per_step": ms2 / K, "h2d_bytes_per_step": 2 * B * T * 8,
                      "d2h_bytes_per_step": 4, "api": "transformers.GPT2LMHeadModel + torch.optim.AdamW", "last_loss": float(host_loss)}
    if world > 1:
        dist.destroy_process_group()
    return out if rank == 0 else {}
