"""Distributed learned mixer: correctness + timing (run under torchrun, one rank per GPU; also works on 1 GPU).

Correctness (``--model gpt2-tiny`` in the tests, any model otherwise), after K sequential meta-steps from w = 1/N:
  * peer back-end (csrc/meta_avg.cu kernels over the symmetric windows)      -- the product
  * collective back-end (the same step sequence with NCCL all_reduce)         -- the baseline
  * one-rank sequential oracle (the reference loop on a single GPU that pulled all N deltas: ops.weighted_avg / multi_dot)
must agree on w[N, P] and on the final average; w must be bit-identical across ranks.

Timing: device time per meta-step (CUDA events, max over ranks, median of the timed steps) for both validation-parallel
modes of the peer back-end, the collective back-end and the one-rank formulation of round 1 (pull of N full deltas +
single-rank multi-dot), plus the per-round costs (delta all-to-all, final average).

    torchrun --nproc-per-node 8 scripts/meta_check.py --model gpt2 --val-batch 8 --val-seq 512 --steps 12
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from distributedtraining_b200 import ops
from distributedtraining_b200.data import SyntheticTokens
from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.parallel.exchange import CollectiveExchange, PeerExchange
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed, max_over_ranks
from distributedtraining_b200.parallel.meta import DistributedMetaLearner


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2-tiny")
    ap.add_argument("--val-batch", type=int, default=8)
    ap.add_argument("--val-seq", type=int, default=64)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--delta-dtype", default="fp32")
    ap.add_argument("--skip-collective", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, dev = init_distributed("nccl")
    tr = Trainer(args.model, device=dev, batch=8, seq=64, lr=1e-3, seed=0, dropout_seed=rank, use_graph=False)
    V = tr.cfg.vocab_size
    g = torch.Generator().manual_seed(1000 + rank)
    for _ in range(args.train_steps):  # rank-specific data -> rank-specific delta
        tr.step(torch.randint(0, V - 1, (8, 64), dtype=torch.int32, generator=g).to(dev))
    Bv, Tv = args.val_batch, min(args.val_seq, tr.cfg.n_positions)
    val = [dict(b) for b in SyntheticTokens(Bv, Tv, V, pad_id=V - 1, seed=7, pool=3, device=str(dev)).pool]
    N, P = world, len(tr.man)
    ex = PeerExchange(tr.man, delta_dtype=args.delta_dtype)
    r = 1
    ex.publish_delta(tr, r)
    barrier_sync(dev)  # fetch_delta below is the non-blocking host API (None until the peer's flag is up): publish everywhere first
    miners = list(range(world))
    K = args.steps
    out = {"world": world, "model": args.model, "val_batch": [Bv, Tv], "steps": K, "delta_dtype": args.delta_dtype, "modes": {}}

    def time_steps(fn, n):
        ts = []
        for k in range(n):
            barrier_sync(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(k); e1.record()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1), dev))
        return ts

    # ---- one-rank sequential oracle: every rank pulls ALL deltas (round-1 formulation) and runs the reference loop ----
    full = [ex.fetch_delta(i, r).float().clone() for i in miners]
    ref = Trainer(tr.cfg, device=dev, batch=Bv, seq=Tv, seed=0, init_flat=tr.base.clone(), use_graph=False)
    w_ref = torch.full((N, P), 1.0 / N, device=dev)
    G = torch.empty(N, P, device=dev)

    def seq_step(k):
        ops.weighted_avg(ref.base, full, w_ref, ref.man, [ref.master], [ref.p16])
        ref.loss_and_grad(val[k % len(val)])
        ops.multi_dot(ref.grad, full, ref.base, ref.master, ref.man, G)
        w_ref.add_(G, alpha=-0.01)
    t_seq = time_steps(seq_step, K)
    want = torch.empty_like(ref.master)
    ops.weighted_avg(ref.base, full, w_ref, ref.man, [want])
    out["one_rank_local_deltas_ms_per_step"] = med(t_seq[2:]) if K > 3 else med(t_seq)
    w_scale = float((w_ref - 1.0 / N).abs().max())
    out["w_moved"] = w_scale

    def run_mode(name, exch, mode, push=False):
        bufs = exch.trainer_buffers() if push else None  # push mode: theta_bar lands in the window-resident bf16 copy by multimem.st
        t2 = Trainer(tr.cfg, device=dev, batch=8, seq=64, seed=0, init_flat=tr.base.clone(), use_graph=False, buffers=bufs)
        ml = DistributedMetaLearner(t2, exch, miners, val, meta_lr=0.01, mode=mode)
        barrier_sync(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if ml.peer:
            ml.begin_round(r)
        else:
            ml.begin_round(r, deltas=full)
        e1.record()
        torch.cuda.synchronize()
        t_prep = max_over_ranks(e0.elapsed_time(e1), dev)
        ts = time_steps(lambda k: ml.step(k), K)
        if ml.peer:
            e0.record(); ml.final_average_shard(r); e1.record()
            torch.cuda.synchronize()
            t_fin = max_over_ranks(e0.elapsed_time(e1), dev)
            sl = slice(ml.e0, ml.e1)
            final = ex.win.local("base", torch.float32)[:tr.man.total]
            err_base = float((final[sl] - want[sl]).abs().max()) if ml.e1 > ml.e0 else 0.0
            ex.win.check_errors()
        else:
            e0.record(); final = ml.final_average_full(torch.empty_like(t2.master)); e1.record()
            torch.cuda.synchronize()
            t_fin = max_over_ranks(e0.elapsed_time(e1), dev)
            err_base = float((final - want).abs().max())
        wsum = torch.tensor([float(ml.w.double().sum())], dtype=torch.float64, device=dev)
        lo, hi = wsum.clone(), wsum.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res = {"mode": ml.mode, "theta_bar_all_gather": ml.describe()["theta_bar_all_gather"], "rows_per_rank": ml.r1 - ml.r0, "ms_per_step": med(ts[2:]) if K > 3 else med(ts), "ms_steps": [round(x, 3) for x in ts],
               "ms_round_prepare_and_transpose": t_prep, "ms_final_average": t_fin,
               "w_err_vs_one_rank": float((ml.w - w_ref).abs().max()), "base_err_vs_one_rank": err_base,
               "w_identical_across_ranks": bool(lo.item() == hi.item()), "last_loss": float(ml.loss_acc[1])}
        res["ok"] = bool(res["w_err_vs_one_rank"] <= 0.05 * w_scale + 1e-6 and res["w_identical_across_ranks"]
                         and err_base <= 1e-4 + 0.05 * float(want.abs().max()) * 1e-3)
        out["modes"][name] = res
        del ml, t2
        torch.cuda.empty_cache()

    run_mode("peer_replicate", ex, "replicate")
    if world > 1:
        run_mode("peer_dp", ex, "dp")
        if ex.win.mc_ptr:
            run_mode("peer_replicate_push", ex, "replicate", push=True)
            run_mode("peer_dp_push", ex, "dp", push=True)
        if not args.skip_collective:
            cex = CollectiveExchange(tr.man)
            run_mode("nccl_replicate", cex, "replicate")
            run_mode("nccl_dp", cex, "dp")
    # ---- round-1 formulation timed for reference: one rank pulls the N full deltas over NVLink every step ----
    if world > 1:
        d_ptrs, s_ptrs = ex._delta_ptrs(r, miners)
        mode_id = {"fp32": 0, "bf16": 1, "fp8": 2}[args.delta_dtype]

        def old_step(k):
            if rank == 0:
                ex.gather_average(ref.base, w_ref, r, miners, ref.master, ref.p16, wait=False)
                ref.loss_and_grad(val[k % len(val)])
                ops.multi_dot(ref.grad, d_ptrs, ref.base, ref.master, ref.man, G, dscales=s_ptrs, mode=mode_id)
        t_old = time_steps(old_step, K)
        out["one_rank_peer_pull_ms_per_step"] = med(t_old[2:]) if K > 3 else med(t_old)
    ok = torch.tensor([int(all(m["ok"] for m in out["modes"].values()))], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(ok.item())
    if rank == 0:
        print("META_CHECK " + json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(args.out or f"gpurun_out/meta_check_n{world}_{args.model}.json", "w"), indent=1)
    if dist.is_initialized():
        dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
