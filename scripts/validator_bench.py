"""BASELINE.json config 3: N-1 miners + 1 validator rank; the validator scores every miner's delta.

Compares, per miner (device-timed on the validator rank):
  fused    -- eval GEMMs read W (local) and dW_i (miner's PEER window) as two accumulating tcgen05 passes; theta_base+delta_i
              is never materialised (small tensors go through the chunk-restricted fused apply kernel)
  applied  -- ONE fused kernel materialises base+delta_i (peer read) into master+bf16, then a plain eval forward
  nccl     -- the reference-style path: NCCL broadcast of delta_i to the validator, torch add, torch cast, eval forward
and checks that all three give the same losses.   torchrun --nproc-per-node N scripts/validator_bench.py [--model gpt2-medium]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.btt_connector import BittensorNetwork, MemoryLedger
from distributedtraining_b200.chain_manager import ChainMultiAddressStore
from distributedtraining_b200.config import Configurator
from distributedtraining_b200.data import SyntheticTokens
from distributedtraining_b200.hf_manager import HFManager
from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.parallel.exchange import PeerExchange
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed
from distributedtraining_b200.validation_logic import DeltaValidator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2-medium")
    ap.add_argument("--miner-steps", type=int, default=10)
    ap.add_argument("--eval-batches", type=int, default=13)   # 100 texts / batch 8 (reference neurons/validator.py:49,98)
    ap.add_argument("--eval-batch", type=int, default=8)
    ap.add_argument("--eval-seq", type=int, default=512)
    ap.add_argument("--eval-rows", type=int, default=52, help="rows per eval batch of the collective validator (100 texts -> 2 batches)")
    a = ap.parse_args()
    rank, world, dev = init_distributed("nccl")
    vrank = world - 1
    miners = list(range(world - 1)) if world > 1 else [0]
    is_val = rank == vrank
    B, T = (a.eval_batch, a.eval_seq) if is_val else (32, 64)
    tr = Trainer(a.model, device=dev, batch=B, seq=T, lr=5e-4, seed=0, use_graph=False)
    ex = PeerExchange(tr.man, delta_dtype="bf16")
    V = tr.cfg.vocab_size
    if (not is_val) or world == 1:
        data = SyntheticTokens(B, T, V, seed=rank, device=str(dev), pool=4)
        if world == 1:
            mtr = Trainer(a.model, device=dev, batch=32, seq=64, lr=5e-4, seed=0, use_graph=False)
            data = SyntheticTokens(32, 64, V, seed=0, device=str(dev), pool=4)
        else:
            mtr = tr
        for i in range(a.miner_steps):
            mtr.step(data.pool[i % 4]["input_ids"])
        ex.publish_delta(mtr, 1)  # to every rank: all ranks of the box score deltas together in the collective mode
    barrier_sync(dev)
    out = None
    # ---- collective mode: N miners + the base = N+1 jobs spread over ALL ranks, large eval batches, CUDA-graph forward ----
    from distributedtraining_b200.validation_logic import CollectiveDeltaValidator
    cfg0 = Configurator.combine_configs([])
    cfg0.wallet.hotkey = f"rank{rank}"
    cfg0.neuron.epoch_length = 0
    BittensorNetwork.initialize(cfg0, ignore_regs=True, ledger=MemoryLedger(), hotkeys=[f"rank{r}" for r in range(world)])
    vloader = list(SyntheticTokens(a.eval_batch, a.eval_seq, V, seed=4242, device=str(dev), pool=a.eval_batches, steps=a.eval_batches))
    cval = CollectiveDeltaValidator(dev, tr, vloader, BittensorNetwork, ex, miners, validator_rank=vrank, eval_rows=a.eval_rows)
    coll = {}
    for rep in range(3):
        barrier_sync(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cval.validate_and_score(round=1)
        e1.record()
        torch.cuda.synchronize()
        from distributedtraining_b200.parallel.launch import max_over_ranks
        ms = max_over_ranks(e0.elapsed_time(e1), dev)
        coll = {"ms_per_miner": ms / len(miners), "ms_round": ms, "losses": [cval.losses[f"rank{r}"] for r in miners],
                "base_loss": cval.base_loss, "scores": [cval.normalized_scores[f"rank{r}"] for r in miners],
                "eval_rows_per_batch": a.eval_rows, "jobs_per_rank": -(-(len(miners) + 1) // world)}
    # the same large-batch graph-captured scoring on the validator rank ALONE (no help from the other ranks)
    solo = {}
    if is_val and world > 1:
        sv = CollectiveDeltaValidator(dev, tr, vloader, BittensorNetwork, ex, miners, validator_rank=vrank, eval_rows=a.eval_rows)
        sv.world, sv.rank = 1, 0
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sv.validate_and_score(round=1, base_changed=False)  # like the other single-rank modes: base loss from the constructor
            e1.record()
            torch.cuda.synchronize()
            solo = {"ms_per_miner": e0.elapsed_time(e1) / len(miners), "losses": [sv.losses[f"rank{r}"] for r in miners]}
    barrier_sync(dev)
    if is_val:
        ex.win.wait(ex.F_DELTA, 1, miners)
        torch.cuda.synchronize()
        cfg = Configurator.combine_configs([])
        cfg.wallet.hotkey = f"rank{rank}"
        cfg.neuron.epoch_length = 0
        hot = [f"rank{r}" for r in miners]
        BittensorNetwork.initialize(cfg, ignore_regs=True, ledger=MemoryLedger(), hotkeys=hot + ([f"rank{rank}"] if world > 1 else []))
        BittensorNetwork.last_set_block -= 10000
        for r in miners:
            BittensorNetwork.ledger.put(f"commit/1/rank{r}", f"peer://{r}")
        chain = ChainMultiAddressStore(BittensorNetwork.ledger, 1, BittensorNetwork.wallet)
        hf = HFManager(local_dir="/tmp/dtb_val", averaged_model_repo_id="avg", exchange=ex, manifest=tr.man)
        loader = list(SyntheticTokens(B, T, V, seed=4242, device=str(dev), pool=a.eval_batches, steps=a.eval_batches))
        res = {}
        for mode in ("applied", "fused"):
            val = DeltaValidator(dev, tr, None, loader, BittensorNetwork, hf, chain_manager=chain, fused_eval=(mode == "fused"))
            val.validate_and_score()  # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            val.validate_and_score()
            e1.record(); torch.cuda.synchronize()
            res[mode] = {"ms_per_miner": e0.elapsed_time(e1) / len(miners), "losses": [val.losses[h] for h in hot],
                         "base_loss": val.base_loss, "scores": [val.normalized_scores[h] for h in hot]}
        out = {"model": a.model, "world": world, "miners": len(miners), "eval_tokens_per_miner": a.eval_batches * B * T,
               "delta_bytes_bf16": tr.man.total * 2, **res, "collective_all_ranks": coll, "large_batch_graph_one_rank": solo}
        out["max_loss_diff_collective_vs_applied"] = max(abs(x - y) for x, y in zip(coll["losses"], res["applied"]["losses"]))
    # ---- NCCL-style baseline: broadcast each delta to the validator, torch add + cast, eval ----
    if world > 1:
        n = tr.man.total
        buf = torch.empty(n, dtype=torch.bfloat16, device=dev)
        losses = []
        times = []
        for rep in range(2):
            barrier_sync(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            losses = []
            for r in miners:
                if rank == r:
                    buf.copy_(ex.delta_buf(1)[:n])
                dist.broadcast(buf, src=r)
                if is_val:
                    tr.master.copy_(tr.base + buf.float())
                    tr.p16.copy_(tr.master)
                    acc = None
                    for b in loader:
                        l = tr.eval_loss(b["input_ids"]).double() * B
                        acc = l if acc is None else acc + l
                    losses.append(float(acc) / (len(loader) * B))
                    tr.master.copy_(tr.base); tr.p16.copy_(tr.master)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / len(miners))
        if is_val:
            out["nccl"] = {"ms_per_miner": times[-1], "losses": losses}
    if is_val:
        f, ap_ = out["fused"]["losses"], out["applied"]["losses"]
        out["max_loss_diff_fused_vs_applied"] = max(abs(x - y) for x, y in zip(f, ap_))
        if "nccl" in out:
            out["max_loss_diff_fused_vs_nccl"] = max(abs(x - y) for x, y in zip(f, out["nccl"]["losses"]))
        print("VALBENCH " + json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/validator_bench_n{world}.json", "w"), indent=1)
    if dist.is_initialized():
        dist.barrier(device_ids=[dev.index])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
