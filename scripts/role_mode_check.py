"""Role mode on the peer plane with the averager at a NON-ZERO rank (regression of the round-1 advisor finding: the base-round
flag of the averager lives in slot F_BASE + <averager rank>; reading slot 0 made miners never see a new base).

rank 0 = miner (Trainer + HFManager.push_changes / check_for_new_submissions / pull / update_model), rank 1 = averager
(ParameterizedAverager: cache_params_locally -> meta_learning -> push_to_hf_hub), lock-stepped by barriers; two rounds.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.averaging_logic import ParameterizedAverager
from distributedtraining_b200.btt_connector import BittensorNetwork, MemoryLedger
from distributedtraining_b200.chain_manager import ChainMultiAddressStore
from distributedtraining_b200.config import Configurator
from distributedtraining_b200.data import SyntheticTokens
from distributedtraining_b200.hf_manager import HFManager
from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.parallel.exchange import PeerExchange
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed


def main():
    rank, world, dev = init_distributed("nccl")
    assert world >= 2
    AVG = world - 1                      # the averager is the LAST rank, miners are 0 .. world-2
    tr = Trainer("gpt2-tiny", device=dev, batch=8, seq=64, lr=1e-3, seed=0, dropout_seed=rank, use_graph=False)
    ex = PeerExchange(tr.man)
    cfg = Configurator.combine_configs([])
    cfg.wallet.hotkey = f"rank{rank}"
    miners = list(range(world - 1))
    BittensorNetwork.initialize(cfg, ignore_regs=True, ledger=MemoryLedger(), hotkeys=[f"rank{r}" for r in miners])
    for r in miners:
        BittensorNetwork.ledger.put(f"commit/{cfg.netuid}/rank{r}", f"peer://{r}")
    chain = ChainMultiAddressStore(BittensorNetwork.ledger, cfg.netuid, BittensorNetwork.wallet)
    hf = HFManager(local_dir="/tmp/dtb_role", my_repo_id=f"peer://{rank}" if rank != AVG else None, averaged_model_repo_id="avg",
                   exchange=ex, manifest=tr.man, model_config=tr.cfg)
    V = tr.cfg.vocab_size
    data = SyntheticTokens(8, 64, V, pad_id=V - 1, seed=10 + rank, device=str(dev), pool=4)
    val = list(SyntheticTokens(4, 64, V, pad_id=V - 1, seed=7, device=str(dev), pool=2, steps=2))
    avg = ParameterizedAverager(tr, dev, hf_manager=hf, local_dir="/tmp/dtb_role/model", chain_manager=chain,
                                bittensor_network=BittensorNetwork, fresh_only=True) if rank == AVG else None
    res = {"rank": rank, "saw_new_base": [], "base_matches_averager": []}
    for rnd in range(1, 3):
        if rank != AVG:
            for i in range(3):
                tr.step(data.pool[i % 4])
            hf.push_changes("weight_diff.pt", trainer=tr)
        barrier_sync(dev)
        if rank == AVG:
            n = avg.cache_params_locally()
            assert n == len(miners), n
            avg.weights = None
            avg.meta_learning(val, 1, 0.01)
            avg._adopt_as_base()
            avg.push_to_hf_hub()
        barrier_sync(dev)
        if rank != AVG:
            seen = hf.check_for_new_submissions(hf.model_repo_id)
            res["saw_new_base"].append(bool(seen))
            if seen:
                hf.pull_latest_model()
                hf.update_model(tr, lr=5e-5)
        torch.cuda.synchronize()
        sums = [None] * world
        dist.all_gather_object(sums, ops.checksum(tr.base))
        res["base_matches_averager"].append(sums[rank] == sums[AVG])
        barrier_sync(dev)
    res["published_round"] = hf._published_round if rank == AVG else None
    allres = [None] * world
    dist.all_gather_object(allres, res)
    ok = all(all(r["saw_new_base"]) and all(r["base_matches_averager"]) for r in allres if r["rank"] != AVG)
    ok = ok and allres[AVG]["published_round"] == 2
    if rank == 0:
        print("ROLE_CHECK " + json.dumps({"ok": ok, "averager_rank": AVG, "ranks": allres}), flush=True)
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
