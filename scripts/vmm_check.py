"""VMM + multicast windows (csrc/symm_runtime.cu): unicast peer mapping, multimem.st landing on every rank, multimem.ld_reduce
summing over ranks -- on the PRODUCT windows (PeerExchange), not on a library-allocated side arena."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import Manifest
from distributedtraining_b200.parallel.exchange import PeerExchange
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed, max_over_ranks


def main():
    rank, world, dev = init_distributed("nccl")
    mb = int(os.environ.get("VMM_CHECK_MB", "64"))
    n = mb * (1 << 20) // 4
    n = n // (4 * world * 32) * (4 * world * 32)
    man = Manifest([("big", (n,), "normal", True)])
    ex = PeerExchange(man, delta_dtype="fp32")
    win = ex.win
    out = {"world": world, "backing": win.backing, "multicast": bool(win.mc_ptr), "numel": n}
    # unicast: every rank writes its rank id into its delta0 region, peers read it
    win.local("delta0", torch.float32)[:n].fill_(float(rank + 1))
    barrier_sync(dev)
    ok_uc = all(float(win.peer("delta0", r, torch.float32)[:n].mean()) == float(r + 1) for r in range(world))
    out["unicast_ok"] = bool(ok_uc)
    ok_mc = True
    if win.mc_ptr and world > 1:
        base = torch.full((n,), 0.5, device=dev)
        per4 = n // 4 // world
        lo4, hi4 = rank * per4, (rank + 1) * per4
        barrier_sync(dev)
        ts = []
        for it in range(6):
            barrier_sync(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            # base region of every rank <- 0.5 + 1.0 * sum_r delta_r over my shard (in-switch reduce + multicast store)
            ops.nvls_avg(win.mc("delta0"), win.mc("base"), base, lo4, hi4, 1.0, 1.0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1), dev))
        barrier_sync(dev)
        want = 0.5 + sum(range(1, world + 1))
        got = win.local("base", torch.float32)[:n]
        ok_mc = bool((got == want).all())
        out["nvls_ms"] = sorted(ts[2:])[len(ts[2:]) // 2]
        out["nvls_GBps_reduced_in_per_rank"] = n * 4 / world / (out["nvls_ms"] * 1e-3) / 1e9
    out["multicast_ok"] = ok_mc
    allok = torch.tensor([int(ok_uc and ok_mc)], device=dev)
    if world > 1:
        dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(allok.item())
    if rank == 0:
        print("VMM_CHECK " + json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/vmm_check_n{world}.json", "w"), indent=1)
    if dist.is_initialized():
        dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
