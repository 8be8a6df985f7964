"""Multi-GPU correctness + bandwidth check of the peer-memory exchange (run under torchrun, one rank per GPU).

Verifies, for fp32 / bf16 / fp8 deltas:
  * pull form   (gather_average on rank 0)           == torch reference of  s*base + sum_i w_i delta_i
  * sharded form (reduce-scatter + all-gather kernel) == same, on EVERY rank
and times both against the NCCL all_gather + torch weighted-sum baseline (device-timed, max over ranks).
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import Manifest
from distributedtraining_b200.parallel.exchange import PeerExchange, _torch_weighted_avg
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed, max_over_ranks


class FakeTrainer:
    def __init__(self, master, base):
        self.master, self.base = master, base
    def emit_delta(self, out, scales=None):
        return ops.delta_emit(self.master, self.base, out, scales)


def main():
    rank, world, dev = init_distributed("nccl")
    numel_mb = int(os.environ.get("PEER_CHECK_MB", "64"))
    # a manifest with a few big and many small tensors
    n_big = numel_mb * (1 << 20) // 4 // 4
    man = Manifest([(f"big{i}", (n_big,), "normal", True) for i in range(4)] + [(f"small{i}", (777 + i,), "normal", False) for i in range(20)])
    n, P = man.total, len(man)
    g = torch.Generator(device="cuda").manual_seed(1234)           # identical base everywhere
    base = torch.randn(n, device=dev, generator=g)
    gl = torch.Generator(device="cuda").manual_seed(100 + rank)     # rank-specific "training"
    master = base + 0.01 * torch.randn(n, device=dev, generator=gl)
    w = (torch.rand(world, P, device=dev, generator=g) - 0.2)       # identical w everywhere (same generator state)
    out = {"world": world, "numel": n, "results": []}
    for dt in ["fp32", "bf16", "fp8"]:
        ex = PeerExchange(man, delta_dtype=dt)
        tr = FakeTrainer(master, base)
        # ---- reference: all_gather the DECODED deltas with NCCL and reduce with torch ----
        mine = torch.empty(n, dtype=ex.delta_dtype, device=dev)
        sc = torch.empty(n // 32, device=dev) if dt == "fp8" else None
        tr.emit_delta(mine, sc)
        dec = ops.dequant_fp8(mine, sc) if dt == "fp8" else mine.float()
        allg = torch.empty(world, n, device=dev)
        dist.all_gather_into_tensor(allg.view(-1), dec)
        ref = torch.empty(n, device=dev)
        _torch_weighted_avg(base, allg, w, man.tensor_ids(dev), ref)
        # ---- ours ----
        r = 1
        ex.publish_delta(tr, r)
        pull = torch.zeros(n, device=dev)
        if rank == 0:
            ex.gather_average(base, w, r, list(range(world)), pull)
        got = ex.sharded_average_broadcast(base, w, r, list(range(world))).clone()
        # pull-only round: reduce-scatter by pull + all-gather by pull fused with the reset
        class PT:
            is_cuda = True
        pt = PT()
        pt.base, pt.master = base.clone(), torch.zeros(n, device=dev)
        pt.p16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        pt.m, pt.v = torch.ones(n, device=dev), torch.ones(n, device=dev)
        per = ex.reduce_scatter_average(pt.base, w, r, list(range(world)))
        ex.all_gather_reset(pt, per)
        torch.cuda.synchronize()
        err_pr = max((pt.base - ref).abs().max().item(), (pt.master - ref).abs().max().item())
        err_pr16 = (pt.p16.float() - ref).abs().max().item()
        moments_cleared = bool(pt.m.abs().max().item() == 0 and pt.v.abs().max().item() == 0)
        torch.cuda.synchronize()
        ex.win.check_errors()
        # NVLS push round on multicast-bound windows: reduce-scatter by pull + multimem.st broadcast from the same kernel
        err_push = err_push16 = 0.0
        if ex.win.mc_ptr and world > 1:
            barrier_sync(dev)
            base_win = ex.win.local("base", torch.float32)[:n]
            base_win.copy_(base)
            barrier_sync(dev)
            d_, s_ = ex._delta_ptrs(r, list(range(world)))
            ex.push_average(base_win, d_, s_, w, r, {"fp32": 0, "bf16": 1, "fp8": 2}[dt],
                            wait_flags=[ex.win.flag_ptr(ex.F_DELTA + q) for q in range(world)])
            ex.wait_base()
            torch.cuda.synchronize()
            barrier_sync(dev)
            err_push = (base_win - ref).abs().max().item()
            err_push16 = (ex.win.local("base16", torch.bfloat16)[:n].float() - ref).abs().max().item()
            ex.win.check_errors()
        err_sh = (got - ref).abs().max().item()
        err_pull = (pull - ref).abs().max().item() if rank == 0 else 0.0
        b16 = ex.win.local("base16", torch.bfloat16)[:n].float()
        err_b16 = (b16 - ref).abs().max().item()
        scale = ref.abs().max().item()
        # ---- timing (3 warm, 5 timed; each iteration is a fresh round so flags advance) ----
        def timed(fn, iters=5):
            ts = []
            for it in range(3 + iters):
                barrier_sync(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(max_over_ranks(e0.elapsed_time(e1), dev))
            return sorted(ts)[len(ts) // 2]
        state = {"r": r}
        def ours():
            state["r"] += 1
            ex.publish_delta(tr, state["r"])
            ex.sharded_average_broadcast(base, w, state["r"], list(range(world)))
        def nccl():
            tr.emit_delta(mine, sc)
            dist.all_gather_into_tensor(allg.view(-1)[: world * n] if dt == "fp32" else allg.view(-1), dec if dt != "fp32" else mine)
            _torch_weighted_avg(base, allg, w, man.tensor_ids(dev), ref)
        t_ours = timed(ours)
        t_nccl = timed(nccl) if dt == "fp32" else None
        esz = {"fp32": 4, "bf16": 2, "fp8": 1}[dt]
        res = {"dtype": dt, "max_err_pull_round": err_pr, "max_err_pull_round_bf16": err_pr16, "moments_cleared": moments_cleared,
               "max_err_sharded": err_sh, "max_err_pull": err_pull, "max_err_push_round_nvls": err_push,
               "max_err_push_round_nvls_bf16": err_push16, "multicast": bool(ex.win.mc_ptr), "max_err_bf16_copy": err_b16, "ref_max": scale,
               "ms_fused_round": t_ours, "ms_nccl_allgather_torch_avg": t_nccl,
               "delta_bytes": n * esz, "nvlink_in_bytes_per_rank": (world - 1) * n * esz // world}
        tol = {"fp32": 1e-5, "bf16": 1e-5, "fp8": 1e-5}[dt] * max(scale, 1.0)
        res["ok"] = bool(err_sh <= tol and err_pull <= tol and err_push <= tol and err_push16 <= 1e-2 * max(scale, 1.0)
                         and err_b16 <= 1e-2 * max(scale, 1.0) and err_pr <= tol
                         and err_pr16 <= 1e-2 * max(scale, 1.0) and moments_cleared)
        out["results"].append(res)
        ex.win.close()
    allok = torch.tensor([int(all(r["ok"] for r in out["results"]))], device=dev)
    dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(allok.item())
    if rank == 0:
        print("PEER_CHECK " + json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/peer_check_n{world}.json", "w"), indent=1)
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
