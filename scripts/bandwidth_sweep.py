"""BASELINE.json config 5: delta all-gather + weighted-average bandwidth sweep (1 MB - 4 GB of fp32 delta per miner) at
N GPUs, fused peer kernel vs NCCL all_gather + torch weighted sum.  Run under torchrun; device-timed, max over ranks.

    torchrun --nproc-per-node N scripts/bandwidth_sweep.py [--sizes-mb 1,4,16,64,256,1024,4096] [--dtype fp32]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import Manifest
from distributedtraining_b200.parallel.exchange import PeerExchange, _torch_weighted_avg
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed, max_over_ranks

NVLINK_GBS = 770.0  # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)


class T:
    is_cuda = True

    def __init__(self, master, base):
        self.master, self.base = master, base
        self.p16 = torch.empty(master.numel(), dtype=torch.bfloat16, device=master.device)
        self.m = torch.zeros_like(master)
        self.v = torch.zeros_like(master)
    def emit_delta(self, out, scales=None):
        return ops.delta_emit(self.master, self.base, out, scales)


def timed(fn, dev, iters=5, warm=3):
    ts = []
    for it in range(warm + iters):
        barrier_sync(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            ts.append(max_over_ranks(e0.elapsed_time(e1), dev))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-mb", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--dtype", default="fp32")
    a = ap.parse_args()
    rank, world, dev = init_distributed("nccl")
    out = {"world": world, "dtype": a.dtype, "rows": []}
    esz = {"fp32": 4, "bf16": 2, "fp8": 1}[a.dtype]
    for mb in [int(x) for x in a.sizes_mb.split(",")]:
        n = mb * (1 << 20) // 4  # elements (sizes are quoted as fp32 delta bytes per miner)
        ntens = 16
        man = Manifest([(f"t{i}", (n // ntens,), "normal", True) for i in range(ntens)])
        n = man.total
        base = torch.randn(n, device=dev)
        master = base + 0.01 * torch.randn(n, device=dev)
        w = torch.full((world, len(man)), 1.0 / world, device=dev)
        ex = PeerExchange(man, delta_dtype=a.dtype, with_base16=True, with_meta=False)
        tr = T(master, base)
        st = {"r": 0}
        miners = list(range(world))
        def fused_sharded():
            st["r"] += 1
            ex.win.publish(ex.F_DELTA, st["r"])          # delta already resident: time the exchange + reduction only
            ex.sharded_average_broadcast(base, w, st["r"], miners)
        def fused_pull_round():  # reduce-scatter by pull + all-gather by pull fused with the base/optimizer reset
            st["r"] += 1
            ex.win.publish(ex.F_DELTA, st["r"])
            per = ex.reduce_scatter_average(base, w, st["r"], miners)
            ex.all_gather_reset(tr, per)
        pull_out = torch.empty(n, device=dev)
        def fused_pull():
            st["r"] += 1
            ex.win.publish(ex.F_DELTA, st["r"])
            if rank == 0:
                ex.gather_average(base, w, st["r"], miners, pull_out)
        tr.emit_delta(ex.delta_buf(0)[:n], ex.scale_buf(0)); tr.emit_delta(ex.delta_buf(1)[:n], ex.scale_buf(1))
        mine = torch.empty(n, dtype=torch.float32 if a.dtype != "bf16" else torch.bfloat16, device=dev)
        tr.emit_delta(mine)
        allg = torch.empty(world, n, dtype=mine.dtype, device=dev)
        ref = torch.empty(n, device=dev)
        tid = man.tensor_ids(dev)
        def nccl_full():
            dist.all_gather_into_tensor(allg.view(-1), mine)
            _torch_weighted_avg(base, allg, w, tid, ref)
        def nccl_allgather_only():
            dist.all_gather_into_tensor(allg.view(-1), mine)
        # NVLS push round: reduce-scatter by pull from the miners' windows + broadcast by multimem.st FROM THE SAME KERNEL into every
        # rank's (window-resident) fp32 base and bf16 copy, then the flag wait -- the product round when the windows are multicast-bound
        base_win = ex.win.local("base", torch.float32)[:n]
        base_win.copy_(base)
        mode_id = {"fp32": 0, "bf16": 1, "fp8": 2}[a.dtype]
        def fused_push_round():
            st["r"] += 1
            ex.win.publish(ex.F_DELTA, st["r"])
            d, s = ex._delta_ptrs(st["r"], miners)
            ex.push_average(base_win, d, s, w, st["r"], mode_id, wait_flags=[ex.win.flag_ptr(ex.F_DELTA + r) for r in miners])
            ex.wait_base()
        # the FAIR NCCL bar: pre-scale the local delta by its row of w, ncclAllReduce, one axpy -- 2 x (N-1)/N x bytes
        # on the wire instead of the all_gather's (N-1) x bytes
        scaled = torch.empty(n, dtype=torch.float32, device=dev)
        ssum = w.sum(0)
        def nccl_allreduce_prescaled():
            torch.mul(mine.float() if mine.dtype != torch.float32 else mine, w[rank][tid], out=scaled)
            dist.all_reduce(scaled)
            torch.addcmul(scaled, base, ssum[tid], out=ref)
        base0 = base.clone()
        t_push = timed(fused_push_round, dev) if (ex.win.mc_ptr and world > 1) else None
        t_ar = timed(nccl_allreduce_prescaled, dev) if world > 1 else None
        t_pr = timed(fused_pull_round, dev)
        base.copy_(base0)  # the pull round adopts the new base in place; restore for the other variants
        t_sh, t_pull = timed(fused_sharded, dev), timed(fused_pull, dev)
        t_nf, t_ag = timed(nccl_full, dev), timed(nccl_allgather_only, dev)
        torch.cuda.synchronize(); ex.win.check_errors()
        bytes_delta = n * esz
        in_sharded = (world - 1) * bytes_delta / world      # per-rank NVLink ingress of the sharded kernel
        row = {"delta_mb_fp32": mb, "delta_bytes": bytes_delta,
               "ms_fused_pull_round_rs_plus_ag_reset": t_pr, "ms_fused_sharded_gather_avg_bcast": t_sh, "ms_fused_pull_gather_avg_rank0": t_pull,
               "ms_nccl_allgather_plus_torch_avg": t_nf, "ms_nccl_allgather_only": t_ag,
               "ms_fused_push_round_nvls": t_push, "ms_nccl_prescale_allreduce_axpy": t_ar, "window_backing": ex.win.backing,
               "fused_sharded_ingress_gbs_per_rank": in_sharded / t_sh / 1e6 if world > 1 else None,
               "fused_pull_ingress_gbs_rank0": (world - 1) * bytes_delta / t_pull / 1e6 if world > 1 else None,
               "nccl_allgather_busbw_gbs": (world - 1) * n * mine.element_size() / t_ag / 1e6 if world > 1 else None,
               "roofline_ms_sharded": max(in_sharded / (NVLINK_GBS * 1e6), (world + 2) * n * 4 / world / (6578.7 * 1e6)),
               "speedup_vs_nccl_path": t_nf / min(t_sh, t_pr)}
        # pull round: NVLink ingress = 2 x (world-1)/world x bytes (delta shards in phase 1 at esz, fp32 base shards in phase 2)
        in_pull = (world - 1) / world * (bytes_delta + n * 4)
        row["pull_round_ingress_gbs_per_rank"] = in_pull / t_pr / 1e6 if world > 1 else None
        row["roofline_ms_pull_round"] = max(in_pull / (NVLINK_GBS * 1e6), 22.0 * n / (6578.7 * 1e6))
        row["fraction_of_roofline_pull_round"] = row["roofline_ms_pull_round"] / t_pr
        row["fraction_of_roofline"] = row["roofline_ms_sharded"] / t_sh
        if t_push:
            # push round per rank: ingress (N-1)/N delta bytes by pull + (N-1)/N x 6 B/elem landing from the switch; egress 6 B/elem x n/N
            in_push = (world - 1) / world * (bytes_delta + n * 6)
            row["push_round_ingress_gbs_per_rank"] = in_push / t_push / 1e6
            row["roofline_ms_push_round"] = in_push / (NVLINK_GBS * 1e6)
            row["fraction_of_roofline_push_round"] = row["roofline_ms_push_round"] / t_push
            row["speedup_push_vs_fair_nccl"] = t_ar / t_push
        out["rows"].append(row)
        if rank == 0:
            print("SWEEP " + json.dumps(row), flush=True)
        ex.win.close()
        del base, master, allg, ref, mine, pull_out
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/bandwidth_sweep_n{world}_{a.dtype}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
