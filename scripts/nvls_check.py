"""NVLS plane check (torchrun, one rank per GPU): in-switch delta sum + multicast of the new base vs
  * the torch reference (all_reduce of the deltas in fp32),
  * NCCL all_reduce + torch axpy (the library baseline for a uniform average; NCCL itself may use NVLS),
  * the peer plane's pull round (reduce-scatter by pull + all-gather by pull).
Device-timed, max over ranks.   torchrun --nproc-per-node N scripts/nvls_check.py [--mb 498]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import Manifest
from distributedtraining_b200.parallel.exchange import NvlsExchange, PeerExchange
from distributedtraining_b200.parallel.launch import barrier_sync, init_distributed, max_over_ranks


class FakeTrainer:
    def __init__(self, master, base):
        self.master, self.base = master, base

    def emit_delta(self, out, scales=None):
        return ops.delta_emit(self.master, self.base, out, scales)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=498, help="fp32 arena size in MB (GPT-2-small = 498)")
    a = ap.parse_args()
    rank, world, dev = init_distributed("nccl")
    n_big = a.mb * (1 << 20) // 4 // 4
    man = Manifest([(f"big{i}", (n_big,), "normal", True) for i in range(4)] + [(f"small{i}", (777 + i,), "normal", False) for i in range(20)])
    n = man.total
    g = torch.Generator(device="cuda").manual_seed(1234)
    base = torch.randn(n, device=dev, generator=g)
    gl = torch.Generator(device="cuda").manual_seed(100 + rank)
    master = base + 0.01 * torch.randn(n, device=dev, generator=gl)
    tr = FakeTrainer(master, base)
    out = {"world": world, "numel": n, "bytes_fp32": 4 * n}
    # ---- reference ----
    d = master - base
    ref = d.clone()
    dist.all_reduce(ref)
    ref = base + ref / world
    # ---- NVLS ----
    try:
        ex = NvlsExchange(man)
    except Exception as e:  # no multicast object on this box
        out["nvls"] = f"unavailable: {e!r}"[:300]
        if rank == 0:
            print("NVLS_CHECK " + json.dumps(out), flush=True)
        dist.destroy_process_group()
        return
    ex.publish_delta(tr, 1)
    nb = ex.average_broadcast(base)
    torch.cuda.synchronize()
    out["max_err_nvls"] = float((nb - ref).abs().max())
    out["ref_max"] = float(ref.abs().max())

    def timed(fn, iters=8, warm=2):
        ts = []
        for it in range(warm + iters):
            barrier_sync(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if it >= warm:
                ts.append(max_over_ranks(e0.elapsed_time(e1), dev))
        return sorted(ts)[len(ts) // 2]

    out["ms_nvls_reduce_plus_multicast"] = timed(lambda: ex.average_broadcast(base))
    buf = torch.empty_like(d)

    def nccl_path():
        buf.copy_(d)
        dist.all_reduce(buf)
        torch.add(base, buf, alpha=1.0 / world, out=buf)

    out["ms_nccl_allreduce_plus_axpy"] = timed(nccl_path)
    # ---- peer plane pull round (no optimizer reset here: reduce-scatter by pull + sharded result) ----
    px = PeerExchange(man, delta_dtype="fp32")
    px.publish_delta(tr, 1)
    w = torch.full((world, len(man)), 1.0 / world, device=dev)
    px.win.device_barrier()
    out["ms_peer_pull_reduce_scatter_only"] = timed(lambda: px.reduce_scatter_average(base, w, 1, list(range(world))))
    out["nvlink_bytes_in_per_rank_nvls"] = int(4 * n * (1.0 / world + (world - 1) / world))
    out["ok"] = out["max_err_nvls"] <= 1e-5 * max(out["ref_max"], 1.0)
    ok = torch.tensor([int(out["ok"])], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(ok.item())
    if rank == 0:
        print("NVLS_CHECK " + json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/nvls_check_n{world}.json", "w"), indent=1)
    dist.barrier(device_ids=[dev.index])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
