"""Worker of tests/test_plumbing.py::test_coordinator_collective_plane_gloo (one process per rank, gloo, CPU)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.parallel.exchange import CollectiveExchange  # noqa: E402
from distributedtraining_b200.parallel.local_sgd import LocalSGDCoordinator  # noqa: E402


def main():
    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(100 + rank)
    tr = Trainer("gpt2-tiny", device="cpu", batch=2, seq=16, lr=1e-2, seed=0)  # same theta_base everywhere
    ids = [torch.randint(0, tr.cfg.vocab_size, (2, 16), dtype=torch.int32) for _ in range(3)]
    for b in ids:
        tr.step(b)                                                              # rank-specific data -> rank-specific delta
    base0 = tr.base.clone()
    my_delta = tr.master - tr.base
    deltas = [torch.empty_like(my_delta) for _ in range(world)]
    dist.all_gather(deltas, my_delta)
    g = torch.Generator().manual_seed(7)
    val = [torch.randint(0, tr.cfg.vocab_size, (2, 16), dtype=torch.int32, generator=g) for _ in range(2)]
    coord = LocalSGDCoordinator(tr, CollectiveExchange(tr.man), mixer="learned", meta_steps=2, meta_lr=0.05, val_batches=val,
                                post_pull_lr=5e-5)
    coord.finish_round()
    w = coord.w.clone()
    tid = tr.man.tensor_ids("cpu")
    want = base0 * w.sum(0)[tid]
    for i in range(world):
        want = want + deltas[i] * w[i][tid]
    res = {"rank": rank, "err_vs_manual": float((tr.base - want).abs().max()), "master_is_base": bool(torch.equal(tr.master, tr.base)),
           "moments_zero": float(tr.m.abs().max()) == 0.0, "lr": tr.opt.host["lr"], "w_moved": float((w - 1.0 / world).abs().max()),
           "base_sum": float(tr.base.double().sum()), "w_sum": float(w.double().sum())}
    json.dump(res, open(os.path.join(out_dir, f"coord_{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
