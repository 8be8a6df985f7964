"""Worker of tests/test_meta.py (one process per rank, gloo, CPU): the distributed learned mixer against the sequential
single-rank formulation of the reference loop (hivetrain/averaging_logic.py:490-541)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedtraining_b200 import ops  # noqa: E402
from distributedtraining_b200.models.trainer import Trainer  # noqa: E402
from distributedtraining_b200.parallel.exchange import CollectiveExchange  # noqa: E402
from distributedtraining_b200.parallel.meta import DistributedMetaLearner  # noqa: E402


def main():
    out_dir, mode = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(100 + rank)
    tr = Trainer("gpt2-tiny", device="cpu", batch=2, seq=16, lr=1e-2, seed=0)
    for _ in range(3):
        tr.step(torch.randint(0, tr.cfg.vocab_size, (2, 16), dtype=torch.int32))
    my_delta = (tr.master - tr.base).clone()
    deltas = [torch.empty_like(my_delta) for _ in range(world)]
    dist.all_gather(deltas, my_delta)
    g = torch.Generator().manual_seed(7)
    Bv, Tv = 5, 32  # 5 rows over 2 ranks: uneven split in dp mode
    val = []
    for _ in range(2):
        ids = torch.randint(0, tr.cfg.vocab_size - 1, (Bv, Tv), dtype=torch.int32, generator=g)
        lens = torch.randint(3, Tv + 1, (Bv,), generator=g)
        am = (torch.arange(Tv)[None] < lens[:, None]).int()
        ids = torch.where(am.bool(), ids, torch.full_like(ids, tr.cfg.vocab_size - 1))
        val.append({"input_ids": ids, "attention_mask": am, "labels": ids.clone()})
    ml = DistributedMetaLearner(tr, CollectiveExchange(tr.man), val_batches=val, meta_lr=0.05, mode=mode)
    ml.begin_round(1, deltas=deltas)
    ml.run(2)  # 2 x 2 passes x 2 batches = 8 sequential steps
    new_base = ml.final_average_full(torch.empty_like(tr.master))
    # ---- sequential single-rank oracle: the reference loop on one trainer holding all deltas ----
    ref = Trainer("gpt2-tiny", device="cpu", batch=Bv, seq=Tv, lr=1e-2, seed=0)
    N, P = world, len(ref.man)
    w = torch.full((N, P), 1.0 / N)
    G = torch.empty(N, P)
    losses = []
    for _ in range(4):
        for b in val:
            ops.weighted_avg(ref.base, deltas, w, ref.man, [ref.master])
            losses.append(float(ref.loss_and_grad(b)))
            ops.multi_dot(ref.grad, deltas, ref.base, ref.master, ref.man, G)
            w.add_(G, alpha=-0.05)
    want = torch.empty_like(ref.master)
    ops.weighted_avg(ref.base, deltas, w, ref.man, [want])
    res = {"rank": rank, "mode": ml.mode, "w_err": float((ml.w - w).abs().max()), "w_moved": float((w - 1.0 / N).abs().max()),
           "base_err": float((new_base - want).abs().max()), "last_loss_err": abs(float(ml.loss_acc[1]) - losses[-1]),
           "w_sum": float(ml.w.double().sum()), "rows": [ml.r0, ml.r1]}
    json.dump(res, open(os.path.join(out_dir, f"meta_{mode}_{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
