"""Distributed learned mixer (parallel/meta.py) == the reference's sequential meta-learning loop
(hivetrain/averaging_logic.py:490-541), in both validation-parallel modes, on 2 gloo ranks."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,port", [("replicate", 29651), ("dp", 29652)])
def test_distributed_meta_learning_matches_sequential(tmp_path, mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_meta_worker.py"), str(tmp_path), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(tmp_path / f"meta_{mode}_{k}.json")) for k in range(2)]
    for x in res:
        assert x["mode"] == mode
        assert x["w_moved"] > 1e-5, x                      # the loop really moved w
        assert x["w_err"] < 2e-5 * max(1.0, x["w_moved"] / 1e-3), x
        assert x["base_err"] < 1e-5 and x["last_loss_err"] < 1e-4, x
    assert res[0]["w_sum"] == res[1]["w_sum"]              # w bit-identical on every rank without a broadcast
    if mode == "dp":
        assert res[0]["rows"] == [0, 3] and res[1]["rows"] == [3, 5]


def test_single_process_meta_learner_and_nan_screen():
    from distributedtraining_b200 import ops
    from distributedtraining_b200.models.trainer import Trainer
    from distributedtraining_b200.parallel.meta import DistributedMetaLearner
    torch.manual_seed(0)
    tr = Trainer("gpt2-tiny", device="cpu", batch=2, seq=16, lr=1e-2, seed=0)
    deltas = [1e-2 * torch.randn_like(tr.master) for _ in range(3)]
    deltas[1][123] = float("nan")  # a diverged miner: must be skipped, w re-initialised over the remaining two
    val = [torch.randint(0, tr.cfg.vocab_size, (2, 16), dtype=torch.int32)]
    ml = DistributedMetaLearner(tr, None, miners=[0, 1, 2], val_batches=val, meta_lr=0.05)
    ml.begin_round(1, deltas=deltas)
    assert ml.active.tolist() == [1, 0, 1]
    assert torch.allclose(ml.w[0], torch.full_like(ml.w[0], 0.5)) and float(ml.w[1].abs().max()) == 0.0
    ml.run_steps(3)
    assert float(ml.w[1].abs().max()) == 0.0 and bool(torch.isfinite(ml.w).all())
    out = ml.final_average_full(torch.empty_like(tr.master))
    good = [deltas[0], deltas[2]]
    want = torch.empty_like(tr.master)
    ops.weighted_avg(tr.base, good, ml.w[[0, 2]], tr.man, [want])
    assert bool(torch.isfinite(out).all()) and float((out - want).abs().max()) < 1e-6
