"""Durability (SURVEY.md 5.4 / VERDICT item 9): periodic ``--save_every`` checkpoints of every role's durable state, kill -9
and ``--resume``; plus the real-text data path of the neurons."""
import glob
import json
import os
import signal
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _common(tmp_path):
    return ["--device", "cpu", "--backend", "disk", "--model", "gpt2-tiny", "--batch_size", "4", "--seq_len", "16",
            "--storage.model_dir", str(tmp_path / "model"), "--storage.gradient_dir", str(tmp_path / "grad"),
            "--checkpoint_dir", str(tmp_path / "ckpt")]


def test_miner_kill_minus_nine_and_resume(tmp_path):
    args = _common(tmp_path) + ["--local_steps", "3", "--rounds", "100000", "--save_every", "1"]
    p = subprocess.Popen([sys.executable, os.path.join(ROOT, "neurons", "miner.py")] + args, cwd=ROOT, stdout=subprocess.DEVNULL,
                         stderr=subprocess.DEVNULL)
    try:
        t0 = time.time()
        while time.time() - t0 < 240:
            done = [f for f in glob.glob(str(tmp_path / "ckpt" / "rank0_round*.pt")) if ".tmp." not in f]
            if len(done) >= 2 or (done and int(done[0][-11:-3]) >= 3):
                break
            time.sleep(0.2)
        else:
            raise AssertionError("no periodic checkpoint appeared")
    finally:
        p.send_signal(signal.SIGKILL)  # no atexit, no final save: only the PERIODIC checkpoints exist
        p.wait()
    files = sorted(f for f in glob.glob(str(tmp_path / "ckpt" / "rank0_round*.pt")) if ".tmp." not in f)
    assert 1 <= len(files) <= 2  # pruned to the newest two
    blob = torch.load(files[-1], weights_only=False)
    r_kill = blob["round"]
    assert r_kill >= 1 and blob["extra"]["global_step"] == 3 * r_kill
    # resume: two more rounds continue the numbering and start from the saved arenas
    r = subprocess.run([sys.executable, os.path.join(ROOT, "neurons", "miner.py")] + _common(tmp_path) +
                       ["--local_steps", "3", "--rounds", "2", "--save_every", "1", "--resume"], cwd=ROOT, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted(f for f in glob.glob(str(tmp_path / "ckpt" / "rank0_round*.pt")) if ".tmp." not in f)
    last = torch.load(files[-1], weights_only=False)
    assert last["round"] == r_kill + 2 and last["extra"]["global_step"] == 3 * (r_kill + 2)
    assert int(open(tmp_path / "model" / "deltas" / "weight_diff_0.pt.round").read()) == r_kill + 2  # published round numbering continues
    assert int(last["trainer"]["step"]) == 3 * (r_kill + 2)  # Adam step counter survived (no pull happened: optimizer not reset)


def test_colocated_job_checkpoints_and_resumes(tmp_path):
    base = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1"]
    args = _common(tmp_path) + ["--local_steps", "2", "--meta_epochs", "1", "--val_batch", "2", "--save_every", "1",
                                "--metrics_jsonl", str(tmp_path / "m.jsonl")]
    r = subprocess.run(base + ["--master-port", "29661", os.path.join(ROOT, "neurons", "colocated.py")] + args + ["--rounds", "2"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ck = [torch.load(sorted(glob.glob(str(tmp_path / "ckpt" / f"colocated_rank{k}_round*.pt")))[-1], weights_only=False) for k in (0, 1)]
    assert ck[0]["round"] == 2 and ck[1]["round"] == 2
    assert torch.equal(ck[0]["trainer"]["base"], ck[1]["trainer"]["base"])          # both ranks adopted the same averaged base
    assert torch.equal(ck[0]["extra"]["coordinator"]["w"], ck[1]["extra"]["coordinator"]["w"])
    assert ck[0]["extra"]["coordinator"]["meta_steps_done"] == 2 * 1 * 2            # rounds x meta_epochs^2 x val batches
    r = subprocess.run(base + ["--master-port", "29662", os.path.join(ROOT, "neurons", "colocated.py")] + args + ["--rounds", "1", "--resume"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ck2 = torch.load(sorted(glob.glob(str(tmp_path / "ckpt" / "colocated_rank0_round*.pt")))[-1], weights_only=False)
    assert ck2["round"] == 3 and ck2["extra"]["coordinator"]["rounds_total"] == 3 and ck2["extra"]["global_step"] == 6
    assert not torch.equal(ck2["trainer"]["base"], ck[0]["trainer"]["base"])        # it kept training from the restored state
    recs = [json.loads(l) for l in open(tmp_path / "m.jsonl")]
    assert any("meta_pass" in x for x in recs)


def test_validator_and_averager_state_roundtrip():
    from distributedtraining_b200.averaging_logic import ParameterizedAverager
    from distributedtraining_b200.btt_connector import LocalBittensorNetwork  # noqa: F401
    from distributedtraining_b200.models.trainer import Trainer
    tr = Trainer("gpt2-tiny", device="cpu", batch=2, seq=16)
    avg = ParameterizedAverager(tr, "cpu")
    avg.weights = torch.rand(3, len(tr.man))
    avg._consumed = {"rank0": 4, "rank1": 5}
    avg.round = 7
    sd = avg.state_dict()
    avg2 = ParameterizedAverager(Trainer("gpt2-tiny", device="cpu", batch=2, seq=16), "cpu")
    avg2.load_state_dict(sd)
    assert avg2.round == 7 and avg2._consumed == {"rank0": 4, "rank1": 5} and torch.equal(avg2.weights, avg.weights)


def test_text_data_path_miner(tmp_path):
    txt = tmp_path / "train.txt"
    txt.write_text("\n".join(["= Heading =", "", "the quick brown fox jumps over the lazy dog " * 3, "short"] * 8))
    from distributedtraining_b200.data import build_text_loader
    ld = build_text_loader(str(txt), "byte", 512, 4, 16, drop_last=True)
    b = next(iter(ld))
    assert b["input_ids"].shape == (4, 16) and b["input_ids"].dtype == torch.int32
    assert int(b["kv_len"][1]) == 1 and int(b["input_ids"][1, 0]) == 511      # empty line -> all PAD (kv_len clamped to 1)
    assert int(b["kv_len"][2]) == 16 and int(b["kv_len"][3]) == 5
    assert torch.equal(b["labels"], b["input_ids"])                          # PAD not masked in the labels
    sys.path.insert(0, os.path.join(ROOT, "neurons"))
    import importlib
    miner = importlib.import_module("miner")
    loop = miner.main(_common(tmp_path) + ["--local_steps", "2", "--rounds", "1", "--data.train_file", str(txt)])
    assert loop.global_step == 2 and loop.rounds_sent == 1
