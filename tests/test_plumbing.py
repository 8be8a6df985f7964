"""Multi-process CPU plumbing (BASELINE.json config 1 shape): miner processes + an averager process, gloo rendezvous,
local-disk hub -- plus the supervision / rendezvous / fault-injection utilities."""
import json
import os
import subprocess
import sys
import threading

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, port, args, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "neurons", "run.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_two_miners_one_averager_gloo_disk(tmp_path):
    args = ["--roles", "miner:0-1,averager:2", "--device", "cpu", "--backend", "disk", "--model", "gpt2-tiny", "--batch_size", "4",
            "--seq_len", "16", "--local_steps", "5", "--rounds", "2", "--meta_epochs", "1", "--storage.model_dir", str(tmp_path / "model"),
            "--storage.gradient_dir", str(tmp_path / "grad"), "--metrics_jsonl", str(tmp_path / "metrics.jsonl")]
    r = _torchrun(3, 29641, args)
    assert r.returncode == 0, r.stderr[-3000:]
    # both miners published 2 rounds, the averager published a new base
    for rk in (0, 1):
        assert int(open(tmp_path / "model" / "deltas" / f"weight_diff_{rk}.pt.round").read()) == 2
    assert int(open(tmp_path / "model" / "base" / "averaged_model.pt.round").read()) >= 1
    sd = torch.load(tmp_path / "model" / "averaged_model.pt", weights_only=False)
    assert len(sd) == 29 and "lm_head.weight" in sd and all(torch.isfinite(v).all() for v in sd.values())
    recs = [json.loads(l) for l in open(tmp_path / "metrics.jsonl")]
    assert any(r.get("role") == "averager" and "loss_averaged" in r for r in recs)
    assert any(r.get("role") == "miner" and "train_loss" in r for r in recs)


def test_fault_injection_nan_and_drop(tmp_path):
    """rank 1 publishes a NaN delta, rank 2 never publishes: the averager must finish on rank 0's delta alone."""
    args = ["--roles", "miner:0-2,averager:3", "--device", "cpu", "--backend", "disk", "--model", "gpt2-tiny", "--batch_size", "2",
            "--seq_len", "16", "--local_steps", "3", "--rounds", "1", "--meta_epochs", "1", "--inject", "nan:1,drop:2",
            "--storage.model_dir", str(tmp_path / "model"), "--storage.gradient_dir", str(tmp_path / "grad"),
            "--metrics_jsonl", str(tmp_path / "metrics.jsonl")]
    r = _torchrun(4, 29642, args)
    assert r.returncode == 0, r.stderr[-3000:]
    assert not (tmp_path / "model" / "deltas" / "weight_diff_2.pt").exists()
    recs = [json.loads(l) for l in open(tmp_path / "metrics.jsonl") if '"w_mean"' in l]
    assert recs and len(recs[-1]["w_mean"]) == 1  # exactly one valid miner was mixed
    sd = torch.load(tmp_path / "model" / "averaged_model.pt", weights_only=False)
    assert all(torch.isfinite(v).all() for v in sd.values())


def test_supervisor_autoupdate_and_rendezvous(tmp_path):
    from distributedtraining_b200.utils.auto_update import get_version_difference, monitor_repo, read_version_value
    from distributedtraining_b200.utils.bootstrap_server import StorePool, make_server
    from distributedtraining_b200.utils.bootstrap_stress import stress_test
    from distributedtraining_b200.utils.supervisor import supervise

    assert supervise([sys.executable, "-c", "pass"]) == 0
    assert supervise([sys.executable, "-c", "import sys; sys.exit(3)"], max_restarts=1, backoff=0.01) == 3
    assert get_version_difference("0.3.2", "0.3.4") == 2 and read_version_value(os.path.join(ROOT, "template", "__init__.py"))
    # auto-update against a local "remote"
    remote, clone = tmp_path / "remote", tmp_path / "clone"
    def git(*a, cwd):
        subprocess.run(["git", "-c", "user.email=t@t", "-c", "user.name=t", *a], cwd=cwd, check=True, capture_output=True)
    os.makedirs(remote / "template")
    git("init", "-b", "main", cwd=remote)
    (remote / "template" / "__init__.py").write_text('__version__ = "0.1.0"\n')
    git("add", "-A", cwd=remote); git("commit", "-m", "v1", cwd=remote)
    git("clone", str(remote), str(clone), cwd=tmp_path)
    (remote / "template" / "__init__.py").write_text('__version__ = "0.2.0"\n')
    git("commit", "-am", "v2", cwd=remote)
    seen = []
    assert monitor_repo(str(clone), interval=0.0, on_update=lambda a, b: seen.append((a, b)), max_checks=1) == "0.2.0"
    assert seen == [("0.1.0", "0.2.0")] and read_version_value(str(clone / "template" / "__init__.py")) == "0.2.0"
    # rendezvous service + stress client
    pool = StorePool(base_port=46211)
    srv = make_server(pool, port=46210)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        st = stress_test("http://127.0.0.1:46210/return_dht_address?job=a&world_size=2", 4, 3, 10)
        assert st["errors"] == 0 and st["requests"] == 12 and pool.check_and_manage() == 0
    finally:
        srv.shutdown()


def test_dummy_miner_and_wallets(tmp_path):
    from distributedtraining_b200.btt_connector import MemoryLedger
    from distributedtraining_b200.models.transformer import build_manifest, get_config
    from distributedtraining_b200.parallel.exchange import DiskExchange
    from distributedtraining_b200.utils.dummy_miner import ValidationCommunicator
    from distributedtraining_b200.utils.generate_wallets import generate_multiple_wallets

    man = build_manifest(get_config("gpt2-tiny"))
    ex = DiskExchange(str(tmp_path / "hub"), 5, man)
    dm = ValidationCommunicator(ex, man, hotkey="rank5", kind="random")
    msg = dm.send()
    assert ValidationCommunicator.verify(msg) and ex.delta_round(5) == 1 and ex.fetch_delta(5, 1).numel() == man.total
    led = MemoryLedger()
    ws = generate_multiple_wallets(4, str(tmp_path / "w"), ledger=led, validators=1)
    assert len(ws) == 4 and led.get("stake/test_hotkey_3") == "10000.0" and len(led.keys("hotkey/")) == 4


def test_coordinator_collective_plane_gloo(tmp_path):
    """Co-located round on the collective plane (gloo): all_gather of the deltas, learned-mixer steps on the averager rank,
    weighted average on every rank.  The averager's master copy is overwritten while it evaluates candidate averages, so the
    round must use the deltas gathered BEFORE that (regression: they used to be re-emitted afterwards)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29644", os.path.join(ROOT, "tests", "_coord_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(tmp_path / f"coord_{k}.json")) for k in range(2)]
    for x in res:
        assert x["err_vs_manual"] < 1e-5 and x["master_is_base"] and x["moments_zero"] and x["lr"] == 5e-5, x
        assert x["w_moved"] > 1e-6, x                       # the mixer really learned something
    assert res[0]["base_sum"] == res[1]["base_sum"] and res[0]["w_sum"] == res[1]["w_sum"]   # identical on every rank


def test_two_miners_one_validator_gloo_disk(tmp_path):
    """Validator role end to end on the disk plane: both miners' deltas are scored against the base, a rank that never
    published (the validator itself) gets score 0 like a failed download upstream."""
    args = ["--roles", "miner:0-1,validator:2", "--device", "cpu", "--backend", "disk", "--model", "gpt2-tiny", "--batch_size", "4",
            "--seq_len", "16", "--local_steps", "5", "--rounds", "1", "--storage.model_dir", str(tmp_path / "model"),
            "--storage.gradient_dir", str(tmp_path / "grad"), "--metrics_jsonl", str(tmp_path / "metrics.jsonl")]
    r = _torchrun(3, 29645, args)
    assert r.returncode == 0, r.stderr[-3000:]
    recs = [json.loads(l) for l in open(tmp_path / "metrics.jsonl")]
    val = {x["hotkey"]: x for x in recs if x.get("role") == "validator" and "hotkey" in x}
    assert val["rank0"]["loss_score"] > 0 and val["rank1"]["loss_score"] > 0 and val["rank2"]["loss_score"] == 0.0, val
