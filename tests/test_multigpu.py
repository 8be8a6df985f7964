"""Multi-GPU tests (need >= 2 devices; launched through torchrun, NCCL control plane, peer-memory data plane)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port() -> int:
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _launch(tag, n, script, args=(), env=None, timeout=900):
    """torchrun ``script`` on n local GPUs; the complete stdout / stderr goes to gpurun_out/multigpu_<tag>.log so that a failure
    on a remote box can be read afterwards."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, **(env or {})))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multigpu_{tag}.log"), "w") as f:
            f.write(f"rc={r.returncode}\n--- stdout ---\n{r.stdout[-20000:]}\n--- stderr ---\n{r.stderr[-20000:]}\n")
    except OSError:
        pass
    return r


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_peer_exchange_matches_nccl_reference():
    n = min(_ngpu(), 4)
    env = dict(os.environ, PEER_CHECK_MB="8")
    r = _launch("peer_check", n, "scripts/peer_check.py", [], env=env, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("PEER_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[-1][len("PEER_CHECK "):])
    assert out["all_ranks_ok"], out


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_bench_two_gpus_runs():
    r = _launch("bench", 2, "bench.py", ["--gpus", "2", "--steps", "6", "--warmup", "3", "--model", "gpt2-tiny", "--batch-size", "8", "--local-steps", "3"], timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["n_gpus"] == 2 and out["value"] > 0 and out["rounds_in_timed_region"] >= 2 and out["e2e"]["value"] > 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nvls_plane_matches_allreduce_reference():
    """In-switch reduction + multicast broadcast (multimem.*) == fp32 all-reduce reference; skipped where the box exposes no
    multicast object."""
    n = min(_ngpu(), 4)
    r = _launch("nvls_check", n, "scripts/nvls_check.py", ["--mb", "16"], timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("NVLS_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[-1][len("NVLS_CHECK "):])
    if "nvls" in out:
        pytest.skip(out["nvls"])
    assert out["all_ranks_ok"], out


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_distributed_meta_learner_peer_vs_nccl_vs_sequential():
    """The sharded learned mixer over peer windows (csrc/meta_avg.cu) == its NCCL formulation == the sequential one-rank
    reference loop; w bit-identical on every rank."""
    n = min(_ngpu(), 4)
    r = _launch("meta_check", n, "scripts/meta_check.py", ["--model", "gpt2-tiny", "--val-batch", "9", "--val-seq", "64", "--steps", "6"], timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("META_CHECK ")]
    assert line, r.stdout[-2000:] + r.stderr[-3000:]
    out = json.loads(line[-1][len("META_CHECK "):])
    brief = {k: {kk: v[kk] for kk in ("ok", "w_err_vs_one_rank", "base_err_vs_one_rank", "w_identical_across_ranks")}
             for k, v in out["modes"].items()}
    assert out["all_ranks_ok"] and out["w_moved"] > 0 and r.returncode == 0, (brief, out["w_moved"], r.stderr[-1500:])


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_collective_validator_matches_serial_scoring():
    """All ranks scoring the deltas together (large-batch, graph-captured EvalModel, verdicts from the publish flags) gives the
    same losses as the serial apply-then-evaluate validator and as the NCCL-broadcast baseline."""
    n = min(_ngpu(), 4)
    r = _launch("validator_bench", n, "scripts/validator_bench.py", ["--model", "gpt2-tiny", "--eval-seq", "64", "--eval-batches", "5", "--eval-rows", "20", "--miner-steps", "4"], timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("VALBENCH ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    out = json.loads(line[-1][len("VALBENCH "):])
    assert out["max_loss_diff_collective_vs_applied"] < 2e-2, out
    assert all(abs(a - b) < 2e-2 for a, b in zip(out["collective_all_ranks"]["losses"], out["nccl"]["losses"])), out


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_role_mode_peer_plane_averager_at_nonzero_rank():
    """Miners must see (and adopt) the base published by an averager that is NOT rank 0 (advisor finding of round 1)."""
    n = min(_ngpu(), 3)
    r = _launch("role_mode_check", n, "scripts/role_mode_check.py", [], timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("ROLE_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    assert json.loads(line[-1][len("ROLE_CHECK "):])["ok"]
