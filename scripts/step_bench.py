"""Single-GPU miner-step micro benchmark (not the contract bench; see bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributedtraining_b200 import ops
from distributedtraining_b200.models.trainer import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="gpt2")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--lm-chunk", type=int, default=32768)
ap.add_argument("--fp8", action="store_true")
ap.add_argument("--fp8-bwd", action="store_true")
ap.add_argument("--dropout", type=float, default=None)
a = ap.parse_args()
torch.manual_seed(0)
tr = Trainer(a.model, device="cuda", batch=a.batch, seq=a.seq, lr=5e-4, use_graph=not a.no_graph, lm_chunk=a.lm_chunk, fp8_forward=a.fp8, fp8_backward=a.fp8_bwd, dropout=a.dropout)
V = tr.cfg.vocab_size
pool = [torch.randint(0, V, (a.batch, a.seq), dtype=torch.int32, device="cuda") for _ in range(4)]
losses = []
for i in range(a.warmup):
    losses.append(float(tr.step(pool[i % 4])))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    l = tr.step(pool[i % 4])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
tok = a.batch * a.seq
nparam = tr.man.num_params
flops = 6.0 * nparam * tok
print(json.dumps({"model": a.model, "batch": a.batch, "seq": a.seq, "graph": not a.no_graph, "fp8_forward": a.fp8, "fp8_dgrad": a.fp8_bwd, "ms_per_step": ms,
                  "tokens_per_s": tok / ms * 1e3, "mfu_vs_1412": flops / ms / 1e9 / 1412.2,
                  "launches_per_step": tr.launches_per_step, "loss_first": losses[0], "loss_last": float(l),
                  "mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
