"""Engine (explicit fwd/bwd over flat arenas) against the autograd oracle; trainer behaviour.  CPU + GPU."""
import pytest
import torch

from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import Arena, Manifest
from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.models.transformer import (TransformerEngine, build_manifest, get_config, make_targets,
                                                         new_model, oracle_loss, to_hf_state_dict)


def test_manifest_counts_match_reference_models():
    # SURVEY/BASELINE: GPT-2-small + PAD = 124 440 576 params in 148 tensors; medium 354 824 192 / 292; llama 1 235 814 400 / 146
    for name, n_params, n_tensors in [("gpt2", 124_440_576, 148), ("gpt2-medium", 354_824_192, 292),
                                      ("llama-3.2-1b", 1_235_814_400, 146)]:
        man = build_manifest(get_config(name))
        assert man.num_params == n_params and len(man) == n_tensors


def test_make_targets_shift_keeps_pad_labels():
    ids = torch.tensor([[5, 6, 7, 50257, 50257]])
    t = make_targets(ids)
    assert t.tolist() == [[6, 7, 50257, 50257, -1]]  # PAD is a real label, as in the reference (miner.py:95-99)


@pytest.mark.parametrize("name", ["gpt2-tiny", "llama-tiny"])
def test_engine_matches_autograd_cpu(name):
    torch.manual_seed(0)
    cfg, man, arena = new_model(name)
    arena.flat.add_(torch.randn_like(arena.flat) * 0.02)
    B, T = 3, 16
    ids = torch.randint(0, cfg.vocab_size, (B, T))
    theta = arena.flat.clone().requires_grad_(True)
    loss = oracle_loss(cfg, man, theta, ids)
    loss.backward()
    grads = torch.zeros_like(arena.flat)
    eng = TransformerEngine(cfg, man, arena.flat, grads, B, T, lm_chunk=32)
    eng.set_batch(ids.int())
    l2 = eng.forward_backward()
    assert abs(float(loss) - float(l2)) < 1e-5
    assert (theta.grad - grads).abs().max().item() < 1e-5
    assert abs(float(eng.forward_loss()) - float(loss)) < 1e-5


def test_trainer_learns_and_delta_roundtrip_cpu():
    torch.manual_seed(0)
    tr = Trainer("gpt2-tiny", device="cpu", batch=4, seq=16, lr=1e-2)
    ids = torch.randint(0, tr.cfg.vocab_size, (4, 16), dtype=torch.int32)
    l0 = float(tr.step(ids))
    for _ in range(10):
        l = float(tr.step(ids))
    assert l < l0
    d = torch.empty_like(tr.master)
    tr.emit_delta(d)
    assert torch.allclose(tr.base + d, tr.master)
    # base pull: theta == base, optimizer re-created, lr switched (reference training_manager.py:365-378)
    new_base = tr.base + 0.5 * d
    tr.load_base(new_base, lr=5e-5)
    assert torch.equal(tr.master, new_base) and tr.m.abs().max() == 0 and tr.opt.host["lr"] == 5e-5 and tr.opt.host_step == 0


def test_hf_state_dict_has_149_keys_and_conv1d_layout():
    cfg, man, arena = new_model("gpt2-tiny")
    sd = to_hf_state_dict(cfg, arena)
    assert len(sd) == len(man) + 1 and sd["lm_head.weight"].data_ptr() == sd["transformer.wte.weight"].data_ptr()
    assert sd["transformer.h.0.attn.c_attn.weight"].shape == (cfg.n_embd, 3 * cfg.n_embd)


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,T", [("gpt2-tiny", 4, 64), ("llama-tiny", 2, 128)])
def test_engine_matches_autograd_gpu(name, B, T):
    torch.manual_seed(0)
    cfg, man, arena = new_model(name)
    arena.flat.add_(torch.randn_like(arena.flat) * 0.02)
    flat = arena.flat.cuda()
    p16 = flat.bfloat16()
    ids = torch.randint(0, cfg.vocab_size, (B, T), device="cuda")
    theta = p16.float().requires_grad_(True)  # oracle sees the same bf16-rounded weights
    loss = oracle_loss(cfg, man, theta, ids)
    loss.backward()
    grads = torch.zeros_like(flat)
    eng = TransformerEngine(cfg, man, p16, grads, B, T, lm_chunk=128)
    eng.set_batch(ids.int())
    l2 = eng.forward_backward()
    assert abs(float(loss) - float(l2)) < 3e-2
    # per-tensor relative error (bf16 activations): compare direction and magnitude
    bad = []
    for s in man:
        a, b = man.view(theta.grad, s.name), man.view(grads, s.name)
        rel = (a - b).norm() / (a.norm() + 1e-8)
        if rel > 6e-2:
            bad.append((s.name, float(rel)))
    assert not bad, bad


@pytest.mark.gpu
def test_trainer_graph_equals_eager_gpu():
    torch.manual_seed(0)
    ids = [torch.randint(0, 512, (8, 64), dtype=torch.int32) for _ in range(4)]
    out = []
    for use_graph in (False, True):
        tr = Trainer("gpt2-tiny", device="cuda", batch=8, seq=64, lr=1e-3, seed=1, use_graph=use_graph)
        losses = [float(tr.step(x.cuda())) for x in ids]
        out.append((losses, tr.master.clone()))
        assert tr.launches_per_step > 10
    assert max(abs(a - b) for a, b in zip(out[0][0], out[1][0])) < 2e-3
    assert (out[0][1] - out[1][1]).abs().max().item() < 1e-3
    assert out[0][0][-1] < out[0][0][0] + 0.5


@pytest.mark.parametrize("name", ["gpt2-tiny", "llama-tiny"])
def test_fused_delta_and_peer_source_forward(name):
    """dual-B (x (W+dW)^T without materialising W+dW) and persist-B (weights pulled from a source arena and persisted as
    a side effect of the forward GEMMs) give the same loss as the materialised model.  CPU: reference ops."""
    torch.manual_seed(0)
    cfg, man, arena = new_model(name)
    base = arena.flat.clone()
    delta = torch.randn_like(base) * 0.01
    B, T = 2, 16
    ids = torch.randint(0, cfg.vocab_size, (B, T), dtype=torch.int32)
    # oracle: materialised base + delta
    eng = TransformerEngine(cfg, man, (base + delta).clone(), None, B, T, lm_chunk=32)
    eng.set_batch(ids)
    want = float(eng.forward_loss())
    # fused delta: matrices stay at base, small tensors get base+delta through the chunk-restricted apply kernel
    p = base.clone()
    eng2 = TransformerEngine(cfg, man, p, None, B, T, lm_chunk=16)
    ops.weighted_avg(base, [delta], torch.ones(1, len(man)), man, [p], chunk_ids=eng2.small_chunk_ids(), unit_base=True)
    big = man[cfg.family == "gpt2" and "transformer.h.0.mlp.c_fc.weight" or "model.layers.0.mlp.down_proj.weight"]
    assert torch.equal(p[big.offset:big.offset + big.numel], base[big.offset:big.offset + big.numel])
    eng2.set_delta(delta)
    eng2.set_batch(ids)
    assert abs(float(eng2.forward_loss()) - want) < 1e-5
    eng2.set_delta(None)
    ops.weighted_avg(base, [delta], torch.zeros(1, len(man)), man, [p], chunk_ids=eng2.small_chunk_ids(), unit_base=True)
    assert torch.allclose(p, base)
    # peer source + persist
    local = torch.zeros_like(base)
    eng3 = TransformerEngine(cfg, man, local, None, B, T, lm_chunk=16)
    eng3.set_source((base + delta).clone())
    eng3.set_batch(ids)
    assert abs(float(eng3.forward_loss()) - want) < 1e-5
    eng3.persist_small_from_source()
    eng3.set_source(None)
    for sp in man:  # every tensor arrived as a side effect of the first forward (arena padding is not transported)
        assert torch.allclose(man.view(local, sp.name), man.view(base + delta, sp.name)), sp.name
    assert abs(float(eng3.forward_loss()) - want) < 1e-5


@pytest.mark.gpu
def test_fp8_forward_trains_close_to_bf16():
    """fp8 (e4m3, delayed scaling) forward GEMMs: losses track the bf16 engine and the model still learns."""
    torch.manual_seed(0)
    ids = [torch.randint(0, 512, (8, 64), dtype=torch.int32, device="cuda") for _ in range(4)]
    res = {}
    for fp8 in (False, True):
        tr = Trainer("gpt2-tiny", device="cuda", batch=8, seq=64, lr=1e-3, seed=1, use_graph=False, fp8_forward=fp8)
        res[fp8] = [float(tr.step(ids[i % 4])) for i in range(12)]
    assert res[True][-1] < res[True][0]
    assert max(abs(a - b) for a, b in zip(res[False], res[True])) < 0.15, (res[False], res[True])


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gpt2-tiny", "llama-tiny"])
def test_fp8_backward_grads_close_to_bf16(model):
    """fp8 dgrad (dY quantised to e5m2 x the TRANSPOSED e4m3 weight copy): the gradient arena stays aligned with the bf16
    engine's (cosine per tensor), and a short run still learns at the same pace."""
    torch.manual_seed(0)
    V = get_config(model).vocab_size
    ids = [torch.randint(0, V - 1, (8, 64), dtype=torch.int32, device="cuda") for _ in range(4)]
    tr = {k: Trainer(model, device="cuda", batch=8, seq=64, lr=1e-3, seed=1, use_graph=False, dropout=0.0, fp8_forward=k > 0, fp8_backward=k > 1)
          for k in (0, 1, 2)}
    for k in tr:  # two passes: the first only collects amax (delayed scaling), the second uses rolled scales
        for _ in range(3):
            tr[k].loss_and_grad(ids[0])
            if k:
                tr[k].engine.roll_fp8_scales()
    man = tr[0].man
    worst = 1.0
    for i in range(len(man)):
        a, b = man.view(tr[0].grad, i).float().flatten(), man.view(tr[2].grad, i).float().flatten()
        if a.norm() > 0:
            worst = min(worst, float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)))
    assert worst > 0.97, worst
    res = {k: [float(tr[k].step(ids[i % 4])) for i in range(12)] for k in tr}
    assert res[2][-1] < res[2][0]
    assert max(abs(a - b) for a, b in zip(res[0], res[2])) < 0.2, (res[0], res[2])


def _with_dropout(name, p=0.1):
    import dataclasses
    return dataclasses.replace(get_config(name), dropout=p)


def test_dropout_masks_are_counter_based_and_unbiased():
    from distributedtraining_b200.ops import reference as ref
    m0 = ref.drop_mult_2d((7, 1), 3, 0.1, 257, 128, "cpu")
    assert torch.equal(m0, ref.drop_mult_2d((7, 1), 3, 0.1, 257, 128, "cpu"))          # pure function of the state
    assert not torch.equal(m0, ref.drop_mult_2d((7, 2), 3, 0.1, 257, 128, "cpu"))      # counter changes the mask
    assert not torch.equal(m0, ref.drop_mult_2d((7, 1), 4, 0.1, 257, 128, "cpu"))      # so does the site id
    keep = (m0 > 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01 and abs(m0.mean().item() - 1.0) < 0.02
    ma = ref.drop_mult_attn((7, 1), 1, 0.1, 3, 32, 4, "cpu")
    assert ma.shape == (3, 4, 32, 32) and abs((ma > 0).float().mean().item() - 0.9) < 0.02


def test_engine_dropout_matches_autograd_cpu():
    """Train-mode dropout (embd / attn / resid, GPT-2's 0.1): the explicit engine and plain autograd agree when the
    oracle applies the same counter-based masks; eval mode ignores dropout; consecutive steps use different masks."""
    torch.manual_seed(0)
    cfg = _with_dropout("gpt2-tiny")
    cfg, man, arena = new_model(cfg)
    arena.flat.add_(torch.randn_like(arena.flat) * 0.02)
    B, T = 3, 16
    ids = torch.randint(0, cfg.vocab_size, (B, T))
    grads = torch.zeros_like(arena.flat)
    eng = TransformerEngine(cfg, man, arena.flat, grads, B, T, lm_chunk=32)
    eng.set_batch(ids.int())
    l_eng = float(eng.forward_backward())
    theta = arena.flat.clone().requires_grad_(True)
    loss = oracle_loss(cfg, man, theta, ids, drop_state=eng.rng.state)
    loss.backward()
    assert abs(float(loss) - l_eng) < 1e-5
    assert (theta.grad - grads).abs().max().item() < 1e-5
    l_eval = float(eng.forward_loss())
    assert abs(l_eval - float(oracle_loss(cfg, man, arena.flat, ids))) < 1e-5 and abs(l_eval - l_eng) > 1e-4
    l_eng2 = float(eng.forward_backward())
    assert abs(l_eng2 - l_eng) > 1e-6  # the counter advanced -> new masks


@pytest.mark.gpu
def test_engine_dropout_matches_autograd_gpu():
    torch.manual_seed(0)
    cfg, man, arena = new_model(_with_dropout("gpt2-tiny"))
    arena.flat.add_(torch.randn_like(arena.flat) * 0.02)
    B, T = 4, 64
    flat = arena.flat.cuda()
    p16 = flat.bfloat16()
    ids = torch.randint(0, cfg.vocab_size, (B, T), device="cuda")
    grads = torch.zeros_like(flat)
    eng = TransformerEngine(cfg, man, p16, grads, B, T, lm_chunk=128)
    eng.set_batch(ids.int())
    l2 = eng.forward_backward()
    theta = p16.float().requires_grad_(True)
    loss = oracle_loss(cfg, man, theta, ids, drop_state=eng.rng.state)
    loss.backward()
    assert abs(float(loss) - float(l2)) < 3e-2
    assert 0.05 < (eng.xs[0] == 0).float().mean().item() < 0.15  # the embedding-dropout site really dropped ~10 %
    bad = []
    for s in man:
        a, b = man.view(theta.grad, s.name), man.view(grads, s.name)
        rel = (a - b).norm() / (a.norm() + 1e-8)
        if rel > 6e-2:
            bad.append((s.name, float(rel)))
    assert not bad, bad


@pytest.mark.gpu
def test_trainer_dropout_graph_replays_fresh_masks_gpu():
    """The captured step regenerates its masks from the device counter: the same batch gives a different train loss on
    every replay, and the eval loss (no dropout) is deterministic."""
    torch.manual_seed(0)
    tr = Trainer(_with_dropout("gpt2-tiny"), device="cuda", batch=8, seq=64, lr=0.0, seed=1, use_graph=True)
    ids = torch.randint(0, 512, (8, 64), dtype=torch.int32, device="cuda")
    losses = [float(tr.step(ids)) for _ in range(4)]   # lr = 0: identical weights, only the masks differ
    assert len({round(l, 6) for l in losses}) == 4, losses
    assert int(tr.engine.rng.state[1]) >= 4
    e1, e2 = float(tr.eval_loss(ids)), float(tr.eval_loss(ids))
    assert e1 == e2


def test_meta_gradient_is_deterministic_unless_asked():
    """Trainer.loss_and_grad (the averager's meta-gradient) ignores dropout by default and honours ``meta_dropout=True``."""
    cfg = _with_dropout("gpt2-tiny")
    ids = torch.randint(0, 512, (2, 16), dtype=torch.int32)
    tr = Trainer(cfg, device="cpu", batch=2, seq=16, lr=1e-2)
    l1 = float(tr.loss_and_grad(ids)); g1 = tr.grad.clone()
    l2 = float(tr.loss_and_grad(ids))
    assert l1 == l2 and torch.equal(g1, tr.grad) and abs(l1 - float(tr.eval_loss(ids))) < 1e-6
    tr2 = Trainer(cfg, device="cpu", batch=2, seq=16, lr=1e-2, meta_dropout=True)
    assert float(tr2.loss_and_grad(ids)) != float(tr2.loss_and_grad(ids))
