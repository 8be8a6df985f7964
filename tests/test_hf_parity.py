"""Model parity against ``transformers`` (the reference trains HF GPT-2: hivetrain/training_manager.py:39-46,380-384):
state-dict import / export incl. the [PAD] resize, loss and gradients WITH the padding mask, PAD rows in the loss."""
import os

import pytest
import torch

transformers = pytest.importorskip("transformers")

from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.models.transformer import (ModelConfig, build_manifest, from_hf_state_dict, oracle_loss,
                                                         to_hf_state_dict)


def _hf_gpt2(vocab=300, d=64, L=2, H=2, npos=64):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    cfg = GPT2Config(vocab_size=vocab, n_positions=npos, n_embd=d, n_layer=L, n_head=H, resid_pdrop=0.0, embd_pdrop=0.0,
                     attn_pdrop=0.0)
    m = GPT2LMHeadModel(cfg).eval()
    with torch.no_grad():  # HF zero-inits biases and ones LN weights: randomise so that layout mistakes cannot hide
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    return m


def _batch(B=4, T=16, vocab=300, pad=299):
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, pad, (B, T), generator=g)
    lens = torch.tensor([T, 10, 5, 1])[:B]
    am = (torch.arange(T)[None] < lens[:, None]).long()
    return torch.where(am.bool(), ids, torch.full_like(ids, pad)), am


def test_state_dict_roundtrip_and_pad_resize():
    m = _hf_gpt2()
    mc = ModelConfig(family="gpt2", vocab_size=300, n_positions=64, n_embd=64, n_layer=2, n_head=2, name="t")
    man = build_manifest(mc)
    flat = from_hf_state_dict(mc, m.state_dict(), man)
    back = to_hf_state_dict(mc, flat)
    sd = {k: v for k, v in m.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))}
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    m2 = _hf_gpt2()
    m2.load_state_dict(back, strict=False)  # loads into HF as is
    # engine-layout dicts are recognised too (c_attn is [3d, d] there)
    assert torch.equal(from_hf_state_dict(mc, man.views(flat), man), flat)
    # [PAD] resize: a 300-row checkpoint into a 301-row model; new row = mean embedding
    mc1 = ModelConfig(family="gpt2", vocab_size=301, n_positions=64, n_embd=64, n_layer=2, n_head=2, name="t")
    man1 = build_manifest(mc1)
    f1 = from_hf_state_dict(mc1, m.state_dict(), man1)
    wte = man1.view(f1, "transformer.wte.weight")
    assert torch.equal(wte[:300], m.state_dict()["transformer.wte.weight"])
    assert torch.allclose(wte[300], m.state_dict()["transformer.wte.weight"].mean(0))
    # a transposed square matrix is NOT silently accepted by Manifest.pack
    bad = dict(man.views(flat))
    bad["transformer.h.0.attn.c_attn.weight"] = bad["transformer.h.0.attn.c_attn.weight"].t()
    with pytest.raises(ValueError):
        man.pack(bad, torch.zeros(man.total))


def test_loss_and_grads_match_hf_with_padding_mask():
    m = _hf_gpt2()
    mc = ModelConfig(family="gpt2", vocab_size=300, n_positions=64, n_embd=64, n_layer=2, n_head=2, name="t")
    man = build_manifest(mc)
    ids, am = _batch()
    out = m(input_ids=ids, attention_mask=am, labels=ids)  # labels = input_ids, PAD not masked (reference miner)
    out.loss.backward()
    # 1) autograd oracle
    flat = from_hf_state_dict(mc, m.state_dict(), man).requires_grad_(True)
    lo = oracle_loss(mc, man, flat, ids, attention_mask=am)
    assert abs(float(lo) - float(out.loss)) < 2e-5
    lo.backward()
    # 2) the engine (explicit forward/backward over the op layer; CPU = fp32 reference ops)
    tr = Trainer(mc, device="cpu", batch=ids.shape[0], seq=ids.shape[1], init_flat=flat.detach())
    le = tr.loss_and_grad({"input_ids": ids.int(), "attention_mask": am.int(), "labels": ids.int()})
    assert abs(float(le) - float(out.loss)) < 2e-5
    ghf = from_hf_state_dict(mc, {k: (p.grad if p.grad is not None else torch.zeros_like(p))
                                  for k, p in m.named_parameters()}, man)
    assert torch.allclose(flat.grad, ghf, atol=2e-5, rtol=1e-4)
    assert torch.allclose(tr.grad, ghf, atol=2e-5, rtol=1e-4)
    # 3) the padding mask matters: causal-only attention gives a different loss on this batch
    assert abs(float(oracle_loss(mc, man, flat.detach(), ids)) - float(out.loss)) > 1e-4


def test_trainer_from_pretrained_directory(tmp_path):
    m = _hf_gpt2()
    m.save_pretrained(str(tmp_path / "ckpt"))
    tr = Trainer.from_pretrained(str(tmp_path / "ckpt"), device="cpu", batch=2, seq=16)
    assert tr.cfg.vocab_size == 301 and tr.cfg.n_embd == 64  # + [PAD]
    ids, am = _batch(B=2)
    m.resize_token_embeddings(301, mean_resizing=False)
    with torch.no_grad():
        m.transformer.wte.weight[300] = m.transformer.wte.weight[:300].mean(0)
    ref = m(input_ids=ids, attention_mask=am, labels=ids).loss
    got = tr.eval_loss({"input_ids": ids.int(), "attention_mask": am.int(), "labels": ids.int()})
    assert abs(float(got) - float(ref)) < 2e-5
    # the same directory works as ``model_name`` of the miner loop (reference TrainingLoop ctor)
    from distributedtraining_b200.training_manager import DeltaLoop
    loop = DeltaLoop("cpu", str(tmp_path / "ckpt"), [], batch_size=2, seq_len=16)
    assert torch.equal(loop.model.master, tr.master)
    # export -> HF
    m3 = _hf_gpt2(vocab=301)
    m3.load_state_dict(tr.hf_state_dict(), strict=False)
    assert abs(float(m3(input_ids=ids, attention_mask=am, labels=ids).loss) - float(ref)) < 2e-5


@pytest.mark.gpu
def test_engine_on_gpu_matches_hf_gpt2_with_padding_mask():
    """The sm_100a engine (bf16 tensor cores, hand-written kernels) against transformers.GPT2LMHeadModel in fp32 on the same
    device: loss and every parameter gradient, with the reference's batch format (attention_mask passed, PAD in the loss)."""
    from distributedtraining_b200 import ops
    assert ops.have_kernels()
    m = _hf_gpt2(vocab=1003, d=128, L=3, H=2, npos=128).cuda()
    mc = ModelConfig(family="gpt2", vocab_size=1003, n_positions=128, n_embd=128, n_layer=3, n_head=2, name="hf-small")
    man = build_manifest(mc)
    B, T = 16, 64
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 1002, (B, T), generator=g)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    am = (torch.arange(T)[None] < lens[:, None]).long()
    ids = torch.where(am.bool(), ids, torch.full_like(ids, 1002)).cuda()
    am = am.cuda()
    out = m(input_ids=ids, attention_mask=am, labels=ids)
    out.loss.backward()
    flat = from_hf_state_dict(mc, {k: v.detach().cpu() for k, v in m.state_dict().items()}, man)
    tr = Trainer(mc, device="cuda", batch=B, seq=T, init_flat=flat, use_graph=False)
    loss = tr.loss_and_grad({"input_ids": ids.int(), "attention_mask": am.int(), "labels": ids.int()})
    assert abs(float(loss) - float(out.loss)) < 2e-2, (float(loss), float(out.loss))
    ghf = from_hf_state_dict(mc, {k: (p.grad if p.grad is not None else torch.zeros_like(p)).cpu()
                                  for k, p in m.named_parameters()}, man).cuda()
    cos = torch.nn.functional.cosine_similarity(tr.grad, ghf, dim=0)
    assert float(cos) > 0.995, float(cos)
    for s in man.specs:  # per tensor: no layout / transpose mistakes hide in the global cosine
        a, b = man.view(tr.grad, s.name).flatten(), man.view(ghf, s.name).flatten()
        if float(b.norm()) > 1e-6:
            assert float(torch.nn.functional.cosine_similarity(a, b, dim=0)) > 0.98, s.name
    # without the mask the loss is measurably different (PAD rows attend differently)
    l2 = tr.loss_and_grad(ids.int())
    assert abs(float(l2) - float(out.loss)) > 1e-4
