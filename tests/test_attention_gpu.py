"""tcgen05 attention kernels vs the fp32 PyTorch reference (forward, LSE, and all three input gradients)."""
import pytest
import torch

from distributedtraining_b200 import ops
from distributedtraining_b200.ops import reference as ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,H,Hkv", [(4, 64, 2, 2), (2, 128, 3, 3), (2, 256, 2, 2), (1, 512, 4, 2), (3, 192, 2, 1),
                                       (5, 64, 12, 12), (1, 1024, 2, 2), (3, 96, 2, 2), (2, 64, 4, 2), (3, 32, 2, 2)])
def test_attention_fwd_bwd(B, T, H, Hkv):
    torch.manual_seed(0)
    hd = 64
    M = B * T
    qkv = (torch.randn(M, (H + 2 * Hkv) * hd, device="cuda") * 0.7).bfloat16()
    out = torch.zeros(M, H * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    ops.attention_fwd(qkv, out, lse, B, T, H, hd, Hkv)
    r_out = torch.empty(M, H * hd, device="cuda")
    r_lse = torch.empty(B, H, T, device="cuda")
    ref.attention_fwd(qkv, r_out, r_lse, B, T, H, hd, Hkv)
    assert (out.float() - r_out).abs().max().item() < 2e-2 * max(1.0, r_out.abs().max().item())
    assert (lse - r_lse).abs().max().item() < 2e-2
    dout = (torch.randn(M, H * hd, device="cuda") * 0.5).bfloat16()
    dqkv = torch.zeros_like(qkv)
    dbias = torch.zeros((H + 2 * Hkv) * hd, device="cuda")
    ops.attention_bwd(dout, qkv, out, lse, dqkv, B, T, H, hd, Hkv, dbias=dbias)
    # qkv bias gradient (folded into the single-block kernel, separate colsum otherwise) = column sums of the stored dqkv
    want = dqkv.float().sum(0)
    assert (dbias - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
    r_dqkv = torch.empty(M, (H + 2 * Hkv) * hd, device="cuda")
    ref.attention_bwd(dout, qkv, r_out.bfloat16(), r_lse, r_dqkv, B, T, H, hd, Hkv)
    names = ["dq", "dk", "dv"]
    parts = [H * hd, Hkv * hd, Hkv * hd]
    for name, a, b in zip(names, dqkv.float().split(parts, dim=1), r_dqkv.split(parts, dim=1)):
        rel = (a - b).norm() / (b.norm() + 1e-8)
        assert rel < 3e-2, (name, float(rel))
        assert (a - b).abs().max().item() < 5e-2 * max(1.0, b.abs().max().item()), name


@pytest.mark.parametrize("B,T,H,Hkv", [(4, 64, 2, 2), (3, 32, 2, 2), (2, 256, 2, 2), (1, 512, 4, 2), (3, 96, 2, 2)])
def test_attention_dropout_fwd_bwd(B, T, H, Hkv):
    """Dropout on the softmax probabilities: all five kernels regenerate the reference's (head, q row, key token) mask."""
    torch.manual_seed(1)
    hd = 64
    M = B * T
    rng = ops.DropoutRng("cuda", seed=77)
    rng.advance()
    drop = ops.Drop(rng, 4, 0.1)
    qkv = (torch.randn(M, (H + 2 * Hkv) * hd, device="cuda") * 0.7).bfloat16()
    out = torch.zeros(M, H * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    ops.attention_fwd(qkv, out, lse, B, T, H, hd, Hkv, drop=drop)
    r_out, r_lse = torch.empty(M, H * hd, device="cuda"), torch.empty(B, H, T, device="cuda")
    ref.attention_fwd(qkv, r_out, r_lse, B, T, H, hd, Hkv, drop)
    plain = torch.empty(M, H * hd, device="cuda")
    ref.attention_fwd(qkv, plain, None, B, T, H, hd, Hkv)
    assert (plain - r_out).abs().max().item() > 0.05  # the mask really changes the output
    assert (out.float() - r_out).abs().max().item() < 2e-2 * max(1.0, r_out.abs().max().item())
    assert (lse - r_lse).abs().max().item() < 2e-2
    dout = (torch.randn(M, H * hd, device="cuda") * 0.5).bfloat16()
    dqkv = torch.zeros_like(qkv)
    ops.attention_bwd(dout, qkv, out, lse, dqkv, B, T, H, hd, Hkv, drop=drop)
    r_dqkv = torch.empty(M, (H + 2 * Hkv) * hd, device="cuda")
    ref.attention_bwd(dout, qkv, r_out.bfloat16(), r_lse, r_dqkv, B, T, H, hd, Hkv, drop)
    parts = [H * hd, Hkv * hd, Hkv * hd]
    for name, a, b in zip(["dq", "dk", "dv"], dqkv.float().split(parts, dim=1), r_dqkv.split(parts, dim=1)):
        rel = (a - b).norm() / (b.norm() + 1e-8)
        assert rel < 3e-2, (name, float(rel))


@pytest.mark.parametrize("B,T,H,Hkv", [(6, 64, 2, 2), (4, 32, 2, 2), (3, 128, 3, 3), (2, 256, 2, 2), (2, 512, 4, 2), (3, 96, 2, 1),
                                       (1, 1024, 2, 2)])
@pytest.mark.parametrize("dropout", [False, True])
def test_attention_padding_mask_fwd_bwd(B, T, H, Hkv, dropout):
    """kv_len = HF attention_mask of a right-padded batch (the reference passes it in all three roles): keys beyond the
    un-padded prefix are invisible to EVERY query row, PAD query rows included; all five kernels vs the fp32 reference."""
    torch.manual_seed(2)
    hd, M = 64, B * T
    kv = torch.randint(1, T + 1, (B,), device="cuda", dtype=torch.int32)
    kv[0] = T          # one full row
    if B > 1:
        kv[1] = 1      # one row with a single real token
    drop = None
    if dropout:
        rng = ops.DropoutRng("cuda", seed=5)
        rng.advance()
        drop = ops.Drop(rng, 7, 0.1)
    qkv = (torch.randn(M, (H + 2 * Hkv) * hd, device="cuda") * 0.7).bfloat16()
    out = torch.zeros(M, H * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    ops.attention_fwd(qkv, out, lse, B, T, H, hd, Hkv, drop=drop, kv_len=kv)
    r_out, r_lse = torch.empty(M, H * hd, device="cuda"), torch.empty(B, H, T, device="cuda")
    ref.attention_fwd(qkv, r_out, r_lse, B, T, H, hd, Hkv, drop, kv)
    causal = torch.empty(M, H * hd, device="cuda")
    ref.attention_fwd(qkv, causal, None, B, T, H, hd, Hkv, drop)
    if int(kv.min()) < T - 1:
        assert (causal - r_out).abs().max().item() > 0.05  # the padding mask really changes PAD rows
    assert (out.float() - r_out).abs().max().item() < 2e-2 * max(1.0, r_out.abs().max().item())
    assert (lse - r_lse).abs().max().item() < 2e-2
    dout = (torch.randn(M, H * hd, device="cuda") * 0.5).bfloat16()
    dqkv = torch.zeros_like(qkv)
    dbias = torch.zeros((H + 2 * Hkv) * hd, device="cuda")
    ops.attention_bwd(dout, qkv, out, lse, dqkv, B, T, H, hd, Hkv, drop=drop, dbias=dbias, kv_len=kv)
    want = dqkv.float().sum(0)
    assert (dbias - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
    r_dqkv = torch.empty(M, (H + 2 * Hkv) * hd, device="cuda")
    ref.attention_bwd(dout, qkv, r_out.bfloat16(), r_lse, r_dqkv, B, T, H, hd, Hkv, drop, kv)
    parts = [H * hd, Hkv * hd, Hkv * hd]
    for name, a, b in zip(["dq", "dk", "dv"], dqkv.float().split(parts, dim=1), r_dqkv.split(parts, dim=1)):
        rel = (a - b).norm() / (b.norm() + 1e-8)
        assert rel < 3e-2, (name, float(rel))
        # (a sequence with ONE real token funnels all T query rows into key 0: up to T x H bf16-rounded terms in one dK row)
        assert (a - b).abs().max().item() < 1e-1 * max(1.0, b.abs().max().item()), name
    # keys of PAD positions receive exactly zero gradient
    dk = dqkv.float()[:, H * hd:(H + Hkv) * hd].view(B, T, -1)
    for b in range(B):
        assert float(dk[b, int(kv[b]):].abs().max()) == 0.0 if int(kv[b]) < T else True
