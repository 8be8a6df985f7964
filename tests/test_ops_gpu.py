"""Numerics of every hand-written sm_100a kernel against the plain-PyTorch fp32 reference of the same op."""
import math

import pytest
import torch

from distributedtraining_b200 import ops
from distributedtraining_b200.ops import reference as ref
from distributedtraining_b200.models.arena import Manifest

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).bfloat16()


def _close(a, b, rtol=2e-2, atol=None):
    a, b = a.float(), b.float()
    atol = atol if atol is not None else rtol * max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= atol, f"max err {err} > {atol}"


def test_kernels_loaded():
    assert ops.have_kernels()


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(256, 512, 192, 0, 0), (300, 1000, 200, 0, 0), (512, 768, 512, 0, 1),
                                             (768, 512, 640, 1, 1), (256, 512, 512, 1, 0)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    torch.manual_seed(0)
    A, B = _bf(M, K, scale=0.5), _bf(N, K, scale=0.5)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, out, a_mn=bool(a_mn), b_mn=bool(b_mn))
    _close(out, A.float() @ B.float().t(), rtol=1e-2)


@pytest.mark.parametrize("epi", ["bias", "bias_gelu", "bias_resid", "resid", "dgelu"])
def test_gemm_epilogues(epi):
    torch.manual_seed(1)
    M, N, K = 384, 768, 256
    A, B = _bf(M, K, scale=0.5), _bf(N, K, scale=0.5)
    bias, aux = _bf(N), _bf(M, N)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    out2 = torch.empty_like(out)
    r = torch.empty(M, N, device=DEV)
    r2 = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, out, epi=epi, bias=bias, aux=aux, out2=out2 if epi == "bias_gelu" else None)
    ref.gemm(A, B, r, epi=epi, bias=bias, aux=aux, out2=r2)
    _close(out, r, rtol=1e-2)
    if epi == "bias_gelu":
        _close(out2, r2, rtol=1e-2)


def test_gemm_f32_accumulate_splitk():
    torch.manual_seed(2)
    M, N, K = 768, 768, 4096
    A, B = _bf(M, K, scale=0.3), _bf(N, K, scale=0.3)
    a, b = A.t().contiguous(), B.t().contiguous()
    out = torch.ones(M, N, device=DEV)
    ops.gemm(a, b, out, a_mn=True, b_mn=True, accumulate=True, splits=8)
    _close(out, 1.0 + A.float() @ B.float().t(), rtol=1e-4)


def test_gemm_strided_views_vocab():
    torch.manual_seed(3)
    M, V, d, ldl = 256, 1003, 128, 1024
    x, wte = _bf(M, d), _bf(V, d)
    lg = torch.zeros(M, ldl, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, wte, lg[:, :V])
    _close(lg[:, :V], x.float() @ wte.float().t(), rtol=1e-2)
    assert lg[:, V:].abs().max().item() == 0.0
    dx = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    ops.gemm(lg[:, :V], wte, dx, b_mn=True)
    _close(dx, lg[:, :V].float() @ wte.float(), rtol=1e-2)
    dw = torch.zeros(V, d, device=DEV)
    ops.gemm(lg[:, :V], x, dw, a_mn=True, b_mn=True, accumulate=True)
    _close(dw, lg[:, :V].float().t() @ x.float(), rtol=1e-3)


def test_embed():
    torch.manual_seed(4)
    B, T, V, d = 4, 32, 1000, 256
    ids = torch.randint(0, V, (B, T), device=DEV, dtype=torch.int32)
    wte, wpe = _bf(V, d), _bf(64, d)
    out = torch.empty(B * T, d, device=DEV, dtype=torch.bfloat16)
    ops.embed_fwd(ids, wte, wpe, out)
    r = torch.empty(B * T, d, device=DEV)
    ref.embed_fwd(ids.long(), wte, wpe, r)
    _close(out, r, rtol=1e-2)
    dx = _bf(B * T, d)
    dwte, dwpe = torch.zeros(V, d, device=DEV), torch.zeros(64, d, device=DEV)
    rwte, rwpe = torch.zeros(V, d, device=DEV), torch.zeros(64, d, device=DEV)
    ops.embed_bwd(dx, ids, dwte, dwpe)
    ref.embed_bwd(dx, ids.long(), rwte, rwpe)
    _close(dwte, rwte, rtol=1e-4)
    _close(dwpe, rwpe, rtol=1e-4)


@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("d", [768, 2048])
def test_norm(rms, d):
    torch.manual_seed(5)
    M = 515
    x, w, b = _bf(M, d) + 0.3, _bf(d) + 1.0, _bf(d)
    out = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    rout, rmean, rrstd = torch.empty(M, d, device=DEV), torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    if rms:
        ops.rmsnorm_fwd(x, w, 1e-5, out, rstd)
        ref.rmsnorm_fwd(x, w, 1e-5, rout, rrstd)
    else:
        ops.layernorm_fwd(x, w, b, 1e-5, out, mean, rstd)
        ref.layernorm_fwd(x, w, b, 1e-5, rout, rmean, rrstd)
        _close(mean, rmean, rtol=1e-3)
    _close(out, rout, rtol=1e-2)
    _close(rstd, rrstd, rtol=1e-3)
    dy, dres = _bf(M, d), _bf(M, d)
    dx = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    dw, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    rdx, rdw, rdb = torch.empty(M, d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    if rms:
        ops.rmsnorm_bwd(dy, x, w, rrstd, dx, dw, dres)
        ref.rmsnorm_bwd(dy, x, w, rrstd, rdx, rdw, dres)
    else:
        ops.layernorm_bwd(dy, x, w, rmean, rrstd, dx, dw, db, dres)
        ref.layernorm_bwd(dy, x, w, rmean, rrstd, rdx, rdw, rdb, dres)
        _close(db, rdb, rtol=1e-3)
    _close(dx, rdx, rtol=1e-2)
    _close(dw, rdw, rtol=1e-3)


def test_cross_entropy():
    torch.manual_seed(6)
    M, V, ldl = 64, 50258, 50304
    lg = torch.zeros(M, ldl, device=DEV, dtype=torch.bfloat16)
    lg[:, :V] = _bf(M, V, scale=2.0)
    tgt = torch.randint(0, V, (M,), device=DEV, dtype=torch.int32)
    tgt[::7] = -1
    r = lg.clone().float()
    losses, rl = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ref.ce_fwd_bwd(r, tgt, V, rl, 1.0 / M)
    ops.ce_fwd_bwd(lg, tgt, V, losses, 1.0 / M)
    _close(losses, rl, rtol=1e-3)
    _close(lg[:, :V], r[:, :V], rtol=2e-2)
    assert lg[:, V:].abs().max().item() == 0.0


def test_colsum_swiglu_rope():
    torch.manual_seed(7)
    x = _bf(1000, 768)
    out, r = torch.ones(768, device=DEV), torch.ones(768, device=DEV)
    ops.colsum(x, out)
    ref.colsum(x, r)
    _close(out, r, rtol=1e-4)
    gu = _bf(333, 512)
    o, ro = torch.empty(333, 256, device=DEV, dtype=torch.bfloat16), torch.empty(333, 256, device=DEV)
    ops.swiglu_fwd(gu, o)
    ref.swiglu_fwd(gu, ro)
    _close(o, ro, rtol=1e-2)
    d = _bf(333, 256)
    dg, rdg = torch.empty_like(gu), torch.empty(333, 512, device=DEV)
    ops.swiglu_bwd(d, gu, dg)
    ref.swiglu_bwd(d, gu, rdg)
    _close(dg, rdg, rtol=1e-2)
    B, T, H, Hkv, hd = 2, 64, 4, 2, 64
    qkv = _bf(B * T, (H + 2 * Hkv) * hd)
    q2 = qkv.clone().float()
    ops.rope_(qkv, B, T, H, Hkv, hd, 10000.0)
    ref.rope_(q2, B, T, H, Hkv, hd, 10000.0)
    _close(qkv, q2, rtol=1e-2)


def _toy_manifest():
    return Manifest([("a", (300, 70), "normal", True), ("b", (70,), "zeros", False), ("c", (1000, 33), "normal", True),
                     ("d", (5,), "ones", False)])


def test_adamw_delta_reset():
    torch.manual_seed(8)
    n = 256 * 40
    master = torch.randn(n, device=DEV)
    base = master.clone()
    grad = torch.randn(n, device=DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    rm, rmm, rv = master.clone(), m.clone(), v.clone()
    st = ops.AdamState(DEV, 5e-4, eps=1e-6)
    for step in range(1, 4):
        ops.adamw_step(master, p16, grad, m, v, st)
        ref.adamw_step(rm, None, grad, rmm, rv, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, step=step)
    _close(master, rm, rtol=1e-5)
    _close(p16, rm, rtol=1e-2)
    assert int(st.step.item()) == 3
    # lazy optimizer re-creation: after reset() the first step ignores whatever the moment arenas hold ("fresh" flag)
    st.reset()
    m.fill_(1e9); v.fill_(1e9)
    rmm.zero_(); rv.zero_()
    ops.adamw_step(master, p16, grad, m, v, st)
    ref.adamw_step(rm, None, grad, rmm, rv, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, step=1)
    _close(master, rm, rtol=1e-5)
    _close(m, rmm, rtol=1e-4)
    _close(v, rv, rtol=1e-4)
    for dt in (torch.float32, torch.bfloat16):
        d = torch.empty(n, device=DEV, dtype=dt)
        ops.delta_emit(master, base, d)
        _close(d, master - base, rtol=1e-2 if dt == torch.bfloat16 else 1e-6)
    q, sc = torch.empty(n, device=DEV, dtype=torch.uint8), torch.empty(n // 32, device=DEV)
    ops.delta_emit(master, base, q, sc)
    _close(ops.dequant_fp8(q, sc), master - base, rtol=7e-2)
    ops.round_reset(base, master, p16, m, v)
    assert torch.equal(master, base) and m.abs().max().item() == 0 and v.abs().max().item() == 0


@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp8"])
def test_weighted_avg_multi_dot(dtype):
    torch.manual_seed(9)
    man = _toy_manifest()
    n, P, N = man.total, len(man), 5
    base = torch.randn(n, device=DEV)
    true = [torch.randn(n, device=DEV) * 0.1 for _ in range(N)]
    w = torch.rand(N, P, device=DEV) - 0.2
    scales = None
    if dtype == "fp32":
        deltas = true
    elif dtype == "bf16":
        deltas = [t.bfloat16() for t in true]
    else:
        deltas, scales = [], []
        for t in true:
            q, sc = torch.empty(n, device=DEV, dtype=torch.uint8), torch.empty(n // 32, device=DEV)
            ops.delta_emit(t, torch.zeros_like(t), q, sc)
            deltas.append(q)
            scales.append(sc)
    dec = [ops.dequant_fp8(d, s) for d, s in zip(deltas, scales)] if scales else [d.float() for d in deltas]
    out, out16 = torch.empty(n, device=DEV), torch.empty(n, device=DEV, dtype=torch.bfloat16)
    nan = torch.zeros(N, device=DEV, dtype=torch.int32)
    ops.weighted_avg(base, deltas, w, man, [out], [out16], dscales=scales, nan_flags=nan)
    r = torch.empty(n, device=DEV)
    ref.weighted_avg(base, dec, w, man.tensor_ids(DEV), r)
    _close(out, r, rtol=1e-5)
    _close(out16, r, rtol=1e-2)
    assert nan.sum().item() == 0
    # == the reference's formulation sum_i w_ij (base + delta_i)
    tid = man.tensor_ids(DEV)
    direct = sum(w[i][tid] * (base + dec[i]) for i in range(N))
    _close(out, direct, rtol=1e-5)
    # shard form: two launches over disjoint chunk ranges give the same result
    cs, _, _ = man.seg_table(DEV)
    out2 = torch.zeros(n, device=DEV)
    half = cs.numel() // 2
    ops.weighted_avg(base, deltas, w, man, [out2], dscales=scales, chunk_range=(0, half))
    ops.weighted_avg(base, deltas, w, man, [out2], dscales=scales, chunk_range=(half, cs.numel()))
    assert torch.equal(out2, out)
    g = torch.randn(n, device=DEV)
    G, rG = torch.empty(N, P, device=DEV), torch.empty(N, P, device=DEV)
    ops.multi_dot(g, deltas, base, out, man, G, dscales=scales)
    ref.multi_dot(g, dec, base, out, tid, P, rG)
    _close(G, rG, rtol=1e-3)


def test_weighted_avg_nan_screen():
    man = _toy_manifest()
    n, P, N = man.total, len(man), 3
    base = torch.zeros(n, device=DEV)
    deltas = [torch.zeros(n, device=DEV) for _ in range(N)]
    deltas[1][man["c"].offset + 17] = float("nan")
    w = torch.full((N, P), 1.0 / N, device=DEV)
    nan = torch.zeros(N, device=DEV, dtype=torch.int32)
    out = torch.empty(n, device=DEV)
    ops.weighted_avg(base, deltas, w, man, [out], nan_flags=nan)
    assert nan.tolist() == [0, 1, 0]


def test_gemm_dual_b_and_persist():
    """C = A (B + B2)^T in two accumulating passes; B tiles persisted to a local copy while being consumed."""
    torch.manual_seed(11)
    M, N, K = 640, 768, 320
    A, B, B2 = _bf(M, K, scale=0.5), _bf(N, K, scale=0.5), _bf(N, K, scale=0.05)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, out, b2=B2)
    _close(out, A.float() @ (B.float() + B2.float()).t(), rtol=1e-2)
    dst = torch.zeros_like(B)
    ops.gemm(A, B, out, b_persist=dst)
    assert torch.equal(dst, B)
    _close(out, A.float() @ B.float().t(), rtol=1e-2)
    # MN-major B (dgrad form) with persist
    Bt = B.t().contiguous()
    dst2 = torch.zeros_like(Bt)
    ops.gemm(A, Bt, out, b_mn=True, b_persist=dst2)
    assert torch.equal(dst2, Bt)


def test_gemm_fp8_delayed_scaling():
    """e4m3 x e4m3 GEMM (kind::f8f6f4) with delayed per-tensor scales against the fp32 product of the DEQUANTISED operands
    (exact up to accumulation order) and against the unquantised product (fp8 rounding error)."""
    torch.manual_seed(12)
    M, N, K = 1024, 768, 1536
    A, B = _bf(M, K, scale=0.7), _bf(N, K, scale=0.05)
    sa, sb = ops.Fp8Scale(DEV), ops.Fp8Scale(DEV)
    a8, b8 = torch.empty(M, K, device=DEV, dtype=torch.uint8), torch.empty(N, K, device=DEV, dtype=torch.uint8)
    for _ in range(2):  # step 0 collects amax with the initial scale, step 1 uses the rolled scale
        ops.quantize_fp8(A, a8, sa)
        ops.quantize_fp8(B, b8, sb)
        assert abs(sa.amax.item() - A.float().abs().max().item()) < 1e-6
        sa.roll(); sb.roll()
    ops.quantize_fp8(A, a8, sa)
    ops.quantize_fp8(B, b8, sb)
    bias = _bf(N)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm_fp8(a8, b8, out, sa, sb, epi="bias", bias=bias)
    Ad = a8.view(torch.float8_e4m3fn).float() * sa.scale
    Bd = b8.view(torch.float8_e4m3fn).float() * sb.scale
    _close(out, Ad @ Bd.t() + bias.float(), rtol=1e-2)
    exact = A.float() @ B.float().t() + bias.float()
    rel = (out.float() - exact).norm() / exact.norm()
    assert rel < 6e-2, float(rel)


def test_gemm_fp8_e5m2_dgrad_and_transposed_quantise():
    """dgrad formulation: dX[M, K] = dY[M, N] (e5m2) @ W[N, K] with W held as the TRANSPOSED e4m3 copy W8T[K, N] produced
    by quantize_fp8_t -- the byte-exact transpose of quantize_fp8(W) -- so both operands stay K-major for kind::f8f6f4."""
    torch.manual_seed(14)
    M, N, K = 1024, 1536, 768
    dY, Wt = _bf(M, N, scale=3e-3), _bf(N, K, scale=0.05)
    sd, sw = ops.Fp8Scale(DEV), ops.Fp8Scale(DEV)
    d8 = torch.empty(M, N, device=DEV, dtype=torch.uint8)
    w8, w8t = torch.empty(N, K, device=DEV, dtype=torch.uint8), torch.empty(K, N, device=DEV, dtype=torch.uint8)
    for _ in range(2):
        ops.quantize_fp8(dY, d8, sd, e5m2=True)
        ops.quantize_fp8(Wt, w8, sw)
        sd.scale.copy_(sd.amax / ops.E5M2_MAX); sd.amax.zero_()
        sw.roll()
    ops.quantize_fp8(dY, d8, sd, e5m2=True)
    ops.quantize_fp8(Wt, w8, sw)
    ops.quantize_fp8_t(Wt, w8t, sw)
    assert torch.equal(w8t, w8.t().contiguous())
    out = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    ops.gemm_fp8(d8, w8t, out, sd, sw, a_e5m2=True)
    Dd = d8.view(torch.float8_e5m2).float() * sd.scale
    Wd = w8.view(torch.float8_e4m3fn).float() * sw.scale
    _close(out, Dd @ Wd, rtol=1e-2)
    exact = dY.float() @ Wt.float()
    rel = (out.float() - exact).norm() / exact.norm()
    assert rel < 0.12, float(rel)   # e5m2 keeps 2 mantissa bits: ~2^-3 worst-case per element, averaged down by the K sum


def test_loss_mean_and_memset_zero():
    """The step's loss reduction (one deterministic block) and the memset-node zeroing that replaced the ATen kernels."""
    torch.manual_seed(15)
    for n in (1, 33, 4096, 32768, 40001):
        x = torch.rand(n, device=DEV) * 11.0
        out = torch.zeros((), device=DEV)
        ops.loss_mean(x, 1.0 / n, out)
        assert abs(out.item() - x.double().mean().item()) < 1e-5 * 11.0
        out2 = torch.zeros((), device=DEV)
        ops.loss_mean(x, 1.0 / n, out2)
        assert out.item() == out2.item()  # fixed reduction order
    g = torch.randn(1 << 20, device=DEV)
    ops.zero_(g[1024:])
    assert float(g[1024:].abs().max()) == 0.0 and float(g[:1024].abs().max()) > 0.0


def test_checksum_matches_cpu_and_detects_change():
    torch.manual_seed(13)
    x = torch.randn(256 * 37, device=DEV)
    c_gpu, c_cpu = ops.checksum(x), ops.checksum(x.cpu())
    assert c_gpu == c_cpu and len(c_gpu) == 32
    y = x.clone()
    y[1234] += 1e-3
    assert ops.checksum(y) != c_gpu


def test_dropout_sites_match_reference_masks():
    """Every kernel that applies (or re-applies) a dropout site regenerates exactly the reference's counter-based mask."""
    torch.manual_seed(11)
    rng = ops.DropoutRng(DEV, seed=123)
    rng.advance(); rng.advance()
    assert rng.state.tolist()[:2] == [123, 2]
    # GEMM residual epilogue: out = aux + dropout(A B^T + bias)
    M, N, K = 384, 768, 256
    a, b, bias, aux = _bf(M, K), _bf(N, K, scale=0.1), _bf(N), _bf(M, N)
    out, r = torch.empty(M, N, device=DEV, dtype=torch.bfloat16), torch.empty(M, N, device=DEV)
    dr = ops.Drop(rng, 5, 0.1)
    ops.gemm(a, b, out, epi="bias_resid", bias=bias, aux=aux, drop=dr)
    ref.gemm(a, b, r, epi="bias_resid", bias=bias, aux=aux, drop=dr)
    _close(out, r)
    dropped = (ref.drop_mult_2d(rng.state, 5, 0.1, M, N, DEV) == 0)
    assert torch.equal(out[dropped], aux[dropped]) and 0.08 < dropped.float().mean().item() < 0.12
    # embedding fwd / bwd
    V, T, d = 1000, 64, 768
    ids = torch.randint(0, V, (4, T), device=DEV, dtype=torch.int32)
    wte, wpe = _bf(V, d), _bf(T, d)
    x, rx = torch.empty(4 * T, d, device=DEV, dtype=torch.bfloat16), torch.empty(4 * T, d, device=DEV)
    ops.embed_fwd(ids, wte, wpe, x, drop=ops.Drop(rng, 0, 0.1))
    ref.embed_fwd(ids.long(), wte, wpe, rx, ops.Drop(rng, 0, 0.1))
    _close(x, rx)
    dx = _bf(4 * T, d)
    g1, g2 = torch.zeros(V, d, device=DEV), torch.zeros(T, d, device=DEV)
    r1, r2 = torch.zeros(V, d, device=DEV), torch.zeros(T, d, device=DEV)
    ops.embed_bwd(dx, ids, g1, g2, drop=ops.Drop(rng, 0, 0.1))
    ref.embed_bwd(dx, ids.long(), r1, r2, ops.Drop(rng, 0, 0.1))
    _close(g1, r1, rtol=1e-4); _close(g2, r2, rtol=1e-4)


@pytest.mark.parametrize("d", [768, 1024, 128])
def test_layernorm_bwd_masked_copy_and_bias_fold(d):
    torch.manual_seed(12)
    M = 515
    rng = ops.DropoutRng(DEV, seed=9)
    rng.advance()
    x, w = _bf(M, d) + 0.3, _bf(d) + 1.0
    mean, rstd = x.float().mean(-1), torch.rsqrt(x.float().var(-1, unbiased=False) + 1e-5)
    dy, dres = _bf(M, d), _bf(M, d)
    for drop in (None, ops.Drop(rng, 7, 0.1)):
        dx, dxm = torch.empty(M, d, device=DEV, dtype=torch.bfloat16), torch.zeros(M, d, device=DEV, dtype=torch.bfloat16)
        dw, db, dcol = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
        ops.layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, dres, dcol, dxm if drop else None, drop)
        rdx, rdw, rdb = torch.empty(M, d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
        ref.layernorm_bwd(dy, x, w, mean, rstd, rdx, rdw, rdb, dres)
        _close(dx, rdx, rtol=1e-2); _close(dw, rdw, rtol=1e-3); _close(db, rdb, rtol=1e-3)
        if drop is None:
            _close(dcol, dx.float().sum(0), rtol=2e-3)
        else:
            want = dx.float() * ref.drop_mult_2d(rng.state, 7, 0.1, M, d, DEV)
            _close(dxm, want, rtol=1e-2)
            _close(dcol, want.sum(0), rtol=5e-3)


def test_gemm_colsum_fold():
    """Bias gradient folded into the producing GEMM's epilogue == column sums of the stored bf16 output."""
    torch.manual_seed(13)
    M, N = 1000, 3072  # dY [M, 768] x W [768, 3072] (B MN-major) -> du [M, 3072], as in the MLP backward
    a, w, u = _bf(M, 768), _bf(768, N, scale=0.05), _bf(M, N)
    out, r = torch.empty(M, N, device=DEV, dtype=torch.bfloat16), torch.empty(M, N, device=DEV)
    cs = torch.zeros(N, device=DEV)
    ops.gemm(a, w, out, b_mn=True, epi="dgelu", aux=u, colsum_out=cs)
    ref.gemm(a, w, r, b_mn=True, epi="dgelu", aux=u)
    _close(out, r)
    _close(cs, out.float().sum(0), rtol=1e-4)
    cs2 = torch.zeros(N, device=DEV)
    ops.gemm(a, w, out, b_mn=True, colsum_out=cs2)   # plain epilogue, no aux
    _close(cs2, out.float().sum(0), rtol=1e-4)


@pytest.mark.parametrize("epi", ["none", "bias", "bias_gelu", "bias_resid", "resid", "dgelu", "bias_resid_dropout", "f32_accumulate"])
def test_gemm_two_sm_pairs_with_odd_tile_count(epi):
    """Shapes large enough for the 2-SM UMMA path (tiles >= 32) with an ODD number of M-tiles (33): the second CTA of the last
    pair owns an out-of-range tile (TMA zero-fills its operands and clips its stores).  Every epilogue, the dropout epilogue
    and the fp32 split-K accumulate path."""
    torch.manual_seed(21)
    M, N, K = 33 * 128 - 5, 768, 320
    A, B = _bf(M, K, scale=0.5), _bf(N, K, scale=0.5)
    bias, aux = _bf(N), _bf(M, N)
    if epi == "f32_accumulate":      # wgrad form: out[K1, K2] += X^T Y over the M tokens
        X, Y = _bf(M, 4352, scale=0.3), _bf(M, 512, scale=0.3)     # 34 x 2 = 68 tiles of the [4352, 512] output
        out = torch.ones(4352, 512, device=DEV)
        ops.gemm(X, Y, out, a_mn=True, b_mn=True, accumulate=True)
        _close(out, 1.0 + X.float().t() @ Y.float(), rtol=1e-2)
        return
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    r = torch.empty(M, N, device=DEV)
    if epi == "bias_resid_dropout":
        rng = ops.DropoutRng(DEV, seed=5)
        rng.advance()
        dr = ops.Drop(rng, 9, 0.1)
        ops.gemm(A, B, out, epi="bias_resid", bias=bias, aux=aux, drop=dr)
        ref.gemm(A, B, r, epi="bias_resid", bias=bias, aux=aux, drop=dr)
        _close(out, r, rtol=1e-2)
        dropped = ref.drop_mult_2d(rng.state, 9, 0.1, M, N, DEV) == 0
        assert torch.equal(out[dropped], aux[dropped])
        return
    out2, r2 = torch.empty_like(out), torch.empty(M, N, device=DEV)
    ops.gemm(A, B, out, epi=epi, bias=bias, aux=aux, out2=out2 if epi == "bias_gelu" else None)
    ref.gemm(A, B, r, epi=epi, bias=bias, aux=aux, out2=r2)
    _close(out, r, rtol=1e-2)
    if epi == "bias_gelu":
        _close(out2, r2, rtol=1e-2)


def test_meta_kernels_single_gpu():
    """csrc/meta_avg.cu on one device against plain torch: delta shard transpose (fp32/bf16/fp8), active-mask weighted
    average, segmented multi-dot with R summed gradient arenas over a chunk range, w update from slot tables."""
    torch.manual_seed(11)
    man = _toy_manifest()
    n, P, N = man.total, len(man), 5
    tid = man.tensor_ids(DEV)
    cs, cl, _ = man.seg_table("cpu")
    nch = cs.numel()
    c0, c1 = nch // 3, nch - 1
    e0, e1 = int(cs[c0]), int(cs[c1 - 1]) + int(cl[c1 - 1])
    base = torch.randn(n, device=DEV)
    true = [torch.randn(n, device=DEV) * 0.1 for _ in range(N)]
    active = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int32, device=DEV)
    # ---- transpose: typed windows -> local fp32 shards (inactive miners -> zeros) ----
    for mode, name in [(0, "fp32"), (1, "bf16"), (2, "fp8")]:
        scales = None
        if name == "fp32":
            srcs, dec = true, true
        elif name == "bf16":
            srcs = [t.bfloat16() for t in true]
            dec = [s.float() for s in srcs]
        else:
            srcs, scales = [], []
            for t in true:
                q, sc = torch.empty(n, device=DEV, dtype=torch.uint8), torch.empty(n // 32, device=DEV)
                ops.delta_emit(t, torch.zeros_like(t), q, sc)
                srcs.append(q); scales.append(sc)
            dec = [ops.dequant_fp8(q, s) for q, s in zip(srcs, scales)]
        dT = torch.full((N, e1 - e0), 7.0, device=DEV)
        ops.shard_transpose(srcs, scales, [dT[i] for i in range(N)], active, e0, e1, mode)
        for i in range(N):
            want = dec[i][e0:e1] if int(active[i]) else torch.zeros(e1 - e0, device=DEV)
            assert torch.equal(dT[i], want), (name, i)
    # ---- active-mask weighted average over the shard, from virtual (global-index) pointers ----
    dT = torch.stack([t[e0:e1].clone() for t in true])
    dT[2] = float("nan")  # an inactive miner's data must never be touched
    vptr = [dT[i].data_ptr() - 4 * e0 for i in range(N)]
    w = torch.rand(N, P, device=DEV) - 0.2
    out = torch.zeros(n, device=DEV)
    out16 = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    ops.weighted_avg(base, vptr, w, man, [out], [out16], chunk_range=(c0, c1), mode=0, active=active)
    keep = [0, 1, 3, 4]
    want = base * w[keep].sum(0)[tid] + sum(w[i][tid] * true[i] for i in keep)
    _close(out[e0:e1], want[e0:e1], rtol=1e-5)
    _close(out16[e0:e1], want[e0:e1], rtol=1e-2)
    assert float(out[:e0].abs().max()) == 0.0 and float(out[e1:].abs().max()) == 0.0
    # ---- segmented multi-dot: 3 gradient arenas summed with scales, chunk range, two destination tables ----
    gs = [torch.randn(n, device=DEV) for _ in range(3)]
    gsc = [0.5, 0.25, 2.0]
    loss = torch.tensor(3.25, device=DEV)
    tabs = [torch.full((N * P + 1,), -1.0, device=DEV) for _ in range(2)]
    dT[2] = 0.0
    ops.seg_dot(gs, vptr, base, out, man, tabs, N=N, gscales=gsc, chunk_range=(c0, c1), active=active, loss=loss, loss_scale=0.5)
    gsum = sum(s * g for s, g in zip(gsc, gs))
    sel = torch.zeros(n, dtype=torch.bool, device=DEV)
    sel[e0:e1] = True
    rG = torch.zeros(N, P, device=DEV)
    for i in range(N):
        if int(active[i]):
            rG[i].index_add_(0, tid[sel], (gsum * (true[i] + base - out))[sel])
    for tb in tabs:
        _close(tb[:N * P].view(N, P), rG, rtol=2e-3)
        assert float(tb[N * P]) == 3.25 * 0.5
    assert torch.equal(tabs[0], tabs[1])
    # ---- w update: fixed-order sum of the slot tables ----
    w0 = w.clone()
    acc = torch.zeros(2, device=DEV)
    ops.w_update(tabs, w, 0.01, loss_acc=acc)
    _close(w, w0 - 0.01 * 2 * tabs[0][:N * P].view(N, P), rtol=1e-5)
    assert float(acc[0]) == 3.25 and float(acc[1]) == 3.25
    # ---- round prepare: flags already satisfied, NaN verdict of miner 1 ----
    flags = torch.zeros(64, dtype=torch.int32, device=DEV)
    flags[:N] = 4            # publish flags: round 4
    flags[32 + 1] = 4        # miner 1 flagged its round-4 delta as bad
    flags[32 + 3] = 3        # an OLD verdict of miner 3 does not count
    act, nact = torch.zeros(N, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    w2 = torch.zeros(N, P, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.round_prepare([flags.data_ptr() + 4 * i for i in range(N)], [flags.data_ptr() + 4 * (32 + i) for i in range(N)], 4, act, nact,
                      w2, True, err)
    assert act.tolist() == [1, 0, 1, 1, 1] and int(nact) == 4 and int(err) == 0
    assert torch.equal(w2[0], torch.full((P,), 0.25, device=DEV)) and float(w2[1].abs().max()) == 0.0


def test_peer_meta_learner_world1_matches_sequential():
    """DistributedMetaLearner on the peer back-end with a single rank (no torch.distributed): every kernel of the step
    sequence runs (windows, flags, transposes, slots) and must reproduce the sequential reference loop."""
    from distributedtraining_b200.models.trainer import Trainer
    from distributedtraining_b200.parallel.exchange import PeerExchange
    from distributedtraining_b200.parallel.meta import DistributedMetaLearner
    torch.manual_seed(3)
    tr = Trainer("gpt2-tiny", device=DEV, batch=8, seq=64, lr=1e-3, seed=0, use_graph=False)
    for _ in range(3):
        tr.step(torch.randint(0, tr.cfg.vocab_size - 1, (8, 64), dtype=torch.int32, device=DEV))
    ex = PeerExchange(tr.man)
    ex.publish_delta(tr, 1)
    delta = (tr.master - tr.base).clone()
    V = tr.cfg.vocab_size
    val = []
    for s in range(2):
        ids = torch.randint(0, V - 1, (4, 64), dtype=torch.int32, device=DEV)
        kv = torch.tensor([64, 20, 1, 50], dtype=torch.int32, device=DEV)
        val.append({"input_ids": ids, "labels": ids.clone(), "kv_len": kv})
    t2 = Trainer(tr.cfg, device=DEV, batch=8, seq=64, seed=0, init_flat=tr.base.clone(), use_graph=False)
    ml = DistributedMetaLearner(t2, ex, [0], val, meta_lr=0.05)
    ml.begin_round(1, reset_w=False)
    ml.w.fill_(0.7)  # a single miner at w = 1 has a zero meta-gradient by construction (theta_bar == theta_1): start off it
    ml.run(2)
    ref_t = Trainer(tr.cfg, device=DEV, batch=4, seq=64, seed=0, init_flat=tr.base.clone(), use_graph=False)
    P = len(tr.man)
    w = torch.full((1, P), 0.7, device=DEV)
    G = torch.empty(1, P, device=DEV)
    for _ in range(4):
        for b in val:
            ops.weighted_avg(ref_t.base, [delta], w, ref_t.man, [ref_t.master], [ref_t.p16])
            ref_t.loss_and_grad(b)
            ops.multi_dot(ref_t.grad, [delta], ref_t.base, ref_t.master, ref_t.man, G)
            w.add_(G, alpha=-0.05)
    moved = float((w - 0.7).abs().max())
    assert moved > 1e-6
    assert float((ml.w - w).abs().max()) < 0.05 * moved + 1e-6
    ml.final_average_shard(1)
    want = torch.empty_like(ref_t.master)
    ops.weighted_avg(ref_t.base, [delta], ml.w, ref_t.man, [want])
    got = ex.win.local("base", torch.float32)[:tr.man.total]
    assert float((got - want).abs().max()) < 1e-6
    ex.win.check_errors()
    ex.win.close()
