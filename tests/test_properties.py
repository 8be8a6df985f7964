"""Property tests (hypothesis) of the layout / algebra invariants every plane relies on.  CPU only."""
import math

import torch
from hypothesis import given, settings, strategies as st

from distributedtraining_b200 import ops
from distributedtraining_b200.models.arena import ALIGN, CHUNK, Manifest
from distributedtraining_b200.ops import reference as ref
from distributedtraining_b200.parallel.launch import parse_roles

shapes = st.lists(st.tuples(st.integers(1, 70), st.integers(1, 40)), min_size=1, max_size=7)


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_manifest_layout_invariants(shs):
    man = Manifest([(f"t{i}", s, "normal", True) for i, s in enumerate(shs)])
    flat = torch.arange(man.total, dtype=torch.float32)
    seen = torch.zeros(man.total, dtype=torch.int32)
    for s in man:
        assert s.offset % ALIGN == 0 and man.view(flat, s.name).shape == s.shape
        seen[s.offset:s.offset + s.numel] += 1
    assert int(seen.max()) == 1 and int(seen.sum()) == man.num_params          # views never overlap
    cs, cl, ct = man.seg_table("cpu")
    assert int(cl.sum()) == man.total and int(cl.max()) <= CHUNK                # chunks tile the arena exactly ...
    tid = man.tensor_ids("cpu")
    for c in range(cs.numel()):                                                  # ... and never straddle two tensors
        seg = tid[int(cs[c]):int(cs[c]) + int(cl[c])]
        assert int(seg.min()) == int(seg.max()) == int(ct[c])
    assert man.same_layout({s.name: s.shape for s in man}) and not man.same_layout({s.name: (1,) for s in man})


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_weighted_average_identity(shs, n_miners, seed):
    """theta_bar = s_j * base + sum_i w_ij * delta_ij  ==  sum_i w_ij * (base + delta_ij)   (SURVEY 2.6-C.2), and the
    meta-gradient dots G_ij = <g_j, base_j + delta_ij - theta_bar_j> match the per-tensor definition."""
    g = torch.Generator().manual_seed(seed)
    man = Manifest([(f"t{i}", s, "normal", True) for i, s in enumerate(shs)])
    P = len(man)
    valid = torch.zeros(man.total)
    for s in man:                                    # arenas keep their alignment padding at zero (init, AdamW, emit)
        valid[s.offset:s.offset + s.numel] = 1.0
    base = torch.randn(man.total, generator=g) * valid
    deltas = [0.1 * torch.randn(man.total, generator=g) * valid for _ in range(n_miners)]
    w = torch.rand(n_miners, P, generator=g)
    out = torch.empty(man.total)
    ops.weighted_avg(base, deltas, w, man, [out])
    grad = torch.randn(man.total, generator=g) * valid
    G = torch.empty(n_miners, P)
    ops.multi_dot(grad, deltas, base, out, man, G)
    for j, s in enumerate(man):
        want = sum(w[i, j] * (man.view(base, j) + man.view(deltas[i], j)) for i in range(n_miners))
        assert torch.allclose(man.view(out, j), want, atol=1e-5)
        for i in range(n_miners):
            gij = (man.view(grad, j) * (man.view(base, j) + man.view(deltas[i], j) - man.view(out, j))).sum()
            assert math.isclose(float(G[i, j]), float(gij), rel_tol=1e-3, abs_tol=1e-3)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 64), st.integers(0, 2 ** 31 - 1))
def test_block_fp8_delta_roundtrip_error_bound(nblocks, seed):
    g = torch.Generator().manual_seed(seed)
    n = 32 * nblocks
    base = torch.randn(n, generator=g)
    master = base + torch.randn(n, generator=g) * torch.rand(1, generator=g) * 0.1
    q, sc = torch.empty(n, dtype=torch.uint8), torch.empty(nblocks)
    ops.delta_emit(master, base, q, sc)
    back = ops.dequant_fp8(q, sc)
    d = master - base
    amax = d.view(-1, 32).abs().amax(1, keepdim=True).expand(-1, 32).reshape(-1)
    assert bool(((back - d).abs() <= amax / 16 + 1e-12).all())       # e4m3: 3 mantissa bits -> half-ulp <= amax / 16 per block


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 1000), st.integers(0, 50), st.floats(0.01, 0.5))
def test_dropout_mask_is_a_pure_function_with_the_right_rate(seed, counter, stream, p):
    m1 = ref.drop_mult_2d((seed, counter), stream, p, 64, 128, "cpu")
    m2 = ref.drop_mult_2d((seed, counter), stream, p, 64, 128, "cpu")
    assert torch.equal(m1, m2)
    vals = set(m1.unique().tolist())
    assert vals <= {0.0, float(torch.tensor(1.0 / (1.0 - p), dtype=torch.float32))}
    keep = (m1 > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.03                                 # 8192 Bernoulli draws: 3 sigma < 0.017


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 16))
def test_parse_roles_covers_requested_ranks(world):
    roles = parse_roles(f"miner:0-{world - 1}", world)
    assert roles["miner"] == list(range(world))
    if world >= 3:
        r = parse_roles(f"miner:0-{world - 3},validator:{world - 2},averager:{world - 1}", world)
        assert r["validator"] == [world - 2] and r["averager"] == [world - 1] and len(r["miner"]) == world - 2


def test_dropout_hash_golden_values():
    """Pins the counter-based mask arithmetic (shared bit for bit with csrc/dropout.cuh; the GPU tests compare the kernels
    against THIS implementation, this test keeps the implementation itself from drifting)."""
    assert ref.drop_key((7, 1), 3) == 3313189177 and ref.drop_key((0x5EED, 5), 0) == 3387426996 and ref.drop_thr(0.1) == 6554
    m = ref.drop_mult_2d((7, 1), 3, 0.1, 8, 16, "cpu")
    assert (m == 0).nonzero().tolist() == [[0, 2], [0, 7], [0, 8], [1, 10], [1, 14], [2, 5], [2, 6], [2, 15], [3, 5], [3, 10],
                                            [4, 1], [4, 3], [6, 4], [6, 14], [6, 15], [7, 2], [7, 9], [7, 11]]
    a = ref.drop_mult_attn((7, 1), 4, 0.1, 1, 8, 2, "cpu")
    assert (a == 0).nonzero().tolist() == [[0, 0, 0, 3], [0, 0, 2, 2], [0, 0, 2, 3], [0, 0, 3, 2], [0, 0, 4, 0], [0, 0, 6, 6],
                                            [0, 1, 0, 3], [0, 1, 2, 4], [0, 1, 2, 5], [0, 1, 4, 7], [0, 1, 5, 3], [0, 1, 7, 6]]


# ---- attention kernel index arithmetic (csrc/sm100_attention.cu), mirrored in Python -------------------------------------
def _chunk_mask(c_lo, c_hi, ch):
    lo, hi = max(c_lo - ch * 32, 0), min(c_hi - ch * 32, 31)
    return 0 if lo > hi else ((0xFFFFFFFF >> (31 - hi)) & ((0xFFFFFFFF << lo) & 0xFFFFFFFF))


@settings(max_examples=300, deadline=None)
@given(st.integers(-700, 300), st.integers(-700, 300), st.integers(0, 3))
def test_attention_chunk_mask_equals_per_element_predicate(c_lo, c_hi, ch):
    """chunk_mask(): bit i set <=> c_lo <= ch*32 + i <= c_hi -- the predicate the kernels evaluated per element before."""
    vm = _chunk_mask(c_lo, c_hi, ch)
    for i in range(32):
        assert ((vm >> i) & 1) == int(c_lo <= ch * 32 + i <= c_hi)


def _lpt_block(L, grid_x, heads, T, M, heavy_last):
    nps = T // 128
    if nps <= 1 or T % 128 or M % T:
        return L % grid_x, L // grid_x
    per_class = (M // T) * heads
    cls, idx = divmod(L, per_class)
    return (idx // heads) * nps + (nps - 1 - cls if heavy_last else cls), idx % heads


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 6), st.sampled_from([64, 128, 256, 512, 1024]), st.integers(1, 12), st.booleans())
def test_attention_lpt_order_is_a_permutation_with_heaviest_blocks_first(nseq, T, heads, heavy_last):
    """lpt_block(): every (block, head) pair is visited exactly once, and the causal cost of the visited blocks never increases."""
    M = nseq * T
    grid_x = (M + 127) // 128
    seen, costs = set(), []
    for L in range(grid_x * heads):
        blk, h = _lpt_block(L, grid_x, heads, T, M, heavy_last)
        assert 0 <= blk < grid_x and 0 <= h < heads
        seen.add((blk, h))
        if T > 128:
            pos = blk % (T // 128)
            costs.append(pos + 1 if heavy_last else T // 128 - pos)
    assert len(seen) == grid_x * heads
    assert costs == sorted(costs, reverse=True)


def test_loss_mean_and_zero_cpu_fallbacks():
    import torch
    from distributedtraining_b200 import ops
    x = torch.arange(1, 11, dtype=torch.float32)
    out = torch.zeros(())
    ops.loss_mean(x, 0.1, out)
    assert abs(out.item() - 5.5) < 1e-6
    g = torch.ones(16)
    ops.zero_(g)
    assert float(g.abs().sum()) == 0.0
