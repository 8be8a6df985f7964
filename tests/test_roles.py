"""Role logic on CPU with the disk exchange -- the analogue of the reference's ``Local*`` simulation harness
(SURVEY.md section 4): delta round trip, validator scoring maths, learned averaging vs a literal port of the reference
algorithm, genetic mixer, fault injection (missing / NaN / wrong-shape miner), EMA, checkpoint/resume."""
import copy
import math
import os

import pytest
import torch
import torch.nn.functional as F

from distributedtraining_b200 import ops
from distributedtraining_b200.averaging_logic import (DeltaAverager, GeneticAverager, LocalParameterizedAverager,
                                                      ParameterizedAverager)
from distributedtraining_b200.btt_connector import BittensorNetwork, LocalBittensorNetwork, MemoryLedger
from distributedtraining_b200.chain_manager import ChainMultiAddressStore, LocalAddressStore
from distributedtraining_b200.config import Configurator
from distributedtraining_b200.data import SyntheticMNIST, SyntheticTokens
from distributedtraining_b200.hf_manager import HFManager, LocalHFManager
from distributedtraining_b200.models.toys import FeedforwardNN, ModuleTrainer
from distributedtraining_b200.models.trainer import Trainer
from distributedtraining_b200.parallel.exchange import DiskExchange
from distributedtraining_b200.training_manager import DeltaLoop, MNISTDeltaTrain, MNISTTrain
from distributedtraining_b200.validation_logic import DeltaValidator, MNISTDeltaValidator


def _net(hotkeys, me="rank0", epoch_length=0, validators=()):
    cfg = Configurator.combine_configs([])
    cfg.wallet.hotkey = me
    cfg.neuron.epoch_length = epoch_length
    stakes = [10000.0 if h in validators else 10.0 for h in hotkeys]
    BittensorNetwork.initialize(cfg, ignore_regs=True, ledger=MemoryLedger(), hotkeys=list(hotkeys), stakes=stakes)
    BittensorNetwork.last_set_block -= 10_000  # allow set_weights immediately
    return BittensorNetwork


def _toy_miner(seed, lr=0.05, steps=30, data_seed=0):
    torch.manual_seed(seed)
    t = ModuleTrainer(FeedforwardNN(784, (64, 32), 10), lr=lr, optimizer="sgd")
    return t


def _sync(trainers):
    for t in trainers[1:]:
        t.master.copy_(trainers[0].master)
        t.base.copy_(trainers[0].master)
    trainers[0].base.copy_(trainers[0].master)


def test_delta_push_pull_roundtrip_disk(tmp_path):
    torch.manual_seed(0)
    tr = Trainer("gpt2-tiny", batch=2, seq=16, lr=1e-2)
    hf = LocalHFManager(my_repo_id=str(tmp_path / "m0"), averaged_model_repo_id=str(tmp_path / "hub"), manifest=tr.man, rank=0)
    ids = torch.randint(0, 512, (2, 16), dtype=torch.int32)
    for _ in range(3):
        tr.step(ids)
    hf.push_changes(trainer=tr)
    got = hf.receive_gradients("disk://0")
    assert set(got) == set(tr.man.names) and len(got) == len(tr.man)
    want = tr.man.views(tr.master - tr.base)
    for k in got:
        assert torch.allclose(got[k], want[k], atol=1e-7)
    assert hf.receive_gradients("disk://7") is None  # missing miner -> None (reference hf_manager.py:196-197)
    # base publish / poll / adopt
    assert not hf.check_for_new_submissions()
    new_base = tr.base + 0.5 * (tr.master - tr.base)
    hf.push_to_hf_hub(base=new_base)
    hf2 = LocalHFManager(my_repo_id=str(tmp_path / "m1"), averaged_model_repo_id=str(tmp_path / "hub"), manifest=tr.man, rank=1)
    assert not hf2.check_for_new_submissions()  # constructed after the push: records the current version first
    hf.last_known_hash = None
    assert hf.check_for_new_submissions()
    hf.pull_latest_model()
    hf.update_model(tr, lr=5e-5)
    assert torch.allclose(tr.master, new_base) and torch.allclose(tr.base, new_base) and tr.opt.host["lr"] == 5e-5


def test_delta_loop_rounds_and_base_adoption(tmp_path):
    torch.manual_seed(0)
    tr = Trainer("gpt2-tiny", batch=2, seq=16, lr=1e-2)
    ex = DiskExchange(str(tmp_path / "hub"), 0, tr.man)
    hf = HFManager(local_dir=str(tmp_path), my_repo_id="disk://0", averaged_model_repo_id="avg", exchange=ex, manifest=tr.man)
    data = SyntheticTokens(2, 16, 512, pad_id=511, steps=12, pin=False)
    loop = DeltaLoop("cpu", "gpt2-tiny", data, learning_rate=1e-2, hf_manager=hf, trainer=tr, local_steps=4, max_steps=12)
    loop.train(1)
    assert loop.rounds_sent == 3 and ex.delta_round(0) == 3
    # deltas are cumulative w.r.t. the last pulled base (theta_base NOT reset on send, reference :405-431)
    d = ex.fetch_delta(0, 3)
    assert torch.allclose(tr.base + d, tr.master, atol=1e-6)


def test_validator_scores_and_guards(tmp_path):
    torch.manual_seed(0)
    hotkeys = ["rank0", "rank1", "rank2", "rank3"]
    net = _net(hotkeys, me="rank3", validators=("rank3",))
    data = SyntheticMNIST(n=256, batch=64, seed=1)
    val = list(SyntheticMNIST(n=128, batch=64, seed=2))
    base = _toy_miner(0)
    man = base.man
    ex = [DiskExchange(str(tmp_path / "hub"), r, man) for r in range(4)]
    chain = ChainMultiAddressStore(net.ledger, 1, net.wallet)
    for r in range(3):
        net.ledger.put(f"commit/1/rank{r}", f"disk://{r}")
    # rank0: a genuinely trained miner; rank1: a harmful delta; rank2: never publishes
    good = _toy_miner(0)
    _sync([base, good])
    for x, y in data:
        good.step((x, y))
    ex[0].publish_delta(good, 1)
    bad = _toy_miner(0)
    _sync([base, bad])
    bad.master.add_(torch.randn_like(bad.master) * 0.5)
    ex[1].publish_delta(bad, 1)
    hf = HFManager(local_dir=str(tmp_path), averaged_model_repo_id="avg", exchange=ex[3], manifest=man)
    v = MNISTDeltaValidator("cpu", base, None, val, net, hf, chain_manager=chain)
    scores = v.validate_and_score()
    assert v.loss_scores["rank0"] > 0 and v.scores["rank1"] == 0.0 and v.scores["rank2"] == 0.0
    assert abs(sum(scores.values()) - 1.0) < 1e-6 and scores["rank0"] == pytest.approx(1.0)
    assert torch.equal(base.master, base.base)  # theta_base restored after every miner (reference :139)
    # EMA 0.333333 written to the ledger (reference btt_connector.py:317-318)
    assert float(net.base_scores[0]) == pytest.approx(0.333333, rel=1e-4)
    # nobody improves -> no division by zero (reference :186-187 would raise)
    ex[0].publish_delta(bad, 2)
    scores = v.validate_and_score()
    assert all(s == 0.0 for s in scores.values())


def _reference_meta_learning(base_sd, deltas, model, val, meta_epochs, lr):
    """Literal dict-of-tensors port of reference averaging_logic.py:422-448, 490-541 (the oracle for our fused maths)."""
    names = [n for n, _ in model.named_parameters()]
    N, P = len(deltas), len(names)
    w = torch.softmax(torch.ones(N, P), dim=0)
    def averaged():
        out = [torch.zeros_like(base_sd[n]) for n in names]
        for i, d in enumerate(deltas):
            for j, n in enumerate(names):
                out[j] += (base_sd[n] + d[n]) * w[i, j]
        return out
    for _ in range(meta_epochs):
        for _ in range(meta_epochs):
            for x, y in val:
                avg = averaged()
                for p, a in zip(model.parameters(), avg):
                    p.data.copy_(a)
                model.zero_grad()
                loss = F.cross_entropy(model(x), y)
                loss.backward()
                G = torch.zeros(N, P)
                for i, d in enumerate(deltas):
                    for j, (n, p) in enumerate(zip(names, model.parameters())):
                        G[i, j] = torch.sum(p.grad * ((base_sd[n] + d[n]) - p.data))
                w = w - lr * G
    return w, averaged()


def test_parameterized_averager_equals_reference_port(tmp_path):
    torch.manual_seed(0)
    hotkeys = ["rank0", "rank1", "rank2"]
    net = _net(hotkeys)
    data = [SyntheticMNIST(n=128, batch=32, seed=10 + i) for i in range(3)]
    val = list(SyntheticMNIST(n=64, batch=32, seed=99))
    avg_t = _toy_miner(0)
    miners = [_toy_miner(0) for _ in range(3)]
    _sync([avg_t] + miners)
    man = avg_t.man
    ex = [DiskExchange(str(tmp_path / "hub"), r, man) for r in range(3)]
    for r, m in enumerate(miners):
        for x, y in data[r]:
            m.step((x, y))
        ex[r].publish_delta(m, 1)
        net.ledger.put(f"commit/1/rank{r}", f"disk://{r}")
    chain = ChainMultiAddressStore(net.ledger, 1, net.wallet)
    hf = HFManager(local_dir=str(tmp_path), averaged_model_repo_id="avg", exchange=ex[0], manifest=man)
    pa = ParameterizedAverager(avg_t, "cpu", hf, str(tmp_path / "model"), str(tmp_path / "grads"), chain, net, cache_to_disk=True)
    assert pa.cache_params_locally() == 3
    assert torch.allclose(pa._ensure_weights(), torch.full((3, len(man)), 1 / 3))
    # oracle
    ref_model = FeedforwardNN(784, (64, 32), 10)
    base_sd = {n: man.view(avg_t.base, n).clone() for n in man.names}
    deltas = [{n: man.view(m.master - m.base, n).clone() for n in man.names} for m in miners]
    w_ref, avg_ref = _reference_meta_learning(base_sd, deltas, ref_model, val, 2, 0.01)
    pa.meta_learning(val, 2, 0.01)
    assert torch.allclose(pa.weights, w_ref, atol=2e-5)
    for n, a in zip(man.names, avg_ref):
        assert torch.allclose(man.view(avg_t.master, n), a, atol=2e-5)
    # lazy_load_params parity + disk cache round trip
    th = list(pa.lazy_load_params())
    assert torch.allclose(th[1], miners[1].master, atol=1e-6)
    assert torch.allclose(pa.load_weight_delta("rank2"), miners[2].master - miners[2].base, atol=1e-7)
    path = pa.save_model()
    assert set(torch.load(path, weights_only=False)) == set(man.names)


def test_averager_fault_injection(tmp_path):
    """missing miner, NaN delta, wrong-shape delta are skipped, never crash the round (reference would: SURVEY 2.7)."""
    torch.manual_seed(0)
    hotkeys = [f"rank{r}" for r in range(4)]
    net = _net(hotkeys)
    t = _toy_miner(0)
    man = t.man
    ex = [DiskExchange(str(tmp_path / "hub"), r, man) for r in range(4)]
    for r in range(4):
        net.ledger.put(f"commit/1/rank{r}", f"disk://{r}")
    good = _toy_miner(0); _sync([t, good]); good.master.add_(0.01)
    ex[0].publish_delta(good, 1)
    nan = _toy_miner(0); _sync([t, nan]); nan.master[5] = float("nan")
    ex[1].publish_delta(nan, 1)
    other = ModuleTrainer(FeedforwardNN(784, (16,), 10))  # wrong architecture
    DiskExchange(str(tmp_path / "hub"), 2, other.man).publish_delta(other, 1)
    # rank3 publishes nothing
    hf = HFManager(local_dir=str(tmp_path), averaged_model_repo_id="avg", exchange=ex[0], manifest=man)
    pa = ParameterizedAverager(t, "cpu", hf, str(tmp_path / "m"), str(tmp_path / "g"), ChainMultiAddressStore(net.ledger, 1, net.wallet), net)
    assert pa.cache_params_locally() == 1 and pa.miner_hotkeys == ["rank0"]
    pa.get_averaged_model()
    assert torch.allclose(t.master, good.master, atol=1e-6)


def test_delta_and_genetic_averagers(tmp_path):
    torch.manual_seed(0)
    hotkeys = ["rank0", "rank1", "rank2"]
    net = _net(hotkeys, me="rank2", validators=("rank2",))
    val = list(SyntheticMNIST(n=64, batch=32, seed=5))
    t = _toy_miner(0)
    man = t.man
    miners = [_toy_miner(0), _toy_miner(0)]
    _sync([t] + miners)
    ex = [DiskExchange(str(tmp_path / "hub"), r, man) for r in range(3)]
    for r, m in enumerate(miners):
        for x, y in SyntheticMNIST(n=128, batch=32, seed=20 + r):
            m.step((x, y))
        ex[r].publish_delta(m, 1)
        net.ledger.put(f"commit/1/rank{r}", f"disk://{r}")
    chain = ChainMultiAddressStore(net.ledger, 1, net.wallet)
    hf = HFManager(local_dir=str(tmp_path), averaged_model_repo_id="avg", exchange=ex[2], manifest=man)
    # v1 score-weighted: validator row says miner0 = 1.0, miner1 = 0.5
    net.set_weights({"rank0": 1.0, "rank1": 0.5})
    da = DeltaAverager(t, str(tmp_path / "m"), net, chain, hf)
    grads, scores = da.receive_and_score_gradients()
    avg = da.average_gradients(grads, scores)
    s0, s1 = float(scores[0]), float(scores[1])
    want = ((t.base + (miners[0].master - miners[0].base)) * s0 + (t.base + (miners[1].master - miners[1].base)) * s1) / 2
    assert torch.allclose(avg, want, atol=1e-6)
    da.apply_averaged_gradients(avg)
    assert torch.allclose(t.master, want, atol=1e-6)  # really applied (reference :244-248 is a no-op)
    t.master.copy_(t.base)
    ga = GeneticAverager(t, "cpu", hf, str(tmp_path / "m"), str(tmp_path / "g"), chain, net, population_size=6,
                         num_generations=3, seed=1)
    assert ga.cache_params_locally() == 2
    uniform_loss = -float(ga.evaluate_population(torch.ones(1, 2), val)[0])
    ga.run_evolution(val)
    best_loss = -float(ga.evaluate_population(ga.best_weights[None], val)[0])
    assert best_loss <= uniform_loss + 1e-6


def test_mnist_simulation_loops(tmp_path):
    torch.manual_seed(0)
    train, test = SyntheticMNIST(n=256, batch=32, seed=1), SyntheticMNIST(n=64, batch=32, seed=2)
    m = MNISTTrain(FeedforwardNN(784, (64,), 10), lr=0.1, train_loader=train, test_loader=test)
    tr_loss, te_loss, te_acc = m.train(epochs=1, n_steps=4)
    assert math.isfinite(tr_loss) and 0.0 <= te_acc <= 1.0
    d = MNISTDeltaTrain(FeedforwardNN(784, (64,), 10), lr=0.1, train_loader=train, test_loader=test, send_every=4)
    hf = LocalHFManager(my_repo_id=str(tmp_path / "m0"), averaged_model_repo_id=str(tmp_path / "hub"), manifest=d.trainer.man)
    losses = d.train(epochs=2, hf_manager=hf, max_steps=16)
    assert losses[-1] < losses[0] and hf.exchange.delta_round(0) == 4
    assert len(d.calculate_model_hash()) == 64


def test_local_network_and_address_store(tmp_path):
    cfg = Configurator.combine_configs([])
    cfg.wallet.hotkey = "simulated_hotkey_3"
    LocalBittensorNetwork.initialize(cfg, n=20, path=str(tmp_path / "bt" / "metagraph.json"))
    net = LocalBittensorNetwork
    assert net.metagraph.n == 20 and net.get_validator_uids(1024) == [19]
    store = LocalAddressStore(None, 1, net.wallet, path=str(tmp_path / "storage.json"))
    store.store_hf_repo(str(tmp_path / "repo3"))
    assert store.retrieve_hf_repo("simulated_hotkey_3") == str(tmp_path / "repo3") and store.retrieve_hf_repo("x") is None
    # anomaly screen + rate limiter (reference btt_connector.py:388-480)
    net.metrics_data = {f"h{i}": {"loss": 2.0 + 0.01 * i} for i in range(8)}
    net.metrics_data["evil"] = {"loss": 50.0}
    flags = net.detect_metric_anomaly()
    assert flags["evil"] and not flags["h3"]
    assert all(net.rate_limiter("addr", n=3, t=60) for _ in range(3)) and not net.rate_limiter("addr", n=3, t=60)
    # the reference's simulation-layer names resolve to the same objects (btt_connector.py:514-585)
    from distributedtraining_b200.btt_connector import LocalHotkey, LocalMetagraph, LocalWallet
    assert isinstance(net.metagraph, LocalMetagraph) and isinstance(net.wallet, LocalWallet)
    assert LocalHotkey("simulated_hotkey_3").ss58_address == net.wallet.hotkey.ss58_address


def test_checkpoint_resume(tmp_path):
    from distributedtraining_b200.utils.checkpoint import latest_checkpoint, load_checkpoint, save_checkpoint
    cfg = Configurator.combine_configs(["--checkpoint_dir", str(tmp_path / "ck")])
    tr = Trainer("gpt2-tiny", batch=2, seq=16, lr=1e-2)
    ids = torch.randint(0, 512, (2, 16), dtype=torch.int32)
    for _ in range(3):
        tr.step(ids)
    save_checkpoint(cfg, tr, 0, 7, extra={"w": torch.ones(2, 3)})
    tr2 = Trainer("gpt2-tiny", batch=2, seq=16, lr=1.0, seed=5)
    blob = load_checkpoint(latest_checkpoint(cfg, 0), tr2)
    assert blob["round"] == 7 and torch.equal(tr2.master, tr.master) and torch.equal(tr2.m, tr.m)
    assert tr2.opt.host_step == 3 and tr2.opt.host["lr"] == pytest.approx(1e-2)
    a, b = float(tr.step(ids)), float(tr2.step(ids))
    assert abs(a - b) < 1e-6


def test_delta_wire_formats_preserve_validator_ranking():
    """bf16 and block-scaled fp8 deltas (the wire formats of the peer plane) must not change how a validator ranks miners:
    score_i = max(0, loss_base - loss(theta_base + delta_i)) computed from fp32, bf16 and fp8 deltas orders the miners alike
    (SURVEY.md 7.4.6)."""
    from distributedtraining_b200 import ops
    torch.manual_seed(0)
    val = torch.randint(0, 512, (4, 16), dtype=torch.int32)
    base_tr = Trainer("gpt2-tiny", batch=4, seq=16, lr=1e-2, seed=3)
    base = base_tr.base.clone()
    loss_base = float(base_tr.eval_loss(val))
    scores = {"fp32": [], "bf16": [], "fp8": []}
    for steps in (1, 4, 12):                                  # three miners that trained for different amounts on the val batch
        m = Trainer("gpt2-tiny", batch=4, seq=16, lr=1e-2, seed=3)
        for _ in range(steps):
            m.step(val)
        n = m.master.numel()
        d32 = torch.empty(n); m.emit_delta(d32)
        d16 = torch.empty(n, dtype=torch.bfloat16); m.emit_delta(d16)
        q8, sc = torch.empty(n, dtype=torch.uint8), torch.empty(n // 32); m.emit_delta(q8, sc)
        for name, d in (("fp32", d32), ("bf16", d16.float()), ("fp8", ops.dequant_fp8(q8, sc))):
            base_tr.master.copy_(base + d)
            scores[name].append(max(0.0, loss_base - float(base_tr.eval_loss(val))))
    order = lambda xs: sorted(range(len(xs)), key=lambda i: xs[i])
    assert order(scores["fp32"]) == order(scores["bf16"]) == order(scores["fp8"]) == [0, 1, 2], scores
    for a, b in zip(scores["fp32"], scores["fp8"]):
        assert abs(a - b) < 0.05 * max(a, 1e-3) + 1e-3, scores
