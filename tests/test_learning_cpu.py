"""Learning step on torch-CPU: the torch mirror of the Flux ResNet against the independent
restatement oracle/netref.py, the loss and optimiser arithmetic against oracle/learnref.py (numpy Float64), and a short
training run.  No golden vectors exist in the reference for any of this (parity unpinned)."""
import numpy as np
import pytest
import torch

import _pkg


@pytest.fixture(scope="module")
def lrn():
    _pkg.load()
    import alphazero_jl_b200.learning as m
    return m


class HP:
    def __init__(self, nb=2):
        self.num_blocks, self.num_filters, self.conv_kernel_size = nb, 16, (3, 3)
        self.num_policy_head_filters, self.num_value_head_filters, self.batch_norm_momentum = 4, 4, 0.6

    def d(self):
        return dict(num_blocks=self.num_blocks, num_filters=self.num_filters, conv_kernel_size=self.conv_kernel_size,
                    num_policy_head_filters=self.num_policy_head_filters, num_value_head_filters=self.num_value_head_filters)


DIM, A = (7, 6, 3), 7


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.random((n, 126)) < 0.3).astype(np.float32)
    Am = (rng.random((n, A)) < 0.8).astype(np.float32)
    Am[:, 0] = 1
    P = rng.random((n, A)).astype(np.float32) * Am
    P /= P.sum(1, keepdims=True)
    V = rng.choice([-1.0, 0.0, 1.0], n).astype(np.float32)
    W = (np.log2(rng.integers(1, 9, n)) + 1).astype(np.float32)
    return dict(W=W, X=X, A=Am, P=P, V=V)


def test_blob_roundtrip_and_forward_vs_netref(lrn):
    from oracle import netref
    hp = HP()
    blob = netref.make_blob(DIM, A, hp.d(), seed=3, randomize=True)
    net = lrn.ResNetTorch(DIM, A, hp).load_blob(blob)
    assert (net.to_blob() == blob).all()
    d = _data(9)
    net.eval()
    with torch.no_grad():
        P, V = net(torch.from_numpy(d["X"]))
    Pr, Vr = netref.forward(blob, DIM, A, hp.d(), d["X"].reshape(-1, 3, 6, 7).transpose(0, 3, 2, 1))  # -> [B, W, H, C]
    assert np.abs(P.numpy() - Pr).max() < 2e-6 and np.abs(V.numpy() - Vr).max() < 2e-6
    with pytest.raises(ValueError):
        net.load_blob(blob[:-1])


def test_losses_match_restatement(lrn):
    from oracle import learnref
    hp = HP(1)
    net = lrn.ResNetTorch(DIM, A, hp).double()
    torch.manual_seed(0)
    d = _data(32, 1)
    t = {k: torch.from_numpy(v).double() for k, v in d.items()}
    params = lrn.LearningParams(lrn.Adam(1e-3), l2_regularization=1e-4, nonvalidity_penalty=1.0, rewards_renormalization=2.0)
    Hp = float(lrn.entropy_wmean(t["P"], t["W"]))
    Wmean = 1.7
    net.eval()
    with torch.no_grad():
        got = [float(x) for x in lrn.losses(net, params, Wmean, Hp, (t["W"], t["X"], t["A"], t["P"], t["V"]))]
        Pn, Vn = net(t["X"])
    want = learnref.losses(Pn.numpy(), Vn.numpy(), [p.detach().numpy() for p in net.parameters()], *(d[k].astype(np.float64) for k in ("W", "A", "P", "V")),
                           1e-4, 1.0, 2.0, Wmean, learnref.entropy_wmean(d["P"].astype(np.float64), d["W"].astype(np.float64)))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-14)
    params0 = lrn.LearningParams(lrn.Adam(1e-3), l2_regularization=0.0, nonvalidity_penalty=0.0)
    with torch.no_grad():
        L0 = lrn.losses(net, params0, Wmean, Hp, (t["W"], t["X"], t["A"], t["P"], t["V"]))
    assert float(L0[3]) == 0.0 and float(L0[4]) == 0.0


def test_optimiser_rules_and_cyclic_schedule(lrn):
    from oracle import learnref
    xs, ys = lrn.cyclic_schedule(1e-3, 1e-2, 1e-4, 40)
    assert (xs, ys) == ([1, 18, 36, 40], [1e-3, 1e-2, 1e-3, 1e-4]) == learnref.cyclic_schedule(1e-3, 1e-2, 1e-4, 40)
    assert [lrn.pl_schedule([0, 10, 20], [0, 10, 30], x) for x in [-1, 0, 2, 10, 11, 20, 25]] == [0, 0, 2, 10, 12, 30, 30]  # src/schedule.jl:82-89

    class Quad(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.x = torch.nn.Parameter(torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64))

    Hm = np.array([[2.0, 0.3, 0.0], [0.3, 1.0, 0.1], [0.0, 0.1, 4.0]])
    loss = lambda net, _: 0.5 * net.x @ torch.from_numpy(Hm) @ net.x
    grad = lambda x: Hm @ x
    n = 25
    q = Quad()
    seen = []
    lrn.train(q, lrn.CyclicNesterov(1e-2, 1e-1, 1e-3, 0.8, 0.9), loss, iter(range(1000)), n, callback=lambda i, l: seen.append(q.x.detach().numpy().copy()))
    want = learnref.nesterov_run([1.0, -2.0, 0.5], grad, n, 1e-2, 1e-1, 1e-3, 0.8, 0.9)
    assert len(seen) == n and np.allclose(np.array(seen), np.array(want), rtol=1e-12, atol=1e-15)
    q = Quad()
    seen = []
    lrn.train(q, lrn.Adam(2e-2), loss, iter(range(1000)), n, callback=lambda i, l: seen.append(q.x.detach().numpy().copy()))
    assert np.allclose(np.array(seen), np.array(learnref.adam_run([1.0, -2.0, 0.5], grad, n, 2e-2)), rtol=1e-12, atol=1e-15)


def test_batchnorm_train_mode_follows_flux(lrn):
    """mu <- (1-m) mu + m mean(x); sigma2 <- (1-m) sigma2 + m * n/(n-1) * var(x) (biased var normalises the batch)."""
    net = lrn.ResNetTorch(DIM, A, HP(0))
    bn = net.stem[1]
    net.train()
    x = torch.from_numpy(_data(8, 5)["X"])
    y = net.stem[0](x.reshape(-1, 3, 6, 7)).detach()
    net(x)
    m, n = 0.6, y.numel() // y.shape[1]
    mean, var = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False)
    assert torch.allclose(bn.running_mean, m * mean, atol=1e-6)
    assert torch.allclose(bn.running_var, (1 - m) * torch.ones_like(var) + m * var * n / (n - 1), atol=1e-6)


def test_trainer_reduces_loss_and_hands_weights_back(lrn):
    torch.manual_seed(1)
    hp = HP(1)
    net = lrn.ResNetTorch(DIM, A, hp)
    d = _data(256, 7)
    params = lrn.LearningParams(lrn.CyclicNesterov(1e-3, 1e-2, 1e-4, 0.8, 0.9), l2_regularization=1e-4, batch_size=64,
                                loss_computation_batch_size=128, use_gpu=False)
    tr = lrn.Trainer(net, d, params, device="cpu", seed=3)
    assert tr.num_batches_total() == 4 and abs(tr.Wmean - d["W"].mean()) < 1e-6
    before = tr.learning_status()
    ls = tr.batch_updates(40)
    after = tr.learning_status()
    assert len(ls) == 40 and after["L"] < before["L"] and after["Lp"] < before["Lp"]
    assert abs(before["Hp"] - after["Hp"]) == 0 and after["Hpnet"] > 0
    blob = tr.get_trained_network_blob()
    net2 = lrn.ResNetTorch(DIM, A, hp).load_blob(blob)
    net.eval(); net2.eval()
    x = torch.from_numpy(d["X"][:16])
    with torch.no_grad():
        assert all(torch.equal(a, b) for a, b in zip(net(x), net2(x)))


@pytest.mark.parametrize("bn", [False, True])
def test_simplenet_mirror_vs_netref(lrn, oz, bn):
    """SimpleNet torch mirror (src/networks/architectures/simplenet.jl:37-64) against oracle/netref.py; blob round trip."""
    from oracle import netref
    az = _pkg.load()
    hp = az.SimpleNetHP(24, 2, depth_phead=2, depth_vhead=1, use_batch_norm=bn)
    d = dict(width=24, depth_common=2, depth_phead=2, depth_vhead=1, use_batch_norm=bn)
    blob = netref.simplenet_make_blob((3, 3, 3), 9, d, seed=8)
    net = lrn.SimpleNetTorch((3, 3, 3), 9, hp).load_blob(blob)
    assert (net.to_blob() == blob).all()
    gid = oz.game_id("tictactoe")
    states = oz.random_positions(gid, 3, 7, 4)
    Xw = np.stack([oz.vectorize_state(gid, s) for s in states])                       # [B, W, H, C]
    Xf = np.stack([x.reshape(-1, order="F") for x in Xw]).astype(np.float32)          # convert_samples rows
    net.eval()
    with torch.no_grad():
        P, V = net(torch.from_numpy(Xf))
    Pr, Vr = netref.simplenet_forward(blob, (3, 3, 3), 9, d, Xw)
    assert np.abs(P.numpy() - Pr).max() < 1e-6 and np.abs(V.numpy() - Vr).max() < 1e-6
    # one optimiser step runs in train mode and changes the weights
    data = dict(W=np.ones(7, np.float32), X=Xf, A=np.stack([oz.GameEnv(gid, s).actions_mask() for s in states]).astype(np.float32),
                P=np.full((7, 9), 0, np.float32), V=np.zeros(7, np.float32))
    data["P"] = data["A"] / data["A"].sum(1, keepdims=True)
    tr = lrn.Trainer(net, data, lrn.LearningParams(lrn.Adam(1e-2), l2_regularization=1e-4, batch_size=7, use_gpu=False), device="cpu")
    before = net.to_blob().copy()
    assert len(tr.batch_updates(2)) == 2 and (net.to_blob() != before).any()
