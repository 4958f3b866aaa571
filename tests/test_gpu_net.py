"""GPU parity of the network forward (tcgen05 tower + heads) and of MCTS driven by it."""
import numpy as np
import pytest

import _pkg
from tests import netcheck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    return _pkg.load()


@pytest.fixture(scope="module")
def ctx(az):
    c = az.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("blocks,batch,seed,randomize", [(0, 37, 1, True), (1, 300, 2, True), (5, 300, 6, True), (7, 1000, 3, True),
                                                         (7, 1000, 1, False), (5, 500, 1, False), (1, 200, 1, False)])
def test_resnet_forward_matches_fp32_reference(az, oz, ctx, blocks, batch, seed, randomize):
    """P, V, Pinvalid within 1e-3 of the fp32 reference (tolerance stated by BASELINE.json north_star); policy logits and
    pre-tanh value within 1e-3 for the freshly initialised network, within the measured fp16 bound (tests/netcheck.py)
    for networks with randomised biases / BatchNorm statistics."""
    gs = az.GameSpec("connect-four")
    hp = netcheck.c4_hp(blocks)
    net, blob = netcheck.make_net(az, ctx, gs, hp, seed=seed, randomize=randomize)
    states = gs.random_positions(11, batch, 38)
    states[0] = gs.init_state()
    r = netcheck.compare(az, oz, gs, net, blob, hp, states)
    assert r["dP"] < netcheck.TOL and r["dV"] < netcheck.TOL and r["dI"] < netcheck.TOL, (r["dP"], r["dV"], r["dI"])
    if not randomize:  # logits (inputs of softmax / tanh) within 1e-3 as well: the contract of BASELINE.json north_star
        assert r["dL"] < netcheck.TOL and r["dVpre"] < netcheck.TOL, (r["dL"], r["dVpre"])
    else:
        assert r["dL"] < netcheck.LOGIT_TOL_PERTURBED and r["dVpre"] < netcheck.LOGIT_TOL_PERTURBED, (r["dL"], r["dVpre"])
        assert r["rmsL"] < netcheck.LOGIT_RMS_PERTURBED and r["rmsVpre"] < netcheck.LOGIT_RMS_PERTURBED, (r["rmsL"], r["rmsVpre"])
    assert (r["P"][~r["mask"]] == 0).all() and np.allclose(r["P"].sum(1), 1, atol=1e-5)
    # batch invariance: a state's output bits do not depend on its position in the batch
    P2, V2, _ = net.evaluate_batch(states[::-1].copy())
    assert (P2[::-1] == r["P"]).all() and (V2[::-1] == r["V"]).all()
    net.close()


def _net_with_env(az, ctx, gs, hp, env):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return netcheck.make_net(az, ctx, gs, hp, seed=5, randomize=True)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("blocks,batch", [(1, 37), (5, 700), (7, 4096)])
def test_persistent_tower_equals_per_layer_kernels(az, oz, ctx, blocks, batch):
    """The persistent whole-tower kernel (one launch, neighbour flags between layers) and the per-layer kernel issue the same
    MMAs in the same order for every output row, so with the same residual format (AZ_LO=16: fp16 low-order part) their
    outputs must be bit-identical -- at a batch that spans every CTA pair (4096 leaves), one that leaves most pairs idle (37)
    and an odd size (700).  The default e4m3 low-order part (8-bit, through kind::f8f6f4 identity MMAs) must agree with the
    fp16 one to well within the network tolerance."""
    gs = az.GameSpec("connect-four")
    hp = netcheck.c4_hp(blocks)
    states = gs.random_positions(17, batch, 38)
    outs, errs = [], []
    for env in ({"AZ_LO": "16"}, {"AZ_TOWER": "layer"}, {}):
        net, blob = _net_with_env(az, ctx, gs, hp, env)
        for rep in range(3):       # repeated launches: the flag counters keep counting across launches
            P, V, _ = net.evaluate_batch(states)
        L, Vp = net.forward_logits(states)
        outs.append((P, V, L, Vp))
        if batch <= 700:
            r = netcheck.compare(az, oz, gs, net, blob, hp, states)
            errs.append((r["rmsL"], r["rmsVpre"], r["dL"], r["dVpre"]))
        net.close()
    for a, b in zip(outs[0], outs[1]):
        assert (a.view(np.uint32) == b.view(np.uint32)).all()
    # e4m3 low-order part: a different (equally valid) rounding of the skip stream, so single outputs move by about the
    # network's own fp16 noise; what must hold is that its error against the fp32 reference is no larger than the fp16 one's
    d = [float(np.abs(a - b).max()) for a, b in zip(outs[0], outs[2])]
    print("lo8 vs lo16: max |dP| %.2e |dV| %.2e |dL| %.2e |dVpre| %.2e" % tuple(d))
    assert max(d) < 2.5e-3, d
    if blocks > 1:
        assert max(d) > 0.0    # the 8-bit path really ran
    if errs:
        print("vs fp32 reference (rmsL, rmsVpre, maxL, maxVpre): lo16 %s  lo8 %s" % (errs[0], errs[2]))
        assert errs[2][0] <= 1.2 * errs[0][0] + 2e-5 and errs[2][1] <= 1.2 * errs[0][1] + 2e-5, (errs[0], errs[2])


def test_resnet_precision_stress(az, oz, ctx):
    """The measured error distribution of 7-block networks with randomised biases / BatchNorm statistics (gamma in
    [0.7, 1.3], sigma2 in [0.6, 1.5], mu, beta ~ N(0, 0.1); oracle/netref.py make_blob) -- the hardest case for the fp16
    tensor-core operands (11 significand bits, like TF32).  P stays within 1e-3; V, the logits and the pre-tanh value
    exceed 1e-3 in the worst position of a few hundred (max up to ~2e-3, RMS 3e-4..6e-4) and the test pins exactly that
    bound (tests/netcheck.py, DESIGN.md "precision").  Freshly initialised networks meet 1e-3 on everything (test above)."""
    gs = az.GameSpec("connect-four")
    hp = netcheck.c4_hp(7)
    states = gs.random_positions(11, 400, 38)
    for seed in (8, 5, 2):
        net, blob = netcheck.make_net(az, ctx, gs, hp, seed=seed, randomize=True)
        r = netcheck.compare(az, oz, gs, net, blob, hp, states)
        rms_v = float(np.sqrt(np.mean((r["V"] - r["Vr"]) ** 2)))
        rms_p = float(np.sqrt(np.mean((r["P"] - r["Pr"]) ** 2)))
        print("seed %d: max dP %.2e dV %.2e dL %.2e dVpre %.2e  rms dP %.2e dV %.2e" % (seed, r["dP"], r["dV"], r["dL"], r["dVpre"], rms_p, rms_v))
        assert r["dP"] < 1e-3 and r["dV"] < 2.5e-3 and rms_v < 8e-4 and rms_p < 3e-4
        assert r["dL"] < netcheck.LOGIT_TOL_PERTURBED and r["dVpre"] < netcheck.LOGIT_TOL_PERTURBED
        assert r["rmsL"] < netcheck.LOGIT_RMS_PERTURBED and r["rmsVpre"] < netcheck.LOGIT_RMS_PERTURBED
        net.close()


@pytest.mark.parametrize("game,plies", [("tictactoe", 5), ("mancala", 30)])
def test_resnet_other_games_generic_tower(az, oz, ctx, game, plies):
    """Geometries whose row stride is not 8 run the generic 9-tap tcgen05 kernel (3x3 and 14x1 boards)."""
    gs = az.GameSpec(game)
    hp = netcheck.c4_hp(2)
    net, blob = netcheck.make_net(az, ctx, gs, hp, seed=4, randomize=True)
    states = gs.random_positions(13, 150, plies)
    r = netcheck.compare(az, oz, gs, net, blob, hp, states)
    assert r["dP"] < netcheck.TOL and r["dV"] < netcheck.TOL and r["dI"] < netcheck.TOL, (r["dP"], r["dV"], r["dI"])
    net.close()


@pytest.mark.parametrize("game,plies,hp", [
    # scripts/profile/inference.jl-style narrow tower with the ResNetHP default heads (src/networks/architectures/resnet.jl:30-37)
    ("connect-four", 30, dict(num_blocks=5, num_filters=64, conv_kernel_size=(3, 3), num_policy_head_filters=2, num_value_head_filters=1)),
    ("connect-four", 30, dict(num_blocks=2, num_filters=100, conv_kernel_size=(3, 3), num_policy_head_filters=7, num_value_head_filters=32)),
    ("tictactoe", 5, dict(num_blocks=2, num_filters=64, conv_kernel_size=(3, 3), num_policy_head_filters=2, num_value_head_filters=1)),
    ("mancala", 30, dict(num_blocks=1, num_filters=32, conv_kernel_size=(3, 3), num_policy_head_filters=16, num_value_head_filters=8))])
def test_resnet_narrower_shapes(az, oz, ctx, game, plies, hp):
    """num_filters < 128 and fewer head filters than 32: the kernels' unused channels carry zero weights (az_net.cu init())."""
    gs = az.GameSpec(game)
    net, blob = netcheck.make_net(az, ctx, gs, hp, seed=9, randomize=True)
    states = gs.random_positions(17, 300, plies)
    r = netcheck.compare(az, oz, gs, net, blob, hp, states)
    assert r["dP"] < netcheck.TOL and r["dV"] < netcheck.TOL and r["dI"] < netcheck.TOL, (r["dP"], r["dV"], r["dI"])
    assert r["dL"] < netcheck.LOGIT_TOL_PERTURBED and r["dVpre"] < netcheck.LOGIT_TOL_PERTURBED, (r["dL"], r["dVpre"])
    net.close()


def test_fresh_flux_init(az, oz, ctx):
    """Freshly constructed model (zero biases, identity BatchNorm statistics), 5 blocks as shipped."""
    gs = az.GameSpec("connect-four")
    hp = netcheck.c4_hp(5)
    net, blob = netcheck.make_net(az, ctx, gs, hp, seed=1, randomize=False)
    r = netcheck.compare(az, oz, gs, net, blob, hp, gs.random_positions(5, 200, 30))
    assert r["dP"] < netcheck.TOL and r["dV"] < netcheck.TOL
    net.close()


def test_resnet_rejects_bad_blob_and_unsupported(az, ctx):
    gs = az.GameSpec("connect-four")
    net = az.ResNet(ctx, gs, az.ResNetHP(1, 128, (3, 3), 32, 32))
    with pytest.raises(az.AzError):
        net.load(np.zeros(10, np.float32))
    with pytest.raises(az.AzError):
        net.evaluate_batch(gs.random_positions(1, 4, 10))  # not loaded
    net.close()
    for bad in (az.ResNetHP(1, 256, (3, 3), 32, 32), az.ResNetHP(1, 128, (5, 5), 32, 32), az.ResNetHP(1, 128, (3, 3), 64, 32)):
        with pytest.raises(az.AzError) as e:
            az.ResNet(ctx, gs, bad)
        assert e.value.status == 5


def test_mcts_with_network_bit_exact(az, oz, ctx):
    """Tree logic with a real network: the CPU oracle replays with the network's own (P, V) for every state it
    asks about (forward is batch invariant), so visit counts and W must match bit for bit."""
    gs = az.GameSpec("connect-four")
    gid = oz.game_id("connect-four")
    hp = netcheck.c4_hp(1)
    net, _ = netcheck.make_net(az, ctx, gs, hp, seed=3)
    roots = gs.random_positions(21, 12, 24)
    nsims = 120
    eta = np.zeros((len(roots), 7))
    for i, r in enumerate(roots):
        n = int(oz.GameEnv(gid, bytes(r)).actions_mask().sum())
        eta[i, :n] = oz.dirichlet(4, i, 0, n, 1.0)
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), 2 * nsims)
    N, W, P = env.explore(roots, nsims, eta)
    cache = {}

    def oracle(state, n):
        if state not in cache:
            p, v, _ = net.evaluate_batch(np.frombuffer(state, np.uint8)[None])
            m = gs.actions_mask(np.frombuffer(state, np.uint8))
            cache[state] = (p[0][m].tolist(), float(v[0]))
        return cache[state]

    for i, r in enumerate(roots):
        e = oz.Env(gid, oracle, cpuct=2.0, noise_eps=0.25)
        g = oz.GameEnv(gid, bytes(r))
        e.explore(g, nsims, eta[i, :int(g.actions_mask().sum())])
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all() and (P[i] == rP).all()
    env.close()
    net.close()


def test_selfplay_with_network_bit_exact(az, oz, ctx):
    """simulate() end to end with the ResNet as oracle (4 workers x 8 games x 24 sims, reset_every 2): every oracle worker
    replays with the network's own (P, V) per state, so traces (states, actions, pi, z, t) must match bit for bit."""
    import ctypes as C
    gs = az.GameSpec("connect-four")
    gid = oz.game_id("connect-four")
    hp = netcheck.c4_hp(1)
    net, _ = netcheck.make_net(az, ctx, gs, hp, seed=5)
    S, NG, nsims, seed = 4, 8, 24, 99
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0, 20, 30], [1.0, 1.0, 0.3]),
                       dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2)), seed=seed)
    cache = {}
    sb = gs.state_bytes

    def cb(ctxp, g, sp, n, P, V):
        key = bytes(sp[:sb])
        if key not in cache:
            p, v, _ = net.evaluate_batch(np.frombuffer(key, np.uint8)[None])
            m = gs.actions_mask(np.frombuffer(key, np.uint8))
            cache[key] = (p[0][m], float(v[0]))
        p, v = cache[key]
        for i in range(n):
            P[i] = p[i]
        V[0] = v

    fn = oz.ORACLE_FN(cb)
    omp = oz.mcts_params(cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=(0, 20, 30), sched_ys=(1.0, 1.0, 0.3))
    from tests import simref
    traces, _ = simref.oracle_simulate(oz, gid, C.cast(fn, C.c_void_p), omp, seed, S, NG, 2)
    simref.assert_same_samples(out, traces, check_mask=False)
    net.close()


def test_duel_of_two_networks_bit_exact(az, oz, ctx):
    """pit_networks (src/training.jl:130-143) with two different ResNets, alternate_colors and flip_probability = 0.5 (the
    shipped connect-four ArenaParams, games/connect-four/params.jl:32-45, at a small size): each oracle-side player replays
    with its own network's (P, V), so traces and rewards must match bit for bit."""
    import ctypes as C
    from tests import simref
    gs = az.GameSpec("connect-four")
    gid = oz.game_id("connect-four")
    nets = [netcheck.make_net(az, ctx, gs, netcheck.c4_hp(1), seed=5)[0], netcheck.make_net(az, ctx, gs, netcheck.c4_hp(2), seed=6)[0]]
    S, NG, nsims, seed = 4, 10, 24, 4242
    mp = az.MctsParams(cpuct=1.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0], [0.2]), dirichlet_noise_eps=0.05,
                       dirichlet_noise_alpha=1.0)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2, flip_probability=0.5, alternate_colors=True)
    out = az.simulate(ctx, gs, nets[0], az.SelfPlayParams(mp, sim), seed=seed, baseline=nets[1], gamma=1.0)
    sb = gs.state_bytes
    fns = []
    for net in nets:
        def cb(ctxp, g, sp, n, P, V, net=net, cache={}):
            key = bytes(sp[:sb])
            if key not in cache:
                p, v, _ = net.evaluate_batch(np.frombuffer(key, np.uint8)[None])
                m = gs.actions_mask(np.frombuffer(key, np.uint8))
                cache[key] = (p[0][m], float(v[0]))
            p, v = cache[key]
            for i in range(n):
                P[i] = p[i]
            V[0] = v
        fns.append(oz.ORACLE_FN(cb))
    omp = oz.mcts_params(cpuct=1.0, noise_eps=0.05, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=(0,), sched_ys=(0.2,))
    traces, _ = simref.oracle_simulate(oz, gid, C.cast(fns[0], C.c_void_p), omp, seed, S, NG, 2, baseline=C.cast(fns[1], C.c_void_p),
                                       alternate_colors=True, flip_probability=0.5)
    simref.assert_same_samples(out, traces, check_mask=False)
    simref.assert_same_outcomes(out, traces)
    for n in nets:
        n.close()


@pytest.mark.parametrize("game,hp", [("tictactoe", dict(width=200, depth_common=6, use_batch_norm=True)),     # games/tictactoe/params.jl:8-12
                                      ("connect-four", dict(width=100, depth_common=4, use_batch_norm=False)),
                                      ("mancala", dict(width=64, depth_common=2, depth_phead=2, depth_vhead=0, use_batch_norm=True))])
def test_simplenet_forward_and_mcts(az, oz, ctx, game, hp):
    """SimpleNet (fp32 fused MLP) against the torch restatement, and MCTS driven by it (config[0]: 64 trees x 50 sims)."""
    from oracle import netref
    gs = az.GameSpec(game)
    gid = oz.game_id(game)
    blob = netref.simplenet_make_blob(gs.state_dim, gs.num_actions, hp, seed=7)
    net = az.SimpleNet(ctx, gs, az.SimpleNetHP(hp["width"], hp["depth_common"], hp.get("depth_phead", 1), hp.get("depth_vhead", 1),
                                               hp.get("use_batch_norm", False)))
    assert net.num_params == len(blob)
    net.load(blob)
    states = gs.random_positions(3, 64, 5 if game == "tictactoe" else 25)
    X = np.stack([oz.vectorize_state(gid, bytes(s)) for s in states])
    mask = np.stack([oz.GameEnv(gid, bytes(s)).actions_mask() for s in states])
    P0, V0 = netref.simplenet_forward(blob, gs.state_dim, gs.num_actions, hp, X)
    Pr, Vr, Ir = netref.forward_normalized(P0, V0, mask)
    P, V, Pinv = net.evaluate_batch(states)
    assert np.abs(P - Pr).max() < 1e-4 and np.abs(V - Vr).max() < 1e-4 and np.abs(Pinv - Ir).max() < 1e-4
    Lr, Vpr = netref.simplenet_forward(blob, gs.state_dim, gs.num_actions, hp, X, logits=True)
    L, Vp = net.forward_logits(states)
    assert np.abs(L - Lr).max() < 1e-4 and np.abs(Vp - Vpr).max() < 1e-4
    # MCTS with the network as oracle, oracle replay with the network's own outputs: bit-exact
    nsims = 50
    mp = az.MctsParams(cpuct=1.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
    env = az.MctsEnv(ctx, gs, net, mp, len(states), 4 * nsims)
    N, W, _ = env.explore(states, nsims)
    cache = {}

    def oracle(state, n):
        if state not in cache:
            p, v, _ = net.evaluate_batch(np.frombuffer(state, np.uint8)[None])
            m = gs.actions_mask(np.frombuffer(state, np.uint8))
            cache[state] = (p[0][m].tolist(), float(v[0]))
        return cache[state]
    for i in range(0, len(states), 8):
        e = oz.Env(gid, oracle, cpuct=1.0)
        g = oz.GameEnv(gid, bytes(states[i]))
        e.explore(g, nsims)
        _, rN, rW, _, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all()
    env.close()
    net.close()


def test_learning_step_hands_weights_to_engine(az, ctx):
    """a few optimiser steps in torch on the GPU (alphazero.jl_b200/learning.py), then the trained blob goes
    straight into the engine's network; the engine's forward_normalized must agree with the torch model in test mode."""
    import torch
    import alphazero_jl_b200.learning as lrn
    gs = az.GameSpec("connect-four")
    hp = az.ResNetHP(1, 128, (3, 3), 32, 32, batch_norm_momentum=0.6)
    torch.manual_seed(0)
    net_t = lrn.ResNetTorch(gs.state_dim, gs.num_actions, hp)
    states = gs.random_positions(11, 512, 30)
    X = np.stack([gs.vectorize_state(s).reshape(-1, order="F") for s in states]).astype(np.float32)
    Am = np.stack([gs.actions_mask(s) for s in states]).astype(np.float32)
    rng = np.random.default_rng(0)
    P = rng.random(Am.shape).astype(np.float32) * Am
    P /= P.sum(1, keepdims=True)
    data = dict(W=np.ones(len(states), np.float32), X=X, A=Am, P=P, V=rng.choice([-1.0, 0.0, 1.0], len(states)).astype(np.float32))
    params = lrn.LearningParams(lrn.Adam(1e-3), l2_regularization=1e-4, batch_size=128, loss_computation_batch_size=256)
    tr = lrn.Trainer(net_t, data, params, device="cuda", seed=1)
    ls = tr.batch_updates(40)
    assert len(ls) == 40 and np.isfinite(ls).all() and np.isfinite(list(tr.learning_status().values())).all()
    assert np.mean(ls[-8:]) < 0.9 * np.mean(ls[:8]), (ls[:8], ls[-8:])       # the loss goes down on the GPU too
    # hand-off WITHOUT a host round trip: the flat parameter tensor on the GPU -> az_net_load_device (device-side BatchNorm fold)
    net = az.ResNet(ctx, gs, hp)
    assert tr.hand_off(net) == net.num_params
    Pe, Ve, Pinv = net.evaluate_batch(states)
    net_t.eval()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the comparison target is true fp32
    try:
        with torch.no_grad():
            Pt, Vt, pinv_t = lrn.forward_normalized(net_t, torch.from_numpy(X).cuda(), torch.from_numpy(Am).cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    # a TRAINED network (BatchNorm statistics have moved): P within 1e-3, V within the measured fp16 bound of tests/netcheck.py
    dP, dV = np.abs(Pe - Pt.cpu().numpy()).max(), np.abs(Ve - Vt.cpu().numpy()).max()
    assert dP < 1e-3 and dV < netcheck.LOGIT_TOL_PERTURBED, (dP, dV)
    assert np.abs(Pinv - pinv_t.cpu().numpy()).max() < 1e-3
    # the host-blob path (az_net_load) goes through the same device fold: identical bits
    blob = tr.get_trained_network_blob()
    net2 = az.ResNet(ctx, gs, hp).load(blob)
    P2, V2, _ = net2.evaluate_batch(states)
    assert (P2 == Pe).all() and (V2 == Ve).all()
    with pytest.raises(az.AzError):          # a host pointer is rejected
        net2.load_device(blob.ctypes.data, len(blob))
    net.close()
    net2.close()


def test_checkpoint_files_round_trip_through_the_engine(az, oz, ctx, tmp_path):
    """SURVEY 8f rank 4 on the GPU: bestnn.azb / mem.azs written from engine state (network blob, device-resident samples of a
    finished self-play run) and read back into a fresh network and sample set: identical outputs and samples."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("az_checkpoint", os.path.join(root, "alphazero.jl_b200", "checkpoint.py"))
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    gs = az.GameSpec("connect-four")
    hp = netcheck.c4_hp(1)
    net, blob = netcheck.make_net(az, ctx, gs, hp, seed=9)
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=24, temperature=az.ConstSchedule(1.0), dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    sp = az.SelfPlay(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=12, num_workers=6, batch_size=6, reset_every=2)), seed=3)
    sp.start()
    sp.wait()
    smp = az.Samples.from_selfplay(sp)
    rows = smp.fetch()
    d = str(tmp_path)
    ck.save_network(os.path.join(d, "bestnn.azb"), "resnet", gs.name, hp, blob)
    ck.save_memory(os.path.join(d, "mem.azs"), gs.name, gs.state_bytes, gs.num_actions, rows)
    ld = ck.load_network(os.path.join(d, "bestnn.azb"))
    blob2 = ld["blob"]
    assert ld["game"] == gs.name and ld["kind"] == "resnet" and (blob2 == blob).all()
    net2 = az.ResNet(ctx, gs, az.ResNetHP(hp["num_blocks"], 128, (3, 3), 32, 32)).load(blob2)
    st = gs.random_positions(2, 50, 30)
    a, b = net.evaluate_batch(st), net2.evaluate_batch(st)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    rows2 = ck.load_memory(os.path.join(d, "mem.azs"))
    smp2 = az.Samples.from_host(ctx, gs, rows2["states"], rows2["pi"], rows2["z"], rows2["t"], n=rows2["n"])
    back = smp2.fetch()
    for k in ("states", "pi", "z", "t", "n"):
        assert (back[k] == rows[k]).all(), k
    for x in (smp, smp2, sp, net, net2):
        x.close()
