"""GPU parity proper (-m gpu): the CUDA tree kernels, called through the C ABI, against the CPU oracle.
Bar: visit counts N, accumulated values W (f64), priors P and selected moves BIT-EXACT."""
import os

import numpy as np
import pytest

import _pkg

pytestmark = pytest.mark.gpu
PONS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pons")


@pytest.fixture(scope="module")
def az():
    return _pkg.load()


@pytest.fixture(scope="module")
def ctx(az):
    c = az.Context(0)
    yield c
    c.close()


def _etas(oz, gid, roots, A, seed, alpha=1.0):
    eta = np.zeros((len(roots), A))
    for i, r in enumerate(roots):
        n = int(oz.GameEnv(gid, r).actions_mask().sum())
        eta[i, :n] = oz.dirichlet(seed, i, 0, n, alpha)
    return eta


def _oracle_explore(oz, gid, kind, roots, eta, nsims, cpuct, eps, rounds=1):
    out = []
    for i, r in enumerate(roots):
        env = oz.Env(gid, kind, cpuct=cpuct, noise_eps=eps)
        g = oz.GameEnv(gid, r)
        n = int(g.actions_mask().sum())
        for _ in range(rounds):
            env.explore(g, nsims, None if eta is None else eta[i, :n])
        _, N, W, P, _ = env.root_stats(g)
        out.append((N, W, P, env.total_simulations, env.total_nodes_traversed, env.num_nodes))
    return out


def _pons_roots(gs, n):
    roots = []
    for name in ["Test_L1_R1", "Test_L2_R1", "Test_L3_R1"]:
        with open(os.path.join(PONS, name)) as f:
            for ln in list(f)[: n // 3]:
                s = gs.init_state()
                for ch in ln.split()[0]:
                    s, _, _ = gs.play(s, int(ch) - 1)
                roots.append(s)
    return np.stack(roots)


@pytest.mark.parametrize("kind", ["uniform", "synth"])
@pytest.mark.parametrize("game,nsims,nroots", [("connect-four", 600, 192), ("tictactoe", 50, 64), ("mancala", 400, 64)])
def test_explore_bit_exact(az, oz, ctx, game, nsims, nroots, kind):
    gs = az.GameSpec(game)
    gid = oz.game_id(game)
    A = gs.num_actions
    roots = gs.random_positions(0xA17A2E80, nroots, 30 if game != "tictactoe" else 5)
    if game == "connect-four":
        roots = np.concatenate([roots, _pons_roots(gs, 63)])
        roots[0] = gs.init_state()
    eta = _etas(oz, gid, roots, A, seed=5)
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    net = az.RandomOracle(ctx, gs) if kind == "uniform" else az.SynthOracle(ctx, gs)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), capacity_nodes_per_tree=2 * nsims)
    N, W, P = env.explore(roots, nsims, eta)
    ts, tn, nn = env.counters()
    ref = _oracle_explore(oz, gid, kind, roots, eta, nsims, 2.0, 0.25)
    for i, (rN, rW, rP, rts, rtn, rnn) in enumerate(ref):
        assert (N[i] == rN).all(), (i, N[i], rN)
        assert (W[i] == rW).all(), (i, W[i], rW)          # f64 bit-exact
        assert (P[i].view(np.uint32) == rP.view(np.uint32)).all()
        assert (ts[i], tn[i], nn[i]) == (rts, rtn, rnn)
        assert N[i].sum() == nsims - 1                      # first simulation only expands the root
    # tree kept across explore! calls (transposition table persists, src/mcts.jl:124-151)
    env.run(nsims)
    N2, W2, _ = env.root_stats()
    ref2 = _oracle_explore(oz, gid, kind, roots[:16], eta[:16], nsims, 2.0, 0.25, rounds=2)
    for i, (rN, rW, *_r) in enumerate(ref2):
        assert (N2[i] == rN).all() and (W2[i] == rW).all()
    pi = env.policy()
    assert np.allclose(pi.sum(1), 1) and (pi[0] == N2[0] / N2[0].sum() / (N2[0] / N2[0].sum()).sum()).all()
    env.reset()
    _, _, nn = env.counters()
    assert (nn == 0).all()
    N3, W3, _ = env.explore(roots, nsims, eta)
    assert (N3 == N).all() and (W3 == W).all()             # reset! really empties the tree
    env.close()
    net.close()


def test_explore_without_noise_and_gamma(az, oz, ctx):
    gs = az.GameSpec("connect-four")
    gid = oz.game_id("connect-four")
    roots = gs.random_positions(3, 32, 20)
    mp = az.MctsParams(gamma=0.9, cpuct=1.0, num_iters_per_turn=200, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
    net = az.SynthOracle(ctx, gs)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), 512)
    N, W, P = env.explore(roots, 200, None)
    for i, r in enumerate(roots):
        e = oz.Env(gid, "synth", gamma=0.9, cpuct=1.0)
        g = oz.GameEnv(gid, r)
        e.explore(g, 200)
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all()
    env.close()
    net.close()


@pytest.mark.parametrize("game,nsims,temp", [("connect-four", 64, ([0, 20, 30], [1.0, 1.0, 0.3])),
                                              ("tictactoe", 50, ([0], [1.0])), ("mancala", 32, ([0, 10], [1.0, 0.0]))])
def test_selfplay_bit_exact(az, oz, ctx, game, nsims, temp):
    """simulate() with 8 workers, 24 games, reset_every 2 against the oracle (dynamic game -> worker map of Util.mapreduce)."""
    gs = az.GameSpec(game)
    gid = oz.game_id(game)
    S, NG, seed = 8, 24, 77
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule(*temp), dirichlet_noise_eps=0.25,
                       dirichlet_noise_alpha=1.0)
    sp = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2)
    net = az.SynthOracle(ctx, gs)
    called = []
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(mp, sp), seed=seed, game_simulated=lambda: called.append(1))
    assert len(called) == NG
    omp = oz.mcts_params(cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=temp[0], sched_ys=temp[1])
    from tests import simref
    traces, slot_of = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, 2)
    assert sorted(traces) == list(range(NG)) and [slot_of[g] for g in range(S)] == list(range(S))
    simref.assert_same_samples(out, traces)
    net.close()


@pytest.mark.parametrize("game,nsims", [("connect-four", 48), ("tictactoe", 40)])
def test_selfplay_flip_probability_bit_exact(az, oz, ctx, game, nsims):
    """play_game(flip_probability = 0.5) (src/play.jl:305-307): trace states before the symmetry, pi / mask / actions in the
    image frame, exactly as the oracle records them."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed = 8, 20, 31
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0, 4], [1.0, 0.5]), dirichlet_noise_eps=0.25,
                       dirichlet_noise_alpha=1.0)
    net = az.SynthOracle(ctx, gs)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2,
                                                                        flip_probability=0.5)), seed=seed, gamma=1.0)
    omp = oz.mcts_params(cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=(0, 4), sched_ys=(1.0, 0.5))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, 2, flip_probability=0.5)
    assert sum(int((t["sym"] != 0).sum()) for t in traces.values()) > NG  # the case is exercised
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    net.close()


@pytest.mark.parametrize("game,nsims,alt,flip,reset", [("connect-four", 48, True, 0.5, 2), ("tictactoe", 40, True, 0.0, 1),
                                                       ("mancala", 24, False, 0.0, 3), ("connect-four", 32, False, 0.0, 0)])
def test_duel_bit_exact(az, oz, ctx, game, nsims, alt, flip, reset):
    """pit_networks-style simulate() with TwoPlayers (src/training.jl:130-143, src/play.jl:248-282): one tree per player per
    worker, per-player oracle, alternate_colors (src/simulations.jl:224-230), rewards_and_redundancy (:292-307)."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed = 6, 17, 1234
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0], [0.8]), dirichlet_noise_eps=0.1,
                       dirichlet_noise_alpha=1.0)
    contender, baseline = az.SynthOracle(ctx, gs), az.RandomOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=reset, flip_probability=flip, alternate_colors=alt)
    called = []
    out = az.simulate(ctx, gs, contender, az.SelfPlayParams(mp, sim), seed=seed, baseline=baseline, gamma=1.0,
                      game_simulated=lambda: called.append(1))
    assert len(called) == NG
    omp = oz.mcts_params(cpuct=2.0, noise_eps=0.1, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=(0,), sched_ys=(0.8,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, reset, baseline="uniform", alternate_colors=alt,
                                       flip_probability=flip)
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    assert (out["colors_flipped"] == [1 if alt and (g + 1) % 2 == 1 else 0 for g in range(NG)]).all()
    rewards, red = az.pit_networks(ctx, gs, contender, baseline, az.SelfPlayParams(mp, sim), seed=seed)
    assert (rewards == out["game_rewards"]).all() and red == out["redundancy"]
    contender.close()
    baseline.close()


def test_full_size_properties(az, ctx):
    """BASELINE config 1 sizes (4096 trees x 600 sims): size-independent invariants instead of an oracle run."""
    gs = az.GameSpec("connect-four")
    S, nsims = 4096, 600
    roots = gs.random_positions(0xA17A2E80, S, 30)
    mp = az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
    net = az.SynthOracle(ctx, gs)
    env = az.MctsEnv(ctx, gs, net, mp, S, capacity_nodes_per_tree=nsims + 8)
    N, W, P = env.explore(roots, nsims)
    ts, tn, nn = env.counters()
    assert (N.sum(1) == nsims - 1).all() and (ts == nsims).all()
    assert (nn <= nsims).all() and (nn >= 1).all()
    assert np.allclose(P.sum(1), 1, atol=1e-6)
    legal = np.stack([gs.actions_mask(r) for r in roots[:256]])
    assert (N[:256][~legal] == 0).all()
    assert (np.abs(W) <= N + 1e-9).all()           # |q| <= 1 per visit (gamma = 1, rewards in [-1, 1])
    t = env.last_timing()
    assert t["expansions"] == nn.sum()
    env.close()
    net.close()


@pytest.mark.parametrize("game,tau", [("connect-four", 0.0), ("connect-four", 0.5), ("mancala", 0.5), ("tictactoe", 2.0)])
def test_explore_prior_temperature_bit_exact(az, oz, ctx, game, tau):
    """prior_temperature != 1 (src/mcts.jl:157-161 -> Util.apply_temperature, src/util.jl:98-110): tau = 0 makes the prior
    one-hot on the first maximal legal action, otherwise P.^(1/tau) renormalised in Float64 and stored as Float32."""
    gs = az.GameSpec(game)
    gid = oz.game_id(game)
    nsims = 150 if game != "tictactoe" else 50
    roots = gs.random_positions(77, 48, 24 if game != "tictactoe" else 4)
    eta = _etas(oz, gid, roots, gs.num_actions, seed=9)
    mp = az.MctsParams(cpuct=1.5, num_iters_per_turn=nsims, dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0, prior_temperature=tau)
    net = az.SynthOracle(ctx, gs)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), 2 * nsims)
    N, W, P = env.explore(roots, nsims, eta)
    ts, tn, nn = env.counters()
    for i, r in enumerate(roots):
        e = oz.Env(gid, "synth", cpuct=1.5, noise_eps=0.25, prior_temperature=tau)
        g = oz.GameEnv(gid, r)
        n = int(g.actions_mask().sum())
        e.explore(g, nsims, eta[i, :n])
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all(), (i, N[i], rN)
        assert (P[i].view(np.uint32) == rP.view(np.uint32)).all(), (i, P[i], rP)
        assert (ts[i], tn[i], nn[i]) == (e.total_simulations, e.total_nodes_traversed, e.num_nodes)
    if tau == 0.0:
        assert ((P == 1.0).sum(1) == 1).all() and ((P == 0.0) | (P == 1.0)).all()
    env.close()
    net.close()


@pytest.mark.parametrize("game,S,nsims,gamma", [("connect-four", 4096, 600, 1.0), ("mancala", 4096, 400, 1.0), ("grid-world", 8192, 200, 0.95)])
def test_full_size_pool_slots_bit_exact(az, oz, ctx, game, S, nsims, gamma):
    """BASELINE configs [1], [3], [4] at their FULL pool size: the whole pool runs on the GPU (every tick batches the leaves
    of all S trees) and 64 randomly chosen slots are compared bit for bit with the CPU oracle -- trees are independent, so
    a slot's statistics must not depend on what the other 4095 / 8191 slots do (batch position, leaf-queue order, table
    placement)."""
    gs = az.GameSpec(game)
    gid = oz.game_id(game)
    A = gs.num_actions
    gw = game == "grid-world"
    roots = gs.random_positions(0xA17A2E80, S, 30) if not gw else gs.random_positions(0xA17A2E80, S)
    rng = np.random.default_rng(2024)
    pick = np.sort(rng.choice(S, 64, replace=False))
    eps = 0.0 if gw else 0.25
    eta = None
    if not gw:
        eta = np.zeros((S, A))
        sub = _etas(oz, gid, roots[pick], A, seed=11)
        e_all = rng.exponential(size=(S, A))           # the other slots: any valid noise vector over their legal actions
        full = np.array([int(gs.actions_mask(r).sum()) for r in roots])
        for i in range(S):
            v = e_all[i, :full[i]]
            eta[i, :full[i]] = v / v.sum()
        eta[pick] = sub
    mp = az.MctsParams(gamma=gamma, cpuct=2.0, num_iters_per_turn=nsims, dirichlet_noise_eps=eps, dirichlet_noise_alpha=1.0)
    net = az.SynthOracle(ctx, gs)
    env = az.MctsEnv(ctx, gs, net, mp, S, capacity_nodes_per_tree=nsims + 8)
    seed = 777
    if gw:
        env.set_noise(seed, np.arange(S) + 5, np.full(S, 2))
    N, W, P = env.explore(roots, nsims, eta)
    ts, tn, nn = env.counters()
    for i in pick:
        e = oz.Env(gid, "synth", gamma=gamma, cpuct=2.0, noise_eps=eps)
        if gw:
            e.set_noise(seed, 5 + int(i), 2)
        g = oz.GameEnv(gid, bytes(roots[i]))
        n = int(g.actions_mask().sum())
        e.explore(g, nsims, None if eta is None else eta[i, :n])
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all(), (i, N[i], rN)
        assert (P[i].view(np.uint32) == rP.view(np.uint32)).all()
        assert (ts[i], tn[i], nn[i]) == (e.total_simulations, e.total_nodes_traversed, e.num_nodes)
    assert (ts == nsims).all()
    env.close()
    net.close()


@pytest.mark.parametrize("game,nsims,gamma", [("connect-four", 200, 1.0), ("tictactoe", 60, 0.9), ("mancala", 100, 1.0)])
def test_rollout_oracle_explore_and_duel_bit_exact(az, oz, ctx, game, nsims, gamma):
    """MCTS.RolloutOracle (src/mcts.jl:27-60), the oracle of Benchmark.MctsRollouts (src/benchmark.jl:134-147): uniform prior,
    value = discounted return of one random playout (draws keyed by (seed, state, ply) on both sides); then a duel of a
    'network' player against the vanilla-MCTS baseline as Benchmark.Duel runs it (src/benchmark.jl:78-99)."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    rseed = 4711
    roots = gs.random_positions(31, 40, 20 if game != "tictactoe" else 4)
    eta = _etas(oz, gid, roots, gs.num_actions, seed=3)
    mp = az.MctsParams(gamma=gamma, cpuct=1.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    net = az.RolloutOracle(ctx, gs, gamma=gamma, seed=rseed)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), 2 * nsims)
    N, W, P = env.explore(roots, nsims, eta)
    for i, r in enumerate(roots):
        e = oz.Env(gid, oz.RolloutOracle(rseed, gamma), gamma=gamma, cpuct=1.0, noise_eps=0.25)
        g = oz.GameEnv(gid, r)
        n = int(g.actions_mask().sum())
        e.explore(g, nsims, eta[i, :n])
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all(), (i, N[i], rN)
        assert (P[i].view(np.uint32) == rP.view(np.uint32)).all()
    env.close()
    assert np.abs(W).max() > 0            # playouts do reach decided games
    # duel: synthetic 'network' vs MCTS with rollouts, alternate colours
    S, NG, seed, ns2 = 5, 12, 99, 24
    mp2 = az.MctsParams(gamma=gamma, cpuct=1.0, num_iters_per_turn=ns2, temperature=az.ConstSchedule(0.5), dirichlet_noise_eps=0.0,
                        dirichlet_noise_alpha=1.0)
    contender = az.SynthOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=1, alternate_colors=True)
    out = az.simulate(ctx, gs, contender, az.SelfPlayParams(mp2, sim), seed=seed, baseline=net, gamma=gamma)
    omp = oz.mcts_params(gamma=gamma, cpuct=1.0, num_iters_per_turn=ns2, sched_xs=(0,), sched_ys=(0.5,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, 1, baseline=oz.RolloutOracle(rseed, gamma), alternate_colors=True)
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    contender.close()
    net.close()
    with pytest.raises(az.AzError) as ei:   # stochastic environments have no deterministic playout
        az.RolloutOracle(ctx, az.GameSpec("grid-world"))
    assert ei.value.status == 5


@pytest.mark.parametrize("game", ["connect-four", "tictactoe"])
def test_duel_of_two_different_players_bit_exact(az, oz, ctx, game):
    """Benchmark.Duel(Benchmark.Full(params), Benchmark.MctsRollouts(params')) (src/benchmark.jl:78-99,134-162): TwoPlayers of
    two MctsPlayers that differ in EVERY MctsParams field (iterations, cpuct, gamma, noise, prior temperature, move
    temperature schedule) and in their oracle; alternate_colors swaps which player moves first."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed, rseed = 5, 14, 2718, 99
    mp_a = az.MctsParams(gamma=1.0, cpuct=2.0, num_iters_per_turn=30, temperature=az.PLSchedule([0, 4], [1.0, 0.3]), dirichlet_noise_eps=0.2,
                         dirichlet_noise_alpha=1.0, prior_temperature=0.7)
    mp_b = az.MctsParams(gamma=0.95, cpuct=1.0, num_iters_per_turn=45, temperature=az.ConstSchedule(0.0), dirichlet_noise_eps=0.0,
                         dirichlet_noise_alpha=0.5)
    a_net, b_net = az.SynthOracle(ctx, gs), az.RolloutOracle(ctx, gs, gamma=0.95, seed=rseed)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2, alternate_colors=True)
    out = az.simulate(ctx, gs, a_net, az.SelfPlayParams(mp_a, sim), seed=seed, baseline=b_net, baseline_mcts=mp_b, gamma=1.0)
    omp_a = oz.mcts_params(gamma=1.0, cpuct=2.0, noise_eps=0.2, noise_alpha=1.0, prior_temperature=0.7, num_iters_per_turn=30,
                           sched_xs=(0, 4), sched_ys=(1.0, 0.3))
    omp_b = oz.mcts_params(gamma=0.95, cpuct=1.0, noise_eps=0.0, noise_alpha=0.5, num_iters_per_turn=45, sched_xs=(0,), sched_ys=(0.0,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp_a, seed, S, NG, 2, baseline=oz.RolloutOracle(rseed, 0.95), alternate_colors=True,
                                       omp_baseline=omp_b)
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    a_net.close()
    b_net.close()


@pytest.mark.parametrize("game,tau", [("connect-four", 0.5), ("tictactoe", 1.0), ("mancala", 0.0)])
def test_duel_against_network_only_player_bit_exact(az, oz, ctx, game, tau):
    """Benchmark.Duel(Benchmark.Full(params), Benchmark.NetworkOnly(τ)) (src/benchmark.jl:78-99,161-176): the baseline is a
    NetworkPlayer under PlayerWithTemperature (src/play.jl:226-235, :112-127) -- no search, the move distribution is the
    oracle's policy -- against an MctsPlayer; alternate_colors and (where the game has symmetries) flip_probability on."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed = 6, 20, 31337
    flip = 0.5 if game != "mancala" else 0.0
    mp_a = az.MctsParams(gamma=1.0, cpuct=2.0, num_iters_per_turn=25, temperature=az.ConstSchedule(0.3), dirichlet_noise_eps=0.2,
                         dirichlet_noise_alpha=1.0)
    a_net, b_net = az.SynthOracle(ctx, gs), az.SynthOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2, alternate_colors=True, flip_probability=flip)
    out = az.simulate(ctx, gs, a_net, az.SelfPlayParams(mp_a, sim), seed=seed, baseline=b_net, baseline_mcts=az.NetworkOnly(tau), gamma=1.0)
    omp_a = oz.mcts_params(gamma=1.0, cpuct=2.0, noise_eps=0.2, noise_alpha=1.0, num_iters_per_turn=25, sched_xs=(0,), sched_ys=(0.3,))
    omp_b = oz.mcts_params(num_iters_per_turn=0, sched_xs=(0,), sched_ys=(tau,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp_a, seed, S, NG, 2, baseline="synth", alternate_colors=True,
                                       flip_probability=flip, omp_baseline=omp_b)
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    # the network-only player's rows carry the oracle's policy itself: float32 values, not visit-count ratios
    a_net.close()
    b_net.close()


@pytest.mark.parametrize("game,depth,amplify,tau,flip", [("connect-four", 4, True, 0.2, 0.5), ("tictactoe", 5, True, 0.0, 0.5),
                                                          ("mancala", 3, False, 0.5, 0.0), ("connect-four", 2, False, 1.0, 0.0)])
def test_duel_against_minmax_player_bit_exact(az, oz, ctx, game, depth, amplify, tau, flip):
    """Benchmark.Duel(Benchmark.Full(params), Benchmark.MinMaxTS(depth, amplify_rewards, τ)) (src/benchmark.jl:78-99,178-196;
    the connect-four benchmark of the reference is this duel at depth 5, games/connect-four/params.jl): the baseline is a
    MinMax.Player (src/minmax.jl) searched on the device, no oracle and no tree on its side."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed = 6, 18, 8086
    mp_a = az.MctsParams(gamma=1.0, cpuct=2.0, num_iters_per_turn=20, temperature=az.ConstSchedule(0.5), dirichlet_noise_eps=0.2,
                         dirichlet_noise_alpha=1.0)
    a_net = az.SynthOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2, alternate_colors=True, flip_probability=flip)
    out = az.simulate(ctx, gs, a_net, az.SelfPlayParams(mp_a, sim), seed=seed, baseline=az.MinMaxTS(depth, amplify, tau), gamma=1.0)
    omp_a = oz.mcts_params(gamma=1.0, cpuct=2.0, noise_eps=0.2, noise_alpha=1.0, num_iters_per_turn=20, sched_xs=(0,), sched_ys=(0.5,))
    omp_b = oz.minmax_params(depth, amplify, tau)
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp_a, seed, S, NG, 2, baseline="uniform", alternate_colors=True,
                                       flip_probability=flip, omp_baseline=omp_b)
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    a_net.close()


def test_reference_script_mcts_vs_minmax(az, oz, ctx):
    """test/mcts_vs_minmax.jl of the reference, the one script it ships for this path: tic-tac-toe, MctsPlayer(RolloutOracle,
    niters = 1000, tau = 0.5) as white against MinMax.Player(depth = 5, amplify_rewards, tau = 0.2), 200 games, trees never
    reset; the script prints the average reward.  Here: the same 200 games on the engine (8 workers), identical to the
    restatement game by game, and the property the script is there to show -- plain MCTS holds its own against minmax."""
    from tests import simref
    gs, gid = az.GameSpec("tictactoe"), oz.game_id("tictactoe")
    S, NG, seed, rseed = 8, 200, 1234, 5
    mp = az.MctsParams(gamma=1.0, cpuct=1.0, num_iters_per_turn=1000, temperature=az.ConstSchedule(0.5), dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
    net = az.RolloutOracle(ctx, gs, gamma=1.0, seed=rseed)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=None)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(mp, sim), seed=seed, baseline=az.MinMaxTS(5, True, 0.2), gamma=1.0)
    omp = oz.mcts_params(gamma=1.0, cpuct=1.0, num_iters_per_turn=1000, sched_xs=(0,), sched_ys=(0.5,))
    traces, _ = simref.oracle_simulate(oz, gid, oz.RolloutOracle(rseed, 1.0), omp, seed, S, NG, 0, baseline="uniform", omp_baseline=oz.minmax_params(5, True, 0.2))
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    r = out["game_rewards"]
    print("average reward %.3f (won %d, drawn %d, lost %d)" % (r.mean(), (r > 0).sum(), (r == 0).sum(), (r < 0).sum()))
    assert (r < 0).sum() <= 4 and r.mean() >= 0.0
    net.close()


def test_network_only_against_minmax_and_errors(az, oz, ctx):
    """Benchmark.Duel(Benchmark.NetworkOnly(), Benchmark.MinMaxTS(...)): neither side searches a tree; plus the parameter
    checks of the MinMax entry point."""
    from tests import simref
    gs, gid = az.GameSpec("connect-four"), oz.game_id("connect-four")
    S, NG, seed = 4, 12, 99
    net = az.SynthOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=1, alternate_colors=True)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(az.NetworkOnly(1.0), sim), seed=seed, baseline=az.MinMaxTS(3, True, 0.2), gamma=1.0)
    traces, _ = simref.oracle_simulate(oz, gid, "synth", oz.mcts_params(num_iters_per_turn=0), seed, S, NG, 1, baseline="uniform",
                                       alternate_colors=True, omp_baseline=oz.minmax_params(3, True, 0.2))
    simref.assert_same_samples(out, traces)
    simref.assert_same_outcomes(out, traces)
    for bad in (az.MinMaxTS(0, True), az.MinMaxTS(9, True), az.MinMaxTS(2, True, -1.0)):
        with pytest.raises(az.AzError) as e:
            az.SelfPlay(ctx, gs, net, az.SelfPlayParams(az.NetworkOnly(1.0), sim), baseline=bad)
        assert e.value.status == 1
    gw = az.GameSpec("grid-world")
    gnet = az.SynthOracle(ctx, gw)
    with pytest.raises(az.AzError) as e:      # no two-player heuristic
        az.SelfPlay(ctx, gw, gnet, az.SelfPlayParams(az.NetworkOnly(1.0), az.SimParams(num_games=4, num_workers=4, batch_size=4)),
                    baseline=az.MinMaxTS(2, True))
    assert e.value.status == 5
    gnet.close()
    net.close()


@pytest.mark.parametrize("game", ["grid-world", "connect-four"])
def test_network_only_player_alone_bit_exact(az, oz, ctx, game):
    """Benchmark.Single(Benchmark.NetworkOnly()) (src/benchmark.jl:101-110): simulate() with a NetworkPlayer on both sides /
    on the single-player environment; every turn is one oracle call."""
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, seed = 8, 30, 77
    net = az.SynthOracle(ctx, gs)
    sim = az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=1)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(az.NetworkOnly(0.7), sim), seed=seed)
    omp = oz.mcts_params(num_iters_per_turn=0, sched_xs=(0,), sched_ys=(0.7,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, 1)
    simref.assert_same_samples(out, traces)
    net.close()


def test_error_paths_mirror_reference_asserts(az, ctx):
    """Precondition violations return AZ_EINVAL / AZ_ESTATE with a message (no exception crosses the ABI)."""
    gs = az.GameSpec("connect-four")
    net = az.SynthOracle(ctx, gs)
    mp = az.MctsParams(cpuct=1.0, num_iters_per_turn=8, dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0)
    with pytest.raises(az.AzError) as e:   # batch_size <= num_workers (src/batchifier.jl:48)
        az.SelfPlay(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=4, num_workers=4, batch_size=8)))
    assert e.value.status == 1 and "batch_size" in str(e.value)
    with pytest.raises(az.AzError) as e:   # niters > 0 (src/play.jl:162)
        az.MctsEnv(ctx, gs, net, az.MctsParams(cpuct=1.0, num_iters_per_turn=0, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0), 4)
    assert e.value.status == 1
    man = az.GameSpec("mancala")
    mnet = az.SynthOracle(ctx, man)
    with pytest.raises(az.AzError) as e:   # flip_probability needs declared symmetries (src/params.jl:377-381)
        az.SelfPlay(ctx, man, mnet, az.SelfPlayParams(mp, az.SimParams(num_games=4, num_workers=4, batch_size=4, flip_probability=0.5)))
    assert e.value.status == 1 and "symmetries" in str(e.value)
    with pytest.raises(az.AzError):        # duel oracles must belong to the game
        az.SelfPlay(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=4, num_workers=4, batch_size=4)), baseline=mnet)
    mnet.close()
    env = az.MctsEnv(ctx, gs, net, mp, 4, 64)
    with pytest.raises(az.AzError):        # eta required when eps != 0
        env.set_roots(gs.random_positions(1, 4, 10), None)
    with pytest.raises(az.AzError) as e:   # MCTS.policy before explore! (src/mcts.jl:262)
        mp0 = az.MctsParams(cpuct=1.0, num_iters_per_turn=8, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
        env0 = az.MctsEnv(ctx, gs, net, mp0, 4, 64)
        env0.set_roots(gs.random_positions(1, 4, 10), None)
        env0.policy()
    assert e.value.status == 4
    ttt = az.GameSpec("tictactoe")
    with pytest.raises(az.AzError):        # oracle built for another game
        az.MctsEnv(ctx, ttt, net, mp, 4, 64)
    # table overflow is detected on device and reported, not silently corrupted
    small = az.MctsEnv(ctx, gs, net, az.MctsParams(cpuct=1.0, num_iters_per_turn=600, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0), 4, 16)
    with pytest.raises(az.AzError) as e:
        small.explore(gs.random_positions(1, 4, 4), 600)
    assert e.value.status == 3
    net.close()


@pytest.mark.parametrize("kind", ["synth", "simplenet"])
def test_grid_world_explore_and_selfplay_bit_exact(az, oz, ctx, kind):
    """Stochastic single-player environment (BASELINE config[4]): in-tree environment noise from the explicit stream,
    `time` outside the state (revisited states, paths up to 201 deep), intermediate rewards with gamma < 1."""
    import ctypes as C
    gs = az.GameSpec("grid-world")
    gid = oz.game_id("grid-world")
    if kind == "synth":
        net = az.SynthOracle(ctx, gs)
        ofn = "synth"
    else:
        from oracle import netref
        hp = dict(width=100, depth_common=4, use_batch_norm=False)  # games/grid-world/params.jl:5-8
        blob = netref.simplenet_make_blob(gs.state_dim, 4, hp, seed=2)
        net = az.SimpleNet(ctx, gs, az.SimpleNetHP(100, 4)).load(blob)
        cache = {}

        def cb(ctxp, g, sp, n, P, V):
            key = bytes(sp[:2])
            if key not in cache:
                p, v, _ = net.evaluate_batch(np.frombuffer(key, np.uint8)[None])
                cache[key] = (p[0], float(v[0]))
            for i in range(n):
                P[i] = cache[key][0][i]
            V[0] = cache[key][1]
        keep = oz.ORACLE_FN(cb)
        ofn = C.cast(keep, C.c_void_p)
    # ---- explore!
    roots = gs.random_positions(9, 48)
    nsims, seed = 200, 4242
    mp = az.MctsParams(gamma=0.95, cpuct=1.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
    env = az.MctsEnv(ctx, gs, net, mp, len(roots), 256)
    env.set_noise(seed, np.arange(len(roots)) + 100, np.full(len(roots), 3))
    N, W, P = env.explore(roots, nsims)
    ts, tn, nn = env.counters()
    for i, r in enumerate(roots):
        e = oz.Env(gid, "synth", gamma=0.95, cpuct=1.0) if kind == "synth" else None
        if e is None:
            e = oz.Env.__new__(oz.Env)
            e.gid, e._fn = gid, ofn
            e.h = oz.lib().oz_env_create(gid, ofn, None, 0.95, 1.0, 0.0, 1.0, 1.0)
        e.set_noise(seed, 100 + i, 3)
        g = oz.GameEnv(gid, bytes(r))
        e.explore(g, nsims)
        _, rN, rW, rP, _ = e.root_stats(g)
        assert (N[i] == rN).all() and (W[i] == rW).all() and (P[i] == rP).all(), i
        assert (ts[i], tn[i], nn[i]) == (e.total_simulations, e.total_nodes_traversed, e.num_nodes)
    assert (nn <= 100).all() and tn.max() > nsims       # at most 100 distinct states; deep paths
    env.close()
    # ---- simulate(): 8 workers x 16 games, 30 sims, argmax moves (ConstSchedule(0), games/grid-world/params.jl:19), reset_every 4
    S, NG, ns2 = 8, 16, 30
    mp2 = az.MctsParams(gamma=0.9, cpuct=1.0, num_iters_per_turn=ns2, temperature=az.ConstSchedule(0.0), dirichlet_noise_eps=0.0,
                        dirichlet_noise_alpha=1.0)
    out = az.simulate(ctx, gs, net, az.SelfPlayParams(mp2, az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=4)), seed=31)
    omp = oz.mcts_params(gamma=0.9, cpuct=1.0, num_iters_per_turn=ns2, sched_xs=(0,), sched_ys=(0.0,))
    from tests import simref
    traces, _ = simref.oracle_simulate(oz, gid, ofn, omp, 31, S, NG, 4)
    assert all(0 <= tr["n_moves"] <= 201 for tr in traces.values())  # 0: the game started on a reward cell
    simref.assert_same_samples(out, traces)
    net.close()
