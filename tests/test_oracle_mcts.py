"""Structural known-answer checks of the oracle's MCTS / schedules / sampling (the reference has no
golden vectors for these: SURVEY 8c).  Invariants are derived from src/mcts.jl, src/schedule.jl,
src/util.jl, src/memory.jl."""
import ctypes as C
import math

import numpy as np
import pytest


def test_pl_schedule_kat(oz):
    """src/schedule.jl:82-89 (commented-out self test of PLSchedule)."""
    xs = (C.c_int * 3)(0, 10, 20)
    ys = (C.c_double * 3)(0, 10, 30)
    got = [oz.lib().oz_pl_schedule(3, xs, ys, x) for x in [-1, 0, 2, 10, 11, 20, 25]]
    assert got == [0, 0, 2, 10, 12, 30, 30]
    xs = (C.c_int * 3)(0, 20, 30)
    ys = (C.c_double * 3)(1.0, 1.0, 0.3)  # games/connect-four/params.jl:28
    assert oz.lib().oz_pl_schedule(3, xs, ys, 5) == 1.0
    assert abs(oz.lib().oz_pl_schedule(3, xs, ys, 25) - 0.65) < 1e-12
    assert oz.lib().oz_pl_schedule(3, xs, ys, 40) == 0.3


def test_det_log_exp(oz):
    L = oz.lib()
    for x in [1e-12, 0.1, 0.5, 0.70710678, 1.0, 1.5, 2.0, 10.0, 12345.678]:
        assert abs(L.oz_det_log(x) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    for x in [-30.0, -1.0, -1e-3, 0.0, 0.3, 1.0, 5.0]:
        assert abs(L.oz_det_exp(x) - math.exp(x)) <= 2e-14 * math.exp(x)


def test_dirichlet_stream(oz):
    for alpha in [1.0, 0.3, 2.5]:
        e = oz.dirichlet(123, 7, 3, 7, alpha)
        assert abs(e.sum() - 1) < 1e-12 and (e > 0).all()
        assert (oz.dirichlet(123, 7, 3, 7, alpha) == e).all()  # counter-based: reproducible
        assert (oz.dirichlet(123, 8, 3, 7, alpha) != e).any()
    # mean of Dirichlet(1) components is 1/n
    m = np.mean([oz.dirichlet(5, g, 0, 4, 1.0) for g in range(4000)], axis=0)
    assert np.abs(m - 0.25).max() < 0.01


@pytest.mark.parametrize("name", ["connect-four", "tictactoe", "mancala"])
@pytest.mark.parametrize("orc", ["uniform", "synth"])
def test_mcts_structural_invariants(oz, name, orc):
    gid = oz.game_id(name)
    env = oz.Env(gid, orc, cpuct=2.0, noise_eps=0.25)
    g = oz.GameEnv(gid)
    n_legal = int(g.actions_mask().sum())
    eta = oz.dirichlet(1, 0, 0, n_legal, 1.0)
    nsims = 200
    env.explore(g, nsims, eta)
    n, N, W, P, V = env.root_stats(g)
    # first simulation only expands the root (src/mcts.jl:205-207)
    assert n == n_legal and N.sum() == nsims - 1
    assert env.total_simulations == nsims
    assert abs(P.sum() - 1) < 1e-6
    acts, pi = env.policy(g)
    assert abs(pi.sum() - 1) < 1e-12 and (pi == N[acts] / N.sum() / (N[acts] / N.sum()).sum()).all()
    # a second explore on the same tree: root exists now
    env.explore(g, nsims, eta)
    _, N2, _, _, _ = env.root_stats(g)
    assert N2.sum() == 2 * nsims - 1
    assert env.num_nodes <= 2 * nsims
    assert env.total_nodes_traversed >= N2.sum()
    env.reset()
    assert env.num_nodes == 0 and env.total_simulations == 2 * nsims  # counters survive reset (mcts.jl:278-281)


def test_zero_prior_sqrt_quirk(oz):
    """With sum(N) == 0 every score is Q = 0, so the first legal action is chosen (argmax = first max)."""
    gid = oz.game_id("connect-four")
    env = oz.Env(gid, "uniform", cpuct=2.0)
    g = oz.GameEnv(gid)
    env.explore(g, 2)
    _, N, _, _, _ = env.root_stats(g)
    assert N.tolist() == [1, 0, 0, 0, 0, 0, 0]


def test_batch_driver_equals_recursive(oz):
    gid = oz.game_id("connect-four")
    A = 7
    roots = oz.random_positions(gid, 0xA17A2E80, 24, 30)
    mp = oz.mcts_params(cpuct=2.0, noise_eps=0.25, num_iters_per_turn=150)
    etas = np.zeros((len(roots), A))
    for i, r in enumerate(roots):
        m = oz.GameEnv(gid, r).actions_mask()
        etas[i, :m.sum()] = oz.dirichlet(9, i, 0, int(m.sum()), 1.0)
    b = oz.Batch(gid, len(roots), mp)
    b.set_roots(roots, etas)
    synth = C.cast(oz.builtin_oracle("synth"), oz.ORACLE_FN)
    while True:
        ls, lt = b.advance()
        if len(ls) == 0:
            break
        P = np.zeros((len(ls), A), np.float32)
        V = np.zeros(len(ls), np.float32)
        for j, s in enumerate(ls):
            m = oz.GameEnv(gid, s).actions_mask()
            buf = np.zeros(48, np.uint8)
            buf[:43] = s
            p = (C.c_float * 9)()
            v = C.c_float()
            synth(None, gid, buf.ctypes.data_as(C.POINTER(C.c_uint8)), int(m.sum()), p, C.byref(v))
            P[j, np.flatnonzero(m)] = list(p)[:m.sum()]
            V[j] = v.value
        b.feed(P, V)
    for i, r in enumerate(roots):
        env = oz.Env(gid, "synth", cpuct=2.0, noise_eps=0.25)
        g = oz.GameEnv(gid, r)
        env.explore(g, 150, etas[i, :int(g.actions_mask().sum())])
        _, N, W, P, _ = env.root_stats(g)
        Nb, Wb, Pb = b.root_stats(i)
        assert (N == Nb).all() and (W == Wb).all() and (P == Pb).all()


def test_self_play_trace_invariants(oz):
    gid = oz.game_id("connect-four")
    mp = oz.mcts_params(cpuct=2.0, noise_eps=0.25, num_iters_per_turn=40, sched_xs=(0, 20, 30), sched_ys=(1.0, 1.0, 0.3))
    traces = oz.worker_run(gid, "synth", mp, seed=11, first=0, stride=1, count=4, reset_every=2)
    for k, tr in enumerate(traces):
        n = tr["n_moves"]
        assert 7 <= n <= 42
        g = oz.GameEnv(gid)
        for i in range(n):
            assert g.state() == bytes(tr["states"][i])
            assert (tr["mask"][i].astype(bool) == g.actions_mask()).all()
            assert abs(tr["pi"][i].sum() - 1) < 1e-5 and (tr["pi"][i][~g.actions_mask()] == 0).all()
            g.play(int(tr["action"][i]))
            assert tr["rewards"][i] == g.white_reward()
        assert g.terminated()
        # z sign convention, src/memory.jl:78-83 (gamma = 1, reward only on the last transition)
        wr = tr["rewards"][-1]
        for i in range(n):
            white = (i % 2 == 0)
            assert tr["z"][i] == (wr if white else -wr) and tr["t"][i] == n - i
        assert tr["edepth"] > 0
    # tree kept between game 0 and 1 (reset_every = 2), dropped before game 2
    assert traces[1]["mem_nodes"] > traces[0]["mem_nodes"] * 0.9
    assert traces[2]["mem_nodes"] < traces[1]["mem_nodes"]
    # determinism
    again = oz.worker_run(gid, "synth", mp, seed=11, first=0, stride=1, count=2, reset_every=2)
    assert (again[1]["action"] == traces[1]["action"]).all()


def test_fix_probvec_and_categorical(oz):
    L = oz.lib()
    pi = np.array([0.2, 0.3, 0.5])
    out = np.zeros(3, np.float32)
    L.oz_fix_probvec(pi.ctypes.data, 3, out.ctypes.data)
    assert (out == pi.astype(np.float32)).all()
    assert L.oz_categorical(out.ctypes.data, 3, 0.0) == 0
    assert L.oz_categorical(out.ctypes.data, 3, 0.2) == 1  # cp <= draw advances
    assert L.oz_categorical(out.ctypes.data, 3, 0.95) == 2
    t = np.zeros(3)
    L.oz_apply_temperature(pi.ctypes.data, 3, 0.0, t.ctypes.data)
    assert t.tolist() == [0, 0, 1]
    L.oz_apply_temperature(pi.ctypes.data, 3, 0.5, t.ctypes.data)
    assert np.allclose(t, pi ** 2 / (pi ** 2).sum(), rtol=1e-14)


# ---- symmetries, flip_probability, TwoPlayers ------------------------------------------------------

def _sym(oz, gid, j, state):
    sb = oz.state_bytes(gid)
    src = np.zeros(oz.STATE_BYTES, np.uint8)
    src[:sb] = state[:sb]
    out = np.zeros(oz.STATE_BYTES, np.uint8)
    oz.lib().oz_apply_symmetry(gid, j, src.ctypes.data, out.ctypes.data)
    return out[:sb].copy()


def test_symmetry_tables(oz):
    """GI.symmetries: connect-four mirror (games/connect-four/game.jl:247-257); tic-tac-toe dihedral group
    (games/tictactoe/game.jl:149-168); checks of src/scripts/test_game.jl:81-96 (same player, image is a valid state,
    legal actions are permuted)."""
    L = oz.lib()
    assert [L.oz_num_symmetries(oz.game_id(n)) for n in ("connect-four", "tictactoe", "mancala", "grid-world")] == [1, 7, 0, 0]
    c4, ttt = oz.game_id("connect-four"), oz.game_id("tictactoe")
    for s in oz.random_positions(c4, 5, 64):
        m = _sym(oz, c4, 0, s)
        assert (m[:42].reshape(6, 7) == s[:42].reshape(6, 7)[:, ::-1]).all() and m[42] == s[42]
        assert (_sym(oz, c4, 0, m) == s).all()  # involution
        assert (oz.GameEnv(c4, m).actions_mask() == oz.GameEnv(c4, s).actions_mask()[::-1]).all()
    # tic-tac-toe: hand-derived tables. rot (x,y)->(y,N-x+1): sym[p] = pos(rot(xy(p)))
    rot = [6, 3, 0, 7, 4, 1, 8, 5, 2]      # p=0 (x=1,y=1) -> (1,3) -> pos 6 (0-based) ...
    flip = [6, 7, 8, 3, 4, 5, 0, 1, 2]     # (x,y) -> (x, N-y+1)
    comp = lambda f, g: [f[g[p]] for p in range(9)]  # (f . g)(p) = f(g(p))
    rot2, rot3 = comp(rot, rot), comp(rot, comp(rot, rot))
    tables = [rot, rot2, rot3, flip, comp(flip, rot), comp(flip, rot2), comp(flip, rot3)]
    base = np.array([1, 2, 0, 0, 1, 0, 2, 0, 0, 2], np.uint8)
    images = set()
    for j, tab in enumerate(tables):
        img = _sym(oz, ttt, j, base)
        assert (img[:9] == base[:9][tab]).all() and img[9] == base[9], j
        images.add(bytes(img))
        # the image is a state with the same number of stones and the same outcome status
        ga, gb = oz.GameEnv(ttt, base), oz.GameEnv(ttt, img)
        assert ga.terminated() == gb.terminated() and ga.actions_mask().sum() == gb.actions_mask().sum()
    assert len(images) == 7 and bytes(base) not in images


@pytest.mark.parametrize("name", ["connect-four", "tictactoe"])
def test_play_game_with_flips(oz, name):
    """play_game(flip_probability=p) (src/play.jl:298-315): the trace keeps the pre-symmetry state, the player thinks on the
    image, and the game continues from the image."""
    gid = oz.game_id(name)
    L = oz.lib()
    mp = oz.mcts_params(num_iters_per_turn=24, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, sched_xs=(0,), sched_ys=(1.0,))
    tr, tr0 = oz.Trace(), oz.Trace()
    sb = oz.state_bytes(gid)
    nflips = 0
    for game in range(6):
        env = oz.Env(gid, "synth", cpuct=2.0, noise_eps=0.25)
        L.oz_play_game2(env.h, env.h, C.byref(mp), 0.5, 77, game, C.byref(tr))
        n = tr.n_moves
        assert n > 0
        states = np.ctypeslib.as_array(tr.states)[:n + 1, :sb]
        think = np.ctypeslib.as_array(tr.think_states)[:n, :sb]
        sym = np.ctypeslib.as_array(tr.sym)[:n]
        acts = np.ctypeslib.as_array(tr.action)[:n]
        mask = np.ctypeslib.as_array(tr.mask)[:n, :oz.num_actions(gid)]
        for i in range(n):
            if sym[i]:
                nflips += 1
                assert (think[i] == _sym(oz, gid, int(sym[i]) - 1, states[i])).all()
            else:
                assert (think[i] == states[i]).all()
            g = oz.GameEnv(gid, think[i])
            assert (g.actions_mask() == mask[i]).all() and mask[i][acts[i]]
            g.play(int(acts[i]))
            assert bytes(g.state()[:sb]) == bytes(states[i + 1])
        # flip_probability = 0 reproduces oz_play_game
        env1, env2 = oz.Env(gid, "synth", cpuct=2.0, noise_eps=0.25), oz.Env(gid, "synth", cpuct=2.0, noise_eps=0.25)
        L.oz_play_game2(env1.h, env1.h, C.byref(mp), 0.0, 77, game, C.byref(tr))
        L.oz_play_game(env2.h, C.byref(mp), 77, game, C.byref(tr0))
        assert bytes(tr) == bytes(tr0)
    assert nflips > 5


def test_two_players_duel(oz):
    """TwoPlayers (src/play.jl:248-282): white's tree only grows on white's turns, black's on black's; with both players
    equal to one env the duel degenerates to a single MctsPlayer."""
    gid = oz.game_id("connect-four")
    L = oz.lib()
    mp = oz.mcts_params(num_iters_per_turn=20, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, sched_xs=(0,), sched_ys=(1.0,))
    w, b = oz.Env(gid, "synth", cpuct=2.0, noise_eps=0.25), oz.Env(gid, "uniform", cpuct=2.0, noise_eps=0.25)
    tr = oz.Trace()
    L.oz_play_game2(w.h, b.h, C.byref(mp), 0.0, 3, 0, C.byref(tr))
    n = tr.n_moves
    nw, nb = (n + 1) // 2, n // 2  # white moves first and colours alternate in connect-four
    assert w.total_simulations == 20 * nw and b.total_simulations == 20 * nb
    assert tr.mem_nodes == w.num_nodes + b.num_nodes
    assert tr.edepth == (w.total_nodes_traversed + b.total_nodes_traversed) / (w.total_simulations + b.total_simulations)
    tot = L.oz_total_reward(C.byref(tr), 1.0)
    assert tot == sum(np.ctypeslib.as_array(tr.rewards)[:n]) and tot in (-1.0, 0.0, 1.0)


@pytest.mark.parametrize("tau", [0.0, 1.0, 0.5])
def test_network_only_player(oz, tau):
    """NetworkPlayer under PlayerWithTemperature (Benchmark.NetworkOnly, src/play.jl:226-235, :112-127; num_iters_per_turn
    = 0 in a parameter block): think() is one oracle call -- its tree stays empty, the recorded policy IS the oracle's
    float32 policy over the available actions, tau = 0 plays the arg-max."""
    gid = oz.game_id("connect-four")
    L = oz.lib()
    mp_w = oz.mcts_params(num_iters_per_turn=20, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, sched_xs=(0,), sched_ys=(1.0,))
    mp_b = oz.mcts_params(num_iters_per_turn=0, sched_xs=(0,), sched_ys=(tau,))
    w, b = oz.Env(gid, "uniform", cpuct=2.0, noise_eps=0.25), oz.Env(gid, "synth")
    synth = C.cast(oz.builtin_oracle("synth"), oz.ORACLE_FN)
    sb, A = oz.state_bytes(gid), oz.num_actions(gid)
    seen_black = 0
    for game in range(6):
        tr = oz.Trace()
        L.oz_play_game2p(w.h, C.byref(mp_w), b.h, C.byref(mp_b), 0.0, 11, game, C.byref(tr))
        n = tr.n_moves
        think = np.ctypeslib.as_array(tr.think_states)[:n]
        pi, mask, act = np.ctypeslib.as_array(tr.pi)[:n, :A], np.ctypeslib.as_array(tr.mask)[:n, :A], np.ctypeslib.as_array(tr.action)[:n]
        for i in range(n):
            if think[i][sb - 1] == 1:          # white (MCTS) to move
                continue
            seen_black += 1
            legal = np.flatnonzero(mask[i])
            P, V = (C.c_float * A)(), C.c_float()
            st = (C.c_uint8 * len(think[i]))(*think[i])
            synth(None, gid, st, len(legal), P, C.byref(V))
            want = np.zeros(A, np.float32)
            want[legal] = np.ctypeslib.as_array(P)[:len(legal)]
            assert (pi[i].view(np.uint32) == want.view(np.uint32)).all()
            assert mask[i][act[i]] == 1
            if tau == 0.0:
                assert act[i] == legal[np.argmax(want[legal])]
    assert seen_black > 10
    assert b.total_simulations == 0 and b.num_nodes == 0 and w.total_simulations > 0


def test_simulate_duel_alternate_colors(oz):
    """simulate() with TwoPlayers and alternate_colors (src/simulations.jl:221-241) + rewards_and_redundancy (:292-307)."""
    from tests import simref
    gid = oz.game_id("tictactoe")
    mp = oz.mcts_params(num_iters_per_turn=16, cpuct=1.0, noise_eps=0.25, noise_alpha=1.0, sched_xs=(0,), sched_ys=(1.0,))
    traces, _ = simref.oracle_simulate(oz, gid, "synth", mp, 9, 4, 12, 2, baseline="uniform", alternate_colors=True, flip_probability=0.5)
    assert [traces[g]["colors_flipped"] for g in range(4)] == [True, False, True, False]
    rewards, red = simref.rewards_and_redundancy(traces)
    assert len(rewards) == 12 and set(rewards) <= {-1.0, 0.0, 1.0} and 0.0 < red < 1.0
    for g, t in traces.items():
        assert rewards[g] == (-1 if t["colors_flipped"] else 1) * t["rewards"].sum()


# ---- replay-buffer side: oracle restatement sanity -------------------------------------------------

def test_samples_ref_merge_augment_convert(oz):
    """merge_by_state / augment_with_symmetries / convert_samples restatement (src/memory.jl:89-130, src/learning.jl:17-51):
    hand-checked small case + invariants of src/scripts/test_game.jl:81-96 for the action permutations."""
    from oracle import samples_ref as sr
    from tests import simref
    gid = oz.game_id("connect-four")
    s0 = bytes(oz.GameEnv(gid).state()[:oz.state_bytes(gid)])
    g = oz.GameEnv(gid); g.play(2)
    s1 = bytes(g.state()[:oz.state_bytes(gid)])
    u = [1 / 7] * 7
    es = [dict(s=s0, pi=[0.1, 0.2, 0.3, 0.4, 0.0, 0.0, 0.0], z=1.0, t=9.0, n=1), dict(s=s1, pi=u, z=-1.0, t=8.0, n=1),
          dict(s=s0, pi=[0.3, 0.2, 0.1, 0.4, 0.0, 0.0, 0.0], z=-1.0, t=5.0, n=3)]
    m = sr.merge_by_state(es)
    assert [e["s"] for e in m] == [s0, s1] and m[0]["n"] == 4 and m[0]["z"] == 0.0 and m[0]["t"] == 7.0
    assert m[0]["pi"] == [(0.1 + 0.3) / 2, (0.2 + 0.2) / 2, (0.3 + 0.1) / 2, 0.4, 0.0, 0.0, 0.0] and m[1] == es[1]
    a = sr.augment_with_symmetries(gid, es[:2])
    assert len(a) == 4 and a[:2] == es[:2]
    assert a[2]["s"] == s0 and a[2]["pi"] == es[0]["pi"][::-1]           # the empty board is its own mirror
    assert a[3]["s"] != s1 and a[3]["z"] == -1.0
    c = sr.convert_samples(gid, 1, m)
    assert c["W"].tolist() == [3.0, 1.0] and c["X"].shape == (2, 126) and c["A"].sum() == 14 and c["V"].tolist() == [0.0, -1.0]
    # tic-tac-toe action permutations are the board permutations of games/tictactoe/game.jl:149-160
    ttt = oz.game_id("tictactoe")
    assert sr._aperm(ttt, 0) == [6, 3, 0, 7, 4, 1, 8, 5, 2] and sr._aperm(ttt, 3) == [6, 7, 8, 3, 4, 5, 0, 1, 2]
    # augment on real traces: image policies are still distributions over the image's legal actions
    mp = oz.mcts_params(num_iters_per_turn=16, cpuct=1.0, noise_eps=0.25, noise_alpha=1.0)
    traces, _ = simref.oracle_simulate(oz, ttt, "synth", mp, 5, 2, 4, 1)
    smp = sr.samples_from_traces(traces)
    aug = sr.augment_with_symmetries(ttt, smp)
    assert len(aug) == 8 * len(smp) and all(abs(sum(e["pi"]) - 1) < 1e-12 for e in aug)
    mg = sr.merge_by_state(aug)
    assert sum(e["n"] for e in mg) == len(aug) and len({e["s"] for e in mg}) == len(mg) < len(aug)


def test_oracle_matches_committed_kats(oz):
    """The oracle reproduces its committed known-answer vectors (tests/golden/oracle_kats.json, generated by
    tests/golden/make_oracle_kats.py): MCTS root statistics bit for bit, self-play / duel / flip traces by hash.  These
    vectors freeze the restatement every GPU parity test is measured against; they are not reference outputs."""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_oracle_kats", os.path.join(here, "make_oracle_kats.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = json.load(open(os.path.join(here, "oracle_kats.json")))
    got = json.loads(json.dumps(mk.build(oz)))
    assert got["format"] == want["format"] and len(got["cases"]) == len(want["cases"])
    for g, w in zip(got["cases"], want["cases"]):
        assert g == w, (w["kind"], w["game"])


def test_netref_matches_committed_kats(oz):
    """oracle/netref.py reproduces tests/golden/netref_kats.json (fp32 torch-CPU; tolerance 2e-6 for thread-count dependent
    summation order in the convolutions)."""
    import importlib.util
    import json
    import os
    from oracle import netref
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_netref_kats", os.path.join(here, "make_netref_kats.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = json.load(open(os.path.join(here, "netref_kats.json")))
    got = mk.build(oz, netref)
    assert got["states"] == want["states"] and got["ttt_states"] == want["ttt_states"]
    for g, w in zip(got["cases"], want["cases"]):
        assert g["num_params"] == w["num_params"] and abs(g["blob_sum"] - w["blob_sum"]) < 1e-9
        for k in ("P", "V", "Pinvalid"):
            if k in w:
                assert np.abs(np.array(g[k]) - np.array(w[k])).max() < 2e-6, (w["net"], k)
