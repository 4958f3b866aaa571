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
