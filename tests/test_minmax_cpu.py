"""MinMax baseline player (src/minmax.jl, Benchmark.MinMaxTS) without a GPU: the heuristics and the search exist three
times -- the oracle's C restatement (oracle/az_oracle.c: recursive, cell arrays, as the reference writes it), the product's
host + device inline code (csrc/az_games.cuh: explicit stack, bitboards; reached on the host through az_game_heuristic_value
/ az_game_minmax_think) and a direct Python transcription of src/minmax.jl below over the oracle's game interface -- and
must agree bit for bit."""
import ctypes as C
import math

import numpy as np
import pytest

import _pkg

GAMES = ["connect-four", "tictactoe", "mancala"]


@pytest.fixture(scope="module")
def az():
    _pkg.build()
    return _pkg.load()


def oz_heuristic(oz, g):
    return oz.lib().oz_heuristic_value(C.byref(g.g))


def py_think(oz, g, depth, amplify, tau, gamma):
    """src/minmax.jl:14-43, :83-114 verbatim over the GameEnv mirror (exp / pow through the oracle's deterministic exp / log)."""
    L = oz.lib()

    def value(game, d):
        if game.terminated():
            return 0.0
        if d == 0:
            return oz_heuristic(oz, game)
        return max(qvalue(game, a, d) for a in np.flatnonzero(game.actions_mask()))

    def qvalue(game, a, d):
        nxt = game.clone()
        nxt.play(int(a))
        wr = nxt.white_reward()
        r = wr if game.white_playing() else -wr
        if amplify and r != 0:
            r = math.copysign(math.inf, r)
        nextv = value(nxt, d - 1)
        if game.white_playing() != nxt.white_playing():
            nextv = -nextv
        return r + gamma * nextv

    actions = np.flatnonzero(g.actions_mask())
    qs = np.array([qvalue(g, a, depth) for a in actions])
    n = len(actions)
    winning = np.flatnonzero(qs == math.inf)
    if len(winning) == 0:
        notlosing = np.flatnonzero(qs > -math.inf)
        best = int(np.argmax(qs))
        if len(notlosing) == 0:
            pi = np.ones(n)
        elif tau == 0:
            pi = (qs == qs[best]).astype(np.float64)
        else:
            Cn = max(abs(qs[a]) for a in notlosing) + np.finfo(np.float64).eps
            pi = np.zeros(n)
            for i in range(n):
                x = (qs[i] - qs[best]) / Cn
                e = 0.0 if x == -math.inf else L.oz_det_exp(x)
                pi[i] = L.oz_det_exp((1.0 / tau) * L.oz_det_log(e)) if e > 0 else 0.0
    else:
        pi = (qs == math.inf).astype(np.float64)
    s = pi[0]
    for v in pi[1:]:
        s = s + v
    return actions, qs, pi / s


def test_heuristic_known_answers(az, oz):
    """Hand-checked values of GI.heuristic_value."""
    c4, ttt, man = oz.game_id("connect-four"), oz.game_id("tictactoe"), oz.game_id("mancala")
    g = oz.GameEnv(c4)
    assert oz_heuristic(oz, g) == 0.0                      # symmetric start
    g.play(3)                                              # white in the centre column, black to move
    # alignments through (col 4, row 1): 4 horizontal, 1 vertical, 1 + 1 diagonal = 7; black (the mover) loses those 7 empty
    # alignments (7 x 0.001), white holds one stone in each (7 x 0.01): mover's view = (69 - 7) e - ((69 - 7) e + 7 x 0.01)
    h = oz_heuristic(oz, g)
    assert abs(h - (-0.07)) < 1e-12 and h < 0
    t = oz.GameEnv(ttt)
    assert oz_heuristic(oz, t) == 0.0
    t.play(4)                                              # centre: 4 alignments with N = 1 for white, 4 empty ones left for black
    ht = oz_heuristic(oz, t)                               # black to move: 4 x 0.09 - (4 x 0.3 + 4 x 0.09)
    assert abs(ht - (-1.2)) < 1e-12
    # mancala: UInt8 stores (games/mancala/game.jl:21) -> the difference wraps modulo 256 (reproduced quirk)
    s = bytearray(oz.GameEnv(man).state())
    s[0], s[1] = 2, 5
    assert oz_heuristic(oz, oz.GameEnv(man, bytes(s))) == 253.0        # white to move: UInt8(2) - UInt8(5)
    s[14] = 2
    assert oz_heuristic(oz, oz.GameEnv(man, bytes(s))) == 3.0          # black: -(UInt8(253))
    s[0], s[1], s[14] = 9, 4, 1
    assert oz_heuristic(oz, oz.GameEnv(man, bytes(s))) == 5.0


@pytest.mark.parametrize("game", GAMES)
def test_heuristic_engine_equals_oracle(az, oz, game):
    gs, gid = az.GameSpec(game), oz.game_id(game)
    states = gs.random_positions(2024, 3000, 9 if game == "tictactoe" else 40)
    for s in states:
        assert gs.heuristic_value(s) == oz_heuristic(oz, oz.GameEnv(gid, bytes(s))), bytes(s)
    with pytest.raises(az.AzError):
        az.GameSpec("grid-world").minmax_think(az.GameSpec("grid-world").init_state(), az.MinMaxTS(2, True))


@pytest.mark.parametrize("game,depth,amplify,tau", [("connect-four", 5, True, 0.2),     # games/connect-four/params.jl benchmark baseline
                                                     ("connect-four", 2, False, 0.0), ("connect-four", 3, True, 1.0),
                                                     ("tictactoe", 5, True, 0.2), ("tictactoe", 3, False, 0.0),
                                                     ("mancala", 4, True, 0.5), ("mancala", 6, False, 0.0)])
def test_minmax_engine_equals_oracle(az, oz, game, depth, amplify, tau):
    """think(::MinMax.Player): product code (explicit stack, bitboards) against the oracle's recursion on random positions."""
    gs, gid = az.GameSpec(game), oz.game_id(game)
    L = oz.lib()
    A = gs.num_actions
    states = gs.random_positions(77 + depth, 120 if depth >= 5 else 400, 9 if game == "tictactoe" else 36)
    ninf = 0
    for s in states:
        q, pi = gs.minmax_think(s, az.MinMaxTS(depth, amplify, tau))
        g = oz.GameEnv(gid, bytes(s))
        acts, opi, oq = (C.c_int * 16)(), (C.c_double * 16)(), (C.c_double * 16)()
        n = L.oz_minmax_think(C.byref(g.g), depth, int(amplify), tau, 1.0, acts, opi, oq)
        wq, wpi = np.zeros(A), np.zeros(A)
        for i in range(n):
            wq[acts[i]], wpi[acts[i]] = oq[i], opi[i]
        assert (q.view(np.uint64) == wq.view(np.uint64)).all() or (q == wq).all(), (bytes(s), q, wq)
        assert (pi.view(np.uint64) == wpi.view(np.uint64)).all(), (bytes(s), pi, wpi)
        assert abs(pi.sum() - 1.0) < 1e-12
        ninf += int(np.isinf(q).any())
    if amplify and game != "mancala":
        assert ninf > 0       # forced wins / losses inside the horizon occur in the sample


@pytest.mark.parametrize("game,depth,amplify,tau", [("tictactoe", 3, True, 0.3), ("connect-four", 2, True, 0.2), ("mancala", 2, False, 0.7),
                                                     ("tictactoe", 9, True, 0.0)])
def test_minmax_oracle_equals_transcription(oz, game, depth, amplify, tau):
    """The C restatement against the line-by-line Python transcription of src/minmax.jl (small depths: Python recursion)."""
    gid = oz.game_id(game)
    L = oz.lib()
    states = oz.random_positions(gid, 5, 12 if depth < 9 else 3, 5 if game == "tictactoe" else 30)
    for s in states:
        g = oz.GameEnv(gid, bytes(s))
        acts, opi, oq = (C.c_int * 16)(), (C.c_double * 16)(), (C.c_double * 16)()
        n = L.oz_minmax_think(C.byref(g.g), depth, int(amplify), tau, 1.0, acts, opi, oq)
        pa, pq, ppi = py_think(oz, g, depth, amplify, tau, 1.0)
        assert n == len(pa) and list(acts[:n]) == list(pa)
        assert (np.array(oq[:n]) == pq).all()
        assert (np.array(opi[:n]).view(np.uint64) == ppi.view(np.uint64)).all()


def test_minmax_plays_the_obvious_move(az, oz):
    """amplify_rewards: an immediate win gets all the mass; with two plies a forced block is the only non-losing move."""
    gs, gid = az.GameSpec("connect-four"), oz.game_id("connect-four")
    g = oz.GameEnv(gid)
    for a in (0, 6, 0, 6, 0, 5):       # white has three in column 1, white to move
        g.play(a)
    q, pi = gs.minmax_think(np.frombuffer(g.state(), np.uint8), az.MinMaxTS(1, True, 0.2))
    assert q[0] == math.inf and pi[0] == 1.0 and pi.sum() == 1.0
    g.play(3)                           # white ignores it; black must now block column 1
    q, pi = gs.minmax_think(np.frombuffer(g.state(), np.uint8), az.MinMaxTS(2, True, 0.2))
    assert (q[1:] == -math.inf).all() and q[0] > -math.inf and pi[0] == 1.0
    # everything loses: uniform (src/minmax.jl:92-93)
    t = oz.GameEnv(oz.game_id("tictactoe"))
    for a in (0, 1, 4, 2):              # white: 0, 4 ; black: 1, 2 ; white to move wins at 8 -- give black the move instead
        t.play(a)
    t.play(3)                           # white 0,4,3: threatens 8 (0-4-8) and 5 (3-4-5): black cannot stop both
    ts = az.GameSpec("tictactoe")
    q, pi = ts.minmax_think(np.frombuffer(t.state(), np.uint8), az.MinMaxTS(2, True, 0.5))
    legal = ts.actions_mask(np.frombuffer(t.state(), np.uint8))
    assert (q[legal] == -math.inf).all() and np.allclose(pi[legal], 1.0 / legal.sum())
