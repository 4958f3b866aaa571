"""Generates tests/golden/oracle_kats.json: known-answer vectors of the CPU oracle (oracle/az_oracle.c) for MCTS statistics and
self-play traces.  They do NOT pin the oracle to the reference (Julia cannot run here: SURVEY 8c, "parity unpinned"); they
freeze the oracle's own behaviour so that an accidental change of the restatement -- which every GPU parity test is
measured against -- is caught by the CPU suite.  Run from the repo root:  python tests/golden/make_oracle_kats.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def explore_case(oz, game, orc, nsims, n_roots, seed, eps):
    gid = oz.game_id(game)
    roots = oz.random_positions(gid, seed, n_roots, 12 if game != "tictactoe" else 4)
    out = []
    for i, r in enumerate(roots):
        g = oz.GameEnv(gid, r)
        nl = int(g.actions_mask().sum())
        eta = oz.dirichlet(seed, i, 0, nl, 1.0)
        env = oz.Env(gid, orc, cpuct=2.0, noise_eps=eps)
        env.explore(g, nsims, eta)
        _, N, W, P, V = env.root_stats(g)
        out.append(dict(root=bytes(r).hex(), N=[int(x) for x in N], W=[float(x).hex() for x in W], P=[float(np.float32(x)).hex() for x in P],
                        nodes=int(env.num_nodes), traversed=int(env.total_nodes_traversed)))
    return dict(kind="explore", game=game, oracle=orc, nsims=nsims, seed=seed, noise_eps=eps, roots=out)


def selfplay_case(oz, game, orc, nsims, workers, games, seed, **kw):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import simref
    gid = oz.game_id(game)
    mp = oz.mcts_params(gamma=0.95, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=nsims, sched_xs=(0, 6), sched_ys=(1.0, 0.3))
    traces, slot_of = simref.oracle_simulate(oz, gid, orc, mp, seed, workers, games, 2, **kw)
    out = []
    for g in sorted(traces):
        t = traces[g]
        h = hashlib.sha256()
        for k in ("states", "pi64", "mask", "action", "rewards", "z", "t", "sym", "think_states"):
            h.update(np.ascontiguousarray(t[k]).tobytes())
        out.append(dict(game=g, worker=int(slot_of[g]), n_moves=int(t["n_moves"]), actions=[int(a) for a in t["action"]], sha256=h.hexdigest(),
                        mem_nodes=int(t["mem_nodes"]), edepth=float(t["edepth"]).hex(), total_reward=float(t["total_reward"]).hex(),
                        colors_flipped=bool(t["colors_flipped"])))
    return dict(kind="selfplay", game=game, oracle=orc, nsims=nsims, workers=workers, seed=seed, kw={k: (v if not isinstance(v, str) else v) for k, v in kw.items()},
                games=out)


def minmax_case(oz, game, depth, amplify, tau, n_pos, seed):
    """think(::MinMax.Player) (src/minmax.jl:83-114): root q-values and move distribution, bit patterns."""
    import ctypes as C
    gid = oz.game_id(game)
    out = []
    for r in oz.random_positions(gid, seed, n_pos, 20 if game != "tictactoe" else 4):
        g = oz.GameEnv(gid, r)
        acts, pi, qs = (C.c_int * 16)(), (C.c_double * 16)(), (C.c_double * 16)()
        n = oz.lib().oz_minmax_think(C.byref(g.g), depth, int(amplify), tau, 1.0, acts, pi, qs)
        out.append(dict(root=bytes(r).hex(), actions=[int(a) for a in acts[:n]], q=[float(x).hex() for x in qs[:n]], pi=[float(x).hex() for x in pi[:n]],
                        heuristic=float(oz.lib().oz_heuristic_value(C.byref(g.g))).hex()))
    return dict(kind="minmax", game=game, depth=depth, amplify=bool(amplify), tau=tau, seed=seed, positions=out)


def baseline_player_case(oz, game, baseline_kind, workers, games, seed, **kw):
    """Duels against the non-MCTS baseline players of src/benchmark.jl: NetworkOnly(tau) and MinMaxTS(depth, amplify, tau)."""
    from tests import simref
    gid = oz.game_id(game)
    mp = oz.mcts_params(gamma=1.0, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=16, sched_xs=(0,), sched_ys=(0.5,))
    if baseline_kind == "network_only":
        ob, borc = oz.mcts_params(num_iters_per_turn=0, sched_xs=(0,), sched_ys=(0.5,)), "synth"
    else:
        ob, borc = oz.minmax_params(3, True, 0.2), "uniform"
    traces, slot_of = simref.oracle_simulate(oz, gid, "synth", mp, seed, workers, games, 2, baseline=borc, omp_baseline=ob, alternate_colors=True, **kw)
    out = []
    for g in sorted(traces):
        t = traces[g]
        h = hashlib.sha256()
        for k in ("states", "pi64", "mask", "action", "rewards", "z", "t", "sym", "think_states"):
            h.update(np.ascontiguousarray(t[k]).tobytes())
        out.append(dict(game=g, worker=int(slot_of[g]), n_moves=int(t["n_moves"]), actions=[int(a) for a in t["action"]], sha256=h.hexdigest(),
                        mem_nodes=int(t["mem_nodes"]), total_reward=float(t["total_reward"]).hex(), colors_flipped=bool(t["colors_flipped"])))
    return dict(kind="baseline_player", game=game, baseline=baseline_kind, workers=workers, seed=seed, kw=kw, games=out)


def build(oz):
    cases = [explore_case(oz, "connect-four", "synth", 200, 6, 11, 0.25), explore_case(oz, "connect-four", "uniform", 120, 4, 12, 0.0),
             explore_case(oz, "tictactoe", "synth", 50, 4, 13, 0.25), explore_case(oz, "mancala", "synth", 100, 4, 14, 0.25),
             selfplay_case(oz, "connect-four", "synth", 32, 3, 6, 21), selfplay_case(oz, "tictactoe", "synth", 24, 2, 5, 22, flip_probability=0.5),
             selfplay_case(oz, "connect-four", "synth", 24, 2, 5, 23, baseline="uniform", alternate_colors=True, flip_probability=0.5),
             selfplay_case(oz, "grid-world", "synth", 20, 3, 6, 24),
             minmax_case(oz, "connect-four", 4, True, 0.2, 6, 31), minmax_case(oz, "tictactoe", 5, False, 0.0, 4, 32),
             minmax_case(oz, "mancala", 3, True, 0.5, 4, 33),
             baseline_player_case(oz, "connect-four", "network_only", 3, 6, 41, flip_probability=0.5),
             baseline_player_case(oz, "connect-four", "minmax", 3, 6, 42, flip_probability=0.5),
             baseline_player_case(oz, "mancala", "minmax", 2, 4, 43)]
    return dict(format=1, note="oracle self-consistency vectors (not reference-pinned); see make_oracle_kats.py", cases=cases)


if __name__ == "__main__":
    from oracle import oracle as oz
    oz.lib()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_kats.json")
    with open(path, "w") as f:
        json.dump(build(oz), f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
