"""Oracle-side simulate() with the dynamic game assignment of Util.mapreduce (src/util.jl:169-200), emulating the engine's
lock-step ticks: a move lasts exactly `nsims` ticks (select runs one simulation per call), workers that finish in the
same tick take the next game indices in slot order, and a game that starts on a terminal state (grid-world) ends at once
and its worker asks again in the next round of the same tick."""
import ctypes as C

import numpy as np


def oracle_simulate(oz, gid, oracle, omp, seed, S, NG, reset_every, baseline=None, alternate_colors=False, flip_probability=0.0,
                    omp_baseline=None):
    """`baseline` given: every worker is TwoPlayers(MctsPlayer(oracle, omp), MctsPlayer(baseline, omp_baseline or omp))
    (src/training.jl:130-143; two different MctsPlayers as in Benchmark duels, src/benchmark.jl:78-99)."""
    L = oz.lib()
    def fn_ctx(o):   # (oracle function pointer, its context pointer)
        if isinstance(o, str):
            return oz.builtin_oracle(o), None
        if isinstance(o, oz.RolloutOracle):
            return o.fn, o.ctx_ptr
        return o, None
    fns = [fn_ctx(o) for o in ([oracle] if baseline is None else [oracle, baseline])]
    omps = [omp, omp_baseline if omp_baseline is not None else omp]
    mk = lambda fc, m: L.oz_env_create(gid, fc[0], fc[1], m.gamma, m.cpuct, m.noise_eps, m.noise_alpha, m.prior_temperature)
    envs = [[mk(fc, omps[i]) for i, fc in enumerate(fns)] for _ in range(S)]
    A, sb = oz.num_actions(gid), oz.state_bytes(gid)
    nsims = omp.num_iters_per_turn
    free_at = {w: 0 for w in range(S)}
    played = [0] * S
    traces, slot_of = {}, {}
    nxt = 0
    tr = oz.Trace()
    while nxt < NG:
        t = min(free_at.values())
        asking = sorted(w for w, ft in free_at.items() if ft == t)
        while asking and nxt < NG:
            again = []
            for w in asking:
                if nxt >= NG:
                    break
                g = nxt
                nxt += 1
                flipped = baseline is not None and alternate_colors and (g + 1) % 2 == 1   # src/simulations.jl:224-226
                white, black = (envs[w][0], envs[w][-1]) if not flipped else (envs[w][-1], envs[w][0])
                mw, mb = (omps[0], omps[-1] if baseline is not None else omps[0]) if not flipped else (omps[-1], omps[0])
                L.oz_play_game2p(white, C.byref(mw), black, C.byref(mb), flip_probability, seed, g, C.byref(tr))
                n = tr.n_moves
                traces[g] = dict(n_moves=n, states=np.ctypeslib.as_array(tr.states)[:n + 1, :sb].copy(),
                                 pi=np.ctypeslib.as_array(tr.pi)[:n, :A].copy(), pi64=np.ctypeslib.as_array(tr.pi64)[:n, :A].copy(), mask=np.ctypeslib.as_array(tr.mask)[:n, :A].copy(),
                                 action=np.ctypeslib.as_array(tr.action)[:n].copy(), rewards=np.ctypeslib.as_array(tr.rewards)[:n].copy(),
                                 z=np.ctypeslib.as_array(tr.z)[:n].copy(), t=np.ctypeslib.as_array(tr.t)[:n].copy(),
                                 mem_nodes=tr.mem_nodes, edepth=tr.edepth, sym=np.ctypeslib.as_array(tr.sym)[:n].copy(),
                                 think_states=np.ctypeslib.as_array(tr.think_states)[:n, :sb].copy(), colors_flipped=flipped,
                                 total_reward=L.oz_total_reward(C.byref(tr), omp.gamma))
                slot_of[g] = w
                played[w] += 1
                if reset_every > 0 and played[w] % reset_every == 0:
                    for e in envs[w]:
                        L.oz_env_reset(e)
                # every move lasts as many ticks as its player's iteration budget (select runs one simulation per call)
                ts = traces[g]["think_states"]
                # (a NetworkPlayer, num_iters_per_turn == 0, takes the one tick of its root evaluation)
                free_at[w] = t + sum(max(1, (mw if (sb == 2 or ts[i][sb - 1] == 1) else mb).num_iters_per_turn) for i in range(n))
                if n == 0:
                    again.append(w)
            asking = again
        if nxt >= NG:
            break
        # workers that did not get a game this tick because the games ran out stay idle
    for ee in envs:
        for e in ee:
            L.oz_env_destroy(e)
    return traces, slot_of


def rewards_and_redundancy(traces):
    """src/simulations.jl:292-307 over the oracle traces."""
    rewards = np.array([(-t["total_reward"] if t["colors_flipped"] else t["total_reward"]) for _, t in sorted(traces.items())])
    states = [bytes(s) for _, t in sorted(traces.items()) for s in t["states"]]
    return rewards, 1.0 - len(set(states)) / len(states)


def assert_same_samples(out, traces, check_mask=True):
    k = 0
    for g in sorted(traces):
        tr = traces[g]
        rows = np.flatnonzero(out["game"] == g)
        n = tr["n_moves"]
        assert len(rows) == n == out["moves"][g], (g, len(rows), n, out["moves"][g])
        assert (out["actions"][rows] == tr["action"]).all(), g
        assert (out["states"][rows] == tr["states"][:n]).all(), g
        assert (out["pi"][rows].view(np.uint32) == tr["pi"].view(np.uint32)).all(), g
        if check_mask:
            assert (out["mask"][rows] == tr["mask"]).all(), g
        assert (out["rewards"][rows] == tr["rewards"]).all(), g
        assert (out["z"][rows] == tr["z"].astype(np.float32)).all() and (out["t"][rows] == tr["t"]).all(), g
        assert out["nodes"][g] == tr["mem_nodes"] and out["edepth"][g] == tr["edepth"], g
        k += n
    assert k == len(out["game"]) == out["samples"]


def assert_same_outcomes(out, traces):
    """rewards_and_redundancy + the final state of every trace."""
    rewards, red = rewards_and_redundancy(traces)
    assert (out["game_rewards"] == rewards).all()
    assert out["redundancy"] == red
    for g, t in sorted(traces.items()):
        assert bytes(out["final_states"][g]) == bytes(t["states"][t["n_moves"]]), g
        assert bool(out["colors_flipped"][g]) == bool(t["colors_flipped"]), g
