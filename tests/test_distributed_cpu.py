"""N>1 host path on CPU: game split (src/simulations.jl:268-277) and the sample all-gather with the gloo backend."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dist():
    import importlib.util
    spec = importlib.util.spec_from_file_location("az_distributed", os.path.join(ROOT, "alphazero.jl_b200", "distributed.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_split_games_matches_divrem():
    d = _load_dist()
    for num_games, world in [(5000, 8), (32768, 8), (4096, 1), (10, 3), (7, 7)]:
        parts = [d.split_games(num_games, world, r) for r in range(world)]
        assert sum(c for c, _ in parts) == num_games
        num_each, rem = divmod(num_games, world)
        assert parts[0][0] == num_each + rem and all(c == num_each for c, _ in parts[1:])
        nxt = 0
        for c, f in parts:  # contiguous, disjoint global game ranges in rank order
            assert f == nxt
            nxt += c
    with pytest.raises(AssertionError):
        d.split_games(3, 4, 0)


def _fake_samples(rank, n):
    rng = np.random.default_rng(rank)
    return dict(states=rng.integers(0, 3, (n, 43)).astype(np.uint8), pi=rng.random((n, 7)).astype(np.float32),
                mask=rng.integers(0, 2, (n, 7)).astype(np.uint8), z=rng.random(n).astype(np.float32),
                t=rng.random(n).astype(np.float32), game=rng.integers(0, 4, n).astype(np.int32),
                rewards=rng.random(n), actions=rng.integers(0, 7, n).astype(np.int32))


def _fake_outcomes(rank):
    rng = np.random.default_rng(100 + rank)
    g = 3 + rank
    return dict(game_rewards=rng.choice([-1.0, 0.0, 1.0], g), colors_flipped=rng.integers(0, 2, g).astype(np.int32),
                final_states=rng.integers(0, 2, (g, 43)).astype(np.uint8), redundancy=0.0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = _load_dist()
    n = 5 + 3 * rank
    smp = _fake_samples(rank, n)
    out = d.allgather_samples(smp, first_game=100 * rank, dist=dist, device="cpu")
    oc = d.allgather_outcomes(smp, _fake_outcomes(rank), dist=dist, device="cpu")
    q.put((rank, {k: v.copy() for k, v in out.items()}, oc))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_samples_gloo_world2():
    import torch.multiprocessing as mp
    d = _load_dist()
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    res = {r: a for r, a, _ in got}
    ocs = {r: o for r, _, o in got}
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = {}
    for k in d.SAMPLE_KEYS:
        exp[k] = np.concatenate([_fake_samples(r, 5 + 3 * r)[k] + (np.int32(100 * r) if k == "game" else 0) for r in range(world)])
    for r in range(world):
        for k in d.SAMPLE_KEYS:
            assert res[r][k].dtype == exp[k].dtype and (res[r][k] == exp[k]).all(), (r, k)
    # arena outcomes: rewards / colours in rank order, redundancy over the union of all ranks' trace states
    rw = np.concatenate([_fake_outcomes(r)["game_rewards"] for r in range(world)])
    fl = np.concatenate([_fake_outcomes(r)["colors_flipped"] for r in range(world)])
    sts = np.concatenate([_fake_samples(r, 5 + 3 * r)["states"] for r in range(world)] + [_fake_outcomes(r)["final_states"] for r in range(world)])
    red = 1.0 - len({bytes(x) for x in sts}) / len(sts)
    for r in range(world):
        assert (ocs[r]["game_rewards"] == rw).all() and (ocs[r]["colors_flipped"] == fl).all() and ocs[r]["redundancy"] == red


def test_allgather_single_rank_passthrough():
    d = _load_dist()
    s = _fake_samples(0, 4)
    out = d.allgather_samples(s, first_game=7)
    assert (out["game"] == s["game"] + 7).all() and (out["pi"] == s["pi"]).all()
    oc = d.allgather_outcomes(s, _fake_outcomes(0))
    assert (oc["game_rewards"] == _fake_outcomes(0)["game_rewards"]).all() and 0.0 <= oc["redundancy"] < 1.0
