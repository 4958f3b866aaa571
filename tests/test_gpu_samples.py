"""Replay-buffer side on the GPU against the CPU restatement oracle/samples_ref.py: samples exported
from a finished self-play run on the device, merge_by_state, augment_with_symmetries, convert_samples
(src/memory.jl:74-130, src/learning.jl:17-51).  Byte / integer fields and Float64 means are compared bit for bit
(sums run in original order on both sides); log2 weights to 1 ulp of Float32."""
import numpy as np
import pytest

import _pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    return _pkg.load()


@pytest.fixture(scope="module")
def ctx(az):
    c = az.Context(0)
    yield c
    c.close()


def _same(gid, oz, got, want):
    from oracle import samples_ref as sr
    st, pi, z, t, n = sr.to_arrays(gid, want)
    assert len(got["z"]) == len(want)
    assert (got["states"] == st).all()
    assert (got["pi"].view(np.uint64) == pi.view(np.uint64)).all()
    assert (got["z"].view(np.uint64) == z.view(np.uint64)).all() and (got["t"] == t).all() and (got["n"] == n).all()


@pytest.mark.parametrize("game,flip", [("connect-four", 0.0), ("tictactoe", 0.0), ("connect-four", 0.5), ("mancala", 0.0)])
def test_export_merge_augment_convert(az, oz, ctx, game, flip):
    from oracle import samples_ref as sr
    from tests import simref
    gs, gid = az.GameSpec(game), oz.game_id(game)
    S, NG, nsims, seed = 8, 24, 32, 2024
    mp = az.MctsParams(gamma=0.9, cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0], [1.0]), dirichlet_noise_eps=0.25,
                       dirichlet_noise_alpha=1.0)
    net = az.SynthOracle(ctx, gs)
    sp = az.SelfPlay(ctx, gs, net, az.SelfPlayParams(mp, az.SimParams(num_games=NG, num_workers=S, batch_size=S, reset_every=2,
                                                                       flip_probability=flip)), seed=seed)
    sp.start()
    sp.wait()
    smp = az.Samples.from_selfplay(sp)
    omp = oz.mcts_params(gamma=0.9, cpuct=2.0, noise_eps=0.25, noise_alpha=1.0, num_iters_per_turn=nsims)
    traces, _ = simref.oracle_simulate(oz, gid, "synth", omp, seed, S, NG, 2, flip_probability=flip)
    ref = sr.samples_from_traces(traces)
    _same(gid, oz, smp.fetch(), ref)                                  # push_trace! rows, Float64 pi and z (gamma = 0.9)
    # merge_by_state: first-occurrence order on both sides
    merged = smp.merge_by_state()
    rmerged = sr.merge_by_state(ref)
    assert len(rmerged) < len(ref)
    _same(gid, oz, merged.fetch(), rmerged)
    # augment_with_symmetries, then merge again (images of different samples collide)
    if gs.name in ("connect-four", "tictactoe"):
        aug = smp.augment_with_symmetries()
        raug = sr.augment_with_symmetries(gid, ref)
        _same(gid, oz, aug.fetch(), raug)
        am = aug.merge_by_state()
        _same(gid, oz, am.fetch(), sr.merge_by_state(raug))
        final, rfinal = am, sr.merge_by_state(raug)
    else:
        aug = smp.augment_with_symmetries()                            # no symmetries declared: identity
        _same(gid, oz, aug.fetch(), ref)
        final, rfinal = merged, rmerged
    for w in (az.CONSTANT_WEIGHT, az.LOG_WEIGHT, az.LINEAR_WEIGHT):
        got, want = final.convert(w), sr.convert_samples(gid, w, rfinal)
        for k in ("X", "A", "P", "V"):
            assert (got[k] == want[k]).all(), k
        assert np.allclose(got["W"], want["W"], rtol=2e-7, atol=0)
    # concat = append!(buf, experience); a host-built set behaves like the exported one
    st, pi, z, t, n = sr.to_arrays(gid, ref)
    host = az.Samples.from_host(ctx, gs, st, pi, z, t, n)
    both = smp.concat(host)
    _same(gid, oz, both.fetch(), ref + ref)
    mm = both.merge_by_state().fetch()
    assert mm["n"].sum() == 2 * len(ref) and (mm["states"] == merged.fetch()["states"]).all()
    for s in (smp, merged, aug, host, both):
        s.close()
    sp.close()
    net.close()


def test_samples_scale_and_errors(az, ctx):
    """A large synthetic set (1 M samples drawn from 50 k distinct positions): counts, n and mass are conserved."""
    gs = az.GameSpec("connect-four")
    rng = np.random.default_rng(0)
    base = gs.random_positions(7, 50000, 30)
    base = np.unique(base, axis=0)
    idx = rng.integers(0, len(base), 1_000_000)
    st = base[idx]
    mask = np.stack([gs.actions_mask(s) for s in base])[idx]
    pi = rng.random((len(idx), 7)) * mask
    pi /= pi.sum(1, keepdims=True)
    z = rng.choice([-1.0, 0.0, 1.0], len(idx))
    t = rng.integers(1, 43, len(idx)).astype(np.float64)
    s = az.Samples.from_host(ctx, gs, st, pi, z, t)
    m = s.merge_by_state()
    out = m.fetch()
    assert len(m) == len(np.unique(idx)) and out["n"].sum() == len(idx)
    assert np.allclose(out["pi"].sum(1), 1, atol=1e-9)
    first = {}
    for k, i in enumerate(idx[:200000]):
        first.setdefault(int(i), k)
    order = sorted(first, key=first.get)[:1000]               # groups come in order of first occurrence
    assert (out["states"][:len(order)] == base[order]).all()
    j = order[0]
    sel = np.flatnonzero(idx == j)
    acc = z[sel[0]]
    for k in sel[1:]:
        acc = acc + z[k]
    assert out["z"][0] == acc / len(sel) and out["n"][0] == len(sel)
    a = m.augment_with_symmetries()
    assert len(a) == 2 * len(m)
    c = a.convert(az.LOG_WEIGHT)
    assert c["X"].shape == (len(a), 126) and (c["A"].sum(1) >= 1).all() and np.allclose(c["P"].sum(1), 1, atol=1e-5)
    with pytest.raises(az.AzError):
        ttt = az.GameSpec("tictactoe")
        other = az.Samples.from_host(ctx, ttt, np.zeros((1, 10), np.uint8) + np.array([0] * 9 + [1], np.uint8), np.full((1, 9), 1 / 9), [0.0], [1.0])
        s.concat(other)
    for x in (s, m, a):
        x.close()
