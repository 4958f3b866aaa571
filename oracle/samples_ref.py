"""TEST INFRASTRUCTURE ONLY (see oracle/az_oracle.h): CPU restatement of the replay-buffer side of the reference --
merge_by_state / augment_with_symmetries (src/memory.jl:89-130) and convert_samples (src/learning.jl:17-51).  Parity
unpinned: the reference's tests hold no golden vectors for these functions.

A sample is a dict(s=bytes, pi=[Float64 over the legal actions of s, ascending], z, t, n) = TrainingSample
(src/memory.jl:20-26).  Python floats are IEEE doubles and every sum below is left to right, like the reference's
`mean(x for x in samples)` (sum of a generator / count)."""
import ctypes as C
import math

import numpy as np

from . import oracle as oz


def _mask(gid, s):
    return oz.GameEnv(gid, np.frombuffer(s, np.uint8)).actions_mask().astype(bool)


def samples_from_traces(traces):
    """push_trace! (src/memory.jl:74-87) over oracle traces (dicts of tests/simref.py), games in index order.  The policy
    vector is the one recorded in the frame the player thought in (mask `mask`), paired with trace.states[i]."""
    out = []
    for _, tr in sorted(traces.items()):
        for i in range(tr["n_moves"]):
            pi = [float(tr["pi64"][i][a]) for a in range(tr["pi64"].shape[1]) if tr["mask"][i][a]]
            out.append(dict(s=bytes(tr["states"][i]), pi=pi, z=float(tr["z"][i]), t=float(tr["t"][i]), n=1))
    return out


def merge_samples(es):  # src/memory.jl:89-96
    k = len(es)
    pi = list(es[0]["pi"])
    z, t, n = es[0]["z"], es[0]["t"], es[0]["n"]
    for e in es[1:]:
        pi = [a + b for a, b in zip(pi, e["pi"])]
        z, t, n = z + e["z"], t + e["t"], n + e["n"]
    return dict(s=es[0]["s"], pi=[p / k for p in pi], z=z / k, t=t / k, n=n)


def merge_by_state(samples):  # src/memory.jl:98-110 (Dict order is unspecified there; here: first occurrence)
    d = {}
    for e in samples:
        d.setdefault(e["s"], []).append(e)
    return [merge_samples(es) for es in d.values()]


_APERM = {}


def _aperm(gid, j):
    """Action permutation of symmetry j, derived from the images of one-stone boards: aperm[p] = q such that a stone on the
    cell of action q lands on the cell of action p (connect-four: bottom cell of column q is byte q; tic-tac-toe: cell q)."""
    if (gid, j) not in _APERM:
        L = oz.lib()
        A, sb = oz.num_actions(gid), oz.state_bytes(gid)
        perm = [None] * A
        for q in range(A):
            probe = np.zeros(oz.STATE_BYTES, np.uint8)
            probe[q] = 1
            probe[sb - 1] = 1
            pim = np.zeros(oz.STATE_BYTES, np.uint8)
            L.oz_apply_symmetry(gid, j, probe.ctypes.data, pim.ctypes.data)
            (p,) = [p for p in range(A) if pim[p] == 1]
            perm[p] = q
        _APERM[(gid, j)] = perm
    return _APERM[(gid, j)]


def symmetries(gid, s):
    """GI.symmetries(gspec, state) -> [(symstate, aperm)] (games/connect-four/game.jl:252-257, games/tictactoe/game.jl:164-168)."""
    L = oz.lib()
    sb = oz.state_bytes(gid)
    out = []
    for j in range(L.oz_num_symmetries(gid)):
        src = np.zeros(oz.STATE_BYTES, np.uint8)
        src[:sb] = np.frombuffer(s, np.uint8)
        img = np.zeros(oz.STATE_BYTES, np.uint8)
        L.oz_apply_symmetry(gid, j, src.ctypes.data, img.ctypes.data)
        out.append((bytes(img[:sb]), _aperm(gid, j)))
    return out


def apply_symmetry(gid, e, sym):  # src/memory.jl:112-124
    symstate, aperm = sym
    mask, symmask = _mask(gid, e["s"]), _mask(gid, symstate)
    full = [0.0] * len(mask)
    it = iter(e["pi"])
    for a in range(len(mask)):
        if mask[a]:
            full[a] = next(it)
    full = [full[aperm[a]] for a in range(len(mask))]
    assert all(full[a] == 0.0 for a in range(len(mask)) if not symmask[a])
    return dict(s=symstate, pi=[full[a] for a in range(len(mask)) if symmask[a]], z=e["z"], t=e["t"], n=e["n"])


def augment_with_symmetries(gid, samples):  # src/memory.jl:126-130
    return list(samples) + [apply_symmetry(gid, e, sym) for e in samples for sym in symmetries(gid, e["s"])]


def convert_samples(gid, weighing, samples):  # src/learning.jl:17-51; weighing 0 constant, 1 log, 2 linear
    A = oz.num_actions(gid)
    W, X, Am, P, V = [], [], [], [], []
    for e in samples:
        W.append(np.float32(1.0 if weighing == 0 else (math.log2(e["n"]) + 1 if weighing == 1 else e["n"])))
        X.append(oz.vectorize_state(gid, np.frombuffer(e["s"], np.uint8)).reshape(-1, order="F"))
        m = _mask(gid, e["s"])
        p = np.zeros(A)
        p[m] = e["pi"]
        Am.append(m.astype(np.float32))
        P.append(p.astype(np.float32))
        V.append(np.float32(e["z"]))
    return dict(W=np.array(W, np.float32), X=np.array(X, np.float32).reshape(len(samples), -1), A=np.array(Am, np.float32).reshape(len(samples), A),
                P=np.array(P, np.float32).reshape(len(samples), A), V=np.array(V, np.float32))


def to_arrays(gid, samples):
    """(states, pi A-wide, z, t, n) arrays in the C-ABI layout of az_samples_from_host / az_samples_fetch."""
    A, sb = oz.num_actions(gid), oz.state_bytes(gid)
    k = len(samples)
    st = np.zeros((k, sb), np.uint8)
    pi = np.zeros((k, A))
    for i, e in enumerate(samples):
        st[i] = np.frombuffer(e["s"], np.uint8)
        pi[i][_mask(gid, e["s"])] = e["pi"]
    return st, pi, np.array([e["z"] for e in samples], np.float64), np.array([e["t"] for e in samples], np.float64), \
        np.array([e["n"] for e in samples], np.int32)
