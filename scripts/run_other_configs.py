#!/usr/bin/env python
"""Runs BASELINE.json configs[3] (mancala, 4096 games x 400 sims, variable action mask) and configs[4] (grid-world,
8192 envs x 200 sims, single-player value backup) through the public mirror at full size, checks size-independent
invariants and prints throughput.  These configs are parity-test cases (tests/), not bench lines."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

az = _pkg.load()
ctx = az.Context(0)
out = {}

# ---- config[3]: mancala, ResNet (5 blocks x 128 as shipped in games/mancala/params.jl), 4096 trees x 400 sims ----
gs = az.GameSpec("mancala")
hp = az.ResNetHP(5, 128, (3, 3), 32, 32)
net = az.ResNet(ctx, gs, hp).load(az.fresh_resnet_blob(gs, hp, seed=1))
S, nsims = 4096, 400
roots = gs.random_positions(11, S, 30)
rng = np.random.default_rng(0)
eta = np.zeros((S, 6))
for i, r in enumerate(roots):
    n = int(gs.actions_mask(r).sum())
    e = rng.exponential(size=n)
    eta[i, :n] = e / e.sum()
mp = az.MctsParams(cpuct=1.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.2, dirichlet_noise_alpha=1.0)  # games/mancala/params.jl
env = az.MctsEnv(ctx, gs, net, mp, S, nsims + 8)
env.explore(roots, nsims, eta)
N, W, P = env.explore(roots, nsims, eta) if False else env.root_stats()
env.reset()
t0 = time.perf_counter()
N, W, P = env.explore(roots, nsims, eta)
dt = time.perf_counter() - t0
t = env.last_timing()
legal = np.stack([gs.actions_mask(r) for r in roots])
assert (N.sum(1) == nsims - 1).all() and (N[~legal] == 0).all() and np.allclose(P.sum(1), 1, atol=1e-5)
out["mancala_4096x400_resnet5"] = dict(expansions_per_s=t["expansions"] / dt, sims_per_s=S * nsims / dt, seconds=dt,
                                       mean_legal_actions=float(legal.sum(1).mean()))
env.close()
net.close()

# ---- config[4]: grid-world, SimpleNet(100, 4) (games/grid-world/params.jl), 8192 envs x 200 sims, eps = 0 ----
gs = az.GameSpec("grid-world")
hp = az.SimpleNetHP(100, 4)
net = az.SimpleNet(ctx, gs, hp).load(az.fresh_simplenet_blob(gs, hp, seed=1))
S, nsims = 8192, 200
roots = gs.random_positions(5, S)
mp = az.MctsParams(gamma=1.0, cpuct=1.0, num_iters_per_turn=nsims, dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)
env = az.MctsEnv(ctx, gs, net, mp, S, 128)
env.set_noise(123, np.arange(S), np.zeros(S, np.int32))
env.explore(roots, nsims)
env.reset()
t0 = time.perf_counter()
N, W, P = env.explore(roots, nsims)
dt = time.perf_counter() - t0
ts, tn, nn = env.counters()
assert (N.sum(1) >= nsims - 1).all() and (nn <= 100).all()  # a path may pass through the root state again (time is not in the key)
out["gridworld_8192x200_simplenet"] = dict(sims_per_s=S * nsims / dt, expansions_per_s=env.last_timing()["expansions"] / dt, seconds=dt,
                                           mean_exploration_depth=float(tn.sum() / ts.sum()), max_nodes_per_tree=int(nn.max()))
env.close()
net.close()
ctx.close()
print(json.dumps(out))
