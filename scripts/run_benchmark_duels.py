#!/usr/bin/env python
"""The evaluation benchmark of the shipped Connect-Four experiment (games/connect-four/params.jl:73-104) at its own
parameters, all four duels (the reference keeps the two MinMax ones commented out): AlphaZero (arena MctsParams: 600
iterations, cpuct 2, tau 0.2, noise 0.05) and the network-only player (tau 0.5) against MctsRollouts (1000 iterations,
cpuct 1) and MinMaxTS (depth 5, tau 0.2, amplify_rewards); 256 games, 256 workers, reset_every 2, flip_probability 0.5,
no colour alternation.  Random-init 5-block ResNet (no checkpoints in the image): the rewards say nothing about playing
strength, the line records that every duel of `Benchmark.run` runs on the engine and how fast."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

az = _pkg.load()
ctx = az.Context(0)
gs = az.GameSpec("connect-four")
hp = az.ResNetHP(5, 128, (3, 3), 32, 32)
net = az.ResNet(ctx, gs, hp).load(az.fresh_resnet_blob(gs, hp, seed=1))
arena_mcts = az.MctsParams(num_iters_per_turn=600, cpuct=2.0, temperature=az.ConstSchedule(0.2), dirichlet_noise_eps=0.05, dirichlet_noise_alpha=1.0)
rollout_mcts = az.MctsParams(num_iters_per_turn=1000, cpuct=1.0, temperature=az.ConstSchedule(0.2), dirichlet_noise_eps=0.05, dirichlet_noise_alpha=1.0)
sim = az.SimParams(num_games=256, num_workers=256, batch_size=256, reset_every=2, flip_probability=0.5, alternate_colors=False)
rollouts = az.RolloutOracle(ctx, gs, gamma=1.0, seed=7)
minmax = az.MinMaxTS(depth=5, amplify_rewards=True, tau=0.2)
duels = [("AlphaZero / MCTS rollouts (1000)", arena_mcts, dict(baseline=rollouts, baseline_mcts=rollout_mcts)),
         ("AlphaZero / MinMax (depth 5)", arena_mcts, dict(baseline=minmax)),
         ("Network Only / MCTS rollouts (1000)", az.NetworkOnly(0.5), dict(baseline=rollouts, baseline_mcts=rollout_mcts)),
         ("Network Only / MinMax (depth 5)", az.NetworkOnly(0.5), dict(baseline=minmax))]
out = {}
for name, mp, kw in duels:
    t0 = time.perf_counter()
    r = az.simulate(ctx, gs, net, az.SelfPlayParams(mp, sim), seed=2024, gamma=1.0, **kw)
    dt = time.perf_counter() - t0
    rew = r["game_rewards"]
    out[name] = dict(games=int(len(rew)), seconds=dt, games_per_s=len(rew) / dt, moves=int(r["samples"]), avg_reward=float(rew.mean()),
                     won=int((rew > 0).sum()), drawn=int((rew == 0).sum()), lost=int((rew < 0).sum()), redundancy=float(r["redundancy"]))
net.close()
rollouts.close()
ctx.close()
print(json.dumps(out))
