"""Multi-GPU host logic: one rank per GPU, games sharded over ranks, ONE all-gather of samples at iteration end.

Mirrors simulate_distributed (src/simulations.jl:252-290): `num_each, rem = divrem(num_games, nworkers)`, the first
worker takes `num_each + rem`; results are concatenated in rank order (reduce(vcat, results), :289).  There is no
collective inside the simulation loop.  Uses torch.distributed only as plumbing (nccl on GPUs, gloo in CPU tests).
"""
import numpy as np

SAMPLE_KEYS = ("states", "pi", "mask", "z", "t", "game", "rewards", "actions")


def split_games(num_games, world, rank):
    """(count, first_global_game_index) for this rank; src/simulations.jl:268,277."""
    num_each, rem = divmod(num_games, world)
    assert num_each >= 1, "simulate_distributed asserts num_each >= 1 (src/simulations.jl:269)"
    count = num_each + rem if rank == 0 else num_each
    first = 0 if rank == 0 else num_each * rank + rem
    return count, first


def _single(dist):
    return dist is None or not dist.is_initialized() or dist.get_world_size() == 1


def allgather_counts(n, dist, device="cpu"):
    import torch
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, cnt)
    return [int(c.item()) for c in counts]


def allgather_rows(a, dist, device="cpu", counts=None):
    """Concatenate the per-rank arrays `a` (rows differ per rank) in rank order on every rank: counts first (unless the
    caller already has them), then ONE all_gather of rows padded to the maximum count."""
    import torch
    a = np.ascontiguousarray(a)
    if _single(dist):
        return a
    world = dist.get_world_size()
    if counts is None:
        counts = allgather_counts(len(a), dist, device)
    pad = np.zeros((max(counts),) + a.shape[1:], a.dtype)
    pad[:len(a)] = a
    t = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)])


def allgather_samples(samples, first_game, dist=None, device="cpu"):
    """All-gather the per-rank sample rows (dict of numpy arrays as returned by SelfPlay.fetch) so that every rank holds
    the whole iteration's samples in rank order (reduce(vcat, results), src/simulations.jl:289)."""
    local = {k: np.ascontiguousarray(samples[k]) for k in SAMPLE_KEYS}
    local["game"] = local["game"] + np.int32(first_game)  # local game index -> global game index
    if _single(dist):
        return local
    counts = allgather_counts(len(local["z"]), dist, device)
    return {k: allgather_rows(local[k], dist, device, counts) for k in SAMPLE_KEYS}


def allgather_outcomes(samples, outcomes, dist=None, device="cpu"):
    """rewards_and_redundancy over the games of ALL ranks (src/simulations.jl:282-307): per-game rewards and colors_flipped
    concatenated in rank order; redundancy = 1 - |unique states| / |states| over every trace state of every rank (the
    per-rank value az_selfplay_outcomes returns only sees its own games).  `samples` = SelfPlay.fetch(), `outcomes` =
    SelfPlay.outcomes()."""
    rewards = allgather_rows(outcomes["game_rewards"], dist, device)
    flipped = allgather_rows(outcomes["colors_flipped"], dist, device)
    states = np.concatenate([allgather_rows(samples["states"], dist, device), allgather_rows(outcomes["final_states"], dist, device)])
    uniq = len(np.unique(states, axis=0)) if len(states) else 0
    return dict(game_rewards=rewards, colors_flipped=flipped, redundancy=(1.0 - uniq / len(states)) if len(states) else 0.0)
