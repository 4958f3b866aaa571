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


def allgather_samples(samples, first_game, dist=None, device="cpu"):
    """All-gather the per-rank sample rows (dict of numpy arrays as returned by SelfPlay.fetch) so that every rank holds
    the whole iteration's samples in rank order.  Row counts differ per rank: counts are gathered first, rows are padded
    to the maximum, gathered with one all_gather per field group, then trimmed."""
    import torch
    n = len(samples["z"])
    local = {k: np.ascontiguousarray(samples[k]) for k in SAMPLE_KEYS}
    local["game"] = local["game"] + np.int32(first_game)  # local game index -> global game index
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    out = {}
    for k in SAMPLE_KEYS:
        a = local[k]
        pad = np.zeros((m,) + a.shape[1:], a.dtype)
        pad[:n] = a
        t = torch.from_numpy(pad).to(device)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        out[k] = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)])
    return out
