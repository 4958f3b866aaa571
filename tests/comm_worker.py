"""One rank of the NCCL communicator test (spawned by tests/test_gpu_comm.py; also usable by hand):
    python tests/comm_worker.py <rank> <world> <dir>
Every rank builds ALL ranks' synthetic sample sets from seeds (so it knows the expected concatenation), uploads its own,
all-gathers through az_samples_allgather and compares bit for bit; then checks az_net_broadcast by comparing the network
outputs with a directly loaded copy of rank 0's blob.  The 128-byte communicator id travels through a file in <dir>."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rank_samples(gs, rank):
    n = 3000 + 1717 * rank + (0 if rank != 1 else -2999)   # rank 1 holds a single sample: padding is exercised
    rng = np.random.default_rng(100 + rank)
    states = gs.random_positions(55 + rank, n, 30)
    pi = rng.random((n, gs.num_actions))
    z = rng.standard_normal(n)
    t = rng.integers(1, 40, n).astype(np.float64)
    cnt = rng.integers(1, 5, n).astype(np.int32)
    return states, pi, z, t, cnt


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import _pkg
    az = _pkg.load()
    ctx = az.Context(rank)
    gs = az.GameSpec("connect-four")
    idfile = os.path.join(d, "nccl_id.bin")

    def exchange(b):
        if b is not None:
            with open(idfile + ".tmp", "wb") as f:
                f.write(b)
            os.replace(idfile + ".tmp", idfile)
            return b
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise TimeoutError("no communicator id from rank 0")
            time.sleep(0.05)
        return open(idfile, "rb").read()

    comm = az.Comm(ctx, rank, world, exchange if world > 1 else None)
    parts = [rank_samples(gs, r) for r in range(world)]
    mine = az.Samples.from_host(ctx, gs, *parts[rank][:4], n=parts[rank][4])
    for rep in range(3):
        allg, counts = comm.allgather_samples(mine)
    assert counts.tolist() == [len(p[2]) for p in parts], counts
    got = allg.fetch()
    want = [np.concatenate([p[k] for p in parts]) for k in range(5)]
    assert (got["states"] == want[0]).all()
    assert (got["pi"].view(np.uint64) == want[1].view(np.uint64)).all()
    assert (got["z"].view(np.uint64) == want[2].view(np.uint64)).all()
    assert (got["t"] == want[3]).all() and (got["n"] == want[4]).all()
    ms_gather = comm.last_ms
    # an empty local set on the last rank
    empty = az.Samples.from_host(ctx, gs, np.zeros((0, gs.state_bytes), np.uint8), np.zeros((0, gs.num_actions)), np.zeros(0), np.zeros(0))
    a2, c2 = comm.allgather_samples(empty if rank == world - 1 else mine)
    assert c2[world - 1] == 0 and len(a2) == sum(c2)
    # weights: rank 0's blob everywhere
    from tests import netcheck
    from oracle import netref
    hp = netcheck.c4_hp(2)
    blob0 = netref.make_blob(gs.state_dim, gs.num_actions, hp, seed=77, randomize=True)
    net = az.ResNet(ctx, gs, az.ResNetHP(hp["num_blocks"], hp["num_filters"], hp["conv_kernel_size"], 32, 32))
    if rank != 0:
        net.load(netref.make_blob(gs.state_dim, gs.num_actions, hp, seed=78 + rank, randomize=True))  # something else first
    comm.broadcast_network(net, blob0 if rank == 0 else None, root=0)
    ref = az.ResNet(ctx, gs, az.ResNetHP(hp["num_blocks"], hp["num_filters"], hp["conv_kernel_size"], 32, 32)).load(blob0)
    st = gs.random_positions(9, 64, 30)
    P, V, _ = net.evaluate_batch(st)
    Pr, Vr, _ = ref.evaluate_batch(st)
    assert (P == Pr).all() and (V == Vr).all()
    print("rank %d/%d OK: gathered %d samples (%.3f ms on device), broadcast %.3f ms" % (rank, world, len(allg), ms_gather, comm.last_ms), flush=True)
    for x in (mine, allg, empty, a2, net, ref, comm, ctx):
        x.close()


if __name__ == "__main__":
    main()
