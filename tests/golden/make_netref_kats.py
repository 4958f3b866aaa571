"""Generates tests/golden/netref_kats.json: outputs of the fp32 torch restatement of the Flux networks (oracle/netref.py) for
fixed seeds and positions.  Like oracle_kats.json these freeze the restatement (not reference-pinned: Flux cannot run here).
Run from the repo root:  python tests/golden/make_netref_kats.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build(oz, netref):
    gid = oz.game_id("connect-four")
    states = oz.random_positions(gid, 77, 6, 20)
    X = np.stack([oz.vectorize_state(gid, s) for s in states])          # [B, W, H, C]
    mask = np.stack([oz.GameEnv(gid, s).actions_mask() for s in states]).astype(np.float32)
    cases = []
    for nb, seed, rnd in [(0, 1, True), (2, 2, True), (5, 3, False)]:
        hp = dict(num_blocks=nb, num_filters=32, conv_kernel_size=(3, 3), num_policy_head_filters=8, num_value_head_filters=8)
        blob = netref.make_blob((7, 6, 3), 7, hp, seed=seed, randomize=rnd)
        P, V = netref.forward(blob, (7, 6, 3), 7, hp, X)
        Pn, Vn, Pinv = netref.forward_normalized(P, V, mask)
        cases.append(dict(net="resnet", hp=dict(hp, conv_kernel_size=[3, 3]), seed=seed, randomize=rnd, num_params=int(len(blob)),
                          blob_sum=float(np.float64(blob.astype(np.float64).sum())), P=[[float(x) for x in r] for r in Pn],
                          V=[float(x) for x in Vn], Pinvalid=[float(x) for x in Pinv]))
    ttt = oz.game_id("tictactoe")
    st = oz.random_positions(ttt, 5, 4, 4)
    Xt = np.stack([oz.vectorize_state(ttt, s) for s in st])
    hp = dict(width=32, depth_common=3, use_batch_norm=True)
    blob = netref.simplenet_make_blob((3, 3, 3), 9, hp, seed=4)
    P, V = netref.simplenet_forward(blob, (3, 3, 3), 9, hp, Xt)
    cases.append(dict(net="simplenet", hp=hp, seed=4, num_params=int(len(blob)), blob_sum=float(np.float64(blob.astype(np.float64).sum())),
                      P=[[float(x) for x in r] for r in P], V=[float(x) for x in V]))
    return dict(format=1, states=[bytes(s).hex() for s in states], ttt_states=[bytes(s).hex() for s in st], cases=cases)


if __name__ == "__main__":
    from oracle import netref
    from oracle import oracle as oz
    oz.lib()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "netref_kats.json")
    with open(path, "w") as f:
        json.dump(build(oz, netref), f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
