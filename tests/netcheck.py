"""Shared helper: CUDA ResNet forward (through the C ABI) against the fp32 torch restatement (oracle/netref.py).
Tolerance 1e-3 on the policy logits, the pre-tanh value, P and V (BASELINE.json north_star: "value/policy logits within 1e-3")."""
import numpy as np

TOL = 1e-3


def c4_hp(num_blocks):
    return dict(num_blocks=num_blocks, num_filters=128, conv_kernel_size=(3, 3), num_policy_head_filters=32,
                num_value_head_filters=32)


def make_net(az, ctx, gs, hp, seed=1, randomize=True):
    from oracle import netref
    blob = netref.make_blob(gs.state_dim, gs.num_actions, hp, seed=seed, randomize=randomize)
    net = az.ResNet(ctx, gs, az.ResNetHP(hp["num_blocks"], hp["num_filters"], hp["conv_kernel_size"],
                                         hp["num_policy_head_filters"], hp["num_value_head_filters"]))
    assert net.num_params == len(blob) == netref.num_params(gs.state_dim, gs.num_actions, hp)
    net.load(blob)
    return net, blob


def compare(az, oz, gs, net, blob, hp, states):
    from oracle import netref
    gid = oz.game_id(gs.name)
    X = np.stack([oz.vectorize_state(gid, bytes(s)) for s in states])
    mask = np.stack([oz.GameEnv(gid, bytes(s)).actions_mask() for s in states])
    P0, V0 = netref.forward(blob, gs.state_dim, gs.num_actions, hp, X)
    Pr, Vr, Ir = netref.forward_normalized(P0, V0, mask)
    P, V, Pinv = net.evaluate_batch(states)
    # the inputs of the output non-linearities ("logits" of BASELINE.json north_star): policy logits before the softmax,
    # value before the tanh
    Lr, Vpr = netref.forward(blob, gs.state_dim, gs.num_actions, hp, X, logits=True)
    L, Vp = net.forward_logits(states)
    return dict(dP=float(np.abs(P - Pr).max()), dV=float(np.abs(V - Vr).max()), dI=float(np.abs(Pinv - Ir).max()),
                dL=float(np.abs(L - Lr).max()), dVpre=float(np.abs(Vp - Vpr).max()),
                P=P, V=V, Pr=Pr, Vr=Vr, mask=mask, L=L, Lr=Lr, Vpre=Vp, Vprer=Vpr)


def smoke(az, ctx, gs):
    from oracle import oracle as oz
    hp = c4_hp(2)
    net, blob = make_net(az, ctx, gs, hp)
    states = gs.random_positions(7, 96, 30)
    r = compare(az, oz, gs, net, blob, hp, states)
    assert r["dP"] < TOL and r["dV"] < TOL and r["dL"] < TOL and r["dVpre"] < TOL, {k: r[k] for k in ("dP", "dV", "dL", "dVpre")}
    net.close()
    print("smoke OK: ResNet forward within %.1e of the fp32 reference (dP=%.2e dV=%.2e)" % (TOL, r["dP"], r["dV"]))
