"""TEST INFRASTRUCTURE ONLY (see oracle/az_oracle.h): numpy Float64 restatement of the reference's loss and optimiser
arithmetic (src/learning.jl:56-90, src/networks/flux.jl:68-95, src/schedule.jl:64-80,132-136).  Parity unpinned: the
reference's tests hold no golden vectors for these, and Optimisers.jl / Flux are un-vendored third-party code."""
import math

import numpy as np

EPS32 = float(np.finfo(np.float32).eps)


def forward_normalized(P, V, A):          # src/networks/network.jl:264-271
    p = P * A
    sp = p.sum(1, keepdims=True)
    return p / (sp + EPS32), V, 1.0 - sp[:, 0]


def losses(P_net, V_net, params_list, W, A, P, V, creg, cinv, rrn, Wmean, Hp):   # src/learning.jl:66-90
    Ph, Vh, pinv = forward_normalized(P_net, V_net, A)
    V = V / rrn
    Vh = Vh / rrn
    sw = W.sum()
    Lp = -(P * np.log(Ph + EPS32) * W[:, None]).sum() / sw - Hp
    Lv = ((Vh - V) * (Vh - V) * W).sum() / sw
    Lreg = creg * sum((w * w).sum() for w in params_list) if creg != 0 else 0.0
    Linv = cinv * (pinv * W).sum() / sw if cinv != 0 else 0.0
    L = (W.mean() / Wmean) * (Lp + Lv + Lreg + Linv)
    return L, Lp, Lv, Lreg, Linv


def entropy_wmean(P, W):
    return -(P * np.log(P + EPS32) * W[:, None]).sum() / W.sum()


def pl_schedule(xs, ys, i):               # src/schedule.jl:64-80
    pt = max([k for k, x in enumerate(xs) if x <= i], default=-1)
    if pt < 0:
        return ys[0]
    if pt == len(xs) - 1:
        return ys[-1]
    return ys[pt] + (ys[pt + 1] - ys[pt]) / (xs[pt + 1] - xs[pt]) * (i - xs[pt])


def cyclic_schedule(base, mid, term, n, xmid=0.45, xback=0.90):   # src/schedule.jl:132-136
    return [1, math.floor(xmid * n), math.floor(xback * n), n], [base, mid, base, term]


def nesterov_run(x0, grad_fn, n, lr_base, lr_high, lr_low, mom_low, mom_high):
    """train!(…, ::CyclicNesterov, …) on a parameter vector: returns the iterates."""
    lr = cyclic_schedule(lr_base, lr_high, lr_low, n)
    mo = cyclic_schedule(mom_high, mom_low, mom_high, n)
    eta, rho = lr_low, mom_high
    x, vel = np.array(x0, np.float64), np.zeros(len(x0))
    out = []
    for i in range(1, n + 1):
        dx = grad_fn(x)
        newdx = -rho * rho * vel + (1 + rho) * eta * dx
        vel = rho * vel - eta * dx
        x = x - newdx
        eta, rho = pl_schedule(*lr, i), pl_schedule(*mo, i)
        out.append(x.copy())
    return out


def adam_run(x0, grad_fn, n, eta, b1=0.9, b2=0.999, eps=1e-8):
    x, m, v = np.array(x0, np.float64), np.zeros(len(x0)), np.zeros(len(x0))
    out = []
    for t in range(1, n + 1):
        dx = grad_fn(x)
        m = b1 * m + (1 - b1) * dx
        v = b2 * v + (1 - b2) * dx * dx
        x = x - eta * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + eps)
        out.append(x.copy())
    return out
