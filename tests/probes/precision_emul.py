"""CPU emulation of the rounding points of the tcgen05 network path (which ones dominate the logit error).
A design probe kept with the test infrastructure because it leans on the oracle's network restatement (DESIGN.md "precision"); not collected by pytest."""
import sys, os
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _pkg
from oracle import netref, oracle as oz

def h16(x):
    return x.half().float()

def hilo(x):  # two-term fp16 split
    hi = x.half().float()
    return hi + (x - hi).half().float()

def run(blob, dim, A, hp, X, cfg):
    """cfg keys: w_stem, w_tower, w_head, w_dense (weight rounding fns), a_tower1, a_tower2 (conv input rounding),
    skip (rounding of the block output kept for the skip path), a_head, a_dense"""
    Wd, Hd, C = dim
    blob = np.asarray(blob, np.float32); q = [0]
    def take(n):
        v = blob[q[0]:q[0] + n]; q[0] += n; return v
    def convbn(x, l, lb, wr, relu=True, add=None):
        kw, kh, ci, co = l[1:]
        w = take(kw*kh*ci*co).reshape((kw, kh, ci, co), order="F"); b = take(co)
        g, be, mu, var = (take(co) for _ in range(4))
        sc = (g / np.sqrt(var + np.float32(1e-5))).astype(np.float32)
        wt = torch.tensor(w.transpose(3, 2, 1, 0)[:, :, ::-1, ::-1].copy()) * torch.tensor(sc).view(-1, 1, 1, 1)
        sh = torch.tensor((b - mu) * sc + be)
        y = F.conv2d(x.double(), wr(wt).double(), None, padding=(kh//2, kw//2)).float() + sh.view(1, -1, 1, 1)
        if add is not None: y = y + add
        return torch.relu(y) if relu else y
    def dense(x, l, wr):
        out, inn = l[1:]
        w = take(out*inn).reshape((out, inn), order="F"); b = take(out)
        return (x.double() @ wr(torch.tensor(w)).double().T).float() + torch.tensor(b)
    L = netref.resnet_layers(dim, A, hp); it = iter(L)
    x = torch.tensor(np.asarray(X, np.float32)).permute(0, 3, 2, 1)
    l = next(it); lb = next(it)
    x = cfg["skip0"](convbn(x, l, lb, cfg["w_stem"]))
    for _ in range(hp["num_blocks"]):
        l = next(it); lb = next(it)
        y = convbn(cfg["a_tower1"](x), l, lb, cfg["w_tower"])
        l = next(it); lb = next(it)
        x = cfg["skip"](convbn(cfg["a_tower2"](y), l, lb, cfg["w_tower"], relu=True, add=x))
    B = x.shape[0]
    l = next(it); lb = next(it)
    v = cfg["a_dense"](convbn(cfg["a_head"](x), l, lb, cfg["w_head"])).reshape(B, -1)
    v = torch.relu(dense(v, next(it), cfg["w_dense"]))
    vpre = dense(v, next(it), lambda w: w)[:, 0]
    l = next(it); lb = next(it)
    p = cfg["a_dense"](convbn(cfg["a_head"](x), l, lb, cfg["w_head"])).reshape(B, -1)
    plog = dense(p, next(it), cfg["w_dense"])
    return plog.numpy(), vpre.numpy()

ident = lambda t: t
BASE = dict(w_stem=h16, w_tower=h16, w_head=h16, w_dense=h16, a_tower1=h16, a_tower2=h16, skip=hilo, skip0=h16, a_head=h16, a_dense=h16)
def main():
    az = _pkg.load()
    gs = az.GameSpec("connect-four")
    gid = oz.game_id("connect-four")
    cases = [(5, 6, 300), (7, 3, 300), (7, 8, 300), (7, 5, 300), (7, 2, 300)]
    variants = {
        "current": {},
        "exact": {k: ident for k in BASE},
        "stem hi/lo + skip0 hilo": dict(w_stem=hilo, skip0=hilo),
        "heads exact": dict(w_head=ident, w_dense=ident, a_head=ident, a_dense=ident),
        "heads a_head hilo only": dict(a_head=hilo),
        "tower weights exact": dict(w_tower=ident),
        "tower acts exact": dict(a_tower1=ident, a_tower2=ident),
        "a_tower1 hilo (conv1 input 2 MMAs)": dict(a_tower1=hilo),
        "a_tower2 hilo": dict(a_tower2=hilo),
        "stem+heads exact": dict(w_stem=ident, skip0=ident, w_head=ident, w_dense=ident, a_head=ident, a_dense=ident),
    }
    for blocks, seed, n in cases:
        hp = dict(num_blocks=blocks, num_filters=128, conv_kernel_size=(3, 3), num_policy_head_filters=32, num_value_head_filters=32)
        blob = netref.make_blob(gs.state_dim, gs.num_actions, hp, seed=seed, randomize=True)
        states = gs.random_positions(11, n, 38)
        X = np.stack([oz.vectorize_state(gid, bytes(s)) for s in states])
        Lr, Vr = netref.forward(blob, gs.state_dim, gs.num_actions, hp, X, logits=True)
        print("blocks %d seed %d: |L|max %.2f |Vpre|max %.2f" % (blocks, seed, np.abs(Lr).max(), np.abs(Vr).max()))
        for name, d in variants.items():
            cfg = dict(BASE); cfg.update(d)
            L, V = run(blob, gs.state_dim, gs.num_actions, hp, X, cfg)
            print("   %-36s dL %.2e  dVpre %.2e   rmsL %.2e" % (name, np.abs(L - Lr).max(), np.abs(V - Vr).max(), np.sqrt(np.mean((L - Lr)**2))))
main()
