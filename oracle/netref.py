"""fp32 PyTorch-CPU restatement of the reference's two-head networks with Flux semantics.

TEST INFRASTRUCTURE ONLY (see az_oracle.h).  "parity unpinned": Flux/NNlib/cuDNN are third-party
dependencies that are not vendored in the reference (Project.toml:6-31) and ship no golden vectors;
this file restates their documented semantics at the reference's call sites:

  ResNet       src/networks/architectures/resnet.jl:53-92
  forward      src/networks/flux.jl:127-132            (common -> vhead, phead)
  forward_normalized / evaluate_batch   src/networks/network.jl:264-271,308-315

Flux facts used: Conv is a TRUE convolution (kernel flipped w.r.t. torch.conv2d) over WHCN arrays;
BatchNorm in test mode is (x - mu) / sqrt(sigma2 + 1e-5) * gamma + beta; Dense is W*x + b with W[out,in];
flatten is column-major over (W,H,C); softmax over the action dimension.

Parameter blob (float32, the layout az_net_load expects), in layer order common, vhead, phead:
  Conv: W[kw,kh,cin,cout] column-major (kw fastest), then b[cout];  BatchNorm: gamma, beta, mu, sigma2;
  Dense: W[out,in] column-major (out fastest), then b[out].
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS32 = float(np.finfo(np.float32).eps)


def resnet_layers(dim, num_actions, hp):
    """[(kind, shape...)] in blob order. dim = (W, H, C)."""
    W, H, C = dim
    nf, nb = hp["num_filters"], hp["num_blocks"]
    k = hp.get("conv_kernel_size", (3, 3))
    npf, nvf = hp["num_policy_head_filters"], hp["num_value_head_filters"]
    L = [("conv", k[0], k[1], C, nf), ("bn", nf)]
    for _ in range(nb):
        L += [("conv", k[0], k[1], nf, nf), ("bn", nf), ("conv", k[0], k[1], nf, nf), ("bn", nf)]
    L += [("conv", 1, 1, nf, nvf), ("bn", nvf), ("dense", nf, W * H * nvf), ("dense", 1, nf)]
    L += [("conv", 1, 1, nf, npf), ("bn", npf), ("dense", num_actions, W * H * npf)]
    return L


def layer_size(l):
    if l[0] == "conv":
        return l[1] * l[2] * l[3] * l[4] + l[4]
    if l[0] == "bn":
        return 4 * l[1]
    return l[1] * l[2] + l[1]


def num_params(dim, num_actions, hp):
    return sum(layer_size(l) for l in resnet_layers(dim, num_actions, hp))


def make_blob(dim, num_actions, hp, seed=1, randomize=True):
    """Glorot-uniform weights (Flux default). randomize=False: zero biases, BN gamma=1 beta=0 mu=0 sigma2=1
    (a freshly constructed Flux model); randomize=True also draws biases / BN statistics so that folding is exercised."""
    rng = np.random.default_rng(seed)
    parts = []
    for l in resnet_layers(dim, num_actions, hp):
        if l[0] == "conv":
            kw, kh, ci, co = l[1:]
            s = np.sqrt(6.0 / (kw * kh * ci + kw * kh * co))
            parts.append(rng.uniform(-s, s, kw * kh * ci * co))
            parts.append(rng.normal(0, 0.05, co) if randomize else np.zeros(co))
        elif l[0] == "bn":
            n = l[1]
            if randomize:
                parts += [rng.uniform(0.7, 1.3, n), rng.normal(0, 0.1, n), rng.normal(0, 0.1, n), rng.uniform(0.6, 1.5, n)]
            else:
                parts += [np.ones(n), np.zeros(n), np.zeros(n), np.ones(n)]
        else:
            out, inn = l[1:]
            s = np.sqrt(6.0 / (inn + out))
            parts.append(rng.uniform(-s, s, out * inn))
            parts.append(rng.normal(0, 0.05, out) if randomize else np.zeros(out))
    return np.concatenate(parts).astype(np.float32)


def forward(blob, dim, num_actions, hp, X, dtype=torch.float32, logits=False):
    """X: [B, W, H, C] (vectorize_state per sample, Flux WHC order).  Returns (P [B,A] softmax, V [B]); with logits=True
    the inputs of the two output non-linearities instead (policy logits [B,A], pre-tanh value [B])."""
    Wd, Hd, C = dim
    q = [0]
    blob = np.asarray(blob, np.float32)

    def take(n):
        v = blob[q[0]:q[0] + n]
        q[0] += n
        return v

    def conv(x, l):
        kw, kh, ci, co = l[1:]
        w = take(kw * kh * ci * co).reshape((kw, kh, ci, co), order="F")
        b = take(co)
        wt = torch.tensor(w.transpose(3, 2, 1, 0)[:, :, ::-1, ::-1].copy(), dtype=dtype)  # [co,ci,kh,kw] flipped
        return F.conv2d(x, wt, torch.tensor(b, dtype=dtype), padding=(kh // 2, kw // 2))

    def bn(x, l, relu):
        n = l[1]
        g, be, mu, var = (torch.tensor(take(n), dtype=dtype).view(1, n, 1, 1) for _ in range(4))
        y = (x - mu) / torch.sqrt(var + 1e-5) * g + be
        return torch.relu(y) if relu else y

    def dense(x, l):
        out, inn = l[1:]
        w = take(out * inn).reshape((out, inn), order="F")
        b = take(out)
        return x @ torch.tensor(w, dtype=dtype).T + torch.tensor(b, dtype=dtype)

    L = resnet_layers(dim, num_actions, hp)
    it = iter(L)
    x = torch.tensor(np.asarray(X, np.float32), dtype=dtype).permute(0, 3, 2, 1)  # [B,C,H,W]
    x = bn(conv(x, next(it)), next(it), True)
    for _ in range(hp["num_blocks"]):
        y = bn(conv(x, next(it)), next(it), True)
        y = bn(conv(y, next(it)), next(it), False)
        x = torch.relu(y + x)
    B = x.shape[0]
    v = bn(conv(x, next(it)), next(it), True).reshape(B, -1)  # (c,h,w) row-major == Flux flatten (w,h,c) column-major
    v = torch.relu(dense(v, next(it)))
    vpre = dense(v, next(it))[:, 0]
    v = torch.tanh(vpre)
    p = bn(conv(x, next(it)), next(it), True).reshape(B, -1)
    plog = dense(p, next(it))
    p = torch.softmax(plog, dim=1)
    assert q[0] == len(blob)
    if logits:
        return plog.numpy().astype(np.float32), vpre.numpy().astype(np.float32)
    return p.numpy().astype(np.float32), v.numpy().astype(np.float32)


def forward_normalized(P, V, mask):
    """src/networks/network.jl:264-271 in float32."""
    P = P.astype(np.float32) * mask.astype(np.float32)
    sp = P.sum(1, dtype=np.float32)
    Pn = P / (sp[:, None] + np.float32(EPS32))
    return Pn.astype(np.float32), V, (np.float32(1) - sp)


# ---- SimpleNet (src/networks/architectures/simplenet.jl:37-64) -------------------------------------------------
def simplenet_layers(dim, num_actions, hp):
    """[(kind, ...)] in blob order: common, vhead, phead.  make_dense = Dense [+ BatchNorm(relu)] | Dense(relu)."""
    indim = dim[0] * dim[1] * dim[2]
    w, bn = hp["width"], hp.get("use_batch_norm", False)

    def hidden(i, o):
        return [("dense", o, i, not bn)] + ([("bn", o)] if bn else [])
    L = hidden(indim, w)
    for _ in range(hp["depth_common"]):
        L += hidden(w, w)
    for _ in range(hp.get("depth_vhead", 1)):
        L += hidden(w, w)
    L += [("dense", 1, w, False)]
    for _ in range(hp.get("depth_phead", 1)):
        L += hidden(w, w)
    L += [("dense", num_actions, w, False)]
    return L


def simplenet_make_blob(dim, num_actions, hp, seed=1, randomize=True):
    rng = np.random.default_rng(seed)
    parts = []
    for l in simplenet_layers(dim, num_actions, hp):
        if l[0] == "dense":
            out, inn = l[1], l[2]
            s = np.sqrt(6.0 / (inn + out))
            parts.append(rng.uniform(-s, s, out * inn))
            parts.append(rng.normal(0, 0.05, out) if randomize else np.zeros(out))
        else:
            n = l[1]
            if randomize:
                parts += [rng.uniform(0.7, 1.3, n), rng.normal(0, 0.1, n), rng.normal(0, 0.1, n), rng.uniform(0.6, 1.5, n)]
            else:
                parts += [np.ones(n), np.zeros(n), np.zeros(n), np.ones(n)]
    return np.concatenate(parts).astype(np.float32)


def simplenet_forward(blob, dim, num_actions, hp, X, logits=False):
    """X: [B, W, H, C]; flatten is column-major over (W,H,C) (Flux.flatten). Returns (P softmax, V), or with logits=True
    (policy logits, pre-tanh value)."""
    blob = np.asarray(blob, np.float32)
    q = [0]

    def take(n):
        v = blob[q[0]:q[0] + n]
        q[0] += n
        return v
    B = X.shape[0]
    x = torch.tensor(np.asarray(X, np.float32).reshape(B, -1, order="F") if False else
                     np.stack([np.asarray(X[i], np.float32).reshape(-1, order="F") for i in range(B)]))
    bn_on = hp.get("use_batch_norm", False)

    def dense(x, out, inn, relu):
        w = take(out * inn).reshape((out, inn), order="F")
        b = take(out)
        y = x @ torch.tensor(w).T + torch.tensor(b)
        return torch.relu(y) if relu else y

    def hidden(x, inn, out):
        y = dense(x, out, inn, not bn_on)
        if bn_on:
            g, be, mu, var = (torch.tensor(take(out)) for _ in range(4))
            y = torch.relu((y - mu) / torch.sqrt(var + 1e-5) * g + be)
        return y
    w = hp["width"]
    x = hidden(x, x.shape[1], w)
    for _ in range(hp["depth_common"]):
        x = hidden(x, w, w)
    v = x
    for _ in range(hp.get("depth_vhead", 1)):
        v = hidden(v, w, w)
    vpre = dense(v, 1, w, False)[:, 0]
    v = torch.tanh(vpre)
    p = x
    for _ in range(hp.get("depth_phead", 1)):
        p = hidden(p, w, w)
    plog = dense(p, num_actions, w, False)
    p = torch.softmax(plog, dim=1)
    assert q[0] == len(blob)
    if logits:
        return plog.numpy().astype(np.float32), vpre.numpy().astype(np.float32)
    return p.numpy().astype(np.float32), v.numpy().astype(np.float32)
