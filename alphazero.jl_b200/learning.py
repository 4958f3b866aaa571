"""Learning step next to the self-play engine: the reference's loss, trainer and optimisers
(src/learning.jl:56-142, src/networks/flux.jl:68-95, src/networks/architectures/resnet.jl:53-92) on the B200.

Scope note: this file is PyTorch (library autograd + cuDNN), not hand-written CUDA -- the backward pass is outside the
hot path this repository rebuilds (self-play).  It exists so that the loop closes on the GPU: samples prepared on the
device (`Samples.convert`) -> `Trainer.batch_updates` -> `to_blob()` -> `Network.load` of the engine, without Flux.

Flux facts restated here (third-party, unpinned -- see DESIGN.md section 6): Conv is a true convolution over WHCN arrays
with weight W[kw,kh,cin,cout]; BatchNorm(momentum m, eps 1e-5) normalises with the biased batch variance in train mode
and tracks mu <- (1-m) mu + m mean, sigma2 <- (1-m) sigma2 + m * n/(n-1) * var; Dense is W x + b; Optimisers.Nesterov
and Optimisers.Adam update rules as documented in Optimisers.jl."""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS32 = float(np.finfo(np.float32).eps)


class ResNetTorch(nn.Module):
    """ResNet(gspec, ResNetHP) of src/networks/architectures/resnet.jl:65-92 with its parameters in Flux blob order
    (common: stem conv + BN, blocks of conv/BN/conv/BN; vhead: conv1x1, BN, Dense, Dense; phead: conv1x1, BN, Dense)."""

    def __init__(self, state_dim, num_actions, hp):
        super().__init__()
        W, H, C = state_dim
        self.dim, self.A = (W, H, C), num_actions
        self.hp = hp
        nf, nb = hp.num_filters, hp.num_blocks
        kw, kh = hp.conv_kernel_size
        mom = hp.batch_norm_momentum

        def conv(ci, co, kw_, kh_):
            return nn.Conv2d(ci, co, (kh_, kw_), padding=(kh_ // 2, kw_ // 2))

        def bn(n):
            return nn.BatchNorm2d(n, eps=1e-5, momentum=mom)

        self.stem = nn.ModuleList([conv(C, nf, kw, kh), bn(nf)])
        self.blocks = nn.ModuleList([nn.ModuleList([conv(nf, nf, kw, kh), bn(nf), conv(nf, nf, kw, kh), bn(nf)]) for _ in range(nb)])
        nvf, npf = hp.num_value_head_filters, hp.num_policy_head_filters
        self.vhead = nn.ModuleList([conv(nf, nvf, 1, 1), bn(nvf), nn.Linear(W * H * nvf, nf), nn.Linear(nf, 1)])
        self.phead = nn.ModuleList([conv(nf, npf, 1, 1), bn(npf), nn.Linear(W * H * npf, num_actions)])
        for m in self.modules():  # Flux defaults: glorot_uniform weights, zero biases
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def layers(self):
        out = list(self.stem)
        for b in self.blocks:
            out += list(b)
        return out + list(self.vhead) + list(self.phead)

    # ---- Flux parameter blob (the layout az_net_load expects; include/azb200.h) --------------------------------------
    def to_blob(self):
        parts = []
        for m in self.layers():
            if isinstance(m, nn.Conv2d):      # torch [co,ci,kh,kw] correlation -> Flux [kw,kh,ci,co] convolution (flipped)
                w = m.weight.detach().cpu().double().numpy()[:, :, ::-1, ::-1].transpose(3, 2, 1, 0)
                parts += [w.reshape(-1, order="F"), m.bias.detach().cpu().numpy()]
            elif isinstance(m, nn.BatchNorm2d):
                parts += [m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy(), m.running_mean.cpu().numpy(), m.running_var.cpu().numpy()]
            else:                              # Dense W[out,in] column-major
                parts += [m.weight.detach().cpu().numpy().reshape(-1, order="F"), m.bias.detach().cpu().numpy()]
        return np.concatenate([np.asarray(p, np.float64).ravel() for p in parts]).astype(np.float32)

    def to_blob_tensor(self):
        """The same blob as ONE contiguous float32 tensor on the model's own device (no host copy): Flux order, column-major
        flattening = torch's row-major flattening of the transposed array."""
        parts = []
        for m in self.layers():
            if isinstance(m, nn.Conv2d):      # torch [co,ci,kh,kw] -> flipped -> [kw,kh,ci,co] column-major == [co,ci,kh,kw] (flipped) row-major
                parts += [m.weight.detach().flip(2, 3).reshape(-1), m.bias.detach()]
            elif isinstance(m, nn.BatchNorm2d):
                parts += [m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var]
            else:                              # Dense W[out,in] column-major == W^T row-major
                parts += [m.weight.detach().t().reshape(-1), m.bias.detach()]
        return torch.cat([p.reshape(-1).float() for p in parts]).contiguous()

    def load_blob(self, blob):
        blob = np.asarray(blob, np.float32)
        want = sum(p.numel() for p in self.parameters()) + sum(2 * m.num_features for m in self.modules() if isinstance(m, nn.BatchNorm2d))
        if len(blob) != want:
            raise ValueError("blob has %d floats, the network has %d parameters" % (len(blob), want))
        q = 0

        def take(n):
            nonlocal q
            v = blob[q:q + n]
            q += n
            return v

        with torch.no_grad():
            for m in self.layers():
                if isinstance(m, nn.Conv2d):
                    co, ci, kh, kw = m.weight.shape
                    w = take(kw * kh * ci * co).reshape((kw, kh, ci, co), order="F").transpose(3, 2, 1, 0)[:, :, ::-1, ::-1]
                    m.weight.copy_(torch.from_numpy(w.copy()))
                    m.bias.copy_(torch.from_numpy(take(co).copy()))
                elif isinstance(m, nn.BatchNorm2d):
                    n = m.num_features
                    for t in (m.weight, m.bias, m.running_mean, m.running_var):
                        t.copy_(torch.from_numpy(take(n).copy()))
                else:
                    out, inn = m.weight.shape
                    m.weight.copy_(torch.from_numpy(take(out * inn).reshape((out, inn), order="F").copy()))
                    m.bias.copy_(torch.from_numpy(take(out).copy()))
        return self

    def forward(self, X):
        """X: [B, W*H*C] Float32 rows of vectorize_state (column-major WHC, as convert_samples lays them out).
        Returns (P [B, A] softmax over all actions, V [B]) = Network.forward (src/networks/flux.jl:127-132)."""
        W, H, C = self.dim
        x = X.reshape(-1, C, H, W)  # column-major (w,h,c) == row-major [c][h][w]

        def cb(x, conv, bn, relu=True):
            y = bn(conv(x))
            return F.relu(y) if relu else y

        x = cb(x, *self.stem)
        for c1, b1, c2, b2 in self.blocks:
            y = cb(cb(x, c1, b1), c2, b2, relu=False)
            x = F.relu(y + x)
        B = x.shape[0]
        v = cb(x, self.vhead[0], self.vhead[1]).reshape(B, -1)
        v = torch.tanh(self.vhead[3](F.relu(self.vhead[2](v))))[:, 0]
        p = cb(x, self.phead[0], self.phead[1]).reshape(B, -1)
        return torch.softmax(self.phead[2](p), dim=1), v


class SimpleNetTorch(nn.Module):
    """SimpleNet(gspec, SimpleNetHP) of src/networks/architectures/simplenet.jl:37-64: flatten -> Dense(width) [+ BatchNorm] + relu,
    `depth_common` hidden layers, value head (`depth_vhead` hidden layers, Dense(1), tanh), policy head (`depth_phead` hidden
    layers, Dense(num_actions), softmax); parameters in Flux blob order (common, vhead, phead)."""

    def __init__(self, state_dim, num_actions, hp):
        super().__init__()
        indim = int(np.prod(state_dim))
        self.hp, self.A = hp, num_actions
        w, bn = hp.width, bool(hp.use_batch_norm)

        def hidden(i, o):
            return [nn.Linear(i, o)] + ([nn.BatchNorm1d(o, eps=1e-5, momentum=hp.batch_norm_momentum)] if bn else [])

        common = hidden(indim, w)
        for _ in range(hp.depth_common):
            common += hidden(w, w)
        vhead = []
        for _ in range(hp.depth_vhead):
            vhead += hidden(w, w)
        vhead.append(nn.Linear(w, 1))
        phead = []
        for _ in range(hp.depth_phead):
            phead += hidden(w, w)
        phead.append(nn.Linear(w, num_actions))
        self.common, self.vhead, self.phead = nn.ModuleList(common), nn.ModuleList(vhead), nn.ModuleList(phead)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def layers(self):
        return list(self.common) + list(self.vhead) + list(self.phead)

    def to_blob(self):
        parts = []
        for m in self.layers():
            if isinstance(m, nn.BatchNorm1d):
                parts += [m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy(), m.running_mean.cpu().numpy(), m.running_var.cpu().numpy()]
            else:
                parts += [m.weight.detach().cpu().numpy().reshape(-1, order="F"), m.bias.detach().cpu().numpy()]
        return np.concatenate([np.asarray(p, np.float32).ravel() for p in parts])

    def load_blob(self, blob):
        blob = np.asarray(blob, np.float32)
        want = sum(p.numel() for p in self.parameters()) + sum(2 * m.num_features for m in self.modules() if isinstance(m, nn.BatchNorm1d))
        if len(blob) != want:
            raise ValueError("blob has %d floats, the network has %d parameters" % (len(blob), want))
        q = 0
        with torch.no_grad():
            for m in self.layers():
                ts = (m.weight, m.bias, m.running_mean, m.running_var) if isinstance(m, nn.BatchNorm1d) else None
                if ts is not None:
                    for t in ts:
                        t.copy_(torch.from_numpy(blob[q:q + m.num_features].copy()))
                        q += m.num_features
                else:
                    out, inn = m.weight.shape
                    m.weight.copy_(torch.from_numpy(blob[q:q + out * inn].reshape((out, inn), order="F").copy()))
                    q += out * inn
                    m.bias.copy_(torch.from_numpy(blob[q:q + out].copy()))
                    q += out
        return self

    @staticmethod
    def _run(mods, x, last_plain):
        n = len(mods)
        i = 0
        while i < n:
            m = mods[i]
            x = m(x)
            i += 1
            if i < n and isinstance(mods[i], nn.BatchNorm1d):
                x = mods[i](x)
                i += 1
            if not (last_plain and i == n):
                x = F.relu(x)
        return x

    def forward(self, X):
        """X: [B, W*H*C] rows of vectorize_state (Flux.flatten is column-major over W,H,C = the row layout of convert_samples)."""
        c = self._run(list(self.common), X, last_plain=False)
        v = torch.tanh(self._run(list(self.vhead), c, last_plain=True))[:, 0]
        p = torch.softmax(self._run(list(self.phead), c, last_plain=True), dim=1)
        return p, v


def forward_normalized(net, X, A):
    """Network.forward_normalized (src/networks/network.jl:264-271): (P masked + renormalised, V, p_invalid)."""
    P, V = net(X)
    P = P * A
    sp = P.sum(1, keepdim=True)
    return P / (sp + EPS32), V, 1.0 - sp[:, 0]


def klloss_wmean(Ph, P, W):   # src/learning.jl:60
    return -(P * torch.log(Ph + EPS32) * W[:, None]).sum() / W.sum()


def entropy_wmean(P, W):      # src/learning.jl:62
    return -(P * torch.log(P + EPS32) * W[:, None]).sum() / W.sum()


def mse_wmean(Yh, Y, W):      # src/learning.jl:58
    return ((Yh - Y) * (Yh - Y) * W).sum() / W.sum()


class LearningParams:  # src/params.jl:196-215
    def __init__(self, optimiser, l2_regularization, nonvalidity_penalty=1.0, batch_size=1024, loss_computation_batch_size=1024,
                 min_checkpoints_per_epoch=1, max_batches_per_checkpoint=1000, num_checkpoints=1, use_gpu=True, use_position_averaging=True,
                 samples_weighing_policy=1, rewards_renormalization=1.0):
        self.optimiser, self.l2_regularization, self.nonvalidity_penalty = optimiser, l2_regularization, nonvalidity_penalty
        self.batch_size, self.loss_computation_batch_size = batch_size, loss_computation_batch_size
        self.min_checkpoints_per_epoch, self.max_batches_per_checkpoint = min_checkpoints_per_epoch, max_batches_per_checkpoint
        self.num_checkpoints, self.use_gpu, self.use_position_averaging = num_checkpoints, use_gpu, use_position_averaging
        self.samples_weighing_policy, self.rewards_renormalization = samples_weighing_policy, rewards_renormalization


class Adam:            # src/params.jl:117-119
    def __init__(self, lr):
        self.lr = lr


class CyclicNesterov:  # src/params.jl:130-136
    def __init__(self, lr_base, lr_high, lr_low, momentum_low, momentum_high):
        self.lr_base, self.lr_high, self.lr_low = lr_base, lr_high, lr_low
        self.momentum_low, self.momentum_high = momentum_low, momentum_high


def pl_schedule(xs, ys, i):   # src/schedule.jl:64-80
    pt = -1
    for k, x in enumerate(xs):
        if x <= i:
            pt = k
    if pt < 0:
        return ys[0]
    if pt == len(xs) - 1:
        return ys[-1]
    return ys[pt] + (ys[pt + 1] - ys[pt]) / (xs[pt + 1] - xs[pt]) * (i - xs[pt])


def cyclic_schedule(base, mid, term, n, xmid=0.45, xback=0.90):   # src/schedule.jl:132-136
    return [1, math.floor(xmid * n), math.floor(xback * n), n], [base, mid, base, term]


def losses(net, params, Wmean, Hp, batch):
    """src/learning.jl:66-90: (L, Lp, Lv, Lreg, Linv).  The L2 penalty runs over ALL trainable parameters, as the
    reference does (its comment at :67-73)."""
    W, X, A, P, V = batch
    creg, cinv = params.l2_regularization, params.nonvalidity_penalty
    Ph, Vh, p_invalid = forward_normalized(net, X, A)
    V = V / params.rewards_renormalization
    Vh = Vh / params.rewards_renormalization
    Lp = klloss_wmean(Ph, P, W) - Hp
    Lv = mse_wmean(Vh, V, W)
    Lreg = creg * sum((w * w).sum() for w in net.parameters()) if creg != 0 else torch.zeros_like(Lv)
    Linv = cinv * (p_invalid * W).sum() / W.sum() if cinv != 0 else torch.zeros_like(Lv)
    L = (W.mean() / Wmean) * (Lp + Lv + Lreg + Linv)
    return L, Lp, Lv, Lreg, Linv


def nesterov_update_(params, grads, vel, eta, rho):
    """Optimisers.Nesterov: newdx = -rho^2 vel + (1+rho) eta dx; vel <- rho vel - eta dx; x <- x - newdx."""
    with torch.no_grad():
        for p, g, v in zip(params, grads, vel):
            newdx = -(rho * rho) * v + (1 + rho) * eta * g
            v.mul_(rho).sub_(eta * g)
            p.sub_(newdx)


def adam_update_(params, grads, state, eta, beta=(0.9, 0.999), eps=1e-8):
    """Optimisers.Adam: mt <- b1 mt + (1-b1) dx; vt <- b2 vt + (1-b2) dx^2; x <- x - eta * mt/(1-b1^t) / (sqrt(vt/(1-b2^t)) + eps)."""
    state["t"] += 1
    t = state["t"]
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, state["m"], state["v"]):
            m.mul_(beta[0]).add_((1 - beta[0]) * g)
            v.mul_(beta[1]).add_((1 - beta[1]) * g * g)
            p.sub_(eta * (m / (1 - beta[0] ** t)) / (torch.sqrt(v / (1 - beta[1] ** t)) + eps))


def train(net, opt, loss_fn, batches, n, callback=None):
    """Network.train! (src/networks/flux.jl:68-95).  CyclicNesterov: step 1 runs with (lr_low, momentum_high), step i+1
    with (lr[i], momentum[i]) -- the reference adjusts the optimiser AFTER each update.  Adam: constant lr (the
    reference's Adam method indexes an undefined `lr` schedule; the shipped intent is a constant rate)."""
    ps = [p for p in net.parameters()]
    if isinstance(opt, CyclicNesterov):
        lr = cyclic_schedule(opt.lr_base, opt.lr_high, opt.lr_low, n)
        mo = cyclic_schedule(opt.momentum_high, opt.momentum_low, opt.momentum_high, n)
        eta, rho = opt.lr_low, opt.momentum_high
        vel = [torch.zeros_like(p) for p in ps]
    else:
        state = dict(t=0, m=[torch.zeros_like(p) for p in ps], v=[torch.zeros_like(p) for p in ps])
    out = []
    for i, d in enumerate(batches, start=1):
        if i > n:
            break
        loss = loss_fn(net, d)
        grads = torch.autograd.grad(loss, ps)
        if isinstance(opt, CyclicNesterov):
            nesterov_update_(ps, grads, vel, eta, rho)
            eta, rho = pl_schedule(*lr, i), pl_schedule(*mo, i)
        else:
            adam_update_(ps, grads, state, opt.lr)
        out.append(float(loss.detach()))
        if callback:
            callback(i, out[-1])
    return out


class Trainer:
    """Trainer (src/learning.jl:96-142).  `data` = dict of Float32 arrays W, X, A, P, V with the sample index first (what
    `Samples.convert` returns; merge_by_state / use_position_averaging happens there, on the device)."""

    def __init__(self, net, data, params, test_mode=False, device=None, seed=0):
        self.device = torch.device(device if device is not None else ("cuda" if params.use_gpu and torch.cuda.is_available() else "cpu"))
        self.net = net.to(self.device)
        self.net.train(not test_mode)
        self.params = params
        self.data = {k: torch.as_tensor(np.asarray(v, np.float32)).to(self.device) for k, v in data.items()}
        W, P = self.data["W"], self.data["P"]
        self.Wmean = float(W.mean())
        self.Hp = float(entropy_wmean(P, W))
        self.gen = torch.Generator(device="cpu").manual_seed(seed)

    def num_samples(self):
        return int(self.data["W"].shape[0])

    def num_batches_total(self):
        return self.num_samples() // self.params.batch_size

    def _batch(self, idx):
        return tuple(self.data[k][idx] for k in ("W", "X", "A", "P", "V"))

    def batches_stream(self):
        """Flux.DataLoader(data; batchsize, partial=false, shuffle=true) |> cycle (src/learning.jl:115-119)."""
        bs = min(self.params.batch_size, self.num_samples())
        while True:
            perm = torch.randperm(self.num_samples(), generator=self.gen).to(self.device)
            for k in range(0, self.num_samples() - bs + 1, bs):
                yield self._batch(perm[k:k + bs])

    def batch_updates(self, n):   # src/learning.jl:130-139
        if not hasattr(self, "_stream"):
            self._stream = self.batches_stream()
        return train(self.net, self.params.optimiser, lambda net, b: losses(net, self.params, self.Wmean, self.Hp, b)[0], self._stream, n)

    def learning_status(self):    # src/learning.jl:157-181: loss terms + policy entropies, weighted over evaluation batches
        bs = min(self.params.loss_computation_batch_size, self.num_samples())
        was = self.net.training
        acc, ws = np.zeros(6), 0.0
        with torch.no_grad():
            for k in range(0, self.num_samples(), bs):
                b = self._batch(slice(k, k + bs))
                Ls = losses(self.net, self.params, self.Wmean, self.Hp, b)
                Pnet, _, _ = forward_normalized(self.net, b[1], b[2])
                w = float(b[0].sum())
                acc += w * np.array([float(x) for x in Ls] + [float(entropy_wmean(Pnet, b[0]))])
                ws += w
        self.net.train(was)
        L, Lp, Lv, Lreg, Linv, Hpnet = acc / ws
        return dict(L=L, Lp=Lp, Lv=Lv, Lreg=Lreg, Linv=Linv, Hp=self.Hp, Hpnet=Hpnet)

    def get_trained_network_blob(self):   # get_trained_network (src/learning.jl:126-128) -> the engine's weight blob
        return self.net.to_blob()

    def hand_off(self, engine_net):
        """get_trained_network without leaving the GPU: the flat parameter tensor of the trained model goes straight into the
        engine's network (az_net_load_device: BatchNorm folded and fp16 layouts written by device kernels).  Falls back to the
        host blob when the model does not live on the engine's GPU."""
        if hasattr(self.net, "to_blob_tensor") and next(self.net.parameters()).is_cuda:
            t = self.net.to_blob_tensor()
            torch.cuda.synchronize(t.device)
            engine_net.load_device(t.data_ptr(), t.numel())
            return t.numel()
        blob = self.net.to_blob()
        engine_net.load(blob)
        return len(blob)
