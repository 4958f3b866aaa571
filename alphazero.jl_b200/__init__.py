"""Host-side mirror of the AlphaZero.jl self-play interface on top of libazb200.so (C ABI, include/azb200.h).

Names follow the reference so that tests read like the reference's own code:
  GameSpec            <- GI.AbstractGameSpec            (src/game.jl)
  MctsParams/SimParams/SelfPlayParams, PLSchedule/ConstSchedule  (src/params.jl, src/schedule.jl)
  RandomOracle / ResNet(gspec, ResNetHP)                 (src/mcts.jl:62-72, src/networks/architectures/resnet.jl)
  MctsEnv.explore / policy / reset                       (src/mcts.jl:239-281)
  simulate(simulator-less)                               (src/simulations.jl:207-244)

There is NO CPU fallback: every compute call goes through the CUDA library and raises if it cannot run.
The directory name contains a dot, so load it with `_pkg.load()` (repo root) rather than `import`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libazb200.so")
MAX_SCHEDULE = 8
AZ_OK = 0
STATUS = {0: "AZ_OK", 1: "AZ_EINVAL", 2: "AZ_ECUDA", 3: "AZ_ENOMEM", 4: "AZ_ESTATE", 5: "AZ_EUNSUPPORTED"}
NET_UNIFORM, NET_SYNTH, NET_RESNET, NET_SIMPLENET = 0, 1, 2, 3


class AzError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS.get(status, status), msg))
        self.status = status


class _MctsParams(C.Structure):
    _fields_ = [("gamma", C.c_double), ("cpuct", C.c_double), ("num_iters_per_turn", C.c_int32), ("temperature_n", C.c_int32),
                ("dirichlet_noise_eps", C.c_double), ("dirichlet_noise_alpha", C.c_double), ("prior_temperature", C.c_double),
                ("temperature_xs", C.c_int32 * MAX_SCHEDULE), ("temperature_ys", C.c_double * MAX_SCHEDULE)]


class _MinMaxParams(C.Structure):
    _fields_ = [("depth", C.c_int32), ("amplify_rewards", C.c_int32), ("tau", C.c_double), ("gamma", C.c_double)]


class _SimParams(C.Structure):
    _fields_ = [("num_games", C.c_int32), ("num_workers", C.c_int32), ("batch_size", C.c_int32), ("fill_batches", C.c_int32),
                ("reset_every", C.c_int32), ("alternate_colors", C.c_int32), ("flip_probability", C.c_double)]


class _ResNetHP(C.Structure):
    _fields_ = [("num_blocks", C.c_int32), ("num_filters", C.c_int32), ("conv_kernel_size", C.c_int32 * 2),
                ("num_policy_head_filters", C.c_int32), ("num_value_head_filters", C.c_int32), ("batch_norm_momentum", C.c_float)]


class _SimpleNetHP(C.Structure):
    _fields_ = [("width", C.c_int32), ("depth_common", C.c_int32), ("depth_phead", C.c_int32), ("depth_vhead", C.c_int32),
                ("use_batch_norm", C.c_int32), ("batch_norm_momentum", C.c_float)]


# every symbol declared in include/azb200.h (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "az_version", "az_ctx_create", "az_ctx_destroy", "az_last_error", "az_ctx_synchronize", "az_ctx_num_launches",
    "az_game_lookup", "az_game_num_actions", "az_game_state_bytes", "az_game_state_dim", "az_game_max_plies",
    "az_game_vectorize_state", "az_game_actions_mask", "az_game_play", "az_game_init_state", "az_game_random_positions",
    "az_game_heuristic_value", "az_game_minmax_think",
    "az_net_create_oracle", "az_net_create_rollout", "az_net_create_resnet", "az_net_create_simplenet", "az_net_num_params", "az_net_load", "az_net_load_device",
    "az_net_forward", "az_net_forward_logits", "az_net_set_profiling", "az_net_get_profile", "az_net_destroy",
    "az_mcts_create", "az_mcts_set_roots", "az_mcts_set_noise", "az_mcts_run", "az_mcts_explore", "az_mcts_root_stats", "az_mcts_policy",
    "az_mcts_reset", "az_mcts_counters", "az_mcts_last_timing", "az_mcts_destroy", "az_mcts_set_profiling", "az_mcts_get_profile",
    "az_selfplay_create", "az_selfplay_start", "az_selfplay_poll", "az_selfplay_wait", "az_selfplay_counts",
    "az_selfplay_fetch", "az_selfplay_stats", "az_selfplay_destroy", "az_selfplay_create_duel", "az_selfplay_create_duel_players", "az_selfplay_create_duel_minmax", "az_selfplay_outcomes",
    "az_selfplay_export_samples", "az_samples_from_host", "az_samples_count", "az_samples_concat", "az_samples_merge_by_state",
    "az_samples_augment_with_symmetries", "az_samples_convert", "az_samples_fetch", "az_samples_destroy",
    "az_comm_unique_id", "az_comm_create", "az_comm_rank", "az_comm_last_ms", "az_comm_destroy", "az_samples_allgather", "az_net_broadcast",
]

_lib = None


def lib():
    """Loads libazb200.so (build it first with alphazero.jl_b200/build.py); fails loudly if missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libazb200.so is not built: run `python alphazero.jl_b200/build.py` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.az_last_error.restype = C.c_char_p
        L.az_last_error.argtypes = [C.c_void_p]
        L.az_ctx_num_launches.restype = C.c_int64
        L.az_ctx_num_launches.argtypes = [C.c_void_p]
        L.az_game_lookup.argtypes = [C.c_char_p]
        vp = C.c_void_p
        sigs = {
            "az_ctx_create": [C.c_int32, C.POINTER(vp)], "az_ctx_destroy": [vp], "az_ctx_synchronize": [vp],
            "az_game_state_dim": [C.c_int32, vp], "az_game_vectorize_state": [C.c_int32, vp, vp],
            "az_game_actions_mask": [C.c_int32, vp, vp], "az_game_play": [C.c_int32, vp, C.c_int32, vp, vp, vp],
            "az_game_heuristic_value": [C.c_int32, vp, vp], "az_game_minmax_think": [C.c_int32, vp, C.POINTER(_MinMaxParams), vp, vp],
            "az_game_init_state": [C.c_int32, vp],
            "az_game_random_positions": [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, vp],
            "az_net_create_oracle": [vp, C.c_int32, C.c_int32, C.POINTER(vp)],
            "az_net_create_rollout": [vp, C.c_int32, C.c_double, C.c_uint64, C.POINTER(vp)],
            "az_net_create_resnet": [vp, C.c_int32, C.POINTER(_ResNetHP), C.POINTER(vp)],
            "az_net_create_simplenet": [vp, C.c_int32, C.POINTER(_SimpleNetHP), C.POINTER(vp)],
            "az_net_num_params": [vp, C.POINTER(C.c_int64)], "az_net_load": [vp, vp, C.c_int64], "az_net_load_device": [vp, vp, C.c_int64],
            "az_net_forward": [vp, vp, C.c_int32, vp, vp, vp], "az_net_destroy": [vp],
            "az_net_forward_logits": [vp, vp, C.c_int32, vp, vp],
            "az_net_set_profiling": [vp, C.c_int32], "az_net_get_profile": [vp, vp, vp, vp, vp],
            "az_mcts_create": [vp, C.c_int32, vp, C.POINTER(_MctsParams), C.c_int32, C.c_int32, C.POINTER(vp)],
            "az_mcts_set_roots": [vp, vp, vp], "az_mcts_run": [vp, C.c_int32], "az_mcts_set_noise": [vp, C.c_uint64, vp, vp],
            "az_mcts_explore": [vp, vp, vp, C.c_int32, vp, vp, vp], "az_mcts_root_stats": [vp, vp, vp, vp],
            "az_mcts_policy": [vp, vp], "az_mcts_reset": [vp], "az_mcts_counters": [vp, vp, vp, vp],
            "az_mcts_last_timing": [vp, vp, vp, vp, vp], "az_mcts_destroy": [vp],
            "az_mcts_set_profiling": [vp, C.c_int32], "az_mcts_get_profile": [vp, vp, vp, vp, vp],
            "az_selfplay_create": [vp, C.c_int32, vp, C.POINTER(_MctsParams), C.POINTER(_SimParams), C.c_uint64, C.POINTER(vp)],
            "az_selfplay_start": [vp, C.c_int32, C.c_int64], "az_selfplay_poll": [vp, vp, vp], "az_selfplay_wait": [vp],
            "az_selfplay_counts": [vp, vp, vp], "az_selfplay_fetch": [vp] + [vp] * 8, "az_selfplay_stats": [vp, vp, vp, vp, vp],
            "az_selfplay_destroy": [vp],
            "az_selfplay_create_duel": [vp, C.c_int32, vp, vp, C.POINTER(_MctsParams), C.POINTER(_SimParams), C.c_uint64, C.POINTER(vp)],
            "az_selfplay_create_duel_players": [vp, C.c_int32, vp, C.POINTER(_MctsParams), vp, C.POINTER(_MctsParams), C.POINTER(_SimParams), C.c_uint64,
                                                C.POINTER(vp)],
            "az_selfplay_create_duel_minmax": [vp, C.c_int32, vp, C.POINTER(_MctsParams), C.POINTER(_MinMaxParams), C.POINTER(_SimParams), C.c_uint64,
                                               C.POINTER(vp)],
            "az_selfplay_outcomes": [vp, C.c_double, vp, vp, vp, vp],
            "az_selfplay_export_samples": [vp, C.POINTER(vp)],
            "az_samples_from_host": [vp, C.c_int32, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp)],
            "az_samples_count": [vp, C.POINTER(C.c_int64)], "az_samples_concat": [vp, vp, C.POINTER(vp)],
            "az_samples_merge_by_state": [vp, C.POINTER(vp)], "az_samples_augment_with_symmetries": [vp, C.POINTER(vp)],
            "az_samples_convert": [vp, C.c_int32, vp, vp, vp, vp, vp], "az_samples_fetch": [vp, vp, vp, vp, vp, vp],
            "az_samples_destroy": [vp],
            "az_comm_unique_id": [vp, vp], "az_comm_create": [vp, vp, C.c_int32, C.c_int32, C.POINTER(vp)],
            "az_comm_rank": [vp, vp, vp], "az_comm_last_ms": [vp, C.POINTER(C.c_double)], "az_comm_destroy": [vp],
            "az_samples_allgather": [vp, vp, C.POINTER(vp), vp], "az_net_broadcast": [vp, vp, vp, C.c_int64, C.c_int32],
        }
        for name, args in sigs.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = C.c_int32
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One context per GPU per process (az_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        st = lib().az_ctx_create(device, C.byref(self.h))
        if st != AZ_OK:
            raise AzError(st, lib().az_last_error(None).decode())

    def check(self, st):
        if st != AZ_OK:
            raise AzError(st, lib().az_last_error(self.h).decode())

    def synchronize(self):
        self.check(lib().az_ctx_synchronize(self.h))

    @property
    def num_launches(self):
        return lib().az_ctx_num_launches(self.h)

    def close(self):
        if self.h:
            lib().az_ctx_destroy(self.h)
            self.h = None


class GameSpec:
    """GI.AbstractGameSpec for the games known to the library (src/examples.jl:17-21)."""

    def __init__(self, name):
        self.name = name
        self.id = lib().az_game_lookup(name.encode())
        if self.id < 0:
            raise KeyError("unknown game %r (known: connect-four, tictactoe, mancala, grid-world)" % name)
        L = lib()
        self.num_actions = L.az_game_num_actions(self.id)
        self.state_bytes = L.az_game_state_bytes(self.id)
        self.max_plies = L.az_game_max_plies(self.id)
        d = (C.c_int32 * 3)()
        L.az_game_state_dim(self.id, d)
        self.state_dim = tuple(d)

    def init_state(self):
        s = np.zeros(self.state_bytes, np.uint8)
        lib().az_game_init_state(self.id, s.ctypes.data)
        return s

    def vectorize_state(self, state):
        s = np.ascontiguousarray(state, np.uint8)
        x = np.zeros(int(np.prod(self.state_dim)), np.float32)
        lib().az_game_vectorize_state(self.id, s.ctypes.data, x.ctypes.data)
        return x.reshape(self.state_dim, order="F")

    def actions_mask(self, state):
        s = np.ascontiguousarray(state, np.uint8)
        m = np.zeros(self.num_actions, np.uint8)
        lib().az_game_actions_mask(self.id, s.ctypes.data, m.ctypes.data)
        return m.astype(bool)

    def play(self, state, action):
        """GI.play! on a copy: returns (next_state, terminated, white_reward)."""
        s = np.ascontiguousarray(state, np.uint8)
        ns = np.zeros(self.state_bytes, np.uint8)
        term, wr = C.c_int32(), C.c_double()
        st = lib().az_game_play(self.id, s.ctypes.data, int(action), ns.ctypes.data, C.byref(term), C.byref(wr))
        if st != AZ_OK:
            raise AzError(st, "illegal action %d" % action)
        return ns, bool(term.value), wr.value

    def heuristic_value(self, state):
        """GI.heuristic_value of the position (host evaluation of the kernels' inline code)."""
        s = np.ascontiguousarray(state, np.uint8)
        v = C.c_double()
        st = lib().az_game_heuristic_value(self.id, s.ctypes.data, C.byref(v))
        if st != AZ_OK:
            raise AzError(st, "az_game_heuristic_value")
        return v.value

    def minmax_think(self, state, player):
        """think(::MinMax.Player, game) (src/minmax.jl:83-114) for a `MinMaxTS` player: (q [A], pi [A]), zero on unavailable actions."""
        s = np.ascontiguousarray(state, np.uint8)
        q, pi = np.zeros(self.num_actions), np.zeros(self.num_actions)
        mm = player.c()
        st = lib().az_game_minmax_think(self.id, s.ctypes.data, C.byref(mm), q.ctypes.data, pi.ctypes.data)
        if st != AZ_OK:
            raise AzError(st, "az_game_minmax_think")
        return q, pi

    def random_positions(self, seed, n, max_plies=30, first_stream=0):
        out = np.zeros((n, self.state_bytes), np.uint8)
        st = lib().az_game_random_positions(self.id, seed, first_stream, n, max_plies, out.ctypes.data)
        if st != AZ_OK:
            raise AzError(st, "az_game_random_positions")
        return out


class ConstSchedule:  # src/schedule.jl:22
    def __init__(self, v):
        self.xs, self.ys = [0], [float(v)]


class PLSchedule:  # src/schedule.jl:37-80
    def __init__(self, xs, ys):
        assert len(xs) == len(ys) and 1 <= len(xs) <= MAX_SCHEDULE
        self.xs, self.ys = [int(x) for x in xs], [float(y) for y in ys]


class MctsParams:
    """src/params.jl:49-57 (keyword names and defaults of the reference)."""

    def __init__(self, gamma=1.0, cpuct=1.0, num_iters_per_turn=None, temperature=None, dirichlet_noise_ϵ=None,
                 dirichlet_noise_α=None, prior_temperature=1.0, **kw):
        eps = kw.pop("dirichlet_noise_eps", dirichlet_noise_ϵ)
        alpha = kw.pop("dirichlet_noise_alpha", dirichlet_noise_α)
        assert not kw, kw
        assert num_iters_per_turn is not None and eps is not None and alpha is not None
        self.gamma, self.cpuct, self.num_iters_per_turn = gamma, cpuct, num_iters_per_turn
        self.temperature = temperature if temperature is not None else ConstSchedule(1.0)
        self.dirichlet_noise_eps, self.dirichlet_noise_alpha, self.prior_temperature = eps, alpha, prior_temperature

    def c(self):
        p = _MctsParams()
        p.gamma, p.cpuct, p.num_iters_per_turn = self.gamma, self.cpuct, self.num_iters_per_turn
        p.dirichlet_noise_eps, p.dirichlet_noise_alpha, p.prior_temperature = \
            self.dirichlet_noise_eps, self.dirichlet_noise_alpha, self.prior_temperature
        p.temperature_n = len(self.temperature.xs)
        for i, (x, y) in enumerate(zip(self.temperature.xs, self.temperature.ys)):
            p.temperature_xs[i], p.temperature_ys[i] = x, y
        return p


def NetworkOnly(τ=1.0, **kw):
    """Benchmark.NetworkOnly(τ) (src/benchmark.jl:161-176) = PlayerWithTemperature(NetworkPlayer(nn), ConstSchedule(τ)): the
    parameter block that makes a self-play / duel player use the network's policy directly (num_iters_per_turn = 0 in the C
    ABI); pass it where an MctsParams goes (`SelfPlayParams(NetworkOnly(0.5), sim)`, `simulate(..., baseline_mcts=NetworkOnly())`)."""
    tau = kw.pop("tau", τ)
    assert not kw, kw
    return MctsParams(num_iters_per_turn=0, temperature=ConstSchedule(tau), dirichlet_noise_eps=0.0, dirichlet_noise_alpha=1.0)


class MinMaxTS:
    """Benchmark.MinMaxTS(depth, amplify_rewards, τ) (src/benchmark.jl:178-196) = MinMax.Player (src/minmax.jl:72-81): pass it
    as the `baseline` of `simulate` / `SelfPlay`; it brings no oracle."""

    def __init__(self, depth, amplify_rewards, τ=0.0, γ=1.0, **kw):
        self.depth, self.amplify_rewards = int(depth), bool(amplify_rewards)
        self.tau, self.gamma = float(kw.pop("tau", τ)), float(kw.pop("gamma", γ))
        assert not kw, kw

    def c(self):
        p = _MinMaxParams()
        p.depth, p.amplify_rewards, p.tau, p.gamma = self.depth, int(self.amplify_rewards), self.tau, self.gamma
        return p


class SimParams:
    """src/params.jl:92-101."""

    def __init__(self, num_games, num_workers, batch_size, use_gpu=True, fill_batches=True, reset_every=1,
                 flip_probability=0.0, alternate_colors=False):
        self.num_games, self.num_workers, self.batch_size = num_games, num_workers, batch_size
        self.use_gpu, self.fill_batches, self.reset_every = use_gpu, fill_batches, reset_every
        self.flip_probability, self.alternate_colors = flip_probability, alternate_colors

    def c(self):
        p = _SimParams()
        p.num_games, p.num_workers, p.batch_size = self.num_games, self.num_workers, self.batch_size
        p.fill_batches, p.alternate_colors = int(self.fill_batches), int(self.alternate_colors)
        p.reset_every = -1 if self.reset_every is None else int(self.reset_every)
        p.flip_probability = self.flip_probability
        return p


class SelfPlayParams:  # src/params.jl:160-163
    def __init__(self, mcts, sim):
        self.mcts, self.sim = mcts, sim


class ResNetHP:  # src/networks/architectures/resnet.jl:30-37
    def __init__(self, num_blocks, num_filters, conv_kernel_size, num_policy_head_filters=2, num_value_head_filters=1,
                 batch_norm_momentum=0.6):
        self.num_blocks, self.num_filters, self.conv_kernel_size = num_blocks, num_filters, tuple(conv_kernel_size)
        self.num_policy_head_filters, self.num_value_head_filters = num_policy_head_filters, num_value_head_filters
        self.batch_norm_momentum = batch_norm_momentum


class SimpleNetHP:  # src/networks/architectures/simplenet.jl:15-22
    def __init__(self, width, depth_common, depth_phead=1, depth_vhead=1, use_batch_norm=False, batch_norm_momentum=0.6):
        self.width, self.depth_common, self.depth_phead, self.depth_vhead = width, depth_common, depth_phead, depth_vhead
        self.use_batch_norm, self.batch_norm_momentum = use_batch_norm, batch_norm_momentum


def fresh_resnet_blob(gspec, hp, seed=1):
    """Parameters of a freshly constructed Flux ResNet (`ResNet(gspec, hp)`, src/networks/architectures/resnet.jl:65-92:
    Glorot-uniform conv / dense weights, zero biases, BatchNorm gamma 1, beta 0, mu 0, sigma2 1) as the float32 blob
    `Network.load` takes (order: include/azb200.h).  No checkpoints travel with the repository: benches and scripts start here."""
    W, H, C = gspec.state_dim
    nf, nb, npf, nvf = hp.num_filters, hp.num_blocks, hp.num_policy_head_filters, hp.num_value_head_filters
    kw, kh = hp.conv_kernel_size
    rng = np.random.default_rng(seed)
    parts = []

    def conv(kw_, kh_, ci, co):
        s = np.sqrt(6.0 / (kw_ * kh_ * ci + kw_ * kh_ * co))
        parts.extend([rng.uniform(-s, s, kw_ * kh_ * ci * co), np.zeros(co)])

    def bn(n):
        parts.extend([np.ones(n), np.zeros(n), np.zeros(n), np.ones(n)])

    def dense(out, inn):
        s = np.sqrt(6.0 / (inn + out))
        parts.extend([rng.uniform(-s, s, out * inn), np.zeros(out)])
    conv(kw, kh, C, nf); bn(nf)
    for _ in range(nb):
        conv(kw, kh, nf, nf); bn(nf); conv(kw, kh, nf, nf); bn(nf)
    conv(1, 1, nf, nvf); bn(nvf); dense(nf, W * H * nvf); dense(1, nf)
    conv(1, 1, nf, npf); bn(npf); dense(gspec.num_actions, W * H * npf)
    return np.concatenate(parts).astype(np.float32)


def fresh_simplenet_blob(gspec, hp, seed=1):
    """The same for `SimpleNet(gspec, hp)` (src/networks/architectures/simplenet.jl:24-49): common = Dense(indim, width) + depth_common
    hidden layers, vhead = depth_vhead hidden layers + Dense(width, 1), phead = depth_phead hidden layers + Dense(width, A); a
    hidden layer is Dense [+ BatchNorm] with zero bias."""
    W, H, C = gspec.state_dim
    w = hp.width
    rng = np.random.default_rng(seed)
    parts = []

    def dense(out, inn):
        s = np.sqrt(6.0 / (inn + out))
        parts.extend([rng.uniform(-s, s, out * inn), np.zeros(out)])

    def hidden(inn, out):
        dense(out, inn)
        if hp.use_batch_norm:
            parts.extend([np.ones(out), np.zeros(out), np.zeros(out), np.ones(out)])
    hidden(W * H * C, w)
    for _ in range(hp.depth_common):
        hidden(w, w)
    for _ in range(hp.depth_vhead):
        hidden(w, w)
    dense(1, w)
    for _ in range(hp.depth_phead):
        hidden(w, w)
    dense(gspec.num_actions, w)
    return np.concatenate(parts).astype(np.float32)


class Network:
    """An MCTS oracle living on the GPU (az_net): RandomOracle, the synthetic hash oracle, ResNet or SimpleNet."""

    def __init__(self, ctx, gspec, handle, kind):
        self.ctx, self.gspec, self.h, self.kind = ctx, gspec, handle, kind

    @property
    def num_params(self):
        n = C.c_int64()
        self.ctx.check(lib().az_net_num_params(self.h, C.byref(n)))
        return n.value

    def load(self, blob):
        b = np.ascontiguousarray(blob, np.float32)
        self.ctx.check(lib().az_net_load(self.h, b.ctypes.data, b.size))
        return self

    def load_device(self, device_ptr, n):
        """Parameters already in device memory of this network's GPU (az_net_load_device): no host round trip."""
        self.ctx.check(lib().az_net_load_device(self.h, C.c_void_p(int(device_ptr)), int(n)))
        return self

    def evaluate_batch(self, states):
        """Network.evaluate_batch (src/networks/network.jl:308-315): returns (P [B,A] masked+renormalised, V [B], Pinvalid [B])."""
        s = np.ascontiguousarray(states, np.uint8).reshape(-1, self.gspec.state_bytes)
        B = s.shape[0]
        P = np.zeros((B, self.gspec.num_actions), np.float32)
        V = np.zeros(B, np.float32)
        Pi = np.zeros(B, np.float32)
        self.ctx.check(lib().az_net_forward(self.h, s.ctypes.data, B, P.ctypes.data, V.ctypes.data, Pi.ctypes.data))
        return P, V, Pi

    def forward_logits(self, states):
        """Parity hook: (policy logits [B,A] before the softmax, value before the tanh [B])."""
        s = np.ascontiguousarray(states, np.uint8).reshape(-1, self.gspec.state_bytes)
        B = s.shape[0]
        L = np.zeros((B, self.gspec.num_actions), np.float32)
        Vp = np.zeros(B, np.float32)
        self.ctx.check(lib().az_net_forward_logits(self.h, s.ctypes.data, B, L.ctypes.data, Vp.ctypes.data))
        return L, Vp

    def set_profiling(self, enable=True):
        self.ctx.check(lib().az_net_set_profiling(self.h, int(enable)))

    def get_profile(self):
        tw, tt = C.c_double(), C.c_double()
        nl, ne = C.c_int64(), C.c_int64()
        self.ctx.check(lib().az_net_get_profile(self.h, C.byref(tw), C.byref(nl), C.byref(tt), C.byref(ne)))
        return dict(tower_ms=tw.value, tower_launches=nl.value, total_ms=tt.value, evals=ne.value)

    def close(self):
        if self.h:
            lib().az_net_destroy(self.h)
            self.h = None


def RandomOracle(ctx, gspec):  # src/mcts.jl:62-72
    h = C.c_void_p()
    ctx.check(lib().az_net_create_oracle(ctx.h, NET_UNIFORM, gspec.id, C.byref(h)))
    return Network(ctx, gspec, h, NET_UNIFORM)


def RolloutOracle(ctx, gspec, gamma=1.0, seed=0):  # src/mcts.jl:27-60 (the oracle of Benchmark.MctsRollouts, src/benchmark.jl:134-147)
    h = C.c_void_p()
    ctx.check(lib().az_net_create_rollout(ctx.h, gspec.id, gamma, seed, C.byref(h)))
    return Network(ctx, gspec, h, 4)


def SynthOracle(ctx, gspec):
    h = C.c_void_p()
    ctx.check(lib().az_net_create_oracle(ctx.h, NET_SYNTH, gspec.id, C.byref(h)))
    return Network(ctx, gspec, h, NET_SYNTH)


def ResNet(ctx, gspec, hp):  # src/networks/architectures/resnet.jl:65-92
    c = _ResNetHP()
    c.num_blocks, c.num_filters = hp.num_blocks, hp.num_filters
    c.conv_kernel_size[0], c.conv_kernel_size[1] = hp.conv_kernel_size
    c.num_policy_head_filters, c.num_value_head_filters = hp.num_policy_head_filters, hp.num_value_head_filters
    c.batch_norm_momentum = hp.batch_norm_momentum
    h = C.c_void_p()
    ctx.check(lib().az_net_create_resnet(ctx.h, gspec.id, C.byref(c), C.byref(h)))
    return Network(ctx, gspec, h, NET_RESNET)


def SimpleNet(ctx, gspec, hp):  # src/networks/architectures/simplenet.jl:37-64
    c = _SimpleNetHP()
    c.width, c.depth_common, c.depth_phead, c.depth_vhead = hp.width, hp.depth_common, hp.depth_phead, hp.depth_vhead
    c.use_batch_norm, c.batch_norm_momentum = int(hp.use_batch_norm), hp.batch_norm_momentum
    h = C.c_void_p()
    ctx.check(lib().az_net_create_simplenet(ctx.h, gspec.id, C.byref(c), C.byref(h)))
    return Network(ctx, gspec, h, NET_SIMPLENET)


class MctsEnv:
    """A pool of n_trees independent MCTS.Env (src/mcts.jl:124-151), one per root, on one GPU."""

    def __init__(self, ctx, gspec, oracle, params, n_trees, capacity_nodes_per_tree=None):
        self.ctx, self.gspec, self.oracle, self.params, self.n = ctx, gspec, oracle, params, n_trees
        cap = capacity_nodes_per_tree or params.num_iters_per_turn * 4
        self.h = C.c_void_p()
        p = params.c()
        ctx.check(lib().az_mcts_create(ctx.h, gspec.id, oracle.h, C.byref(p), n_trees, cap, C.byref(self.h)))

    def set_roots(self, states, eta=None):
        s = np.ascontiguousarray(states, np.uint8).reshape(self.n, self.gspec.state_bytes)
        e = None if eta is None else np.ascontiguousarray(eta, np.float64).reshape(self.n, self.gspec.num_actions)
        self.ctx.check(lib().az_mcts_set_roots(self.h, s.ctypes.data, _ptr(e)))

    def set_noise(self, seed, games, moves):
        """Stochastic environments: in-tree noise stream ids per tree (see az_mcts_set_noise)."""
        g = np.ascontiguousarray(games, np.int64)
        m = np.ascontiguousarray(moves, np.int32)
        self.ctx.check(lib().az_mcts_set_noise(self.h, seed, g.ctypes.data, m.ctypes.data))

    def run(self, nsims=None):
        self.ctx.check(lib().az_mcts_run(self.h, nsims or self.params.num_iters_per_turn))

    def explore(self, states, nsims=None, eta=None):
        """MCTS.explore! on every tree with HOST buffers in and out: returns (N, W, P), each [n_trees, A]."""
        A = self.gspec.num_actions
        s = np.ascontiguousarray(states, np.uint8).reshape(self.n, self.gspec.state_bytes)
        e = None if eta is None else np.ascontiguousarray(eta, np.float64).reshape(self.n, A)
        N = np.zeros((self.n, A), np.int64)
        W = np.zeros((self.n, A), np.float64)
        P = np.zeros((self.n, A), np.float32)
        self.ctx.check(lib().az_mcts_explore(self.h, s.ctypes.data, _ptr(e), nsims or self.params.num_iters_per_turn,
                                             N.ctypes.data, W.ctypes.data, P.ctypes.data))
        return N, W, P

    def root_stats(self):
        A = self.gspec.num_actions
        N = np.zeros((self.n, A), np.int64)
        W = np.zeros((self.n, A), np.float64)
        P = np.zeros((self.n, A), np.float32)
        self.ctx.check(lib().az_mcts_root_stats(self.h, N.ctypes.data, W.ctypes.data, P.ctypes.data))
        return N, W, P

    def policy(self):
        pi = np.zeros((self.n, self.gspec.num_actions), np.float64)
        self.ctx.check(lib().az_mcts_policy(self.h, pi.ctypes.data))
        return pi

    def reset(self):
        self.ctx.check(lib().az_mcts_reset(self.h))

    def counters(self):
        ts = np.zeros(self.n, np.int64)
        tn = np.zeros(self.n, np.int64)
        nn = np.zeros(self.n, np.int64)
        self.ctx.check(lib().az_mcts_counters(self.h, ts.ctypes.data, tn.ctypes.data, nn.ctypes.data))
        return ts, tn, nn

    def last_timing(self):
        a, b = C.c_double(), C.c_double()
        t, e = C.c_int64(), C.c_int64()
        lib().az_mcts_last_timing(self.h, C.byref(a), C.byref(b), C.byref(t), C.byref(e))
        return dict(ms_total=a.value, ms_network=b.value, ticks=t.value, expansions=e.value)

    def set_profiling(self, enable=True):
        self.ctx.check(lib().az_mcts_set_profiling(self.h, 1 if enable else 0))

    def get_profile(self):
        """Device time (ms) inside az_k_select / the network / az_k_expand_backup over the profiled runs since the last call."""
        a, b, c, t = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self.ctx.check(lib().az_mcts_get_profile(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(t)))
        return dict(select_ms=a.value, expand_ms=b.value, net_ms=c.value, ticks=t.value)

    def close(self):
        if self.h:
            lib().az_mcts_destroy(self.h)
            self.h = None


class SelfPlay:
    """simulate() for self-play (src/simulations.jl:207-244, src/training.jl:275-300)."""

    def __init__(self, ctx, gspec, oracle, params, seed=0, baseline=None, baseline_mcts=None):
        """`baseline` given: a duel TwoPlayers(MctsPlayer(oracle, params.mcts), MctsPlayer(baseline, baseline_mcts or params.mcts))
        (src/training.jl:130-143; two different MctsPlayers as in Benchmark duels, src/benchmark.jl:78-99)."""
        self.ctx, self.gspec, self.params = ctx, gspec, params
        self.h = C.c_void_p()
        mp, sp = params.mcts.c(), params.sim.c()
        if baseline is None:
            ctx.check(lib().az_selfplay_create(ctx.h, gspec.id, oracle.h, C.byref(mp), C.byref(sp), seed, C.byref(self.h)))
        elif isinstance(baseline, MinMaxTS):
            mm = baseline.c()
            ctx.check(lib().az_selfplay_create_duel_minmax(ctx.h, gspec.id, oracle.h, C.byref(mp), C.byref(mm), C.byref(sp), seed, C.byref(self.h)))
        elif baseline_mcts is not None:
            mb = baseline_mcts.c()
            ctx.check(lib().az_selfplay_create_duel_players(ctx.h, gspec.id, oracle.h, C.byref(mp), baseline.h, C.byref(mb), C.byref(sp), seed,
                                                            C.byref(self.h)))
        else:
            ctx.check(lib().az_selfplay_create_duel(ctx.h, gspec.id, oracle.h, baseline.h, C.byref(mp), C.byref(sp), seed,
                                                    C.byref(self.h)))

    def start(self, num_games=None, first_game_index=0):
        self.ctx.check(lib().az_selfplay_start(self.h, num_games or self.params.sim.num_games, first_game_index))

    def poll(self):
        d, f = C.c_int32(), C.c_int32()
        self.ctx.check(lib().az_selfplay_poll(self.h, C.byref(d), C.byref(f)))
        return d.value, bool(f.value)

    def wait(self):
        self.ctx.check(lib().az_selfplay_wait(self.h))

    def fetch(self):
        ns, ng = C.c_int64(), C.c_int64()
        self.ctx.check(lib().az_selfplay_counts(self.h, C.byref(ns), C.byref(ng)))
        n, g, A = ns.value, ng.value, self.gspec.num_actions
        out = dict(states=np.zeros((n, self.gspec.state_bytes), np.uint8), pi=np.zeros((n, A), np.float32),
                   mask=np.zeros((n, A), np.uint8), z=np.zeros(n, np.float32), t=np.zeros(n, np.float32),
                   game=np.zeros(n, np.int32), rewards=np.zeros(n, np.float64), actions=np.zeros(n, np.int32))
        self.ctx.check(lib().az_selfplay_fetch(self.h, *[out[k].ctypes.data for k in
                                                         ("states", "pi", "mask", "z", "t", "game", "rewards", "actions")]))
        ed, nodes, moves, tot = np.zeros(g), np.zeros(g, np.int64), np.zeros(g, np.int32), np.zeros(4)
        self.ctx.check(lib().az_selfplay_stats(self.h, ed.ctypes.data, nodes.ctypes.data, moves.ctypes.data, tot.ctypes.data))
        out.update(edepth=ed, nodes=nodes, moves=moves, seconds=tot[0], simulations=tot[1], expansions=tot[2], samples=tot[3])
        return out

    def outcomes(self, gamma=1.0):
        """rewards_and_redundancy (src/simulations.jl:292-307): (game_rewards = total reward per game w.r.t. the first player, colors_flipped,
        final states, redundancy)."""
        ns, ng = C.c_int64(), C.c_int64()
        self.ctx.check(lib().az_selfplay_counts(self.h, C.byref(ns), C.byref(ng)))
        g = ng.value
        rew, fl = np.zeros(g, np.float64), np.zeros(g, np.int32)
        fin = np.zeros((g, self.gspec.state_bytes), np.uint8)
        red = C.c_double()
        self.ctx.check(lib().az_selfplay_outcomes(self.h, gamma, rew.ctypes.data, fl.ctypes.data, fin.ctypes.data, C.byref(red)))
        return dict(game_rewards=rew, colors_flipped=fl, final_states=fin, redundancy=red.value)

    def close(self):
        if self.h:
            lib().az_selfplay_destroy(self.h)
            self.h = None


CONSTANT_WEIGHT, LOG_WEIGHT, LINEAR_WEIGHT = 0, 1, 2  # SamplesWeighingPolicy, src/params.jl:104-108


class Samples:
    """A device-resident Vector{TrainingSample} (src/memory.jl:20-26): the replay-buffer side of the wire.  Every operation
    returns a new set; `close()` frees the device memory."""

    def __init__(self, ctx, gspec, handle):
        self.ctx, self.gspec, self.h = ctx, gspec, handle

    @classmethod
    def from_selfplay(cls, sp):
        """push_trace! rows of a finished SelfPlay run, ordered by (game, ply), no host round trip (src/memory.jl:74-87)."""
        h = C.c_void_p()
        sp.ctx.check(lib().az_selfplay_export_samples(sp.h, C.byref(h)))
        return cls(sp.ctx, sp.gspec, h)

    @classmethod
    def from_host(cls, ctx, gspec, states, pi, z, t, n=None):
        s = np.ascontiguousarray(states, np.uint8).reshape(-1, gspec.state_bytes)
        k = s.shape[0]
        pi = np.ascontiguousarray(pi, np.float64).reshape(k, gspec.num_actions)
        z, t = np.ascontiguousarray(z, np.float64), np.ascontiguousarray(t, np.float64)
        nn = None if n is None else np.ascontiguousarray(n, np.int32)
        h = C.c_void_p()
        ctx.check(lib().az_samples_from_host(ctx.h, gspec.id, k, s.ctypes.data, pi.ctypes.data, z.ctypes.data, t.ctypes.data,
                                             None if nn is None else nn.ctypes.data, C.byref(h)))
        return cls(ctx, gspec, h)

    def __len__(self):
        n = C.c_int64()
        self.ctx.check(lib().az_samples_count(self.h, C.byref(n)))
        return n.value

    def _new(self, fn, *args):
        h = C.c_void_p()
        self.ctx.check(fn(self.h, *args, C.byref(h)))
        return Samples(self.ctx, self.gspec, h)

    def concat(self, other):
        return self._new(lib().az_samples_concat, other.h)

    def merge_by_state(self):  # src/memory.jl:98-110 (groups in order of first occurrence)
        return self._new(lib().az_samples_merge_by_state)

    def augment_with_symmetries(self):  # src/memory.jl:126-130
        return self._new(lib().az_samples_augment_with_symmetries)

    def convert_buffers(self):
        """Host arrays of the right shapes for convert(out=...), already touched (a fresh np.zeros array is lazily mapped: the first
        write to it pays one page fault per 4 KB, which would be charged to whoever fills it)."""
        n, a = len(self), self.gspec.num_actions
        xd = int(np.prod(self.gspec.state_dim))
        out = dict(W=np.empty(n, np.float32), X=np.empty((n, xd), np.float32), A=np.empty((n, a), np.float32),
                   P=np.empty((n, a), np.float32), V=np.empty(n, np.float32))
        for v in out.values():
            v.fill(0)
        return out

    def convert(self, weighing=LOG_WEIGHT, out=None):
        """convert_samples (src/learning.jl:38-51): dict of Float32 arrays W [n], X [n, state_dim], A [n, a], P [n, a], V [n]."""
        n, a = len(self), self.gspec.num_actions
        xd = int(np.prod(self.gspec.state_dim))
        if out is None:
            out = dict(W=np.zeros(n, np.float32), X=np.zeros((n, xd), np.float32), A=np.zeros((n, a), np.float32),
                       P=np.zeros((n, a), np.float32), V=np.zeros(n, np.float32))
        self.ctx.check(lib().az_samples_convert(self.h, weighing, *[out[k].ctypes.data for k in ("W", "X", "A", "P", "V")]))
        return out

    def fetch(self):
        n, a = len(self), self.gspec.num_actions
        out = dict(states=np.zeros((n, self.gspec.state_bytes), np.uint8), pi=np.zeros((n, a), np.float64), z=np.zeros(n, np.float64),
                   t=np.zeros(n, np.float64), n=np.zeros(n, np.int32))
        self.ctx.check(lib().az_samples_fetch(self.h, *[out[k].ctypes.data for k in ("states", "pi", "z", "t", "n")]))
        return out

    def close(self):
        if self.h:
            lib().az_samples_destroy(self.h)
            self.h = None


def _preload_nccl():
    """libazb200.so binds NCCL at run time (dlopen "libnccl.so.2", reusing a copy already in the process).  If PyTorch is
    imported LATER in the same process it must find its own, newer NCCL under that soname, so the newest copy that ships
    with the Python environment (nvidia-nccl wheel) is loaded first; the system library is the fallback inside the C code."""
    import sys
    if "torch" in sys.modules:
        return  # torch has already loaded the NCCL it was built against
    for d in sys.path:
        cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                pass


class Comm:
    """One rank of the engine's NCCL communicator (one process per GPU, src/simulations.jl:252-290).  `exchange(id_bytes)`
    is the caller's own channel for the 128-byte id: it receives rank 0's bytes (None on the other ranks) and returns
    rank 0's bytes on every rank -- e.g. a torch.distributed broadcast, an MPI bcast or the Julia `Distributed` call that
    starts the workers."""
    ID_BYTES = 128

    def __init__(self, ctx, rank, world, exchange=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        _preload_nccl()
        ident = None
        if rank == 0:
            ident = np.zeros(self.ID_BYTES, np.uint8)
            ctx.check(lib().az_comm_unique_id(ctx.h, ident.ctypes.data))
        if world > 1:
            if exchange is None:
                raise ValueError("Comm: world > 1 needs an `exchange` callable for the communicator id")
            ident = np.frombuffer(bytes(exchange(None if ident is None else ident.tobytes())), np.uint8).copy()
        self.h = C.c_void_p()
        ctx.check(lib().az_comm_create(ctx.h, ident.ctypes.data, rank, world, C.byref(self.h)))

    @classmethod
    def from_torch(cls, ctx, dist):
        """Id exchange over an initialised torch.distributed group (any backend); None / uninitialised -> single rank."""
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(ctx, 0, 1)

        def exchange(b):
            box = [b]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        return cls(ctx, dist.get_rank(), dist.get_world_size(), exchange)

    def allgather_samples(self, samples):
        """`reduce(vcat, results)` over ranks (src/simulations.jl:289) for device-resident samples; returns (Samples, counts)."""
        h = C.c_void_p()
        counts = np.zeros(self.world, np.int64)
        self.ctx.check(lib().az_samples_allgather(self.h, samples.h, C.byref(h), counts.ctypes.data))
        return Samples(samples.ctx, samples.gspec, h), counts

    def broadcast_network(self, net, blob=None, root=0):
        """Every rank loads rank `root`'s parameter blob into `net` (same architecture everywhere)."""
        b = None if blob is None else np.ascontiguousarray(blob, np.float32)
        self.ctx.check(lib().az_net_broadcast(self.h, net.h, _ptr(b), net.num_params, root))

    @property
    def last_ms(self):
        ms = C.c_double()
        self.ctx.check(lib().az_comm_last_ms(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            lib().az_comm_destroy(self.h)
            self.h = None


def simulate(ctx, gspec, oracle, params, seed=0, game_simulated=None, first_game_index=0, baseline=None, gamma=None, baseline_mcts=None):
    """simulate(simulator, gspec, p; game_simulated): returns the fetched samples + measurements (self-play simulator,
    src/training.jl:275-300); with `gamma` also the rewards_and_redundancy outputs of the record_trace simulators."""
    import time
    sp = SelfPlay(ctx, gspec, oracle, params, seed, baseline=baseline, baseline_mcts=baseline_mcts)
    try:
        sp.start(params.sim.num_games, first_game_index)
        seen = 0
        while True:
            done, fin = sp.poll()
            if game_simulated:
                for _ in range(done - seen):
                    game_simulated()
            seen = max(seen, done)
            if fin:
                break
            time.sleep(0.005)
        sp.wait()
        out = sp.fetch()
        if gamma is not None:
            out.update(sp.outcomes(gamma))
        return out
    finally:
        sp.close()


def evaluate_network(ctx, gspec, net, params, seed=0, game_simulated=None):
    """evaluate_network (src/training.jl:146-155): a single network on a one-player game with the arena parameters; returns
    (rewards per game, redundancy)."""
    out = simulate(ctx, gspec, net, params, seed, game_simulated, gamma=params.mcts.gamma)
    return out["game_rewards"], out["redundancy"]


class Evaluation:  # Report.Evaluation, src/report.jl:73-80
    def __init__(self, legend, avgr, redundancy, rewards, baseline_rewards, time):
        self.legend, self.avgr, self.redundancy = legend, float(avgr), float(redundancy)
        self.rewards, self.baseline_rewards, self.time = rewards, baseline_rewards, time


def compare_networks(ctx, gspec, contender, baseline, params, seed=0, game_simulated=None, two_players=True):
    """compare_networks (src/training.jl:157-174): two-player games pit the networks against each other; single-player
    games evaluate both separately and compare the mean rewards."""
    import time
    t0 = time.perf_counter()
    legend = "Most recent NN versus best NN so far"
    if two_players:
        rewards_c, red = pit_networks(ctx, gspec, contender, baseline, params, seed, game_simulated)
        avgr, rewards_b = float(np.mean(rewards_c)), None
    else:
        rewards_c, red_c = evaluate_network(ctx, gspec, contender, params, seed, game_simulated)
        rewards_b, red_b = evaluate_network(ctx, gspec, baseline, params, seed, game_simulated)
        avgr, red = float(np.mean(rewards_c)) - float(np.mean(rewards_b)), float(np.mean([red_c, red_b]))
    return Evaluation(legend, avgr, red, rewards_c, rewards_b, time.perf_counter() - t0)


class TernaryOutcomeStatistics:  # src/benchmark.jl:104-121
    def __init__(self, rewards):
        r = np.asarray(rewards.rewards if isinstance(rewards, Evaluation) else rewards)
        self.num_won, self.num_draw, self.num_lost = int((r > 0).sum()), int((r == 0).sum()), int((r < 0).sum())
        assert self.num_won + self.num_draw + self.num_lost == len(r)


def pit_networks(ctx, gspec, contender, baseline, params, seed=0, game_simulated=None):
    """pit_networks (src/training.jl:130-143): `params` = ArenaParams-like object with .mcts and .sim; returns
    (rewards of the contender per game, redundancy)."""
    out = simulate(ctx, gspec, contender, params, seed, game_simulated, baseline=baseline, gamma=params.mcts.gamma)
    return out["game_rewards"], out["redundancy"]
