"""ctypes binding of the CPU oracle (oracle/az_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the product
package.  "parity unpinned" for MCTS statistics (see az_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libaz_oracle.so")

MAX_ACTIONS = 9
STATE_BYTES = 48
MAX_PLIES = 512
GAMES = ["connect-four", "tictactoe", "mancala", "grid-world"]
PURPOSE_DIRICHLET, PURPOSE_CATEGORICAL, PURPOSE_SYMMETRY, PURPOSE_ENV, PURPOSE_POSITION = range(5)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("az_oracle.c", "az_oracle.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaz_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


class MctsParams(C.Structure):
    _fields_ = [("gamma", C.c_double), ("cpuct", C.c_double), ("noise_eps", C.c_double), ("noise_alpha", C.c_double),
                ("prior_temperature", C.c_double), ("num_iters_per_turn", C.c_int), ("sched_n", C.c_int),
                ("sched_xs", C.c_int * 8), ("sched_ys", C.c_double * 8),
                ("player_kind", C.c_int), ("minmax_depth", C.c_int), ("minmax_amplify", C.c_int), ("minmax_tau", C.c_double)]


def minmax_params(depth, amplify_rewards, tau=0.0, gamma=1.0):
    """MinMax.Player (src/minmax.jl:72-81; Benchmark.MinMaxTS, src/benchmark.jl:178-196) as an oz_mcts_params block."""
    p = mcts_params(gamma=gamma, num_iters_per_turn=0)
    p.player_kind, p.minmax_depth, p.minmax_amplify, p.minmax_tau = 1, int(depth), int(bool(amplify_rewards)), float(tau)
    return p


def mcts_params(gamma=1.0, cpuct=1.0, noise_eps=0.0, noise_alpha=1.0, prior_temperature=1.0, num_iters_per_turn=50,
                sched_xs=(0,), sched_ys=(1.0,)):
    p = MctsParams()
    p.gamma, p.cpuct, p.noise_eps, p.noise_alpha, p.prior_temperature = gamma, cpuct, noise_eps, noise_alpha, prior_temperature
    p.num_iters_per_turn = num_iters_per_turn
    p.sched_n = len(sched_xs)
    for i, (x, y) in enumerate(zip(sched_xs, sched_ys)):
        p.sched_xs[i] = int(x)
        p.sched_ys[i] = float(y)
    return p


class Game(C.Structure):
    _fields_ = [("game_id", C.c_int), ("cells", C.c_uint8 * 44), ("curplayer", C.c_uint8), ("finished", C.c_uint8),
                ("winner", C.c_uint8), ("time", C.c_int32), ("last_reward", C.c_double)]


class Trace(C.Structure):
    _fields_ = [("n_moves", C.c_int), ("states", (C.c_uint8 * STATE_BYTES) * (MAX_PLIES + 1)),
                ("pi", (C.c_float * MAX_ACTIONS) * MAX_PLIES), ("mask", (C.c_uint8 * MAX_ACTIONS) * MAX_PLIES),
                ("action", C.c_int32 * MAX_PLIES), ("rewards", C.c_double * MAX_PLIES), ("z", C.c_double * MAX_PLIES),
                ("t", C.c_double * MAX_PLIES), ("mem_nodes", C.c_int64), ("edepth", C.c_double), ("sym", C.c_int32 * MAX_PLIES),
                ("think_states", (C.c_uint8 * STATE_BYTES) * MAX_PLIES), ("pi64", (C.c_double * MAX_ACTIONS) * MAX_PLIES)]


ORACLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float))

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.oz_game_lookup.argtypes = [C.c_char_p]
        L.oz_env_create.restype = C.c_void_p
        L.oz_env_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 5
        for f in ("oz_env_destroy", "oz_env_reset"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("oz_env_num_nodes", "oz_env_total_simulations", "oz_env_total_nodes_traversed"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int64
        L.oz_explore.argtypes = [C.c_void_p, C.POINTER(Game), C.c_int, C.c_void_p]
        L.oz_env_set_noise.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.oz_root_stats.argtypes = [C.c_void_p, C.POINTER(Game)] + [C.c_void_p] * 4
        L.oz_policy.argtypes = [C.c_void_p, C.POINTER(Game), C.c_void_p, C.c_void_p]
        L.oz_game_white_reward.restype = C.c_double
        L.oz_heuristic_value.restype = C.c_double
        L.oz_heuristic_value.argtypes = [C.POINTER(Game)]
        L.oz_minmax_think.argtypes = [C.POINTER(Game), C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oz_det_log.restype = C.c_double
        L.oz_det_log.argtypes = [C.c_double]
        L.oz_det_exp.restype = C.c_double
        L.oz_det_exp.argtypes = [C.c_double]
        L.oz_pl_schedule.restype = C.c_double
        L.oz_pl_schedule.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oz_uniform_f32.restype = C.c_float
        L.oz_uniform_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32]
        L.oz_dirichlet.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_void_p]
        L.oz_categorical.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.oz_apply_temperature.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.oz_fix_probvec.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oz_play_game.argtypes = [C.c_void_p, C.POINTER(MctsParams), C.c_uint64, C.c_uint64, C.POINTER(Trace)]
        L.oz_play_game2.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(MctsParams), C.c_double, C.c_uint64, C.c_uint64, C.POINTER(Trace)]
        L.oz_play_game2p.restype = None
        L.oz_play_game2p.argtypes = [C.c_void_p, C.POINTER(MctsParams), C.c_void_p, C.POINTER(MctsParams), C.c_double, C.c_uint64, C.c_uint64,
                                     C.POINTER(Trace)]
        L.oz_apply_symmetry.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oz_total_reward.restype = C.c_double
        L.oz_total_reward.argtypes = [C.POINTER(Trace), C.c_double]
        L.oz_worker_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(MctsParams), C.c_uint64, C.c_uint64,
                                    C.c_uint64, C.c_int, C.c_int, C.POINTER(Trace)]
        L.oz_random_position.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.oz_philox.argtypes = [C.c_uint64] + [C.c_uint32] * 4 + [C.c_void_p]
        L.oz_state_key.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.oz_vectorize_state.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.oz_game_set_state.argtypes = [C.POINTER(Game), C.c_int, C.c_void_p]
        L.oz_game_get_state.argtypes = [C.POINTER(Game), C.c_void_p]
        L.oz_game_init.argtypes = [C.POINTER(Game), C.c_int]
        L.oz_game_play.argtypes = [C.POINTER(Game), C.c_int, C.c_void_p]
        L.oz_game_actions_mask.argtypes = [C.POINTER(Game), C.c_void_p]
        L.oz_game_terminated.argtypes = [C.POINTER(Game)]
        L.oz_game_white_playing.argtypes = [C.POINTER(Game)]
        L.oz_game_white_reward.argtypes = [C.POINTER(Game)]
        L.oz_batch_create.restype = C.c_void_p
        L.oz_batch_create.argtypes = [C.c_int, C.c_int, C.POINTER(MctsParams)]
        L.oz_batch_destroy.argtypes = [C.c_void_p]
        L.oz_batch_reset_trees.argtypes = [C.c_void_p]
        L.oz_batch_set_roots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oz_batch_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oz_batch_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oz_batch_vectorize.restype = None
        L.oz_batch_vectorize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oz_batch_root_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oz_batch_total_expansions.restype = C.c_int64
        L.oz_batch_total_expansions.argtypes = [C.c_void_p]
        L.oz_batch_total_simulations.restype = C.c_int64
        L.oz_batch_total_simulations.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def game_id(name):
    return lib().oz_game_lookup(name.encode())


def num_actions(gid):
    return lib().oz_num_actions(gid)


def state_bytes(gid):
    return lib().oz_state_bytes(gid)


def state_dim(gid):
    d = (C.c_int * 3)()
    lib().oz_state_dim(gid, d)
    return tuple(d)


def builtin_oracle(name):
    L = lib()
    return C.cast(getattr(L, {"uniform": "oz_uniform_oracle", "synth": "oz_synth_oracle"}[name]), C.c_void_p)


class RolloutCtx(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("gamma", C.c_double)]


class RolloutOracle:
    """MCTS.RolloutOracle(gspec, gamma) (src/mcts.jl:27-60) with the playout draws keyed by (seed, state, ply)."""

    def __init__(self, seed, gamma=1.0):
        self.ctx = RolloutCtx(seed, gamma)
        self.fn = C.cast(lib().oz_rollout_oracle, C.c_void_p)
        self.ctx_ptr = C.cast(C.pointer(self.ctx), C.c_void_p)


class GameEnv:
    """Mirror of a GI.AbstractGameEnv (src/game.jl) on top of the C oracle."""

    def __init__(self, gid, state=None):
        self.gid = gid
        self.g = Game()
        if state is None:
            lib().oz_game_init(C.byref(self.g), gid)
        else:
            self.set_state(state)

    def set_state(self, state):
        s = np.zeros(STATE_BYTES, np.uint8)
        st = np.frombuffer(bytes(state), np.uint8)
        s[:len(st)] = st
        lib().oz_game_set_state(C.byref(self.g), self.gid, s.ctypes.data)

    def clone(self):
        o = GameEnv.__new__(GameEnv)
        o.gid = self.gid
        o.g = Game()
        C.memmove(C.byref(o.g), C.byref(self.g), C.sizeof(Game))
        return o

    def state(self):
        s = np.zeros(STATE_BYTES, np.uint8)
        lib().oz_game_get_state(C.byref(self.g), s.ctypes.data)
        return bytes(s[:state_bytes(self.gid)])

    def terminated(self):
        return bool(lib().oz_game_terminated(C.byref(self.g)))

    def white_playing(self):
        return bool(lib().oz_game_white_playing(C.byref(self.g)))

    def actions_mask(self):
        m = np.zeros(MAX_ACTIONS, np.uint8)
        lib().oz_game_actions_mask(C.byref(self.g), m.ctypes.data)
        return m[:num_actions(self.gid)].astype(bool)

    def play(self, a, env_u=None):
        u = None if env_u is None else np.ascontiguousarray(env_u, np.float64).ctypes.data
        lib().oz_game_play(C.byref(self.g), int(a), u)

    def white_reward(self):
        return lib().oz_game_white_reward(C.byref(self.g))


def vectorize_state(gid, state):
    d = state_dim(gid)
    x = np.zeros(d[0] * d[1] * d[2], np.float32)
    s = np.zeros(STATE_BYTES, np.uint8)
    s[:len(state)] = np.frombuffer(bytes(state), np.uint8)
    lib().oz_vectorize_state(gid, s.ctypes.data, x.ctypes.data)
    return x.reshape(d, order="F")  # Julia column-major [w, h, c]


def state_key(gid, state):
    s = np.zeros(STATE_BYTES, np.uint8)
    s[:len(state)] = np.frombuffer(bytes(state), np.uint8)
    k = np.zeros(2, np.uint64)
    lib().oz_state_key(gid, s.ctypes.data, k.ctypes.data)
    return int(k[0]), int(k[1])


class Env:
    """MCTS.Env (src/mcts.jl:124-151)."""

    def __init__(self, gid, oracle="uniform", gamma=1.0, cpuct=1.0, noise_eps=0.0, noise_alpha=1.0, prior_temperature=1.0):
        self.gid = gid
        octx = None
        if isinstance(oracle, str):
            self._fn = builtin_oracle(oracle)
        elif isinstance(oracle, RolloutOracle):
            self._keep, self._fn, octx = oracle, oracle.fn, oracle.ctx_ptr
        else:  # python callable (state_bytes, n_legal) -> (P list, V)
            sb = state_bytes(gid)

            def cb(ctx, g, sp, n, P, V):
                p, v = oracle(bytes(sp[:sb]), n)
                for i in range(n):
                    P[i] = p[i]
                V[0] = v
            self._cb = ORACLE_FN(cb)
            self._fn = C.cast(self._cb, C.c_void_p)
        self.h = lib().oz_env_create(gid, self._fn, octx, gamma, cpuct, noise_eps, noise_alpha, prior_temperature)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oz_env_destroy(self.h)
            self.h = None

    def set_noise(self, seed, game, move):
        lib().oz_env_set_noise(self.h, seed, game, move)

    def explore(self, game, nsims, eta=None):
        e = None if eta is None else np.ascontiguousarray(eta, np.float64)
        lib().oz_explore(self.h, C.byref(game.g), nsims, None if e is None else e.ctypes.data)

    def root_stats(self, game):
        A = num_actions(self.gid)
        N = np.zeros(A, np.int64)
        W = np.zeros(A, np.float64)
        P = np.zeros(A, np.float32)
        V = np.zeros(1, np.float32)
        n = lib().oz_root_stats(self.h, C.byref(game.g), N.ctypes.data, W.ctypes.data, P.ctypes.data, V.ctypes.data)
        return n, N, W, P, float(V[0])

    def policy(self, game):
        acts = np.zeros(MAX_ACTIONS, np.int32)
        pi = np.zeros(MAX_ACTIONS, np.float64)
        n = lib().oz_policy(self.h, C.byref(game.g), acts.ctypes.data, pi.ctypes.data)
        return acts[:n].copy(), pi[:n].copy()

    def reset(self):
        lib().oz_env_reset(self.h)

    num_nodes = property(lambda self: lib().oz_env_num_nodes(self.h))
    total_simulations = property(lambda self: lib().oz_env_total_simulations(self.h))
    total_nodes_traversed = property(lambda self: lib().oz_env_total_nodes_traversed(self.h))


def worker_run(gid, oracle, mp, seed, first, stride, count, reset_every):
    """One worker of simulate() (src/simulations.jl:221-241); returns a list of trace dicts."""
    fn = builtin_oracle(oracle) if isinstance(oracle, str) else oracle
    arr = (Trace * count)()
    lib().oz_worker_run(gid, fn, None, C.byref(mp), seed, first, stride, count, reset_every, arr)
    A, sb = num_actions(gid), state_bytes(gid)
    out = []
    for tr in arr:
        n = tr.n_moves
        out.append(dict(
            n_moves=n,
            states=np.ctypeslib.as_array(tr.states)[:n + 1, :sb].copy(),
            pi=np.ctypeslib.as_array(tr.pi)[:n, :A].copy(),
            mask=np.ctypeslib.as_array(tr.mask)[:n, :A].copy(),
            action=np.ctypeslib.as_array(tr.action)[:n].copy(),
            rewards=np.ctypeslib.as_array(tr.rewards)[:n].copy(),
            z=np.ctypeslib.as_array(tr.z)[:n].copy(),
            t=np.ctypeslib.as_array(tr.t)[:n].copy(),
            mem_nodes=tr.mem_nodes, edepth=tr.edepth))
    return out


def random_positions(gid, seed, n, max_plies=30, first_stream=0):
    sb = state_bytes(gid)
    out = np.zeros((n, sb), np.uint8)
    buf = np.zeros(STATE_BYTES, np.uint8)
    for i in range(n):
        lib().oz_random_position(gid, seed, first_stream + i, max_plies, buf.ctypes.data)
        out[i] = buf[:sb]
    return out


def dirichlet(seed, game, move, n, alpha):
    eta = np.zeros(n, np.float64)
    lib().oz_dirichlet(seed, game, move, n, alpha, eta.ctypes.data)
    return eta


def set_threads(n):
    """Host threads of the batch driver (pthread parallel-for over independent trees)."""
    lib().oz_set_threads(int(n))


class Batch:
    """Lock-step batched driver (CPU baseline): explore nsims on n fixed roots with a batched evaluator."""

    def __init__(self, gid, n, mp):
        self.gid, self.n = gid, n
        self.h = lib().oz_batch_create(gid, n, C.byref(mp))
        self._ls = np.zeros((n, state_bytes(gid)), np.uint8)
        self._lt = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oz_batch_destroy(self.h)
            self.h = None

    def set_roots(self, states, eta=None):
        s = np.ascontiguousarray(states, np.uint8)
        e = None if eta is None else np.ascontiguousarray(eta, np.float64)
        lib().oz_batch_set_roots(self.h, s.ctypes.data, None if e is None else e.ctypes.data)

    def reset_trees(self):
        lib().oz_batch_reset_trees(self.h)

    def advance(self):
        k = lib().oz_batch_advance(self.h, self._ls.ctypes.data, self._lt.ctypes.data)
        return self._ls[:k], self._lt[:k]

    def vectorize(self, leaf_states, xshape):
        """(X [k, *xshape] float32, mask [k, A] bool) of the pending leaves returned by advance()."""
        k = len(leaf_states)
        A = num_actions(self.gid)
        X = np.zeros((k,) + tuple(xshape), np.float32)
        m = np.zeros((k, A), np.uint8)
        ls = np.ascontiguousarray(leaf_states, np.uint8)
        lib().oz_batch_vectorize(self.h, ls.ctypes.data, k, int(np.prod(xshape)), X.ctypes.data, m.ctypes.data)
        return X, m.astype(bool)

    def feed(self, P, V):
        P = np.ascontiguousarray(P, np.float32)
        V = np.ascontiguousarray(V, np.float32)
        lib().oz_batch_feed(self.h, P.ctypes.data, V.ctypes.data)

    def root_stats(self, i):
        A = num_actions(self.gid)
        N = np.zeros(A, np.int64)
        W = np.zeros(A, np.float64)
        P = np.zeros(A, np.float32)
        lib().oz_batch_root_stats(self.h, i, N.ctypes.data, W.ctypes.data, P.ctypes.data)
        return N, W, P

    expansions = property(lambda self: lib().oz_batch_total_expansions(self.h))
    simulations = property(lambda self: lib().oz_batch_total_simulations(self.h))
