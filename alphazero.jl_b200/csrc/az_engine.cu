// az_engine.cu -- tree-pool engine + C ABI (compiled with -fmad=false: bit-exact f64 PUCT / backup).
//
// Replaces, behind include/azb200.h:
//   MCTS.Env / explore! / policy / reset!                     src/mcts.jl:124-151,239-281
//   Batchifier inference server + worker tasks                 src/batchifier.jl:47-127, src/simulations.jl:23-155
//   simulate(): worker loop over games, reset_every            src/simulations.jl:207-244
// by a per-GPU "tick" scheduler:  select -> leaf batch -> oracle/network -> expand+backup -> per-move kernel.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_set>
#include <thread>
#include <vector>

#include "az_internal.h"
#include "az_tree.cuh"

static std::string g_create_err;

#define AZ_TRY(ctx, expr)                 \
  do {                                    \
    int s__ = (expr);                     \
    if (s__ != AZ_OK) return s__;         \
  } while (0)
#define AZ_FAIL(ctx, code, msg) \
  do {                          \
    (ctx)->err = (msg);         \
    return (code);              \
  } while (0)

template <class T>
static int az_dalloc(az_ctx* ctx, T** p, size_t n, bool zero = true) {
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc ") + std::to_string(n * sizeof(T)) + " B: " + cudaGetErrorString(e);
    cudaGetLastError();
    return AZ_ENOMEM;
  }
  if (zero) AZ_CUDA(ctx, cudaMemsetAsync(*p, 0, n * sizeof(T), ctx->stream));
  return AZ_OK;
}

// ------------------------------------------------------------------------------------------------
// game dispatch
// ------------------------------------------------------------------------------------------------
#define AZ_DISPATCH_GAME(game, F, ...)                     \
  switch (game) {                                          \
    case 0: return F<GameC4>(__VA_ARGS__);                 \
    case 1: return F<GameTTT>(__VA_ARGS__);                \
    case 2: return F<GameMancala>(__VA_ARGS__);            \
    case 3: return F<GameGW>(__VA_ARGS__);                 \
    default: return AZ_EINVAL;                             \
  }

static const char* AZ_GAME_NAMES[4] = {"connect-four", "tictactoe", "mancala", "grid-world"};

template <class G> static int g_num_actions() { return G::A; }
template <class G> static int g_state_bytes() { return G::STATE_BYTES; }
template <class G> static int g_max_plies() { return G::MAX_PLIES; }
template <class G> static int g_state_dim(int32_t* d) { d[0] = G::XW; d[1] = G::XH; d[2] = G::XC; return AZ_OK; }
template <class G> static int g_vectorize(const uint8_t* s, float* x) { G::vectorize(G::from_bytes(s), x); return AZ_OK; }
template <class G> static int g_heuristic(const uint8_t* s, double* v) { *v = G::heuristic_value(G::from_bytes(s)); return AZ_OK; }
// think(::MinMax.Player) on the host with the code the kernels run (az_minmax_qvalue / az_minmax_policy are host + device)
template <class G> static int g_minmax_think(const uint8_t* s, const az_minmax_params* mm, double* q, double* pi) {
  if (!G::HAS_HEURISTIC) return AZ_EUNSUPPORTED;
  if (mm->depth < 1 || mm->depth > AZ_MINMAX_MAX_DEPTH || !(mm->tau >= 0.0)) return AZ_EINVAL;
  const AzEnv e = G::from_bytes(s);
  if (G::terminated(e)) return AZ_EINVAL;
  const uint32_t legal = G::legal_mask(e);
  double qs[G::A], ps[G::A];
  int acts[G::A], n = 0;
  for (int a = 0; a < G::A; a++)
    if ((legal >> a) & 1u) { acts[n] = a; qs[n] = az_minmax_qvalue<G>(e, a, mm->depth, mm->amplify_rewards != 0, mm->gamma); n++; }
  az_minmax_policy(qs, n, mm->tau, ps);
  for (int a = 0; a < G::A; a++) { if (q) q[a] = 0.0; if (pi) pi[a] = 0.0; }
  for (int i = 0; i < n; i++) { if (q) q[acts[i]] = qs[i]; if (pi) pi[acts[i]] = ps[i]; }
  return AZ_OK;
}
template <class G> static int g_mask(const uint8_t* s, uint8_t* m) {
  AzEnv e = G::from_bytes(s);
  uint32_t l = G::terminated(e) ? G::legal_mask(e) : G::legal_mask(e);
  for (int i = 0; i < G::A; i++) m[i] = (l >> i) & 1;
  return AZ_OK;
}
template <class G> static int g_play(const uint8_t* s, int a, uint8_t* ns, int32_t* term, double* wr) {
  AzEnv e = G::from_bytes(s);
  if (a < 0 || a >= G::A || !((G::legal_mask(e) >> a) & 1) || G::terminated(e)) return AZ_EINVAL;
  AzEnv n = G::play(e, a);
  G::to_bytes(n, ns);
  if (term) *term = G::terminated(n);
  if (wr) *wr = G::white_reward(n);
  return AZ_OK;
}
template <class G> static int g_init_state(uint8_t* s) { G::to_bytes(G::init(), s); return AZ_OK; }
template <class G> static int g_random_positions(uint64_t seed, uint64_t first, int n, int max_plies, uint8_t* out) {
  if (G::STOCHASTIC) {  // random non-terminal start cells (RL.reset!, games/grid-world/game.jl:36)
    for (int i = 0; i < n; i++)
      for (uint64_t attempt = 0;; attempt++) {
        AzEnv e = G::init_game(seed, first + (uint64_t)i + (attempt << 32));
        if (!G::terminated(e)) { G::to_bytes(e, out + (size_t)i * G::STATE_BYTES); break; }
      }
    return AZ_OK;
  }
  for (int i = 0; i < n; i++) {
    for (uint64_t attempt = 0;; attempt++) {
      uint64_t st = first + (uint64_t)i + (attempt << 32);
      AzEnv e = G::init();
      uint32_t o[4];
      az_philox(seed, 0, AZ_PURPOSE_POSITION, (uint32_t)st, (uint32_t)(st >> 32), o);
      int k = (int)(o[0] % (uint32_t)(max_plies + 1));
      bool ok = true;
      for (int ply = 0; ply < k; ply++) {
        uint32_t legal = G::legal_mask(e);
        int acts[AZ_MAX_ACTIONS], na = 0;
        for (int a = 0; a < G::A; a++) if ((legal >> a) & 1) acts[na++] = a;
        az_philox(seed, (uint32_t)(ply + 1), AZ_PURPOSE_POSITION, (uint32_t)st, (uint32_t)(st >> 32), o);
        e = G::play(e, acts[o[0] % (uint32_t)na]);
        if (G::terminated(e)) { ok = false; break; }
      }
      if (ok) { G::to_bytes(e, out + (size_t)i * G::STATE_BYTES); break; }
    }
  }
  return AZ_OK;
}

// ------------------------------------------------------------------------------------------------
// built-in oracles as networks
// ------------------------------------------------------------------------------------------------
template <class G, int KIND>
struct OracleNet : az_net {
  int eval(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V) override {
    az_k_oracle<G, KIND><<<(max_rows + 127) / 128, 128, 0, ctx->stream>>>(envs, n_rows, P, V);
    ctx->launches++;
    return AZ_OK;
  }
};
// MCTS.RolloutOracle (src/mcts.jl:27-60): the "vanilla MCTS" baseline of the benchmark duels (Benchmark.MctsRollouts,
// src/benchmark.jl:134-147)
template <class G>
struct RolloutNet : az_net {
  double gamma = 1.0;
  uint64_t seed = 0;
  int eval(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V) override {
    az_k_rollout<G><<<(max_rows + 63) / 64, 64, 0, ctx->stream>>>(envs, n_rows, P, V, gamma, seed);
    ctx->launches++;
    return AZ_OK;
  }
};
template <class G> static int g_make_rollout(az_ctx* ctx, double gamma, uint64_t seed, az_net** out) {
  if (G::STOCHASTIC) AZ_FAIL(ctx, AZ_EUNSUPPORTED, "az_net_create_rollout: stochastic environments are not supported");
  RolloutNet<G>* n = new RolloutNet<G>();
  n->ctx = ctx; n->kind = AZ_NET_ROLLOUT; n->game = G::ID; n->gamma = gamma; n->seed = seed;
  *out = n;
  return AZ_OK;
}
template <class G> static int g_make_oracle(az_ctx* ctx, int kind, az_net** out) {
  az_net* n = nullptr;
  if (kind == AZ_NET_UNIFORM) n = new OracleNet<G, 0>();
  else if (kind == AZ_NET_SYNTH) n = new OracleNet<G, 1>();
  else AZ_FAIL(ctx, AZ_EINVAL, "az_net_create_oracle: kind must be AZ_NET_UNIFORM or AZ_NET_SYNTH");
  n->ctx = ctx; n->kind = kind; n->game = G::ID;
  *out = n;
  return AZ_OK;
}

// ------------------------------------------------------------------------------------------------
// MCTS pool
// ------------------------------------------------------------------------------------------------
struct az_mcts {
  az_ctx* ctx = nullptr;
  az_net* net = nullptr;
  az_net* net2 = nullptr;  // duel mode: oracle of the odd trees (player 1)
  int game = 0;
  virtual ~az_mcts() {}
  virtual int set_roots(const uint8_t* states, const double* eta) = 0;
  virtual int run(int nsims) = 0;
  virtual int root_stats(int64_t* N, double* W, float* P) = 0;
  virtual int policy(double* pi) = 0;
  virtual int reset() = 0;
  virtual int counters(int64_t* ts, int64_t* tn, int64_t* nn) = 0;
  virtual int set_noise(uint64_t seed, const int64_t* games, const int32_t* moves) = 0;
  double ms_total = 0, ms_net = 0;
  int64_t ticks = 0, expansions = 0;
  // per-kernel device timing of the tree kernels (bench.py's roofline_tree): see Mcts::tick_profiled
  virtual int set_profiling(int enable) = 0;
  virtual void drain_profile() = 0;
  double prof_select_ms = 0, prof_expand_ms = 0, prof_net_ms = 0;
  int64_t prof_ticks = 0;
};

template <class G>
struct Mcts : az_mcts {
  AzPool p{};
  az_mcts_params mp{};
  size_t cap = 0;
  std::vector<void*> allocs;
  int64_t* d_N = nullptr; double* d_W = nullptr; float* d_P = nullptr;
  int32_t* h_pin = nullptr;  // pinned: [0] n_leaves, [1..4] flags
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool profile_net = false;

  template <class T> int alloc(T** ptr, size_t n) {
    int s = az_dalloc(ctx, ptr, n);
    if (s == AZ_OK) allocs.push_back(*ptr);
    return s;
  }
  int create(az_ctx* c, az_net* n, const az_mcts_params* params, int S, int cap_nodes, az_net* n2 = nullptr, const az_mcts_params* params1 = nullptr) {
    ctx = c; net = n; net2 = n2; game = G::ID; mp = *params;
    if (n2 && (S & 1)) AZ_FAIL(ctx, AZ_EINVAL, "duel pool: the number of trees must be even");
    if (n2 && n2->game != G::ID) AZ_FAIL(ctx, AZ_EINVAL, "az_mcts_create: second oracle was built for another game");
    p.duel = n2 ? 1 : 0; p.row_base1 = n2 ? S / 2 : 0;
    if (S <= 0 || cap_nodes <= 0) AZ_FAIL(ctx, AZ_EINVAL, "az_mcts_create: n_trees and capacity must be positive");
    if (params->num_iters_per_turn <= 0) AZ_FAIL(ctx, AZ_EINVAL, "MctsPlayer: niters > 0 (src/play.jl:162)");
    if (n->game != G::ID) AZ_FAIL(ctx, AZ_EINVAL, "az_mcts_create: oracle was built for another game");
    cap = 64;
    while (cap < (size_t)cap_nodes + (size_t)cap_nodes / 4 + 8) cap <<= 1;
    p.S = S; p.cap_mask = (uint32_t)(cap - 1); p.maxd = G::MAX_PLIES + 1;
    p.max_sims_per_call = 1;
    if (const char* e = getenv("AZ_NO_GRAPH")) use_graph = !(e[0] == '1');
    if (const char* e = getenv("AZ_FUSE_TREE")) fuse_tree = !(e[0] == '0');
    if (const char* e = getenv("AZ_DEVICE_LOOP")) use_device_loop = !(e[0] == '0');
    if (const char* e = getenv("AZ_MAX_SIMS_PER_CALL")) p.max_sims_per_call = std::max(1, atoi(e));
    p.c.gamma = params->gamma; p.c.cpuct = params->cpuct; p.c.eps = params->dirichlet_noise_eps;
    p.c.alpha = params->dirichlet_noise_alpha; p.c.prior_temp = params->prior_temperature;
    const az_mcts_params* q1 = params1 ? params1 : params;   // duel: player 1's MctsPlayer may differ (src/benchmark.jl:78-99)
    p.c1.gamma = q1->gamma; p.c1.cpuct = q1->cpuct; p.c1.eps = q1->dirichlet_noise_eps;
    p.c1.alpha = q1->dirichlet_noise_alpha; p.c1.prior_temp = q1->prior_temperature;
    AZ_TRY(ctx, alloc(&p.nodes, (size_t)S * cap * G::LANES));
    AZ_TRY(ctx, alloc(&p.root, S)); AZ_TRY(ctx, alloc(&p.tag, S)); AZ_TRY(ctx, alloc(&p.node_count, S));
    AZ_TRY(ctx, alloc(&p.total_sims, S)); AZ_TRY(ctx, alloc(&p.total_nodes, S)); AZ_TRY(ctx, alloc(&p.sims_done, S));
    AZ_TRY(ctx, alloc(&p.sims_target, S)); AZ_TRY(ctx, alloc(&p.status, S)); AZ_TRY(ctx, alloc(&p.eta, (size_t)S * G::A));
    AZ_TRY(ctx, alloc(&p.pending, S)); AZ_TRY(ctx, alloc(&p.leaf_pos, S)); AZ_TRY(ctx, alloc(&p.leaf_env, S));
    AZ_TRY(ctx, alloc(&p.leaf_row, S)); AZ_TRY(ctx, alloc(&p.depth, S));
    AZ_TRY(ctx, alloc(&p.path_node, (size_t)S * p.maxd)); AZ_TRY(ctx, alloc(&p.path_meta, (size_t)S * p.maxd));
    AZ_TRY(ctx, alloc(&p.path_r, (size_t)S * p.maxd));
    AZ_TRY(ctx, alloc(&p.n_leaves, 4)); AZ_TRY(ctx, alloc(&p.batch_env, S));
    AZ_TRY(ctx, alloc(&p.batch_P, (size_t)S * G::A)); AZ_TRY(ctx, alloc(&p.batch_V, S));
    AZ_TRY(ctx, alloc(&p.flags, 4)); AZ_TRY(ctx, alloc(&p.expansions, 1)); AZ_TRY(ctx, alloc(&d_tick_count, 1));
    AZ_TRY(ctx, alloc(&p.noise_game, S)); AZ_TRY(ctx, alloc(&p.noise_move, S));
    AZ_TRY(ctx, alloc(&d_N, (size_t)S * G::A)); AZ_TRY(ctx, alloc(&d_W, (size_t)S * G::A)); AZ_TRY(ctx, alloc(&d_P, (size_t)S * G::A));
    std::vector<uint32_t> tags(S, 64u | 1u);
    AZ_CUDA(ctx, cudaMemcpyAsync(p.tag, tags.data(), S * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    AZ_CUDA(ctx, cudaMallocHost((void**)&h_pin, 64));
    for (auto& e : ev) AZ_CUDA(ctx, cudaEventCreate(&e));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AZ_OK;
  }
  ~Mcts() override {
    for (void* q : allocs) cudaFree(q);
    if (h_pin) cudaFreeHost(h_pin);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : tev) if (e) cudaEventDestroy(e);
    drop_graph();
    drop_loop_graph();
  }
  int groups_grid() const { return (int)(((size_t)p.S * 32 + 127) / 128); }  // one warp per tree

  int set_roots(const uint8_t* states, const double* eta) override {
    if (!eta && p.c.eps != 0.0) AZ_FAIL(ctx, AZ_EINVAL, "az_mcts_set_roots: eta is required when dirichlet_noise_eps != 0");
    std::vector<AzEnv> r(p.S);
    for (int i = 0; i < p.S; i++) r[i] = G::from_bytes(states + (size_t)i * G::STATE_BYTES);
    AZ_CUDA(ctx, cudaMemcpyAsync(p.root, r.data(), p.S * sizeof(AzEnv), cudaMemcpyHostToDevice, ctx->stream));
    if (eta) AZ_CUDA(ctx, cudaMemcpyAsync(p.eta, eta, (size_t)p.S * G::A * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // r is a stack-owned staging buffer
    return AZ_OK;
  }
  // ---- CUDA-graph replay of a tick (optionally followed by an extra kernel, e.g. the self-play move kernel) ----
  cudaGraphExec_t gexec = nullptr;
  int64_t graph_launches = 0;   // kernels per replay
  uint64_t graph_net_gen = 0;
  bool use_graph = true, graph_broken = false;
  bool run_mode = false;        // inside az_mcts_run (explore-only ticks)
  // device-driven explore loop: ONE graph launch = a WHILE node whose body is a tick + the loop-condition kernel
  bool use_device_loop = true;  // AZ_DEVICE_LOOP=0: the host replays the tick graph and polls the flags
  bool device_loop_broken = false;
  cudaGraphExec_t lexec = nullptr;
  uint64_t loop_net_gen = 0;
  int loop_max_ticks = 0;
  int64_t loop_launches = 0;    // kernels per iteration of the WHILE body
  int32_t* d_tick_count = nullptr;
  void drop_loop_graph() { if (lexec) { cudaGraphExecDestroy(lexec); lexec = nullptr; } }
  bool graph_is_fused = false;  // which tick body the captured graph holds
  void drop_graph() { if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; } }
  template <class Extra>
  int tick_graphed(Extra extra) {
    // (re)allocations must happen before the capture begins (cudaMalloc / cudaFree / stream sync are illegal inside it)
    AZ_TRY(ctx, net->reserve(net2 ? p.row_base1 : p.S));
    if (net2) AZ_TRY(ctx, net2->reserve(p.row_base1));
    if (!use_graph || graph_broken || !net->capturable() || (net2 && !net2->capturable())) {
      drop_graph();
      AZ_TRY(ctx, tick(false));
      extra();
      return AZ_OK;
    }
    const uint64_t gen_now = net->generation() + (net2 ? 0x100000000ull * net2->generation() : 0);
    if (gexec && graph_net_gen != gen_now) drop_graph();  // a network reallocated buffers / reloaded weights
    if (gexec && graph_is_fused != (run_mode && fuse_tree)) drop_graph();
    graph_is_fused = run_mode && fuse_tree;
    if (!gexec) {
      graph_net_gen = gen_now;
      const int64_t l0 = ctx->launches;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { graph_broken = true; cudaGetLastError(); return tick_graphed(extra); }
      int st = tick(false);
      extra();
      cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
      if (st != AZ_OK || e != cudaSuccess || cudaGraphInstantiate(&gexec, g, 0) != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        gexec = nullptr; graph_broken = true; cudaGetLastError();
        ctx->launches = l0;
        return tick_graphed(extra);
      }
      cudaGraphDestroy(g);
      graph_launches = ctx->launches - l0;
      ctx->launches = l0;
    }
    AZ_CUDA(ctx, cudaGraphLaunch(gexec, ctx->stream));
    ctx->launches += graph_launches;
    return AZ_OK;
  }
  // one tick: select -> oracle -> expand+backup
  // ---- profiled ticks: CUDA events around select, the network and expand+backup on the context's stream (graph replay off);
  //      the event ring is drained whenever it fills so every tick of a pass is counted ----
  static constexpr int TPROF_SLOTS = 1024;
  bool profiling = false;
  std::vector<cudaEvent_t> tev;
  int tprof_n = 0;
  void tprof_drain() {
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < tprof_n; i++) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, tev[i * 4 + 0], tev[i * 4 + 1]);
      cudaEventElapsedTime(&b, tev[i * 4 + 1], tev[i * 4 + 2]);
      cudaEventElapsedTime(&c, tev[i * 4 + 2], tev[i * 4 + 3]);
      prof_select_ms += a; prof_net_ms += b; prof_expand_ms += c;
    }
    prof_ticks += tprof_n;
    tprof_n = 0;
  }
  void drain_profile() override { if (profiling) tprof_drain(); }
  int set_profiling(int enable) override {
    if (enable && tev.empty()) {
      tev.resize((size_t)TPROF_SLOTS * 4);
      for (auto& e : tev) AZ_CUDA(ctx, cudaEventCreate(&e));
    }
    if (profiling) tprof_drain();
    profiling = enable != 0;
    if (enable) { prof_select_ms = prof_expand_ms = prof_net_ms = 0; prof_ticks = 0; tprof_n = 0; }
    return AZ_OK;
  }
  int tick_profiled() {
    if (tprof_n == TPROF_SLOTS) tprof_drain();
    cudaEvent_t* e = &tev[(size_t)tprof_n * 4];
    AZ_TRY(ctx, net->reserve(net2 ? p.row_base1 : p.S));
    AZ_CUDA(ctx, cudaMemsetAsync(p.n_leaves, 0, 4 * sizeof(int32_t), ctx->stream));
    cudaEventRecord(e[0], ctx->stream);
    az_k_select<G><<<groups_grid(), 128, 0, ctx->stream>>>(p);
    cudaEventRecord(e[1], ctx->stream);
    if (!net2) {
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, p.S, p.batch_P, p.batch_V));
    } else {
      const int b1 = p.row_base1;
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, b1, p.batch_P, p.batch_V));
      AZ_TRY(ctx, net2->eval(p.batch_env + b1, p.n_leaves + 2, b1, p.batch_P + (size_t)b1 * G::A, p.batch_V + b1));
    }
    cudaEventRecord(e[2], ctx->stream);
    az_k_expand_backup<G><<<groups_grid(), 128, 0, ctx->stream>>>(p);
    cudaEventRecord(e[3], ctx->stream);
    ctx->launches += 2;
    tprof_n++;
    return AZ_OK;
  }
  // explore-only tick (az_mcts_run): [expand + backup of the previous tick's leaves, fused with the next select] -> oracle.
  // On the first tick nothing is pending, so the kernel is just the select; the run ends on the tick whose select finds
  // no tree with simulations left, i.e. after the last leaves have been expanded.
  bool fuse_tree = true;   // AZ_FUSE_TREE=0: separate select / expand kernels (the self-play order)
  int tick_fused() {
    AZ_CUDA(ctx, cudaMemsetAsync(p.n_leaves, 0, 4 * sizeof(int32_t), ctx->stream));
    az_k_expand_select<G><<<groups_grid(), 128, 0, ctx->stream>>>(p);
    if (!net2) {
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, p.S, p.batch_P, p.batch_V));
    } else {
      const int b1 = p.row_base1;
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, b1, p.batch_P, p.batch_V));
      AZ_TRY(ctx, net2->eval(p.batch_env + b1, p.n_leaves + 2, b1, p.batch_P + (size_t)b1 * G::A, p.batch_V + b1));
    }
    ctx->launches += 1;
    return AZ_OK;
  }
  int tick(bool time_net) {
    if (run_mode && fuse_tree) return tick_fused();
    AZ_CUDA(ctx, cudaMemsetAsync(p.n_leaves, 0, 4 * sizeof(int32_t), ctx->stream));
    az_k_select<G><<<groups_grid(), 128, 0, ctx->stream>>>(p);
    if (time_net) cudaEventRecord(ev[2], ctx->stream);
    if (!net2) {
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, p.S, p.batch_P, p.batch_V));
    } else {  // one leaf queue per player's oracle (at most one tree per worker is thinking)
      const int b1 = p.row_base1;
      AZ_TRY(ctx, net->eval(p.batch_env, p.n_leaves, b1, p.batch_P, p.batch_V));
      AZ_TRY(ctx, net2->eval(p.batch_env + b1, p.n_leaves + 2, b1, p.batch_P + (size_t)b1 * G::A, p.batch_V + b1));
    }
    if (time_net) cudaEventRecord(ev[3], ctx->stream);
    az_k_expand_backup<G><<<groups_grid(), 128, 0, ctx->stream>>>(p);
    ctx->launches += 2;
    return AZ_OK;
  }
  int check_flags() {
    AZ_CUDA(ctx, cudaMemcpyAsync(h_pin + 8, p.n_leaves, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    AZ_CUDA(ctx, cudaMemcpyAsync(h_pin + 1, p.flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_pin[1]) AZ_FAIL(ctx, AZ_ENOMEM, "MCTS table overflow: raise capacity_nodes_per_tree");
    if (h_pin[3]) AZ_FAIL(ctx, AZ_ESTATE, "simulation path exceeded the per-game ply bound");
    if (h_pin[4]) AZ_FAIL(ctx, AZ_ESTATE, "MCTS.policy: explore! must be called before policy (src/mcts.jl:262)");
    return AZ_OK;
  }
  // Build (once per network generation / tick budget) and launch the WHILE graph.  Returns AZ_OK, or AZ_EUNSUPPORTED when
  // the driver refuses the construction (then the caller falls back to the host-driven loop for good).
  int run_device_loop(int max_ticks) {
    AZ_TRY(ctx, net->reserve(net2 ? p.row_base1 : p.S));
    if (net2) AZ_TRY(ctx, net2->reserve(p.row_base1));
    const uint64_t gen_now = net->generation() + (net2 ? 0x100000000ull * net2->generation() : 0);
    if (lexec && (loop_net_gen != gen_now || loop_max_ticks != max_ticks)) drop_loop_graph();
    if (!lexec) {
      cudaGraph_t g = nullptr, body = nullptr;
      cudaGraphConditionalHandle h;
      bool ok = cudaGraphCreate(&g, 0) == cudaSuccess && cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
      cudaGraphNode_t cnode;
      if (ok) {
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = h;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        ok = cudaGraphAddNode(&cnode, g, nullptr, 0, &cp) == cudaSuccess;
        if (ok) body = cp.conditional.phGraph_out[0];
      }
      const int64_t l0 = ctx->launches;
      int st = AZ_OK;
      if (ok) ok = cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        st = tick(false);
        az_k_loop_cond<<<1, 1, 0, ctx->stream>>>(p, h, d_tick_count, max_ticks);
        ok = cudaStreamEndCapture(ctx->stream, nullptr) == cudaSuccess && st == AZ_OK;
      }
      loop_launches = ctx->launches - l0 + 1;
      ctx->launches = l0;
      if (ok) ok = cudaGraphInstantiate(&lexec, g, 0) == cudaSuccess;
      if (g) cudaGraphDestroy(g);
      if (!ok) { lexec = nullptr; cudaGetLastError(); return AZ_EUNSUPPORTED; }
      loop_net_gen = gen_now; loop_max_ticks = max_ticks;
    }
    AZ_CUDA(ctx, cudaMemsetAsync(d_tick_count, 0, sizeof(int32_t), ctx->stream));
    AZ_CUDA(ctx, cudaGraphLaunch(lexec, ctx->stream));
    int32_t t = 0;
    AZ_CUDA(ctx, cudaMemcpyAsync(&t, d_tick_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    AZ_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ticks = t;
    ctx->launches += (int64_t)t * loop_launches;
    return AZ_OK;
  }
  int run(int nsims) override {
    if (nsims <= 0) AZ_FAIL(ctx, AZ_EINVAL, "az_mcts_run: nsims must be positive");
    std::vector<int32_t> tgt(p.S, nsims);
    std::vector<uint8_t> st(p.S, 1);
    AZ_CUDA(ctx, cudaMemsetAsync(p.sims_done, 0, p.S * sizeof(int32_t), ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(p.expansions, 0, sizeof(int64_t), ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(p.flags, 0, 4 * sizeof(int32_t), ctx->stream));  // an earlier failed call must not poison this one
    AZ_CUDA(ctx, cudaMemcpyAsync(p.sims_target, tgt.data(), p.S * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    AZ_CUDA(ctx, cudaMemcpyAsync(p.status, st.data(), p.S, cudaMemcpyHostToDevice, ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AZ_CUDA(ctx, cudaEventRecord(ev[0], ctx->stream));
    ticks = 0; ms_net = 0;
    run_mode = true;
    struct RunModeGuard { bool& f; ~RunModeGuard() { f = false; } } run_mode_guard{run_mode};
    // every unfinished tree completes >= 1 simulation per tick, so nsims + 1 ticks always suffice
    bool looped = false;
    if (use_device_loop && !device_loop_broken && use_graph && !profiling && fuse_tree && net->capturable() && (!net2 || net2->capturable())) {
      const int st = run_device_loop(nsims + 1);
      if (st == AZ_OK) looped = true;
      else if (st == AZ_EUNSUPPORTED) device_loop_broken = true;
      else return st;
    }
    for (int t = 0; !looped && t < nsims + 1; t++) {
      if (profiling) AZ_TRY(ctx, tick_profiled());
      else AZ_TRY(ctx, tick_graphed([] {}));
      ticks++;
      if (t + 1 >= nsims / 2 && ((t + 1) % 8 == 0 || t + 1 >= nsims)) {
        AZ_TRY(ctx, check_flags());
        if (h_pin[8] == 0 && h_pin[9] == 0 && h_pin[10] == 0) break;  // no leaf pending and no tree with simulations left
      }
    }
    if (!looped) AZ_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    AZ_TRY(ctx, check_flags());
    int64_t ex = 0;
    AZ_CUDA(ctx, cudaMemcpy(&ex, p.expansions, sizeof(int64_t), cudaMemcpyDeviceToHost));
    expansions = ex;
    float ms = 0;
    AZ_CUDA(ctx, cudaEventElapsedTime(&ms, ev[0], ev[1]));
    ms_total = ms;
    return AZ_OK;
  }
  int root_stats(int64_t* N, double* W, float* P) override {
    az_k_root_stats<G><<<groups_grid(), 128, 0, ctx->stream>>>(p, d_N, d_W, d_P);
    ctx->launches++;
    size_t n = (size_t)p.S * G::A;
    if (N) AZ_CUDA(ctx, cudaMemcpyAsync(N, d_N, n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (W) AZ_CUDA(ctx, cudaMemcpyAsync(W, d_W, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (P) AZ_CUDA(ctx, cudaMemcpyAsync(P, d_P, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AZ_OK;
  }
  int policy(double* pi) override {  // src/mcts.jl:255-271 (host arithmetic on the fetched integer counts)
    std::vector<int64_t> N((size_t)p.S * G::A);
    AZ_TRY(ctx, root_stats(N.data(), nullptr, nullptr));
    for (int s = 0; s < p.S; s++) {
      int64_t ntot = 0;
      for (int a = 0; a < G::A; a++) ntot += N[(size_t)s * G::A + a];
      if (ntot == 0) AZ_FAIL(ctx, AZ_ESTATE, "MCTS.explore! must be called before MCTS.policy (src/mcts.jl:262)");
      double sum = 0.0; bool first = true;
      for (int a = 0; a < G::A; a++) {
        double v = (double)N[(size_t)s * G::A + a] / (double)ntot;
        pi[(size_t)s * G::A + a] = v;
        if (N[(size_t)s * G::A + a] || true) { sum = first ? v : sum + v; first = false; }
      }
      for (int a = 0; a < G::A; a++) pi[(size_t)s * G::A + a] = pi[(size_t)s * G::A + a] / sum;
    }
    return AZ_OK;
  }
  int reset() override {
    az_k_reset<G><<<p.S, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    AZ_CUDA(ctx, cudaMemsetAsync(p.flags, 0, 4 * sizeof(int32_t), ctx->stream));
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AZ_OK;
  }
  int set_noise(uint64_t seed, const int64_t* games, const int32_t* moves) override {
    p.noise_seed = seed;
    drop_graph();  // the seed is a by-value kernel argument
    AZ_CUDA(ctx, cudaMemcpy(p.noise_game, games, p.S * sizeof(int64_t), cudaMemcpyHostToDevice));
    AZ_CUDA(ctx, cudaMemcpy(p.noise_move, moves, p.S * sizeof(int32_t), cudaMemcpyHostToDevice));
    return AZ_OK;
  }
  int counters(int64_t* ts, int64_t* tn, int64_t* nn) override {
    if (ts) AZ_CUDA(ctx, cudaMemcpy(ts, p.total_sims, p.S * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (tn) AZ_CUDA(ctx, cudaMemcpy(tn, p.total_nodes, p.S * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (nn) {
      std::vector<int32_t> c(p.S);
      AZ_CUDA(ctx, cudaMemcpy(c.data(), p.node_count, p.S * sizeof(int32_t), cudaMemcpyDeviceToHost));
      for (int i = 0; i < p.S; i++) nn[i] = c[i];
    }
    return AZ_OK;
  }
};

template <class G>
static int g_make_mcts(az_ctx* ctx, az_net* net, const az_mcts_params* p, int S, int cap, az_mcts** out) {
  auto* m = new Mcts<G>();
  int s = m->create(ctx, net, p, S, cap);
  if (s != AZ_OK) { m->ctx = nullptr; delete m; return s; }
  *out = m;
  return AZ_OK;
}

// ------------------------------------------------------------------------------------------------
// self-play
// ------------------------------------------------------------------------------------------------
struct az_selfplay {
  az_ctx* ctx = nullptr;
  int game = 0;
  virtual ~az_selfplay() {}
  virtual int start(int num_games, int64_t first) = 0;
  virtual int poll(int32_t* done, int32_t* fin) = 0;
  virtual int wait() = 0;
  virtual int counts(int64_t* ns, int64_t* ng) = 0;
  virtual int fetch(uint8_t* states, float* pi, uint8_t* mask, float* z, float* t, int32_t* gos, double* rew, int32_t* act) = 0;
  virtual int stats(double* ed, int64_t* nodes, int32_t* moves, double* totals) = 0;
  virtual int outcomes(double gamma, double* rewards, int32_t* colors_flipped, uint8_t* final_states, double* redundancy) = 0;
  virtual int export_samples(az_samples** out) = 0;
};

template <class G>
struct SelfPlay : az_selfplay {
  std::unique_ptr<Mcts<G>> pool;
  AzSelfPlay sp{};
  az_sim_params simp{};
  std::vector<void*> allocs;
  int games_cap = 0;
  std::thread worker;
  std::atomic<int> a_done{0}, a_finished{1}, a_status{AZ_OK};
  std::string werr;
  double seconds = 0;
  int64_t total_expansions = 0;
  int32_t* h_pin = nullptr;
  std::vector<int32_t> h_moves;

  template <class T> int alloc(T** ptr, size_t n) {
    int s = az_dalloc(ctx, ptr, n);
    if (s == AZ_OK) allocs.push_back(*ptr);
    return s;
  }
  int create(az_ctx* c, az_net* net, az_net* net2, const az_mcts_params* mp_in, const az_sim_params* s, uint64_t seed, const az_mcts_params* mp1_in = nullptr,
             const az_minmax_params* mm = nullptr) {
    if (!mp1_in) mp1_in = mp_in;
    if (mm) {
      // TwoPlayers(player, MinMax.Player) (Benchmark.Duel against Benchmark.MinMaxTS, src/benchmark.jl:178-196): the second
      // tree of every worker is only a seat (root state, turn flags) -- no table entries, no oracle calls; net2 is a stand-in
      if (!G::HAS_HEURISTIC) { c->err = "MinMax player: the game defines no heuristic_value for a two-player search"; return AZ_EUNSUPPORTED; }
      if (mm->depth < 1 || mm->depth > AZ_MINMAX_MAX_DEPTH) { c->err = "MinMax player: 1 <= depth <= 8"; return AZ_EINVAL; }
      if (!(mm->tau >= 0.0)) { c->err = "MinMax player: tau >= 0"; return AZ_EINVAL; }
      sp.minmax1 = 1; sp.mm_depth = mm->depth; sp.mm_amplify = mm->amplify_rewards ? 1 : 0; sp.mm_tau = mm->tau; sp.mm_gamma = mm->gamma;
    }
    // num_iters_per_turn == 0: NetworkPlayer under PlayerWithTemperature (Benchmark.NetworkOnly, src/benchmark.jl:166-176,
    // src/play.jl:226-235): the engine runs ONE simulation per turn -- on a root that is not in the table yet that is exactly
    // the oracle call think() makes -- and az_k_move reads the move distribution from the root's priors instead of the visit
    // counts.  The search parameters of such a player are inert; the priors must be the raw network output.
    az_mcts_params mp_l = *mp_in, mp1_l = *mp1_in;
    const az_mcts_params *mp = &mp_l, *mp1 = &mp1_l;
    for (az_mcts_params* q : {&mp_l, &mp1_l}) {
      if (q->num_iters_per_turn < 0) { c->err = "MctsPlayer: niters > 0 (src/play.jl:162); 0 selects NetworkPlayer"; return AZ_EINVAL; }
      if (q->num_iters_per_turn == 0) { q->num_iters_per_turn = 1; q->prior_temperature = 1.0; q->dirichlet_noise_eps = 0.0; }
    }
    sp.netonly = mp_in->num_iters_per_turn == 0; sp.netonly1 = !mm && mp1_in->num_iters_per_turn == 0;
    if (mm) mp1_l.gamma = mm->gamma;   // the z rows of a game in which the MinMax player held white
    if (mp1->temperature_n < 1 || mp1->temperature_n > AZ_MAX_SCHEDULE) { c->err = "temperature schedule: 1..8 points"; return AZ_EINVAL; }
    ctx = c; game = G::ID; simp = *s;
    if (s->num_workers <= 0) AZ_FAIL(ctx, AZ_EINVAL, "SimParams: num_workers must be positive");
    if (s->batch_size > s->num_workers) AZ_FAIL(ctx, AZ_EINVAL, "batch_size <= num_workers (src/batchifier.jl:48)");
    if (s->flip_probability != 0.0 && G::NSYM == 0)
      AZ_FAIL(ctx, AZ_EINVAL, "You must specify some game symmetries to use flip_probability>0. (src/params.jl:377-381)");
    if (!(s->flip_probability >= 0.0 && s->flip_probability <= 1.0)) AZ_FAIL(ctx, AZ_EINVAL, "flip_probability must be in [0, 1]");
    if (mp->temperature_n < 1 || mp->temperature_n > AZ_MAX_SCHEDULE) AZ_FAIL(ctx, AZ_EINVAL, "temperature schedule: 1..8 points");
    const int W = s->num_workers;
    const int S = net2 ? 2 * W : W;  // duel: one tree per player per worker (TwoPlayers, src/play.jl:248-252)
    int reset = s->reset_every > 0 ? s->reset_every : std::max(1, (s->num_games + W - 1) / W);
    size_t bound = (size_t)std::min<long long>((long long)std::max(mp->num_iters_per_turn, mp1->num_iters_per_turn) * G::MAX_PLIES * (long long)reset, G::MAX_STATES);
    // table budget: 60 % of the memory that is free right now; Mcts::create rounds (1.25 x cap_nodes) up to a power of
    // two, so ask for at most the largest power-of-two capacity that fits the budget, less the 1.25 head room
    size_t free_b = 0, total_b = 0;
    AZ_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = free_b / 10 * 6;
    size_t cap_fit = 64;
    while (cap_fit * 2 * (size_t)S * G::LANES * 16 <= budget) cap_fit <<= 1;
    if (cap_fit * (size_t)S * G::LANES * 16 > budget) AZ_FAIL(ctx, AZ_ENOMEM, "self-play: not enough free device memory for the tree tables");
    const size_t maxnodes = cap_fit * 4 / 5 - 8;
    int cap_nodes = (int)std::min<size_t>(std::min(bound, maxnodes), (size_t)1 << 28);
    pool.reset(new Mcts<G>());
    int st = pool->create(ctx, net, mp, S, cap_nodes, net2, mp1);
    if (st != AZ_OK) { pool.reset(); return st; }
    pool->p.noise_seed = seed;
    sp.W = W; sp.duel = net2 ? 1 : 0; sp.alternate = s->alternate_colors ? 1 : 0; sp.flip_p = s->flip_probability;
    sp.seed = seed; sp.nsims = mp->num_iters_per_turn; sp.reset_every = s->reset_every; sp.max_plies = G::MAX_PLIES;
    sp.sched_n = mp->temperature_n;
    for (int i = 0; i < mp->temperature_n; i++) { sp.sched_xs[i] = mp->temperature_xs[i]; sp.sched_ys[i] = mp->temperature_ys[i]; }
    sp.nsims1 = mp1->num_iters_per_turn; sp.sched1_n = mp1->temperature_n;
    for (int i = 0; i < mp1->temperature_n; i++) { sp.sched1_xs[i] = mp1->temperature_xs[i]; sp.sched1_ys[i] = mp1->temperature_ys[i]; }
    AZ_TRY(ctx, alloc(&sp.game_of_slot, W)); AZ_TRY(ctx, alloc(&sp.move_of_slot, W)); AZ_TRY(ctx, alloc(&sp.games_on_slot, W));
    AZ_TRY(ctx, alloc(&sp.games_done, 1)); AZ_TRY(ctx, alloc(&sp.active_slots, 1));
    AZ_TRY(ctx, alloc(&sp.next_game, 1)); AZ_TRY(ctx, alloc(&sp.want_game, W));
    AZ_CUDA(ctx, cudaMallocHost((void**)&h_pin, 64));
    return AZ_OK;
  }
  ~SelfPlay() override {
    if (worker.joinable()) worker.join();
    for (void* q : allocs) cudaFree(q);
    for (void* q : game_allocs) cudaFree(q);
    if (h_pin) cudaFreeHost(h_pin);
    if (h_progress) cudaFreeHost(h_progress);
  }
  std::vector<void*> game_allocs;
  template <class T> int galloc(T** ptr, size_t n) {
    int s = az_dalloc(ctx, ptr, n);
    if (s == AZ_OK) game_allocs.push_back(*ptr);
    return s;
  }
  int ensure_game_buffers(int ng) {
    if (ng <= games_cap) return AZ_OK;
    for (void* q : game_allocs) cudaFree(q);
    game_allocs.clear();
    size_t rows = (size_t)ng * G::MAX_PLIES;
    AZ_TRY(ctx, galloc(&sp.s_env, rows)); AZ_TRY(ctx, galloc(&sp.s_root, rows)); AZ_TRY(ctx, galloc(&sp.g_final, ng));
    AZ_TRY(ctx, galloc(&sp.s_pi, rows * G::A)); AZ_TRY(ctx, galloc(&sp.s_action, rows));
    AZ_TRY(ctx, galloc(&sp.s_reward, rows)); AZ_TRY(ctx, galloc(&sp.s_z, rows)); AZ_TRY(ctx, galloc(&sp.s_t, rows));
    AZ_TRY(ctx, galloc(&sp.g_moves, ng)); AZ_TRY(ctx, galloc(&sp.g_edepth, ng)); AZ_TRY(ctx, galloc(&sp.g_nodes, ng));
    games_cap = ng;
    return AZ_OK;
  }
  // the per-move kernel, preceded by the MinMax player's search when the baseline is one (its trees move one tick after
  // their turn began; the q-values travel in p.eta)
  void launch_move(Mcts<G>& m, int grid1) {
    if (sp.minmax1) az_k_minmax_think<G><<<(m.p.S * G::A + 127) / 128, 128, 0, ctx->stream>>>(m.p, sp);
    az_k_move<G><<<grid1, 128, 0, ctx->stream>>>(m.p, sp, 0);
  }
  // WHILE graph whose body is one self-play tick: select -> oracle(s) -> expand + backup -> move -> assign -> loop condition
  int32_t* h_progress = nullptr;   // pinned, device-visible: [0] games finished
  int run_device_loop(Mcts<G>& m, int grid1) {
    if (!h_progress) AZ_CUDA(ctx, cudaHostAlloc((void**)&h_progress, 64, cudaHostAllocMapped));
    h_progress[0] = 0;
    AZ_TRY(ctx, m.net->reserve(m.net2 ? m.p.row_base1 : m.p.S));
    if (m.net2) AZ_TRY(ctx, m.net2->reserve(m.p.row_base1));
    cudaGraph_t g = nullptr, body = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphConditionalHandle h;
    bool ok = cudaGraphCreate(&g, 0) == cudaSuccess && cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) {
      cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
      cp.conditional.handle = h;
      cp.conditional.type = cudaGraphCondTypeWhile;
      cp.conditional.size = 1;
      cudaGraphNode_t cnode;
      ok = cudaGraphAddNode(&cnode, g, nullptr, 0, &cp) == cudaSuccess;
      if (ok) body = cp.conditional.phGraph_out[0];
    }
    const int64_t l0 = ctx->launches;
    int st = AZ_OK;
    if (ok) ok = cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      st = m.tick(false);
      launch_move(m, grid1);
      az_k_assign<G><<<1, 1024, 0, ctx->stream>>>(m.p, sp);
      az_k_selfplay_cond<<<1, 1, 0, ctx->stream>>>(m.p, sp, h, h_progress);
      ok = cudaStreamEndCapture(ctx->stream, nullptr) == cudaSuccess && st == AZ_OK;
    }
    const int64_t per_tick = ctx->launches - l0 + 3 + (sp.minmax1 ? 1 : 0);
    ctx->launches = l0;
    if (ok) ok = cudaGraphInstantiate(&exec, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (!ok) { cudaGetLastError(); return AZ_EUNSUPPORTED; }
    cudaError_t e = cudaGraphLaunch(exec, ctx->stream);
    if (e != cudaSuccess) { cudaGraphExecDestroy(exec); ctx->err = std::string("self-play graph launch: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    // the GPU plays; the host relays progress (game_simulated callbacks are driven by az_selfplay_poll)
    while ((e = cudaStreamQuery(ctx->stream)) == cudaErrorNotReady) {
      a_done.store(((volatile int32_t*)h_progress)[0]);
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    cudaGraphExecDestroy(exec);
    if (e != cudaSuccess) { ctx->err = std::string("self-play graph: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    a_done.store(((volatile int32_t*)h_progress)[0]);
    (void)per_tick;
    return AZ_OK;
  }
  int loop() {
    auto t0 = std::chrono::steady_clock::now();
    Mcts<G>& m = *pool;
    cudaSetDevice(ctx->device);
    AZ_CUDA(ctx, cudaMemsetAsync(sp.games_done, 0, 4, ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(sp.active_slots, 0, 4, ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(sp.next_game, 0, 4, ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(m.p.expansions, 0, 8, ctx->stream));
    AZ_CUDA(ctx, cudaMemsetAsync(m.p.flags, 0, 4 * sizeof(int32_t), ctx->stream));
    AZ_TRY(ctx, m.reset());
    const int grid1 = (m.p.S + 127) / 128;
    az_k_move<G><<<grid1, 128, 0, ctx->stream>>>(m.p, sp, 1);
    az_k_assign<G><<<1, 1024, 0, ctx->stream>>>(m.p, sp);
    ctx->launches += 2;
    int64_t tick = 0;
    m.drop_graph();  // `sp` (game range, buffers) is baked into the captured move-kernel launch
    // device-driven loop: ONE graph launch plays the whole run; the host only watches the progress counter in pinned memory
    bool looped = false;
    if (m.use_device_loop && !m.device_loop_broken && m.use_graph && m.net->capturable() && (!m.net2 || m.net2->capturable())) {
      const int st = run_device_loop(m, grid1);
      if (st == AZ_OK) looped = true;
      else if (st == AZ_EUNSUPPORTED) m.device_loop_broken = true;
      else return st;
    }
    for (; !looped;) {
      AZ_TRY(ctx, m.tick_graphed([&] {
        launch_move(m, grid1);
        az_k_assign<G><<<1, 1024, 0, ctx->stream>>>(m.p, sp);
        ctx->launches += 2 + (sp.minmax1 ? 1 : 0);
      }));
      tick++;
      if (tick % 32 == 0) {
        AZ_CUDA(ctx, cudaMemcpyAsync(h_pin, sp.games_done, 4, cudaMemcpyDeviceToHost, ctx->stream));
        AZ_CUDA(ctx, cudaMemcpyAsync(h_pin + 1, sp.active_slots, 4, cudaMemcpyDeviceToHost, ctx->stream));
        AZ_CUDA(ctx, cudaMemcpyAsync(h_pin + 2, m.p.flags, 16, cudaMemcpyDeviceToHost, ctx->stream));
        AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        a_done.store(h_pin[0]);
        if (h_pin[2]) AZ_FAIL(ctx, AZ_ENOMEM, "MCTS table overflow during self-play");
        if (h_pin[4]) AZ_FAIL(ctx, AZ_ESTATE, "simulation path exceeded the per-game ply bound");
        if (h_pin[5]) AZ_FAIL(ctx, AZ_ESTATE, "root missing at move selection");
        if (h_pin[0] >= sp.num_games) break;
      }
    }
    m.drop_graph();
    if (looped) {  // errors raised on the device end the loop: report them like the polled path does
      AZ_CUDA(ctx, cudaMemcpy(h_pin + 2, m.p.flags, 16, cudaMemcpyDeviceToHost));
      if (h_pin[2]) AZ_FAIL(ctx, AZ_ENOMEM, "MCTS table overflow during self-play");
      if (h_pin[4]) AZ_FAIL(ctx, AZ_ESTATE, "simulation path exceeded the per-game ply bound");
      if (h_pin[5]) AZ_FAIL(ctx, AZ_ESTATE, "root missing at move selection");
    }
    AZ_CUDA(ctx, cudaMemcpy(&total_expansions, m.p.expansions, 8, cudaMemcpyDeviceToHost));
    h_moves.resize(sp.num_games);
    AZ_CUDA(ctx, cudaMemcpy(h_moves.data(), sp.g_moves, sp.num_games * sizeof(int32_t), cudaMemcpyDeviceToHost));
    seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return AZ_OK;
  }
  int start(int num_games, int64_t first) override {
    if (!a_finished.load()) AZ_FAIL(ctx, AZ_ESTATE, "az_selfplay_start: a run is already in flight");
    if (worker.joinable()) worker.join();
    if (num_games <= 0) AZ_FAIL(ctx, AZ_EINVAL, "num_games must be positive");
    if (simp.reset_every <= 0) {
      size_t need = (size_t)std::max(sp.nsims, sp.nsims1) * G::MAX_PLIES * (size_t)((num_games + sp.W - 1) / sp.W);
      (void)need;  // overflow is detected on device and reported as AZ_ENOMEM
    }
    AZ_TRY(ctx, ensure_game_buffers(num_games));
    sp.num_games = num_games; sp.first_game = first;
    a_done.store(0); a_finished.store(0); a_status.store(AZ_OK);
    worker = std::thread([this]() {
      int s = loop();
      if (s != AZ_OK) werr = ctx->err;
      a_status.store(s);
      a_finished.store(1);
    });
    return AZ_OK;
  }
  int poll(int32_t* done, int32_t* fin) override {
    if (done) *done = a_finished.load() && a_status.load() == AZ_OK ? sp.num_games : a_done.load();
    if (fin) *fin = a_finished.load();
    return a_status.load();
  }
  int wait() override {
    if (worker.joinable()) worker.join();
    if (a_status.load() != AZ_OK) ctx->err = werr;
    return a_status.load();
  }
  int counts(int64_t* ns, int64_t* ng) override {
    AZ_TRY(ctx, wait());
    int64_t n = 0;
    for (int v : h_moves) n += v;
    if (ns) *ns = n;
    if (ng) *ng = (int64_t)h_moves.size();
    return AZ_OK;
  }
  int fetch(uint8_t* states, float* pi, uint8_t* mask, float* z, float* t, int32_t* gos, double* rew, int32_t* act) override {
    AZ_TRY(ctx, wait());
    const int ng = sp.num_games;
    size_t rows = (size_t)ng * G::MAX_PLIES;
    std::vector<AzEnv> env(rows), think(rows);
    std::vector<double> hpi(rows * G::A), hz(rows);
    std::vector<float> ht(rows);
    std::vector<int32_t> hact(rows);
    std::vector<double> hrew(rows);
    AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AZ_TRY(ctx, az_d2h(ctx, env.data(), sp.s_env, rows * sizeof(AzEnv)));
    AZ_TRY(ctx, az_d2h(ctx, think.data(), sp.s_root, rows * sizeof(AzEnv)));
    AZ_TRY(ctx, az_d2h(ctx, hpi.data(), sp.s_pi, rows * G::A * sizeof(double)));
    AZ_TRY(ctx, az_d2h(ctx, hz.data(), sp.s_z, rows * sizeof(double)));
    AZ_TRY(ctx, az_d2h(ctx, ht.data(), sp.s_t, rows * sizeof(float)));
    AZ_TRY(ctx, az_d2h(ctx, hact.data(), sp.s_action, rows * sizeof(int32_t)));
    AZ_TRY(ctx, az_d2h(ctx, hrew.data(), sp.s_reward, rows * sizeof(double)));
    size_t k = 0;
    for (int g = 0; g < ng; g++)
      for (int i = 0; i < h_moves[g]; i++, k++) {
        size_t r = (size_t)g * G::MAX_PLIES + i;
        if (states) G::to_bytes(env[r], states + k * G::STATE_BYTES);
        uint32_t legal = G::legal_mask(think[r]);  // pi and the mask are in the frame the player thought in
        for (int a = 0; a < G::A; a++) {
          if (pi) pi[k * G::A + a] = (float)hpi[r * G::A + a];
          if (mask) mask[k * G::A + a] = (legal >> a) & 1;
        }
        if (z) z[k] = (float)hz[r];
        if (t) t[k] = ht[r];
        if (gos) gos[k] = g;
        if (rew) rew[k] = hrew[r];
        if (act) act[k] = hact[r];
      }
    return AZ_OK;
  }
  int stats(double* ed, int64_t* nodes, int32_t* moves, double* totals) override {
    AZ_TRY(ctx, wait());
    const int ng = sp.num_games;
    if (ed) AZ_CUDA(ctx, cudaMemcpy(ed, sp.g_edepth, ng * sizeof(double), cudaMemcpyDeviceToHost));
    if (nodes) AZ_CUDA(ctx, cudaMemcpy(nodes, sp.g_nodes, ng * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (moves) memcpy(moves, h_moves.data(), ng * sizeof(int32_t));
    if (totals) {
      int64_t ns = 0;
      for (int v : h_moves) ns += v;
      // simulations: the device counters (MCTS.Env.total_simulations of every tree, kept across reset!, src/mcts.jl:142,278-281)
      std::vector<int64_t> tsims((size_t)pool->p.S);
      AZ_CUDA(ctx, cudaMemcpy(tsims.data(), pool->p.total_sims, tsims.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
      int64_t sims = 0;
      for (int64_t v : tsims) sims += v;
      totals[0] = seconds; totals[1] = (double)sims; totals[2] = (double)total_expansions; totals[3] = (double)ns;
    }
    return AZ_OK;
  }
  // the finished run's samples as a device-resident set (no host round trip): feeds az_samples_* (merge / augment / convert)
  int export_samples(az_samples** out) override {
    AZ_TRY(ctx, wait());
    const int ng = sp.num_games;
    std::vector<int64_t> off((size_t)ng + 1, 0);
    for (int g = 0; g < ng; g++) off[g + 1] = off[g] + h_moves[g];
    az_samples* o = nullptr;
    AZ_TRY(ctx, az_samples_alloc(ctx, G::ID, off[ng], &o));
    int64_t* d_off = nullptr;
    if (cudaMalloc((void**)&d_off, off.size() * 8) != cudaSuccess) { az_samples_destroy(o); AZ_FAIL(ctx, AZ_ENOMEM, "export_samples: cudaMalloc failed"); }
    cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    const int64_t rows = (int64_t)ng * G::MAX_PLIES;
    az_k_export_samples<G><<<(int)((rows + 255) / 256), 256, 0, ctx->stream>>>(sp, ng, d_off, az_samples_env(o), az_samples_pi(o), az_samples_z(o),
                                                                            az_samples_t(o), az_samples_cnt(o));
    ctx->launches += 1;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_off);
    if (e != cudaSuccess) { az_samples_destroy(o); AZ_FAIL(ctx, AZ_ECUDA, std::string("export_samples: ") + cudaGetErrorString(e)); }
    *out = o;
    return AZ_OK;
  }
  // rewards_and_redundancy (src/simulations.jl:292-307): per game total_reward(trace, gamma) (src/trace.jl:45-47), negated
  // when the colours were flipped; redundancy = 1 - |unique states| / |states| over every trace state (final ones included)
  int outcomes(double gamma, double* rewards, int32_t* flipped, uint8_t* finals, double* redundancy) override {
    AZ_TRY(ctx, wait());
    const int ng = sp.num_games;
    size_t rows = (size_t)ng * G::MAX_PLIES;
    std::vector<AzEnv> env(rows), fin(ng);
    std::vector<double> hrew(rows);
    AZ_CUDA(ctx, cudaMemcpy(env.data(), sp.s_env, rows * sizeof(AzEnv), cudaMemcpyDeviceToHost));
    AZ_CUDA(ctx, cudaMemcpy(fin.data(), sp.g_final, (size_t)ng * sizeof(AzEnv), cudaMemcpyDeviceToHost));
    AZ_CUDA(ctx, cudaMemcpy(hrew.data(), sp.s_reward, rows * sizeof(double), cudaMemcpyDeviceToHost));
    std::unordered_set<std::string> uniq;
    size_t total = 0;
    uint8_t buf[G::STATE_BYTES];
    for (int g = 0; g < ng; g++) {
      const bool fl = sp.duel && sp.alternate && (((sp.first_game + g + 1) & 1) == 1);
      double s = 0.0, gp = 1.0;
      for (int i = 0; i < h_moves[g]; i++) {
        const size_t r = (size_t)g * G::MAX_PLIES + i;
        s = (i == 0) ? gp * hrew[r] : s + gp * hrew[r];
        gp = gp * gamma;
        G::to_bytes(env[r], buf);
        uniq.emplace((const char*)buf, (size_t)G::STATE_BYTES);
        total++;
      }
      G::to_bytes(fin[g], buf);
      uniq.emplace((const char*)buf, (size_t)G::STATE_BYTES);
      total++;
      if (finals) memcpy(finals + (size_t)g * G::STATE_BYTES, buf, G::STATE_BYTES);
      if (rewards) rewards[g] = fl ? -s : s;
      if (flipped) flipped[g] = fl ? 1 : 0;
    }
    if (redundancy) *redundancy = 1.0 - (double)uniq.size() / (double)total;
    return AZ_OK;
  }
};
template <class G>
static int g_make_selfplay(az_ctx* ctx, az_net* net, az_net* net2, const az_mcts_params* mp, const az_sim_params* sp, uint64_t seed, az_selfplay** out,
                           const az_mcts_params* mp1 = nullptr, const az_minmax_params* mm = nullptr) {
  auto* s = new SelfPlay<G>();
  int st = s->create(ctx, net, net2, mp, sp, seed, mp1, mm);
  if (st != AZ_OK) { delete s; return st; }
  *out = s;
  return AZ_OK;
}

// forward_normalized hook for oracle nets / networks: host states -> device envs -> eval -> host
template <class G>
static int g_net_forward(az_net* net, const uint8_t* states, int B, float* P, float* V, float* Pinv, float* logits, float* vpre) {
  az_ctx* ctx = net->ctx;
  std::vector<AzEnv> envs(B);
  for (int i = 0; i < B; i++) envs[i] = G::from_bytes(states + (size_t)i * G::STATE_BYTES);
  AzEnv* d_env; int32_t* d_n; float *d_P, *d_V, *d_Pi, *d_L = nullptr, *d_Vp = nullptr;
  AZ_TRY(ctx, az_dalloc(ctx, &d_env, B)); AZ_TRY(ctx, az_dalloc(ctx, &d_n, 1));
  AZ_TRY(ctx, az_dalloc(ctx, &d_P, (size_t)B * G::A)); AZ_TRY(ctx, az_dalloc(ctx, &d_V, B)); AZ_TRY(ctx, az_dalloc(ctx, &d_Pi, B));
  if (logits) AZ_TRY(ctx, az_dalloc(ctx, &d_L, (size_t)B * G::A));
  if (vpre) AZ_TRY(ctx, az_dalloc(ctx, &d_Vp, B));
  int32_t n = B;
  AZ_CUDA(ctx, cudaMemcpyAsync(d_env, envs.data(), B * sizeof(AzEnv), cudaMemcpyHostToDevice, ctx->stream));
  AZ_CUDA(ctx, cudaMemcpyAsync(d_n, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
  net->dbg_logit = d_L; net->dbg_vpre = d_Vp;
  int st = net->eval_with_pinv(d_env, d_n, B, d_P, d_V, d_Pi);
  net->dbg_logit = nullptr; net->dbg_vpre = nullptr;
  if (st == AZ_OK) {
    if (P) cudaMemcpyAsync(P, d_P, (size_t)B * G::A * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (V) cudaMemcpyAsync(V, d_V, B * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (Pinv) cudaMemcpyAsync(Pinv, d_Pi, B * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (logits) cudaMemcpyAsync(logits, d_L, (size_t)B * G::A * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (vpre) cudaMemcpyAsync(vpre, d_Vp, B * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_env); cudaFree(d_n); cudaFree(d_P); cudaFree(d_V); cudaFree(d_Pi); cudaFree(d_L); cudaFree(d_Vp);
  if (st != AZ_OK) return st;
  if (e != cudaSuccess) AZ_FAIL(ctx, AZ_ECUDA, std::string("az_net_forward: ") + cudaGetErrorString(e));
  return AZ_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define AZ_GUARD_BEGIN try {
#define AZ_GUARD_END(ctx)                                         \
  } catch (const std::exception& ex) {                            \
    if (ctx) (ctx)->err = std::string("exception: ") + ex.what(); \
    return AZ_ESTATE;                                             \
  } catch (...) {                                                 \
    if (ctx) (ctx)->err = "unknown exception";                    \
    return AZ_ESTATE;                                             \
  }

int az_d2h(az_ctx* ctx, void* dst, const void* src, size_t bytes) {
  constexpr size_t CHUNK = (size_t)8 << 20;
  if (bytes == 0) return AZ_OK;
  for (int i = 0; i < 2; i++) {
    if (!ctx->pin[i]) AZ_CUDA(ctx, cudaMallocHost(&ctx->pin[i], CHUNK));
    if (!ctx->pin_ev[i]) AZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
  }
  const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
  for (size_t i = 0; i <= nchunks; i++) {
    if (i < nchunks) {
      const size_t off = i * CHUNK, len = std::min(CHUNK, bytes - off);
      AZ_CUDA(ctx, cudaMemcpyAsync(ctx->pin[i & 1], (const char*)src + off, len, cudaMemcpyDeviceToHost, ctx->stream));
      AZ_CUDA(ctx, cudaEventRecord(ctx->pin_ev[i & 1], ctx->stream));
    }
    if (i > 0) {  // drain the previous chunk while the copy engine works on this one
      const size_t j = i - 1, off = j * CHUNK, len = std::min(CHUNK, bytes - off);
      AZ_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[j & 1]));
      memcpy((char*)dst + off, ctx->pin[j & 1], len);
    }
  }
  return AZ_OK;
}

extern "C" {

int32_t az_version(void) { return AZ_ABI_VERSION; }

int32_t az_ctx_create(int32_t device, az_ctx** out) {
  if (!out) return AZ_EINVAL;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libazb200 has no CPU fallback)";
    cudaGetLastError();
    return AZ_ECUDA;
  }
  if (device < 0 || device >= n) { g_create_err = "az_ctx_create: bad device index"; return AZ_EINVAL; }
  az_ctx* c = new az_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "az_ctx_create: cudaSetDevice/cudaStreamCreate failed";
    delete c;
    return AZ_ECUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
  *out = c;
  return AZ_OK;
}
int32_t az_ctx_destroy(az_ctx* ctx) {
  if (!ctx) return AZ_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 2; i++) { if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]); if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AZ_OK;
}
const char* az_last_error(az_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int32_t az_ctx_synchronize(az_ctx* ctx) {
  if (!ctx) return AZ_EINVAL;
  AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AZ_OK;
}
int64_t az_ctx_num_launches(az_ctx* ctx) { return ctx ? ctx->launches : -1; }

int32_t az_game_lookup(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < 4; i++) if (strcmp(name, AZ_GAME_NAMES[i]) == 0) return i;
  return -1;
}
int32_t az_game_num_actions(int32_t game) { switch (game) { case 0: return GameC4::A; case 1: return GameTTT::A; case 2: return GameMancala::A; case 3: return GameGW::A; } return -1; }
int32_t az_game_state_bytes(int32_t game) { switch (game) { case 0: return GameC4::STATE_BYTES; case 1: return GameTTT::STATE_BYTES; case 2: return GameMancala::STATE_BYTES; case 3: return GameGW::STATE_BYTES; } return -1; }
int32_t az_game_max_plies(int32_t game) { switch (game) { case 0: return GameC4::MAX_PLIES; case 1: return GameTTT::MAX_PLIES; case 2: return GameMancala::MAX_PLIES; case 3: return GameGW::MAX_PLIES; } return -1; }
int32_t az_game_state_dim(int32_t game, int32_t dim[3]) { if (!dim) return AZ_EINVAL; AZ_DISPATCH_GAME(game, g_state_dim, dim) }
int32_t az_game_vectorize_state(int32_t game, const uint8_t* s, float* x) { if (!s || !x) return AZ_EINVAL; AZ_DISPATCH_GAME(game, g_vectorize, s, x) }
int32_t az_game_actions_mask(int32_t game, const uint8_t* s, uint8_t* m) { if (!s || !m) return AZ_EINVAL; AZ_DISPATCH_GAME(game, g_mask, s, m) }
int32_t az_game_play(int32_t game, const uint8_t* s, int32_t a, uint8_t* ns, int32_t* term, double* wr) {
  if (!s || !ns) return AZ_EINVAL;
  AZ_DISPATCH_GAME(game, g_play, s, a, ns, term, wr)
}
int32_t az_game_heuristic_value(int32_t game, const uint8_t* s, double* v) { if (!s || !v) return AZ_EINVAL; AZ_DISPATCH_GAME(game, g_heuristic, s, v) }
int32_t az_game_minmax_think(int32_t game, const uint8_t* s, const az_minmax_params* mm, double* q, double* pi) {
  if (!s || !mm) return AZ_EINVAL;
  AZ_DISPATCH_GAME(game, g_minmax_think, s, mm, q, pi)
}
int32_t az_game_init_state(int32_t game, uint8_t* s) { if (!s) return AZ_EINVAL; AZ_DISPATCH_GAME(game, g_init_state, s) }
int32_t az_game_random_positions(int32_t game, uint64_t seed, uint64_t first, int32_t n, int32_t max_plies, uint8_t* out) {
  if (!out || n < 0 || max_plies < 0) return AZ_EINVAL;
  AZ_DISPATCH_GAME(game, g_random_positions, seed, first, n, max_plies, out)
}

int32_t az_net_create_oracle(az_ctx* ctx, int32_t kind, int32_t game, az_net** out) {
  if (!ctx || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  AZ_DISPATCH_GAME(game, g_make_oracle, ctx, kind, out)
  AZ_GUARD_END(ctx)
}
int32_t az_net_create_rollout(az_ctx* ctx, int32_t game, double gamma, uint64_t seed, az_net** out) {
  if (!ctx || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  AZ_DISPATCH_GAME(game, g_make_rollout, ctx, gamma, seed, out)
  AZ_GUARD_END(ctx)
}
int32_t az_net_create_resnet(az_ctx* ctx, int32_t game, const az_resnet_hp* hp, az_net** out) {
  if (!ctx || !hp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  int st = AZ_OK;
  az_net* n = az_make_resnet(ctx, game, hp, &st);
  if (st != AZ_OK) return st;
  *out = n;
  return AZ_OK;
  AZ_GUARD_END(ctx)
}
int32_t az_net_create_simplenet(az_ctx* ctx, int32_t game, const az_simplenet_hp* hp, az_net** out) {
  if (!ctx || !hp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  int st = AZ_OK;
  az_net* n = az_make_simplenet(ctx, game, hp, &st);
  if (st != AZ_OK) return st;
  *out = n;
  return AZ_OK;
  AZ_GUARD_END(ctx)
}
int32_t az_net_num_params(az_net* net, int64_t* n) { if (!net || !n) return AZ_EINVAL; *n = net->num_params(); return AZ_OK; }
int32_t az_net_load(az_net* net, const float* blob, int64_t n) {
  if (!net || !blob) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(net->ctx->device);
  return net->load(blob, n);
  AZ_GUARD_END(net->ctx)
}
int32_t az_net_load_device(az_net* net, const float* d_blob, int64_t n) {
  if (!net || !d_blob) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(net->ctx->device);
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, d_blob) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != net->ctx->device) {
    cudaGetLastError();
    net->ctx->err = "az_net_load_device: the blob must be device memory of the network's GPU";
    return AZ_EINVAL;
  }
  return net->load_device(d_blob, n);
  AZ_GUARD_END(net->ctx)
}
int32_t az_net_forward(az_net* net, const uint8_t* states, int32_t B, float* P, float* V, float* Pinv) {
  if (!net || !states || !P || !V || B <= 0) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(net->ctx->device);
  AZ_DISPATCH_GAME(net->game, g_net_forward, net, states, B, P, V, Pinv, nullptr, nullptr)
  AZ_GUARD_END(net->ctx)
}
int32_t az_net_forward_logits(az_net* net, const uint8_t* states, int32_t B, float* policy_logits, float* value_pre) {
  if (!net || !states || B <= 0 || (!policy_logits && !value_pre)) return AZ_EINVAL;
  if (net->kind != AZ_NET_RESNET && net->kind != AZ_NET_SIMPLENET) { net->ctx->err = "az_net_forward_logits: networks only"; return AZ_EINVAL; }
  AZ_GUARD_BEGIN
  cudaSetDevice(net->ctx->device);
  AZ_DISPATCH_GAME(net->game, g_net_forward, net, states, B, nullptr, nullptr, nullptr, policy_logits, value_pre)
  AZ_GUARD_END(net->ctx)
}
int32_t az_net_set_profiling(az_net* net, int32_t enable) { if (!net) return AZ_EINVAL; cudaSetDevice(net->ctx->device); return net->set_profiling(enable); }
int32_t az_net_get_profile(az_net* net, double* tower_ms, int64_t* tower_launches, double* total_ms, int64_t* evals) {
  if (!net) return AZ_EINVAL;
  cudaSetDevice(net->ctx->device);
  return net->get_profile(tower_ms, tower_launches, total_ms, evals);
}
int32_t az_net_destroy(az_net* net) { if (!net) return AZ_EINVAL; cudaSetDevice(net->ctx->device); delete net; return AZ_OK; }

int32_t az_mcts_create(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* p, int32_t n_trees, int32_t cap, az_mcts** out) {
  if (!ctx || !oracle || !p || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  AZ_DISPATCH_GAME(game, g_make_mcts, ctx, oracle, p, n_trees, cap, out)
  AZ_GUARD_END(ctx)
}
#define AZ_M(m) if (!(m)) return AZ_EINVAL; cudaSetDevice((m)->ctx->device);
int32_t az_mcts_set_roots(az_mcts* m, const uint8_t* s, const double* eta) { AZ_M(m) if (!s) return AZ_EINVAL; AZ_GUARD_BEGIN return m->set_roots(s, eta); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_run(az_mcts* m, int32_t nsims) { AZ_M(m) AZ_GUARD_BEGIN return m->run(nsims); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_explore(az_mcts* m, const uint8_t* s, const double* eta, int32_t nsims, int64_t* N, double* W, float* P) {
  AZ_M(m) if (!s) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  AZ_TRY(m->ctx, m->set_roots(s, eta));
  AZ_TRY(m->ctx, m->run(nsims));
  return m->root_stats(N, W, P);
  AZ_GUARD_END(m->ctx)
}
int32_t az_mcts_root_stats(az_mcts* m, int64_t* N, double* W, float* P) { AZ_M(m) AZ_GUARD_BEGIN return m->root_stats(N, W, P); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_policy(az_mcts* m, double* pi) { AZ_M(m) if (!pi) return AZ_EINVAL; AZ_GUARD_BEGIN return m->policy(pi); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_reset(az_mcts* m) { AZ_M(m) AZ_GUARD_BEGIN return m->reset(); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_counters(az_mcts* m, int64_t* ts, int64_t* tn, int64_t* nn) { AZ_M(m) AZ_GUARD_BEGIN return m->counters(ts, tn, nn); AZ_GUARD_END(m->ctx) }
int32_t az_mcts_set_noise(az_mcts* m, uint64_t seed, const int64_t* games, const int32_t* moves) {
  AZ_M(m) if (!games || !moves) return AZ_EINVAL;
  AZ_GUARD_BEGIN return m->set_noise(seed, games, moves); AZ_GUARD_END(m->ctx)
}
int32_t az_mcts_last_timing(az_mcts* m, double* ms_total, double* ms_net, int64_t* ticks, int64_t* ex) {
  if (!m) return AZ_EINVAL;
  if (ms_total) *ms_total = m->ms_total;
  if (ms_net) *ms_net = m->ms_net;
  if (ticks) *ticks = m->ticks;
  if (ex) *ex = m->expansions;
  return AZ_OK;
}
int32_t az_mcts_set_profiling(az_mcts* m, int32_t enable) {
  if (!m) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(m->ctx->device);
  return m->set_profiling(enable);
  AZ_GUARD_END(m->ctx)
}
int32_t az_mcts_get_profile(az_mcts* m, double* select_ms, double* expand_ms, double* net_ms, int64_t* ticks) {
  if (!m) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(m->ctx->device);
  m->drain_profile();
  if (select_ms) *select_ms = m->prof_select_ms;
  if (expand_ms) *expand_ms = m->prof_expand_ms;
  if (net_ms) *net_ms = m->prof_net_ms;
  if (ticks) *ticks = m->prof_ticks;
  m->prof_select_ms = m->prof_expand_ms = m->prof_net_ms = 0;
  m->prof_ticks = 0;
  return AZ_OK;
  AZ_GUARD_END(m->ctx)
}
int32_t az_mcts_destroy(az_mcts* m) { AZ_M(m) delete m; return AZ_OK; }

int32_t az_selfplay_create(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* mp, const az_sim_params* sp, uint64_t seed, az_selfplay** out) {
  if (!ctx || !oracle || !mp || !sp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  if (oracle->game != game) AZ_FAIL(ctx, AZ_EINVAL, "az_selfplay_create: oracle was built for another game");
  AZ_DISPATCH_GAME(game, g_make_selfplay, ctx, oracle, nullptr, mp, sp, seed, out)
  AZ_GUARD_END(ctx)
}
int32_t az_selfplay_create_duel(az_ctx* ctx, int32_t game, az_net* white, az_net* black, const az_mcts_params* mp, const az_sim_params* sp,
                                uint64_t seed, az_selfplay** out) {
  if (!ctx || !white || !black || !mp || !sp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  if (white->game != game || black->game != game) AZ_FAIL(ctx, AZ_EINVAL, "az_selfplay_create_duel: oracle was built for another game");
  AZ_DISPATCH_GAME(game, g_make_selfplay, ctx, white, black, mp, sp, seed, out)
  AZ_GUARD_END(ctx)
}
int32_t az_selfplay_create_duel_players(az_ctx* ctx, int32_t game, az_net* white, const az_mcts_params* mp_white, az_net* black,
                                        const az_mcts_params* mp_black, const az_sim_params* sp, uint64_t seed, az_selfplay** out) {
  if (!ctx || !white || !black || !mp_white || !mp_black || !sp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  if (white->game != game || black->game != game) AZ_FAIL(ctx, AZ_EINVAL, "az_selfplay_create_duel_players: oracle was built for another game");
  AZ_DISPATCH_GAME(game, g_make_selfplay, ctx, white, black, mp_white, sp, seed, out, mp_black)
  AZ_GUARD_END(ctx)
}
int32_t az_selfplay_create_duel_minmax(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* mp, const az_minmax_params* baseline,
                                       const az_sim_params* sp, uint64_t seed, az_selfplay** out) {
  if (!ctx || !oracle || !mp || !baseline || !sp || !out) return AZ_EINVAL;
  AZ_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  if (oracle->game != game) AZ_FAIL(ctx, AZ_EINVAL, "az_selfplay_create_duel_minmax: oracle was built for another game");
  AZ_DISPATCH_GAME(game, g_make_selfplay, ctx, oracle, oracle, mp, sp, seed, out, mp, baseline)
  AZ_GUARD_END(ctx)
}
int32_t az_selfplay_export_samples(az_selfplay* s, az_samples** out) {
  AZ_M(s) if (!out) return AZ_EINVAL;
  AZ_GUARD_BEGIN cudaSetDevice(s->ctx->device); return s->export_samples(out); AZ_GUARD_END(s->ctx)
}
int32_t az_selfplay_outcomes(az_selfplay* s, double gamma, double* rewards, int32_t* flipped, uint8_t* finals, double* red) {
  AZ_M(s) AZ_GUARD_BEGIN return s->outcomes(gamma, rewards, flipped, finals, red); AZ_GUARD_END(s->ctx)
}
int32_t az_selfplay_start(az_selfplay* s, int32_t ng, int64_t first) { AZ_M(s) AZ_GUARD_BEGIN return s->start(ng, first); AZ_GUARD_END(s->ctx) }
int32_t az_selfplay_poll(az_selfplay* s, int32_t* done, int32_t* fin) { if (!s) return AZ_EINVAL; return s->poll(done, fin); }
int32_t az_selfplay_wait(az_selfplay* s) { AZ_M(s) AZ_GUARD_BEGIN return s->wait(); AZ_GUARD_END(s->ctx) }
int32_t az_selfplay_counts(az_selfplay* s, int64_t* ns, int64_t* ng) { AZ_M(s) AZ_GUARD_BEGIN return s->counts(ns, ng); AZ_GUARD_END(s->ctx) }
int32_t az_selfplay_fetch(az_selfplay* s, uint8_t* states, float* pi, uint8_t* mask, float* z, float* t, int32_t* gos, double* rew, int32_t* act) {
  AZ_M(s) AZ_GUARD_BEGIN return s->fetch(states, pi, mask, z, t, gos, rew, act); AZ_GUARD_END(s->ctx)
}
int32_t az_selfplay_stats(az_selfplay* s, double* ed, int64_t* nodes, int32_t* moves, double* totals) { AZ_M(s) AZ_GUARD_BEGIN return s->stats(ed, nodes, moves, totals); AZ_GUARD_END(s->ctx) }
int32_t az_selfplay_destroy(az_selfplay* s) { AZ_M(s) delete s; return AZ_OK; }

}  // extern "C"
