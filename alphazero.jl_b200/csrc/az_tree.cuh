// az_tree.cuh -- on-device MCTS: select / expand / backup / per-move kernels over a pool of independent trees.
//
// One tree ("slot") = one MCTS.Env of the reference (src/mcts.jl:124-151): a state-keyed transposition table,
// here an open-addressing hash table in HBM.  A node is one LANES*16-byte line:
//     lane 0      : key  {a, b | tag<<57}                (16 B)
//     lane 1 + i  : edge {W f64, P f32, N i32} of action i (16 B)      -> Connect-Four: exactly 128 B
// A group of LANES threads owns a slot: one coalesced line load fetches key + all edges of a node, the PUCT
// argmax is a shuffle reduction across the group, and only one writer ever touches a table (no atomics).
//
// Numerics follow src/mcts.jl exactly (W, scores, q in f64 with one rounding per operation -- this file is
// compiled with -fmad=false; P and Vest f32; N integer); see DESIGN.md "bit-exact select".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "az_games.cuh"
#include "az_rng.cuh"

union AzLine16 {
  uint4 u;
  struct { uint64_t a, b; } key;
  struct { double W; float P; int32_t N; } e;
};

struct AzMctsConst {
  double gamma, cpuct, eps, alpha, prior_temp;
};

struct AzPool {
  int S;              // number of trees (slots)
  uint32_t cap_mask;  // table capacity - 1 (lines per slot)
  int maxd;           // path capacity
  uint4* nodes;       // [S][cap][LANES]
  AzEnv* root;        // [S]
  uint32_t* tag;      // [S] 7-bit tag: bit6 = occupied, bits0-5 = generation
  int32_t* node_count;
  int64_t* total_sims;
  int64_t* total_nodes;
  int32_t* sims_done;
  int32_t* sims_target;
  uint8_t* status;    // 0 idle, 1 active
  double* eta;        // [S][A] compact over legal actions
  int32_t* pending;
  uint32_t* leaf_pos;
  AzEnv* leaf_env;
  int32_t* leaf_row;
  int32_t* depth;
  uint32_t* path_node;  // [S][maxd]
  uint16_t* path_meta;  // [S][maxd]  action | pswitch << 8
  double* path_r;       // [S][maxd]
  int32_t* n_leaves;    // [4]: [0] leaves emitted this tick, [1] trees that still have simulations to run,
                        //      [2] leaves of the second oracle's queue (duel mode)
  int32_t duel;         // duel mode (TwoPlayers, src/play.jl:248-282): tree 2*w + k belongs to player k of worker w and its
                        // leaves go to oracle k; queue k starts at row k * row_base1 of the batch arrays
  int32_t row_base1;
  int32_t max_sims_per_call;  // cap on simulations one select call runs for a tree (bounds the kernel's tail)
  AzEnv* batch_env;     // [S]
  float* batch_P;       // [S][A]
  float* batch_V;       // [S]
  int32_t* flags;       // [4]: 0 overflow, 1 active slots not finished (select), 2 path overflow
  int64_t* expansions;  // [1]
  uint64_t noise_seed;  // stochastic environments: in-tree noise stream = (noise_seed, noise_game[slot], noise_move[slot], sim, depth)
  int64_t* noise_game;  // [S]
  int32_t* noise_move;  // [S]
  AzMctsConst c;        // MCTS constants of player 0 (every tree outside a duel)
  AzMctsConst c1;       // duel: constants of player 1's trees (odd slots); TwoPlayers may pair two different MctsPlayers
};
__device__ __forceinline__ const AzMctsConst& az_c(const AzPool& p, int slot) { return (p.duel && (slot & 1)) ? p.c1 : p.c; }

#define AZ_KEYB_MASK ((1ull << 57) - 1)

template <int L>
__device__ __forceinline__ unsigned az_group_mask() {  // the first L lanes of the warp
  return L == 32 ? 0xffffffffu : ((1u << L) - 1u);
}
__device__ __forceinline__ uint32_t az_hash(uint64_t a, uint64_t b) {
  uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15F39CC060ull) * 0xC2B2AE3D27D4EB4Full;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 29;
  return (uint32_t)x;
}

// find the line of `env` in the slot's table.  Returns 1 found / 2 empty (h = insert position); ln = the line.
template <class G>
__device__ __forceinline__ int az_probe(const uint4* tab, uint32_t cap_mask, uint32_t tag, const AzEnv& env, int lane,
                                        unsigned gm, uint32_t& h, AzLine16& ln) {
  constexpr int L = G::LANES;
  h = az_hash(env.a, env.b) & cap_mask;
  for (;;) {
    ln.u = tab[(size_t)h * L + lane];
    int st = 0;
    if (lane == 0) {
      if ((uint32_t)(ln.key.b >> 57) != tag) st = 2;
      else if (ln.key.a == env.a && (ln.key.b & AZ_KEYB_MASK) == env.b) st = 1;
    }
    st = __shfl_sync(gm, st, 0, L);
    if (st) return st;
    h = (h + 1) & cap_mask;
  }
}

// backup (update_state_info!, src/mcts.jl:190-194,216-221) of the recorded path with leaf value q
template <class G>
__device__ __forceinline__ void az_backup(const AzPool& p, int slot, int lane, unsigned gm, uint4* tab, int depth, double q) {
  constexpr int L = G::LANES;
  __syncwarp(gm);
  const uint32_t* pn = p.path_node + (size_t)slot * p.maxd;
  const uint16_t* pm = p.path_meta + (size_t)slot * p.maxd;
  const double* pr = p.path_r + (size_t)slot * p.maxd;
  for (int j = depth - 1; j >= 0; j--) {
    uint32_t meta = pm[j];
    if (meta >> 8) q = -q;
    q = pr[j] + az_c(p, slot).gamma * q;
    bool mine = G::ACYCLIC ? ((j % L) == lane) : (lane == 0);
    if (mine) {
      uint4* addr = tab + (size_t)pn[j] * L + 1 + (meta & 0xFF);
      AzLine16 e;
      e.u = *addr;
      e.e.W = e.e.W + q;
      e.e.N += 1;
      *addr = e.u;
    }
  }
  __syncwarp(gm);
}

// ---- select: run simulations until a leaf needs the oracle or the budget is spent (src/mcts.jl:199-226,239-245) ----
template <class G>
__device__ __forceinline__ void az_select_slot(const AzPool& p, const int slot, const int lane) {
  constexpr int L = G::LANES;
  constexpr int A = G::A;
  unsigned gm = az_group_mask<L>();
  if (!p.status[slot] || p.pending[slot]) return;
  int sims_done = p.sims_done[slot];
  const int target = p.sims_target[slot];
  if (sims_done >= target) return;
  const AzEnv root = p.root[slot];
  const uint32_t tag = p.tag[slot];
  uint4* tab = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
  uint32_t* pn = p.path_node + (size_t)slot * p.maxd;
  uint16_t* pm = p.path_meta + (size_t)slot * p.maxd;
  double* pr = p.path_r + (size_t)slot * p.maxd;
  int64_t tsims = 0, tnodes = 0;
  const int a = lane - 1;
  AzNoiseKey nkey = {0, 0, 0};
  if (G::STOCHASTIC) { nkey.seed = p.noise_seed; nkey.game = (uint64_t)p.noise_game[slot]; nkey.move = (uint32_t)p.noise_move[slot]; }
  int budget = p.max_sims_per_call;
  while (sims_done < target && budget-- > 0) {
    tsims++;
    AzEnv env = root;
    int depth = 0;
    bool isroot = true, need_leaf = false;
    for (;;) {
      if (G::terminated(env) || depth >= p.maxd) {
        if (depth >= p.maxd && lane == 0) p.flags[2] = 1;
        az_backup<G>(p, slot, lane, gm, tab, depth, 0.0);
        tnodes += depth;
        break;
      }
      uint32_t h;
      AzLine16 ln;
      int st = az_probe<G>(tab, p.cap_mask, tag, env, lane, gm, h, ln);
      if (st == 2) {  // new node: ask the oracle (state_info, src/mcts.jl:165-174)
        if (lane == 0) {
          const int q = p.duel ? (slot & 1) : 0;
          int row = atomicAdd(p.n_leaves + 2 * q, 1) + q * p.row_base1;
          p.batch_env[row] = env;
          p.leaf_row[slot] = row;
          p.leaf_pos[slot] = h;
          p.leaf_env[slot] = env;
          p.depth[slot] = depth;
          p.pending[slot] = 1;
        }
        need_leaf = true;
        break;
      }
      // uct_scores (src/mcts.jl:180-188) + argmax (first maximal legal action)
      const uint32_t legal = G::legal_mask(env);
      const bool is_edge = (a >= 0) && (a < A) && ((legal >> a) & 1u);
      int n = is_edge ? ln.e.N : 0;
      int ntot = n;
#pragma unroll
      for (int off = L / 2; off >= 1; off >>= 1) ntot += __shfl_xor_sync(gm, ntot, off, L);
      double score = __longlong_as_double((long long)0xFFF0000000000000ull);  // -inf
      if (is_edge) {
        double Pd = (double)ln.e.P;
        if (isroot && az_c(p, slot).eps != 0.0) {
          int idx = __popc(legal & ((1u << a) - 1u));
          Pd = (1.0 - az_c(p, slot).eps) * Pd + az_c(p, slot).eps * p.eta[(size_t)slot * A + idx];
        }
        double sq = sqrt((double)ntot);
        score = ln.e.W / (double)(n > 1 ? n : 1) + ((az_c(p, slot).cpuct * Pd) * sq) / (double)(n + 1);
      }
      int best = lane;
#pragma unroll
      for (int off = L / 2; off >= 1; off >>= 1) {
        double os = __shfl_xor_sync(gm, score, off, L);
        int ol = __shfl_xor_sync(gm, best, off, L);
        if (os > score || (os == score && ol < best)) { score = os; best = ol; }
      }
      const int act = best - 1;
      const bool wp = G::white_playing(env);
      AzNoise nz = {1.0, 0.0};
      if (G::STOCHASTIC) nz = az_env_noise<AzNoise>(nkey, (uint32_t)sims_done, (uint32_t)depth);
      const AzEnv nx = G::play(env, act, nz);
      const double wr = G::white_reward(nx);
      if (lane == 0) {
        pn[depth] = h;
        pm[depth] = (uint16_t)(act | ((wp != G::white_playing(nx)) ? 0x100 : 0));
        pr[depth] = wp ? wr : -wr;
      }
      depth++;
      env = nx;
      isroot = false;
    }
    if (need_leaf) break;
    sims_done++;
  }
  if (lane == 0) {
    p.sims_done[slot] = sims_done;
    p.total_sims[slot] += tsims;
    p.total_nodes[slot] += tnodes;
    if (sims_done < target) p.n_leaves[1] = 1;  // this tree is not finished (benign race: everyone writes 1)
  }
}

template <class G>
__global__ void __launch_bounds__(128) az_k_select(AzPool p) {
  // one slot per WARP: slots in the same warp would serialise on their divergent loop trip counts (4 dependent
  // pointer chains back to back instead of overlapped); lanes >= L of each warp retire immediately.
  // Tried in round 2 and reverted: keeping lanes [L, L + A) alive as helpers that compute every child's state and
  // prefetch its home line into L2 while the group waits for the parent's line -- 29.2 instead of 26.7 us per tick
  // (profiles/r02d): the chain is bound by the f64 divide / sqrt / play / hash work per level, not by the line loads.
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= p.S || lane >= G::LANES) return;
  az_select_slot<G>(p, slot, lane);
}

// Util.apply_temperature on the prior (src/util.jl:98-110, src/mcts.jl:157-161); sequential like the reference
template <int A>
__device__ __forceinline__ void az_prior_temperature(float* P, uint32_t legal, double tau) {
  if (tau == 1.0) return;
  if (tau == 0.0) {
    int k = -1;
    for (int i = 0; i < A; i++)
      if ((legal >> i) & 1u) if (k < 0 || P[i] > P[k]) k = i;
    for (int i = 0; i < A; i++) P[i] = (i == k) ? 1.0f : 0.0f;
    return;
  }
  double r[A], s = 0.0, it = 1.0 / tau;
  bool first = true;
  for (int i = 0; i < A; i++) {
    r[i] = 0.0;
    if (!((legal >> i) & 1u)) continue;
    r[i] = (P[i] > 0.0f) ? az_det_exp(it * az_det_log((double)P[i])) : 0.0;
    s = first ? r[i] : s + r[i];
    first = false;
  }
  for (int i = 0; i < A; i++) P[i] = ((legal >> i) & 1u) ? (float)(r[i] / s) : 0.0f;
}

// ---- expand + backup: insert the evaluated leaf (init_state_info, src/mcts.jl:157-174) and back its value up ----
template <class G>
__device__ __forceinline__ void az_expand_slot(const AzPool& p, const int slot, const int lane) {
  constexpr int L = G::LANES;
  constexpr int A = G::A;
  unsigned gm = az_group_mask<L>();
  if (!p.pending[slot]) return;
  const int row = p.leaf_row[slot];
  const AzEnv env = p.leaf_env[slot];
  const uint32_t pos = p.leaf_pos[slot];
  const int depth = p.depth[slot];
  const uint32_t tag = p.tag[slot];
  uint4* tab = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
  const uint32_t legal = G::legal_mask(env);
  const int a = lane - 1;
  float P = 0.0f;
  if (a >= 0 && a < A && ((legal >> a) & 1u)) P = p.batch_P[(size_t)row * A + a];
  if (az_c(p, slot).prior_temp != 1.0) {
    float Pv[A];
    for (int i = 0; i < A; i++) Pv[i] = __shfl_sync(gm, P, i + 1, L);
    az_prior_temperature<A>(Pv, legal, az_c(p, slot).prior_temp);
    if (a >= 0 && a < A) P = Pv[a];
  }
  const float V = p.batch_V[row];
  AzLine16 ln;
  ln.u = make_uint4(0, 0, 0, 0);
  if (lane == 0) { ln.key.a = env.a; ln.key.b = env.b | ((uint64_t)tag << 57); }
  else if (a < A) { ln.e.W = 0.0; ln.e.P = P; ln.e.N = 0; }
  // Never fill the table beyond 7/8: the overflow is reported to the host (AZ_ENOMEM) and the insert is skipped, so that
  // linear probing always finds an empty line and the kernels cannot spin on a full table.
  const int nc_new = p.node_count[slot] + 1;
  const bool room = (uint64_t)nc_new * 8 <= ((uint64_t)p.cap_mask + 1) * 7;
  if (room) tab[(size_t)pos * L + lane] = ln.u;
  az_backup<G>(p, slot, lane, gm, tab, depth, (double)V);
  if (lane == 0) {
    if (room) p.node_count[slot] = nc_new;
    else p.flags[0] = 1;
    p.total_nodes[slot] += depth;
    p.sims_done[slot] += 1;
    p.pending[slot] = 0;
    atomicAdd((unsigned long long*)p.expansions, 1ull);
  }
}

template <class G>
__global__ void __launch_bounds__(128) az_k_expand_backup(AzPool p) {
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= p.S || lane >= G::LANES) return;
  az_expand_slot<G>(p, slot, lane);
}
// expand + backup of the evaluated leaf, then straight on to the tree's next simulation (explore! loop, src/mcts.jl:239-245).
// Both kernels are bound by their LONGEST per-tree chain; fused, a launch lasts max_i(expand_i + select_i) instead of
// max_i(expand_i) + max_i(select_i), and one launch per tick disappears.  Used by az_mcts_run; the self-play loop keeps the
// two kernels apart because az_k_move runs between them.
template <class G>
__global__ void __launch_bounds__(128) az_k_expand_select(AzPool p) {
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= p.S || lane >= G::LANES) return;
  az_expand_slot<G>(p, slot, lane);
  __syncwarp(az_group_mask<G::LANES>());   // lane 0's updates of the slot's scalars (pending, sims_done) are visible to the group
  az_select_slot<G>(p, slot, lane);
}

// ---- loop condition of the device-driven explore loop (a CUDA-graph WHILE node, csrc/az_engine.cu Mcts::run_device_loop):
// after a tick, continue while a leaf is pending or a tree still has simulations to run, up to max_ticks.  No host code runs
// between the first select and the last backup of an explore!: this replaces the Julia scheduler (worker tasks + inference
// server, src/simulations.jl:207-244, src/batchifier.jl:47-81) by a loop the GPU drives itself.
__global__ void az_k_loop_cond(AzPool p, cudaGraphConditionalHandle handle, int32_t* __restrict__ tick_count, int max_ticks) {
  const int t = *tick_count + 1;
  *tick_count = t;
  const bool more = (p.n_leaves[0] | p.n_leaves[1] | p.n_leaves[2]) != 0;
  cudaGraphSetConditional(handle, (more && t < max_ticks) ? 1u : 0u);
}

// ---- MCTS.RolloutOracle (src/mcts.jl:27-60): uniform prior, value = discounted return of ONE random playout from the state.
// Julia's global rand() cannot be reproduced, so the playout's action draws come from the Philox stream keyed by
// (seed, hash of the state, ply of the playout): the oracle is a deterministic function of the state, identical on the CPU
// oracle (oracle/az_oracle.c oz_rollout_oracle) and here.  Rewards are folded from the end exactly like the recursion
// `wr + gamma * rollout!(game, gamma)` (:42-50).  One thread per leaf (divergent playouts of <= MAX_PLIES steps).
template <class G>
__global__ void az_k_rollout(const AzEnv* __restrict__ envs, const int32_t* __restrict__ n_rows, float* __restrict__ P,
                             float* __restrict__ V, double gamma, uint64_t seed) {
  constexpr int A = G::A;
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= *n_rows) return;
  AzEnv env = envs[row];
  {
    const uint32_t legal = G::legal_mask(env);
    const int n = __popc(legal);
    const float pu = (float)(1.0 / (double)n);
    for (int i = 0; i < A; i++) P[(size_t)row * A + i] = ((legal >> i) & 1u) ? pu : 0.0f;
  }
  const bool wp = G::white_playing(env);
  const uint64_t h0 = az_splitmix(env.a ^ az_splitmix(env.b));
  double r[G::MAX_PLIES + 1];
  int steps = 0;
  const AzNoise nz = {1.0, 0.0};
  while (steps <= G::MAX_PLIES) {
    const uint32_t legal = G::legal_mask(env);
    const int n = __popc(legal);
    uint32_t o[4];
    az_philox(seed, (uint32_t)steps, AZ_PURPOSE_ROLLOUT, (uint32_t)h0, (uint32_t)(h0 >> 32), o);
    int k = (int)(((uint64_t)o[0] * (uint64_t)n) >> 32);   // uniform over the n available actions (rand(available_actions), :43)
    int act = 0;
    for (int i = 0; i < A; i++)
      if ((legal >> i) & 1u) { if (k == 0) { act = i; break; } k--; }
    env = G::play(env, act, nz);
    r[steps++] = G::white_reward(env);
    if (G::terminated(env)) break;
  }
  double wr = r[steps - 1];
  for (int i = steps - 2; i >= 0; i--) wr = r[i] + gamma * wr;
  V[row] = (float)(wp ? wr : -wr);
}

// ---- built-in oracles: MCTS.RandomOracle (src/mcts.jl:62-72) and the deterministic hash pseudo-network ----
template <class G, int KIND>
__global__ void az_k_oracle(const AzEnv* __restrict__ envs, const int32_t* __restrict__ n_rows, float* __restrict__ P,
                            float* __restrict__ V) {
  constexpr int A = G::A;
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= *n_rows) return;
  const AzEnv env = envs[row];
  const uint32_t legal = G::legal_mask(env);
  if (KIND == 0) {
    int n = __popc(legal);
    float pu = (float)(1.0 / (double)n);
    for (int i = 0; i < A; i++) P[(size_t)row * A + i] = ((legal >> i) & 1u) ? pu : 0.0f;
    V[row] = 0.0f;
  } else {
    uint64_t h0 = az_splitmix(env.a ^ az_splitmix(env.b));
    uint32_t raw[A], sum = 0;
    for (int i = 0; i < A; i++) {
      raw[i] = 0;
      if (!((legal >> i) & 1u)) continue;
      uint64_t hi = az_splitmix(h0 + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull);
      raw[i] = 1u + (uint32_t)(hi >> 48);
      sum += raw[i];
    }
    for (int i = 0; i < A; i++) P[(size_t)row * A + i] = ((legal >> i) & 1u) ? (float)raw[i] / (float)sum : 0.0f;
    V[row] = ((float)(int)(h0 >> 48) - 32768.0f) / 32768.0f;
  }
}

// ---- root statistics (action-indexed) -------------------------------------------------------------------
template <class G>
__global__ void az_k_root_stats(AzPool p, int64_t* N, double* W, float* P) {
  constexpr int L = G::LANES;
  constexpr int A = G::A;
  // one slot per WARP: slots in the same warp would serialise on their divergent loop trip counts (4 dependent
  // pointer chains back to back instead of overlapped); lanes >= L of each warp retire immediately
  int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (slot >= p.S || lane >= L) return;
  unsigned gm = az_group_mask<L>();
  const AzEnv root = p.root[slot];
  const uint4* tab = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
  uint32_t h;
  AzLine16 ln;
  int st = G::terminated(root) ? 2 : az_probe<G>(tab, p.cap_mask, p.tag[slot], root, lane, gm, h, ln);
  int a = lane - 1;
  if (a >= 0 && a < A) {
    bool ok = (st == 1) && ((G::legal_mask(root) >> a) & 1u);
    N[(size_t)slot * A + a] = ok ? (int64_t)ln.e.N : 0;
    W[(size_t)slot * A + a] = ok ? ln.e.W : 0.0;
    P[(size_t)slot * A + a] = ok ? ln.e.P : 0.0f;
  }
}

// MCTS.reset! (src/mcts.jl:278-281): bump the generation tag; clear the table only when the 6-bit generation wraps
template <class G>
__global__ void az_k_reset(AzPool p) {
  constexpr int L = G::LANES;
  int slot = blockIdx.x;
  if (slot >= p.S) return;
  uint32_t tag = p.tag[slot];
  uint32_t gen = (tag & 63u) + 1u;
  if (gen == 64u) {
    uint4* tab = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
    size_t n = ((size_t)p.cap_mask + 1) * L;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) tab[i] = make_uint4(0, 0, 0, 0);
    gen = 1u;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.tag[slot] = 64u | gen;
    p.node_count[slot] = 0;
    p.pending[slot] = 0;
  }
}

// =====================================================================================================
// Self-play: per-move kernel = MCTS.policy + temperature + categorical sample + play! + trace record
// (src/play.jl:298-315, src/mcts.jl:255-271, src/util.jl:68-110, src/schedule.jl:64-80) and, at game end,
// push_trace! (src/memory.jl:74-87), self_play_measurements (src/training.jl:269-273) and the worker's
// reset_every / next-game logic (src/simulations.jl:221-241).
// =====================================================================================================
struct AzSelfPlay {
  uint64_t seed;
  int64_t first_game;     // global index of local game 0
  int32_t num_games;      // local games to play
  int32_t nsims;          // MctsPlayer.niters of player 0 (src/play.jl:156-165)
  int32_t nsims1;         // duel: of player 1
  int32_t netonly;        // player 0 is a NetworkPlayer (src/play.jl:226-235): nsims = 1 (the root evaluation), pi = the root's priors
  int32_t netonly1;       // duel: player 1 is a NetworkPlayer
  int32_t minmax1;        // duel: player 1 is a MinMax.Player (src/minmax.jl:72-81): no tree, no oracle
  int32_t mm_depth;       // MinMax.Player.depth
  int32_t mm_amplify;     // MinMax.Player.amplify_rewards
  double mm_tau;          // MinMax.Player.τ (inside think; the move temperature of such a player is 1, src/play.jl:37-39)
  double mm_gamma;        // MinMax.Player.gamma
  int32_t reset_every;
  int32_t max_plies;
  int32_t sched_n;        // temperature schedule of player 0 (MctsPlayer.τ)
  int32_t sched_xs[8];
  double sched_ys[8];
  int32_t sched1_n;       // duel: of player 1
  int32_t sched1_xs[8];
  double sched1_ys[8];
  double flip_p;          // SimParams.flip_probability (src/play.jl:305-307)
  int32_t duel;           // TwoPlayers: two trees per worker (player 0 = `white` argument of TwoPlayers, 1 = `black`)
  int32_t alternate;      // SimParams.alternate_colors (src/simulations.jl:224-230)
  int32_t W;              // number of workers (game slots); trees = W * (duel ? 2 : 1)
  // per worker
  int32_t* game_of_slot;   // local game index or -1
  int32_t* move_of_slot;
  int32_t* games_on_slot;
  // per (game, ply) rows, stride max_plies
  AzEnv* s_env;            // trace.states[i] (before the optional symmetry of move i)
  AzEnv* s_root;           // the state the player thought on (s_pi and the mask are in its frame)
  AzEnv* g_final;          // per game: last state of the trace
  double* s_pi;            // [A] MCTS.policy output (Float64 in the reference Trace), zero on illegal actions
  int32_t* s_action;
  double* s_reward;
  double* s_z;
  float* s_t;
  // per game
  int32_t* g_moves;
  double* g_edepth;
  int64_t* g_nodes;
  int32_t* games_done;     // [1]
  int32_t* active_slots;   // [1] (unused)
  int32_t* next_game;      // [1] shared game counter of Util.mapreduce (src/util.jl:172-186)
  uint8_t* want_game;      // [S] slot finished its game this tick and asks for the next one
};

__device__ __forceinline__ double az_schedule(const AzSelfPlay& sp, int player, int i) {  // src/schedule.jl:64-80
  const int n = player ? sp.sched1_n : sp.sched_n;
  const int32_t* xs = player ? sp.sched1_xs : sp.sched_xs;
  const double* ys = player ? sp.sched1_ys : sp.sched_ys;
  int pt = -1;
  for (int k = 0; k < n; k++) if (xs[k] <= i) pt = k;
  if (pt < 0) return ys[0];
  if (pt == n - 1) return ys[n - 1];
  double x0 = xs[pt], y0 = ys[pt], x1 = xs[pt + 1], y1 = ys[pt + 1];
  return y0 + ((y1 - y0) / (x1 - x0)) * ((double)i - x0);
}

// colors_flipped of game `game` (0-based global index; sim_id = game + 1): src/simulations.jl:224-226
__device__ __forceinline__ bool az_colors_flipped(const AzSelfPlay& sp, int64_t game) {
  return sp.duel && sp.alternate && (((game + 1) & 1) == 1);
}
// the tree that thinks on `e` for worker w (think(::TwoPlayers), src/play.jl:258-264; flipped_colors :256)
template <class G>
__device__ __forceinline__ int az_tree_of(const AzSelfPlay& sp, int w, const AzEnv& e, int64_t game) {
  if (!sp.duel) return w;
  const bool white = G::white_playing(e);
  const bool fl = az_colors_flipped(sp, game);
  return 2 * w + ((white != fl) ? 0 : 1);
}

// Start of a turn of play_game (src/play.jl:301-308): record the trace state, apply the optional random symmetry
// (GI.apply_random_symmetry!, src/game.jl:329-336), pick the thinking player's tree and set up explore!.
template <class G>
__device__ void az_begin_move(const AzPool& p, const AzSelfPlay& sp, int w, const AzEnv& state, int g, int64_t game, int move) {
  constexpr int A = G::A;
  sp.s_env[(size_t)g * sp.max_plies + move] = state;
  AzEnv root = state;
  if (G::NSYM > 0 && sp.flip_p != 0.0) {
    const double u = az_u01(az_stream_u64(sp.seed, (uint64_t)game, (uint32_t)move, AZ_PURPOSE_SYMMETRY, 0));
    if (u < sp.flip_p) {
      const int j = (int)(az_stream_u64(sp.seed, (uint64_t)game, (uint32_t)move, AZ_PURPOSE_SYMMETRY, 1) % (uint64_t)(G::NSYM > 0 ? G::NSYM : 1));
      root = G::symmetry(state, j);
    }
  }
  const int slot = az_tree_of<G>(sp, w, root, game);
  p.root[slot] = root;
  p.sims_done[slot] = 0;
  p.sims_target[slot] = (sp.duel && (slot & 1)) ? sp.nsims1 : sp.nsims;
  p.status[slot] = 1;
  p.pending[slot] = 0;
  if (sp.duel && (slot & 1) && sp.minmax1) { p.sims_target[slot] = 0; return; }   // thinks in az_k_minmax_think of the next tick
  if (G::STOCHASTIC) { p.noise_game[slot] = game; p.noise_move[slot] = move; }
  double eta[A];
  int n = __popc(G::legal_mask(root));
  az_dirichlet(sp.seed, (uint64_t)game, (uint32_t)move, n, az_c(p, slot).alpha, eta);  // drawn even if eps == 0 (src/mcts.jl:240)
  for (int i = 0; i < n; i++) p.eta[(size_t)slot * A + i] = eta[i];
}

// End of a game on a worker: self_play_measurements (src/training.jl:269-273, measured before the reset), reset_every
// (src/simulations.jl:235-237; reset!(::TwoPlayers) resets both trees, src/play.jl:274-277).  The worker then asks
// az_k_assign for its next game.
template <class G>
__device__ void az_record_game_end(AzPool& p, AzSelfPlay& sp, int w, int g, int n_moves, const AzEnv& last) {
  constexpr int L = G::LANES;
  const int nt = sp.duel ? 2 : 1, t0 = sp.duel ? 2 * w : w;
  sp.g_moves[g] = n_moves;
  sp.g_final[g] = last;
  int64_t nodes = 0, ts = 0, tn = 0;
  for (int k = 0; k < nt; k++) {
    if (k ? (sp.netonly1 || sp.minmax1) : sp.netonly) continue;   // no MCTS.Env to measure (src/training.jl:269-273)
    nodes += p.node_count[t0 + k]; ts += p.total_sims[t0 + k]; tn += p.total_nodes[t0 + k];
  }
  sp.g_nodes[g] = nodes;
  sp.g_edepth[g] = ts == 0 ? 0.0 : (double)tn / (double)ts;
  atomicAdd(sp.games_done, 1);
  const int gos = sp.games_on_slot[w] + 1;
  sp.games_on_slot[w] = gos;
  for (int k = 0; k < nt; k++) {
    const int slot = t0 + k;
    if (sp.reset_every > 0 && gos % sp.reset_every == 0) {
      uint32_t gen = (p.tag[slot] & 63u) + 1u;
      if (gen == 64u) {
        uint4* wt = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
        size_t cnt = ((size_t)p.cap_mask + 1) * L;
        for (size_t i = 0; i < cnt; i++) wt[i] = make_uint4(0, 0, 0, 0);
        gen = 1u;
      }
      p.tag[slot] = 64u | gen;
      p.node_count[slot] = 0;
    }
    p.status[slot] = 0;
    p.sims_target[slot] = 0;
  }
  sp.game_of_slot[w] = -1;
  sp.want_game[w] = 1;
}

// Dynamic game assignment = the shared counter of Util.mapreduce (src/util.jl:169-200): a worker that finishes a game
// takes the next unplayed game index.  Workers that finish in the same tick are served in slot order, which makes the
// game -> worker map deterministic (every move lasts exactly nsims ticks when select runs one simulation per call).
// One block; loops because a game whose initial state is already terminal (grid-world: RL.reset! may pick a reward
// cell) ends at once with an empty trace, exactly as play_game returns immediately (src/play.jl:301-304).
template <class G>
__global__ void __launch_bounds__(1024) az_k_assign(AzPool p, AzSelfPlay sp) {
  __shared__ int s_scan[1024];
  __shared__ int s_base, s_total, s_again;
  for (;;) {
    if (threadIdx.x == 0) { s_base = *sp.next_game; s_total = 0; s_again = 0; }
    __syncthreads();
    for (int start = 0; start < sp.W; start += blockDim.x) {
      const int slot = start + threadIdx.x;  // worker index
      const int want = (slot < sp.W && sp.want_game[slot]) ? 1 : 0;
      // inclusive scan of `want` over the block (slot order)
      s_scan[threadIdx.x] = want;
      __syncthreads();
      for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        int v = threadIdx.x >= (unsigned)off ? s_scan[threadIdx.x - off] : 0;
        __syncthreads();
        s_scan[threadIdx.x] += v;
        __syncthreads();
      }
      const int rank = s_total + s_scan[threadIdx.x] - want;
      const int chunk_total = s_scan[blockDim.x - 1];
      if (want) {
        const int ng = s_base + rank;
        sp.want_game[slot] = 0;
        if (ng < sp.num_games) {
          const AzEnv first = G::init_game(sp.seed, (uint64_t)(sp.first_game + ng));
          if (!G::terminated(first)) {
            sp.game_of_slot[slot] = ng;
            sp.move_of_slot[slot] = 0;
            az_begin_move<G>(p, sp, slot, first, ng, sp.first_game + ng, 0);
          } else {
            az_record_game_end<G>(p, sp, slot, ng, 0, first);  // empty game; asks again in the next round
            s_again = 1;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) s_total += chunk_total;
      __syncthreads();
    }
    if (threadIdx.x == 0) *sp.next_game = min(s_base + s_total, sp.num_games);
    __syncthreads();
    if (!s_again) break;
    __syncthreads();
  }
}

// =====================================================================================================
// MinMax baseline player (src/minmax.jl; Benchmark.MinMaxTS, src/benchmark.jl:178-196): exhaustive depth-limited search
// with GI.heuristic_value at the horizon.  The search itself (az_minmax_qvalue, az_minmax_policy) is host + device code in
// az_games.cuh, next to the heuristics.
// =====================================================================================================
// one thread per (tree, action): the root q-values of the MinMax player's trees whose turn it is, into p.eta (the Dirichlet
// buffer such a tree does not use); az_k_move turns them into the move distribution
template <class G>
__global__ void az_k_minmax_think(AzPool p, AzSelfPlay sp) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = t / G::A, a = t % G::A;
  if (slot >= p.S || !(slot & 1)) return;
  if (!(p.status[slot] && !p.pending[slot] && p.sims_done[slot] >= p.sims_target[slot])) return;
  const AzEnv root = p.root[slot];
  if (!((G::legal_mask(root) >> a) & 1u)) return;
  p.eta[(size_t)slot * G::A + a] = az_minmax_qvalue<G>(root, a, sp.mm_depth, sp.mm_amplify != 0, sp.mm_gamma);
}

// one thread per slot (runs once per move: scalar code, not performance critical)
template <class G>
__global__ void az_k_move(AzPool p, AzSelfPlay sp, int start_games) {
  constexpr int L = G::LANES;
  constexpr int A = G::A;
  int slot = blockIdx.x * blockDim.x + threadIdx.x;  // tree index; the thread of the tree that just finished thinking acts
  const int w = sp.duel ? (slot >> 1) : slot;        // worker
  if (start_games) {
    if (slot >= p.S) return;  // every worker starts by asking for a game (az_k_assign serves them in slot order)
    p.total_sims[slot] = 0;
    p.total_nodes[slot] = 0;
    p.status[slot] = 0;
    p.sims_target[slot] = 0;
    p.pending[slot] = 0;
    if (!sp.duel || (slot & 1) == 0) {
      sp.games_on_slot[w] = 0;
      sp.game_of_slot[w] = -1;
      sp.want_game[w] = 1;
    }
    return;
  }
  // In a duel the acting thread hands the turn to the sibling tree (lane ^ 1 of the same warp): every lane takes its
  // decision before any lane writes.
  const bool my_turn = slot < p.S && p.status[slot] && !p.pending[slot] && p.sims_done[slot] >= p.sims_target[slot];
  __syncwarp();
  if (!my_turn) return;
  const int g = sp.game_of_slot[w];
  const int move = sp.move_of_slot[w];
  const int64_t game = sp.first_game + g;
  const AzEnv root = p.root[slot];
  const uint32_t legal = G::legal_mask(root);
  int acts[A];
  double pi[A], pis[A];
  float pf[A];
  int n = 0;
  const bool minmax = sp.duel && (slot & 1) && sp.minmax1;
  if (minmax) {
    // think(::MinMax.Player): the q-values az_k_minmax_think left in p.eta
    double qs[A];
    for (int i = 0; i < A; i++)
      if ((legal >> i) & 1u) { acts[n] = i; qs[n] = p.eta[(size_t)slot * A + i]; n++; }
    az_minmax_policy(qs, n, sp.mm_tau, pi);
  } else {
    // MCTS.policy (src/mcts.jl:255-271): find the root line (single thread: read the L lanes one by one)
    const uint4* tab = p.nodes + (size_t)slot * ((size_t)p.cap_mask + 1) * L;
    uint32_t h = az_hash(root.a, root.b) & p.cap_mask;
    const uint32_t tag = p.tag[slot];
    for (;;) {
      AzLine16 k;
      k.u = tab[(size_t)h * L];
      if ((uint32_t)(k.key.b >> 57) == tag && k.key.a == root.a && (k.key.b & AZ_KEYB_MASK) == root.b) break;
      if ((uint32_t)(k.key.b >> 57) != tag) { p.flags[3] = 1; break; }  // cannot happen after explore!
      h = (h + 1) & p.cap_mask;
    }
    int64_t ntot = 0;
    // think(::NetworkPlayer) (src/play.jl:230-235): the oracle's policy over the available actions = the priors the root
    // line holds since this turn's single simulation (or an earlier visit) evaluated it; prior_temperature is 1 for such a
    // player
    const bool netonly = (sp.duel && (slot & 1)) ? sp.netonly1 != 0 : sp.netonly != 0;
    for (int i = 0; i < A; i++)
      if ((legal >> i) & 1u) {
        AzLine16 e;
        e.u = tab[(size_t)h * L + 1 + i];
        acts[n] = i;
        pi[n] = netonly ? (double)e.e.P : (double)e.e.N;
        ntot += e.e.N;
        n++;
      }
    if (!netonly) {
      double sum = 0.0;
      for (int i = 0; i < n; i++) { pi[i] = pi[i] / (double)ntot; sum = (i == 0) ? pi[0] : sum + pi[i]; }
      for (int i = 0; i < n; i++) pi[i] = pi[i] / sum;
    }
  }
  // temperature (src/play.jl:208-210,309-310; src/util.jl:98-110); a MinMax.Player keeps the default of 1 (src/play.jl:37-39)
  const double tau = minmax ? 1.0 : az_schedule(sp, (sp.duel && (slot & 1)) ? 1 : 0, move);
  if (tau == 1.0) { for (int i = 0; i < n; i++) pis[i] = pi[i]; }
  else if (tau == 0.0) {
    int k = 0;
    for (int i = 1; i < n; i++) if (pi[i] > pi[k]) k = i;
    for (int i = 0; i < n; i++) pis[i] = (i == k) ? 1.0 : 0.0;
  } else {
    double it = 1.0 / tau, s = 0.0;
    for (int i = 0; i < n; i++) pis[i] = (pi[i] > 0.0) ? az_det_exp(it * az_det_log(pi[i])) : 0.0;
    for (int i = 0; i < n; i++) s = (i == 0) ? pis[0] : s + pis[i];
    for (int i = 0; i < n; i++) pis[i] = pis[i] / s;
  }
  // fix_probvec + rand(Categorical) (src/util.jl:68-90)
  float fs = 0.0f;
  for (int i = 0; i < n; i++) { pf[i] = (float)pis[i]; fs = (i == 0) ? pf[0] : fs + pf[i]; }
  {
    const float rtol = 3.4526698e-4f;
    float d = fabsf(fs - 1.0f), m = fabsf(fs) > 1.0f ? fabsf(fs) : 1.0f;
    bool approx = (fs == 1.0f) || (isfinite(fs) && d <= rtol * m);
    if (!approx) {
      if (fs == 0.0f) for (int i = 0; i < n; i++) pf[i] = 1.0f / (float)n;
      else for (int i = 0; i < n; i++) pf[i] = pf[i] / fs;
    }
  }
  const float u = az_uniform_f32(sp.seed, (uint64_t)game, (uint32_t)move, AZ_PURPOSE_CATEGORICAL, 0);
  float cp = pf[0];
  int k = 0;
  while (cp <= u && k < n - 1) { k++; cp = cp + pf[k]; }
  const int act = acts[k];
  // record (trace.jl:35-39) and play
  const size_t rowi = (size_t)g * sp.max_plies + move;
  sp.s_root[rowi] = root;
  for (int i = 0; i < A; i++) sp.s_pi[rowi * A + i] = 0.0;
  for (int i = 0; i < n; i++) sp.s_pi[rowi * A + acts[i]] = pi[i];
  sp.s_action[rowi] = act;
  AzNoise rnz = {1.0, 0.0};
  if (G::STOCHASTIC) { AzNoiseKey rk = {sp.seed, (uint64_t)game, (uint32_t)move}; rnz = az_env_noise<AzNoise>(rk, AZ_REAL_MOVE, 0u); }
  const AzEnv nx = G::play(root, act, rnz);
  sp.s_reward[rowi] = G::white_reward(nx);
  const int nm = move + 1;
  if (G::terminated(nx) || nm >= sp.max_plies) {
    // push_trace! (src/memory.jl:74-87).  The reference only pushes self-play traces (one player, its mcts gamma); in a duel
    // of two different players the z rows are an extra and use the gamma of the player who held white in this game.
    const double zg = (sp.duel && az_colors_flipped(sp, game)) ? p.c1.gamma : p.c.gamma;
    double wr = 0.0;
    for (int i = nm - 1; i >= 0; i--) {
      const size_t ri = (size_t)g * sp.max_plies + i;
      wr = zg * wr + sp.s_reward[ri];
      sp.s_z[ri] = G::white_playing(sp.s_env[ri]) ? wr : -wr;
      sp.s_t[ri] = (float)(nm - i);
    }
    az_record_game_end<G>(p, sp, w, g, nm, nx);
  } else {
    sp.move_of_slot[w] = nm;
    p.status[slot] = 0;       // this player's turn is over; az_begin_move activates the tree of the player to move
    p.sims_target[slot] = 0;
    az_begin_move<G>(p, sp, w, nx, g, game, nm);
  }
}

// push_trace! rows of a finished run -> device-resident TrainingSamples (src/memory.jl:74-87), ordered by (game, ply).
// The reference pairs trace.states[i] with trace.policies[i], a vector over the legal actions of the state the player
// THOUGHT on; convert_sample later scatters it over the mask of trace.states[i] (src/learning.jl:31-34).  With
// flip_probability > 0 the two frames differ, so the compact policy is re-scattered here the same way.
// Loop condition of the device-driven self-play loop (WHILE node, SelfPlay::loop): continue until every game of the run has
// finished or a kernel raised an error flag.  Progress (games finished so far) goes to host-mapped pinned memory every tick so
// that az_selfplay_poll / the `game_simulated` callback read it without touching the stream.
__global__ void az_k_selfplay_cond(AzPool p, AzSelfPlay sp, cudaGraphConditionalHandle handle, volatile int32_t* __restrict__ host_progress) {
  const int done = *sp.games_done;
  const int err = p.flags[0] | p.flags[2] | p.flags[3];
  host_progress[0] = done;
  __threadfence_system();
  cudaGraphSetConditional(handle, (done < sp.num_games && !err) ? 1u : 0u);
}

template <class G>
__global__ void az_k_export_samples(AzSelfPlay sp, int num_games, const int64_t* __restrict__ goff, AzEnv* oenv, double* opi, double* oz,
                                    double* ot, int32_t* ocnt) {
  constexpr int A = G::A;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)num_games * sp.max_plies) return;
  const int g = (int)(r / sp.max_plies), i = (int)(r % sp.max_plies);
  if (i >= sp.g_moves[g]) return;
  const int64_t k = goff[g] + i;
  const AzEnv st = sp.s_env[r], th = sp.s_root[r];
  double c[A];
  int n = 0;
  const uint32_t lt = G::legal_mask(th), ls = G::legal_mask(st);
  for (int a = 0; a < A; a++) if ((lt >> a) & 1u) c[n++] = sp.s_pi[r * A + a];
  int m = 0;
  for (int a = 0; a < A; a++) opi[k * A + a] = ((ls >> a) & 1u) && m < n ? c[m++] : 0.0;
  oenv[k] = st;
  oz[k] = sp.s_z[r];
  ot[k] = (double)sp.s_t[r];
  ocnt[k] = 1;
}
