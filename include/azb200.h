/*
 * azb200.h -- C ABI of the B200-native self-play engine (libazb200.so).
 *
 * Drop-in boundary for the self-play hot path of AlphaZero.jl.  Every entry point
 * names the reference interface it replaces (paths relative to the reference root);
 * INTEGRATION.md shows the Julia `ccall` stub for each.  Plain C: opaque handles,
 * caller-allocated HOST buffers, int32 status codes (0 = OK), no exception crosses.
 *
 * State byte formats (what the Julia shim passes; one state = az_game_state_bytes):
 *   connect-four (43 B): cells[col + 7*row] in {0 empty, 1 white, 2 black}, curplayer {1,2}
 *                        -- the in-memory layout of (board::SMatrix{7,6,UInt8}, curplayer::UInt8),
 *                        games/connect-four/game.jl:19-22
 *   tictactoe    (10 B): cells[pos], pos = (y-1)*3 + (x-1), {0,1 white,2 black}; curplayer {1,2}
 *                        (games/tictactoe/game.jl:9-14: Bool/Nothing cells mapped to bytes)
 *   mancala      (15 B): stores[2], houses[(player-1) + 2*(num-1)] (12 B), curplayer {1,2}
 *                        (games/mancala/game.jl:22-25)
 *   grid-world   ( 2 B): x, y in 1..10 (games/grid-world/game.jl:20,57; the step counter is not part of the state)
 * Actions are 0-based here (Julia action a  <->  a-1).
 */
#ifndef AZB200_H
#define AZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZ_ABI_VERSION 1
#define AZ_MAX_ACTIONS 9
#define AZ_MAX_SCHEDULE 8

enum az_status { AZ_OK = 0, AZ_EINVAL = 1, AZ_ECUDA = 2, AZ_ENOMEM = 3, AZ_ESTATE = 4, AZ_EUNSUPPORTED = 5 };

typedef struct az_ctx az_ctx;
typedef struct az_net az_net;
typedef struct az_mcts az_mcts;
typedef struct az_selfplay az_selfplay;
typedef struct az_samples az_samples;
typedef struct az_comm az_comm;

/* ---- library / context ---------------------------------------------------------------- */
int32_t az_version(void);
/* one context per GPU per process (the rank-per-GPU analogue of a Distributed worker, src/simulations.jl:268-281) */
int32_t az_ctx_create(int32_t device, az_ctx** out);
int32_t az_ctx_destroy(az_ctx* ctx);
/* message of the last failing call on this context (or of az_ctx_create when ctx == NULL) */
const char* az_last_error(az_ctx* ctx);
int32_t az_ctx_synchronize(az_ctx* ctx);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
int64_t az_ctx_num_launches(az_ctx* ctx);

/* ---- games: GI.AbstractGameSpec queries (src/game.jl:34-120; names = src/examples.jl:17-21) ---- */
int32_t az_game_lookup(const char* name); /* "connect-four" | "tictactoe" | "mancala" | "grid-world"; < 0 if unknown */
int32_t az_game_num_actions(int32_t game);          /* GI.num_actions */
int32_t az_game_state_bytes(int32_t game);
int32_t az_game_state_dim(int32_t game, int32_t dim[3]); /* GI.state_dim */
int32_t az_game_max_plies(int32_t game);
/* GI.vectorize_state, GI.actions_mask(GI.init(spec, s)), GI.game_terminated, GI.white_reward, GI.play! on host
   bytes, evaluated by the SAME inline functions the kernels use (host build of csrc/az_games.cuh) */
int32_t az_game_vectorize_state(int32_t game, const uint8_t* state, float* x);
int32_t az_game_actions_mask(int32_t game, const uint8_t* state, uint8_t* mask);
int32_t az_game_play(int32_t game, const uint8_t* state, int32_t action, uint8_t* next_state, int32_t* terminated,
                     double* white_reward);
int32_t az_game_init_state(int32_t game, uint8_t* state);
/* synthetic random positions: k ~ U{0..max_plies} random legal plies, terminal positions rejected */
int32_t az_game_random_positions(int32_t game, uint64_t seed, uint64_t first_stream, int32_t n, int32_t max_plies,
                                 uint8_t* states);

/* ---- parameters: MctsParams (src/params.jl:49-57), SimParams (src/params.jl:92-101) ------ */
typedef struct {
  double gamma;
  double cpuct;
  int32_t num_iters_per_turn; /* > 0: MctsPlayer.niters.  0 (az_selfplay_create* only): the player is a NetworkPlayer under
                                 PlayerWithTemperature = Benchmark.NetworkOnly (src/play.jl:226-235, :112-127,
                                 src/benchmark.jl:161-176): no search, the move distribution and the recorded policy are the
                                 oracle's policy at the root; only the temperature schedule of this block is used */
  int32_t temperature_n; /* PLSchedule points (ConstSchedule = 1 point), src/schedule.jl:64-80 */
  double dirichlet_noise_eps;
  double dirichlet_noise_alpha;
  double prior_temperature;
  int32_t temperature_xs[AZ_MAX_SCHEDULE];
  double temperature_ys[AZ_MAX_SCHEDULE];
} az_mcts_params;

/* MinMax.Player (src/minmax.jl:72-81; Benchmark.MinMaxTS, src/benchmark.jl:178-196) */
typedef struct {
  int32_t depth;           /* 1..8 */
  int32_t amplify_rewards; /* non-zero rewards become +-Inf (src/minmax.jl:14) */
  double tau;              /* temperature inside think (src/minmax.jl:95-107) */
  double gamma;            /* MinMax.Player's gamma (1 in Benchmark.MinMaxTS) */
} az_minmax_params;
/* GI.heuristic_value(GI.init(spec, s)) (games/<name>/game.jl) and think(::MinMax.Player, game) (src/minmax.jl:83-114) on host bytes,
   evaluated by the same inline functions the kernels use: q[A] root q-values, pi[A] the move distribution (both zero on
   unavailable actions; either may be NULL).  AZ_EUNSUPPORTED for grid-world. */
int32_t az_game_heuristic_value(int32_t game, const uint8_t* state, double* value);
int32_t az_game_minmax_think(int32_t game, const uint8_t* state, const az_minmax_params* player, double* q, double* pi);

typedef struct {
  int32_t num_games;
  int32_t num_workers; /* concurrent game slots on this GPU (= worker tasks, src/simulations.jl:217) */
  int32_t batch_size;  /* accepted for API parity (must be <= num_workers, src/batchifier.jl:48); the engine batches every pending
                          leaf of a tick, which is what batch_size = num_workers gives the reference */
  int32_t fill_batches; /* accepted and irrelevant: the reference pads a short batch with copies of its first state so that the
                           network always sees one batch shape (src/batchifier.jl:66-70, results unchanged); the device
                           network evaluates exactly the pending leaves */
  int32_t reset_every; /* <= 0: never (Julia `nothing`) */
  int32_t alternate_colors; /* duels only: odd sim_id (1-based game index) swaps the players' colours (src/simulations.jl:224-230) */
  double flip_probability;  /* random GI.symmetries image before each turn (src/play.jl:305-307); needs a game that declares
                               symmetries (connect-four, tictactoe), else AZ_EINVAL (src/params.jl:377-381) */
} az_sim_params;

/* ---- networks: Network interface (src/networks/network.jl), ResNet (architectures/resnet.jl) ---- */
enum az_net_kind {
  AZ_NET_UNIFORM = 0, /* MCTS.RandomOracle, src/mcts.jl:62-72 */
  AZ_NET_SYNTH = 1,   /* deterministic hash pseudo-network (parity tests: bit-exact on CPU and GPU) */
  AZ_NET_RESNET = 2,
  AZ_NET_SIMPLENET = 3,
  AZ_NET_ROLLOUT = 4  /* MCTS.RolloutOracle, src/mcts.jl:27-60 */
};
typedef struct { /* ResNetHP, src/networks/architectures/resnet.jl:30-37 */
  int32_t num_blocks;
  int32_t num_filters;
  int32_t conv_kernel_size[2];
  int32_t num_policy_head_filters;
  int32_t num_value_head_filters;
  float batch_norm_momentum;
} az_resnet_hp;
typedef struct { /* SimpleNetHP, src/networks/architectures/simplenet.jl:15-22 */
  int32_t width;
  int32_t depth_common;
  int32_t depth_phead;
  int32_t depth_vhead;
  int32_t use_batch_norm;
  float batch_norm_momentum;
} az_simplenet_hp;

int32_t az_net_create_oracle(az_ctx* ctx, int32_t kind, int32_t game, az_net** out);
/* MCTS.RolloutOracle(gspec, gamma) (src/mcts.jl:27-60): uniform prior over the available actions, value = discounted return of
   one random playout from the state (the oracle of Benchmark.MctsRollouts, src/benchmark.jl:134-147).  The playout's action
   draws come from the engine's Philox stream keyed by (seed, state, ply) -- a deterministic function of the state, since
   Julia's global rand() cannot be reproduced.  Deterministic games only (AZ_EUNSUPPORTED for grid-world). */
int32_t az_net_create_rollout(az_ctx* ctx, int32_t game, double gamma, uint64_t seed, az_net** out);
int32_t az_net_create_resnet(az_ctx* ctx, int32_t game, const az_resnet_hp* hp, az_net** out);
int32_t az_net_create_simplenet(az_ctx* ctx, int32_t game, const az_simplenet_hp* hp, az_net** out);
/* number of float32 values of the parameter blob, in Flux order (see DESIGN.md "weight blob"):
   per Conv: W[kw,kh,cin,cout] (column-major) then b[cout]; per BatchNorm: gamma, beta, mu, sigma2;
   per Dense: W[out,in] (column-major) then b[out]; layer order = common, vhead, phead */
int32_t az_net_num_params(az_net* net, int64_t* n);
/* replaces Network.copy(bestnn; on_gpu=true, test_mode=true) (src/training.jl:278-279): uploads + folds BN */
int32_t az_net_load(az_net* net, const float* blob, int64_t n);
/* the same with the blob already in DEVICE memory of the network's GPU (e.g. the flat parameter vector of the model that was
   just trained there: Flux through CUDA.jl, or alphazero.jl_b200/learning.py): BatchNorm is folded and the kernels' fp16
   layouts are written by device kernels, no host round trip; the caller's buffer is not retained */
int32_t az_net_load_device(az_net* net, const float* device_blob, int64_t n);
/* Network.evaluate_batch / forward_normalized (src/networks/network.jl:264-271,308-315):
   P is A-wide (masked, renormalised, zero on illegal), V[B], Pinvalid[B] (may be NULL) */
int32_t az_net_forward(az_net* net, const uint8_t* states, int32_t B, float* P, float* V, float* Pinvalid);
/* parity hook for Network.forward (src/networks/network.jl:119-132, architectures/resnet.jl:83-90) BEFORE its output
   non-linearities: policy_logits[B*A] = input of the policy head's softmax (every action, no mask), value_pre[B] = input
   of the value head's tanh; either may be NULL.  ResNet and SimpleNet only. */
int32_t az_net_forward_logits(az_net* net, const uint8_t* states, int32_t B, float* policy_logits, float* value_pre);
/* device-side timing (CUDA events on the context's stream) of the network launches, for bench.py's roofline:
   tower_ms = time inside the conv-tower kernels, tower_launches = number of tower kernel launches,
   total_ms = stem + tower + heads, evals = number of batched evaluations since profiling was enabled */
int32_t az_net_set_profiling(az_net* net, int32_t enable);
int32_t az_net_get_profile(az_net* net, double* tower_ms, int64_t* tower_launches, double* total_ms, int64_t* evals);
int32_t az_net_destroy(az_net* net);

/* ---- MCTS.Env pool: n_trees independent MCTS.Env (src/mcts.jl:124-151) on one GPU ---------- */
int32_t az_mcts_create(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* p, int32_t n_trees,
                       int32_t capacity_nodes_per_tree, az_mcts** out);
/* roots[i] = GI.current_state of tree i's game; eta: n_trees x A doubles, compact over legal actions in ascending
   order (src/mcts.jl:228-232), or NULL (then dirichlet_noise_eps must be 0) */
int32_t az_mcts_set_roots(az_mcts* m, const uint8_t* root_states, const double* eta);
/* stochastic environments only (grid-world): the in-tree environment noise of tree i is the stream
   (seed, games[i], moves[i], simulation index, depth) -- replaces Julia's global rand() in act! (games/grid-world/game.jl:45-46) */
int32_t az_mcts_set_noise(az_mcts* m, uint64_t seed, const int64_t* games, const int32_t* moves);
/* MCTS.explore!(env, game, nsims) on every tree (src/mcts.jl:239-245); roots resident on device */
int32_t az_mcts_run(az_mcts* m, int32_t nsims);
/* set_roots + run + root_stats in one call with host buffers (the `think` seam, src/play.jl:196-206) */
int32_t az_mcts_explore(az_mcts* m, const uint8_t* root_states, const double* eta, int32_t nsims, int64_t* N, double* W,
                        float* P);
/* per-tree root ActionStats, action-indexed (A wide, zeros on illegal) */
int32_t az_mcts_root_stats(az_mcts* m, int64_t* N, double* W, float* P);
/* MCTS.policy (src/mcts.jl:255-271), A-wide float64 */
int32_t az_mcts_policy(az_mcts* m, double* pi);
/* MCTS.reset! (src/mcts.jl:278-281) */
int32_t az_mcts_reset(az_mcts* m);
/* total_simulations, total_nodes_traversed, length(tree) per tree (src/mcts.jl:142-143,293-321); any may be NULL */
int32_t az_mcts_counters(az_mcts* m, int64_t* total_simulations, int64_t* total_nodes_traversed, int64_t* num_nodes);
/* device-side timing of the last az_mcts_run: total ms, ms inside the network forward, ticks, expansions */
int32_t az_mcts_last_timing(az_mcts* m, double* ms_total, double* ms_network, int64_t* ticks, int64_t* expansions);
/* per-kernel device timing of the tree kernels for bench.py's roofline_tree: while enabled, az_mcts_run records CUDA events
   around az_k_select, the network evaluation and az_k_expand_backup of every tick (graph replay off); get_profile returns
   the sums since the last call: select_ms, expand_ms (expand + backup), net_ms, ticks */
int32_t az_mcts_set_profiling(az_mcts* m, int32_t enable);
int32_t az_mcts_get_profile(az_mcts* m, double* select_ms, double* expand_ms, double* net_ms, int64_t* ticks);
int32_t az_mcts_destroy(az_mcts* m);

/* ---- self-play: simulate(simulator, gspec, SimParams) (src/simulations.jl:207-244) ------------ */
int32_t az_selfplay_create(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* mp, const az_sim_params* sp,
                           uint64_t seed, az_selfplay** out);
/* pit_networks / Benchmark duels: simulate() with TwoPlayers(MctsPlayer(white_oracle), MctsPlayer(black_oracle))
   (src/training.jl:130-143, src/benchmark.jl:78-99, src/play.jl:248-282).  Each worker owns one tree per player; both
   players use `mp`; leaves of the two players are evaluated by their own oracle each tick.  With alternate_colors,
   `white_oracle`'s player takes black in every game whose 1-based index is odd. */
int32_t az_selfplay_create_duel(az_ctx* ctx, int32_t game, az_net* white_oracle, az_net* black_oracle, const az_mcts_params* mp,
                                const az_sim_params* sp, uint64_t seed, az_selfplay** out);
/* Benchmark.Duel of two DIFFERENT players (src/benchmark.jl:78-99: e.g. Benchmark.Full(params) against
   Benchmark.MctsRollouts(params'), :134-162, or against Benchmark.NetworkOnly(tau), :161-176): as az_selfplay_create_duel,
   but each player brings its own MctsParams (gamma, cpuct, number of iterations -- 0 = network-only player --, Dirichlet
   noise, prior temperature, move temperature schedule) next to its own oracle. */
int32_t az_selfplay_create_duel_players(az_ctx* ctx, int32_t game, az_net* white_oracle, const az_mcts_params* white_params,
                                        az_net* black_oracle, const az_mcts_params* black_params, const az_sim_params* sp,
                                        uint64_t seed, az_selfplay** out);
/* Benchmark.Duel(player, Benchmark.MinMaxTS(depth, amplify_rewards, tau)) (src/benchmark.jl:78-99, :178-196): the baseline
   is a MinMax.Player (src/minmax.jl: exhaustive search `depth` plies deep with GI.heuristic_value at the horizon, rewards
   optionally amplified to +-Inf, move distribution exp(q / (C tau)) over the non-losing moves, :83-114); it needs no oracle
   and no tree, and its move temperature is the default 1 (src/play.jl:37-39).  `mp` / `oracle` describe the other player
   (num_iters_per_turn = 0: a network-only player).  Games with a two-player heuristic only (connect-four, tictactoe,
   mancala); 1 <= depth <= 8. */
int32_t az_selfplay_create_duel_minmax(az_ctx* ctx, int32_t game, az_net* oracle, const az_mcts_params* mp,
                                       const az_minmax_params* baseline, const az_sim_params* sp, uint64_t seed, az_selfplay** out);
/* plays games first_game_index .. first_game_index + num_games - 1 (global indices key the RNG streams);
   returns immediately, the engine runs on its own host thread + CUDA stream */
int32_t az_selfplay_start(az_selfplay* s, int32_t num_games, int64_t first_game_index);
/* non-blocking progress (drives the `game_simulated` callback, src/simulations.jl:238) */
int32_t az_selfplay_poll(az_selfplay* s, int32_t* games_done, int32_t* finished);
int32_t az_selfplay_wait(az_selfplay* s);
int32_t az_selfplay_counts(az_selfplay* s, int64_t* nsamples, int64_t* ngames);
/* samples ordered by (game, ply) = Trace rows (src/trace.jl:17-24) + push_trace! targets (src/memory.jl:74-87):
   states[nsamples*state_bytes], pi[nsamples*A] (zero on illegal), mask[nsamples*A], z, t, game_of_sample,
   rewards (white_reward after the move), actions; any pointer may be NULL.  With flip_probability > 0 `states` holds
   trace.states[i] (the state before the symmetry of turn i, as the reference records it) while pi, mask and actions
   are in the frame the player thought in (the image state), so that policies[i] == pi[mask] exactly as in the reference. */
int32_t az_selfplay_fetch(az_selfplay* s, uint8_t* states, float* pi, uint8_t* mask, float* z, float* t,
                          int32_t* game_of_sample, double* rewards, int32_t* actions);
/* self_play_measurements (src/training.jl:269-273): per game edepth, node count (mem = nodes x
   MCTS.memory_footprint_per_node), number of moves; totals[4] = {seconds, simulations, expansions, samples} */
int32_t az_selfplay_stats(az_selfplay* s, double* edepth_per_game, int64_t* nodes_per_game, int32_t* moves_per_game,
                          double* totals);
/* rewards_and_redundancy (src/simulations.jl:292-307): rewards[ngames] = total_reward(trace, gamma) (src/trace.jl:45-47)
   negated when colors_flipped; colors_flipped[ngames]; final_states[ngames*state_bytes] = last state of each trace;
   *redundancy = 1 - |unique states| / |states| over all trace states; any pointer may be NULL */
int32_t az_selfplay_outcomes(az_selfplay* s, double gamma, double* rewards, int32_t* colors_flipped, uint8_t* final_states,
                             double* redundancy);
int32_t az_selfplay_destroy(az_selfplay* s);

/* ---- replay-buffer side (src/memory.jl:20-130, src/learning.jl:17-51): device-resident TrainingSamples -------------
   A sample set is a device-side struct of arrays (state, pi[A] Float64 zero on illegal actions, z, t Float64, n Int32)
   = Vector{TrainingSample}; every operation returns a NEW set (inputs stay valid until destroyed). */
/* push_trace! rows of the finished run, ordered by (game, ply), without a host round trip (src/memory.jl:74-87) */
int32_t az_selfplay_export_samples(az_selfplay* s, az_samples** out);
/* samples from host arrays (e.g. the Julia MemoryBuffer): states[n*state_bytes], pi[n*A], z[n], t[n], n_rec[n] or NULL (=1) */
int32_t az_samples_from_host(az_ctx* ctx, int32_t game, int64_t n, const uint8_t* states, const double* pi, const double* z,
                             const double* t, const int32_t* n_rec, az_samples** out);
int32_t az_samples_count(az_samples* s, int64_t* n);
/* [a ; b] (append!(buf, experience), src/memory.jl:41) */
int32_t az_samples_concat(az_samples* a, az_samples* b, az_samples** out);
/* merge_by_state (src/memory.jl:98-110): one sample per distinct state, pi / z / t = mean in original order (left-to-right
   Float64 sum / count), n = sum.  The reference returns the groups in Dict iteration order (unspecified); here they come
   in the order of each state's first occurrence. */
int32_t az_samples_merge_by_state(az_samples* in, az_samples** out);
/* augment_with_symmetries (src/memory.jl:112-130): [samples ; apply_symmetry(s, sym) for s in samples for sym in symmetries(s)] */
int32_t az_samples_augment_with_symmetries(az_samples* in, az_samples** out);
/* convert_samples (src/learning.jl:17-51): Float32 tensors with the sample index as the slowest dimension (= Julia's last):
   W[n] (weighing: 0 CONSTANT_WEIGHT, 1 LOG_WEIGHT = log2(n)+1, 2 LINEAR_WEIGHT), X[n*state_dim], A[n*num_actions] (mask),
   P[n*num_actions], V[n]; computed on the GPU, written to host buffers; any pointer may be NULL */
int32_t az_samples_convert(az_samples* s, int32_t weighing, float* W, float* X, float* A, float* P, float* V);
int32_t az_samples_fetch(az_samples* s, uint8_t* states, double* pi, double* z, double* t, int32_t* n_rec);
int32_t az_samples_destroy(az_samples* s);

/* ---- multi-GPU: simulate_distributed (src/simulations.jl:252-290) with one rank (process) per GPU ---------------------
   Games shard over ranks (num_each / rem split, :268,277: done by the caller); nothing crosses GPUs
   inside the simulation loop.  The two exchanges of an iteration run over NCCL (NVLink / NVSwitch) on the context's
   stream with every row staying in HBM.  NCCL is bound at run time (libnccl.so.2); AZ_EUNSUPPORTED if it is absent. */
#define AZ_COMM_ID_BYTES 128
/* rank 0 creates the id and hands the 128 bytes to the other ranks through the caller's own channel (Julia: the
   `Distributed` remotecall that already starts the workers, :271-281; Python mirror: the torch.distributed store) */
int32_t az_comm_unique_id(az_ctx* ctx, uint8_t id[AZ_COMM_ID_BYTES]);
/* collective over all `world` ranks (ncclCommInitRank); the communicator uses the context's device and stream */
int32_t az_comm_create(az_ctx* ctx, const uint8_t id[AZ_COMM_ID_BYTES], int32_t rank, int32_t world, az_comm** out);
int32_t az_comm_rank(az_comm* c, int32_t* rank, int32_t* world);
/* device time (ms, CUDA events on the context's stream) of the last az_samples_allgather / az_net_broadcast */
int32_t az_comm_last_ms(az_comm* c, double* ms);
int32_t az_comm_destroy(az_comm* c);
/* `fetch.(tasks)` + `reduce(vcat, results)` (src/simulations.jl:282-289) for device-resident samples: *out = the sample
   sets of ranks 0 .. world-1 concatenated in rank order, on every rank.  One all-gather of the per-rank counts, then ONE
   all-gather of packed rows padded to the largest count; counts_out[world] (may be NULL) receives the per-rank counts. */
int32_t az_samples_allgather(az_comm* c, az_samples* local, az_samples** out, int64_t* counts_out);
/* every rank ends up with rank `root`'s parameters loaded into `net` (same architecture on all ranks): replaces the
   serialised closure that carries the network to the workers (src/simulations.jl:271-281) and
   Network.copy(bestnn; on_gpu=true, test_mode=true) (src/training.jl:278-279).  blob[n] = Flux-order float32 parameters
   (az_net_load's format) on the root rank, ignored (may be NULL) elsewhere; n = az_net_num_params on every rank. */
int32_t az_net_broadcast(az_comm* c, az_net* net, const float* blob, int64_t n, int32_t root);

#ifdef __cplusplus
}
#endif
#endif
