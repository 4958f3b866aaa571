/*
 * az_oracle.h -- CPU restatement of the AlphaZero.jl self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (alphazero.jl_b200/) never links,
 * imports or calls this file.
 *
 * PARITY STATUS: "parity unpinned" for MCTS statistics and network outputs --
 * the reference (100 % Julia, no Julia toolchain in this image) ships no golden
 * vector for visit counts / policies / network outputs (test/runtests.jl:11-26
 * checks interface invariants and "no exception" only).  What IS pinned, and
 * checked in tests/test_oracle_*.py:
 *   - Connect-Four rules against the 6000 solver positions of
 *     games/connect-four/benchmark/Test_L*_R* (tests/golden/pons/),
 *   - the GameInterface invariants of src/scripts/test_game.jl:37-110,
 *   - the PLSchedule vector of src/schedule.jl:82-89,
 *   - structural MCTS invariants derived from src/mcts.jl.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).
 *
 * Randomness: Julia's Xoshiro + Distributions.jl cannot be reproduced, so all
 * stochastic inputs come from an explicit counter-based stream (Philox4x32-10)
 * keyed by (seed, game index, move index, purpose, draw index); see oz_rng_*.
 */
#ifndef AZ_ORACLE_H
#define AZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OZ_MAX_ACTIONS 9
#define OZ_STATE_BYTES 48 /* fixed-size, zero padded container for a state */
#define OZ_MAX_PLIES 512

enum { OZ_CONNECT_FOUR = 0, OZ_TICTACTOE = 1, OZ_MANCALA = 2, OZ_GRID_WORLD = 3, OZ_NUM_GAMES = 4 };

/* State byte formats (identical to the C-ABI formats in include/azb200.h):
 *  connect-four (43 B): cells[col + 7*row] in {0 empty,1 white,2 black}, curplayer {1,2}
 *                       == memory layout of (board::SMatrix{7,6,UInt8}, curplayer::UInt8),
 *                       games/connect-four/game.jl:19-22
 *  tictactoe   (10 B): cells[pos] pos=(y-1)*3+(x-1), {0,1,2}; curplayer {1 white,2 black}
 *  mancala     (15 B): stores[2], houses[(player-1)+2*(num-1)] (12), curplayer {1,2}
 *  grid-world  ( 2 B): x, y in 1..10
 */
typedef struct { uint8_t b[OZ_STATE_BYTES]; } oz_state;

typedef struct {
  int game_id;
  uint8_t cells[44];
  uint8_t curplayer; /* 1 = white, 2 = black */
  uint8_t finished;
  uint8_t winner;
  int32_t time;        /* grid-world only (not part of the state) */
  double last_reward;  /* grid-world only (common_rl_intf.jl:154-160) */
} oz_game;

int oz_game_lookup(const char* name);
int oz_num_actions(int game_id);
int oz_state_bytes(int game_id);
void oz_state_dim(int game_id, int dim[3]);

void oz_game_init(oz_game* g, int game_id);
void oz_game_set_state(oz_game* g, int game_id, const uint8_t* state);
void oz_game_get_state(const oz_game* g, uint8_t* state);
int oz_game_terminated(const oz_game* g);
int oz_game_white_playing(const oz_game* g);
void oz_game_actions_mask(const oz_game* g, uint8_t* mask);
/* env_u: stream of uniform draws for stochastic environments (grid-world); may be NULL */
void oz_game_play(oz_game* g, int action, const double* env_u);
double oz_game_white_reward(const oz_game* g);
void oz_vectorize_state(int game_id, const uint8_t* state, float* x);

/* ---- explicit RNG stream ------------------------------------------------ */
void oz_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
enum { OZ_PURPOSE_DIRICHLET = 0, OZ_PURPOSE_CATEGORICAL = 1, OZ_PURPOSE_SYMMETRY = 2, OZ_PURPOSE_ENV = 3, OZ_PURPOSE_ROLLOUT = 5,
       OZ_PURPOSE_POSITION = 4 };
double oz_det_log(double x);
double oz_det_exp(double x);
/* eta[n] ~ Dirichlet(n, alpha) from stream (seed, game, move, DIRICHLET) */
void oz_dirichlet(uint64_t seed, uint64_t game, uint32_t move, int n, double alpha, double* eta);
float oz_uniform_f32(uint64_t seed, uint64_t game, uint32_t move, int purpose, uint32_t idx);

/* ---- MCTS ----------------------------------------------------------------- */
/* oracle(state) -> (P over legal actions in ascending action order, V); src/mcts.jl:6-17 */
typedef void (*oz_oracle_fn)(void* ctx, int game_id, const uint8_t* state, int n_legal, float* P, float* V);
void oz_uniform_oracle(void* ctx, int game_id, const uint8_t* state, int n_legal, float* P, float* V);
/* deterministic hash pseudo-network, bit-reproducible on the GPU */
void oz_synth_oracle(void* ctx, int game_id, const uint8_t* state, int n_legal, float* P, float* V);
/* MCTS.RolloutOracle (src/mcts.jl:27-60); ctx points to an oz_rollout_ctx (the playout's draws are keyed by (seed, state, ply)) */
typedef struct { uint64_t seed; double gamma; } oz_rollout_ctx;
void oz_rollout_oracle(void* ctx, int game_id, const uint8_t* state, int n_legal, float* P, float* V);
void oz_state_key(int game_id, const uint8_t* state, uint64_t key[2]);

typedef struct oz_env oz_env;
oz_env* oz_env_create(int game_id, oz_oracle_fn oracle, void* octx, double gamma, double cpuct, double noise_eps,
                      double noise_alpha, double prior_temperature);
void oz_env_destroy(oz_env*);
void oz_env_reset(oz_env*);
int64_t oz_env_num_nodes(const oz_env*);
int64_t oz_env_total_simulations(const oz_env*);
int64_t oz_env_total_nodes_traversed(const oz_env*);
/* stochastic environments (grid-world): ids of the in-tree environment-noise stream used by the next oz_explore */
void oz_env_set_noise(oz_env*, uint64_t seed, uint64_t game, uint32_t move);
/* explore!(env, game, nsims) with an explicit eta (length n_legal; may be NULL iff noise_eps == 0) */
void oz_explore(oz_env*, const oz_game* root, int nsims, const double* eta);
/* root statistics in action-indexed form (A wide, zeros on illegal); returns n_legal or -1 if root unknown */
int oz_root_stats(const oz_env*, const oz_game* root, int64_t* N, double* W, float* P, float* Vest);
/* policy(env, game): pi over legal actions (compact), returns n_legal */
int oz_policy(const oz_env*, const oz_game* root, int* actions, double* pi);

/* ---- schedules, temperature, sampling ------------------------------------ */
double oz_pl_schedule(int n, const int* xs, const double* ys, int i);
void oz_apply_temperature(const double* pi, int n, double tau, double* out);
void oz_fix_probvec(const double* pi, int n, float* out);
int oz_categorical(const float* p, int n, float u);

/* ---- self-play ------------------------------------------------------------- */
typedef struct {
  double gamma, cpuct, noise_eps, noise_alpha, prior_temperature;
  int num_iters_per_turn; /* 0: NetworkPlayer under PlayerWithTemperature (Benchmark.NetworkOnly) */
  int sched_n;
  int sched_xs[8];
  double sched_ys[8];
  /* player_kind 1: MinMax.Player (src/minmax.jl:72-81; Benchmark.MinMaxTS) with the four fields below and `gamma`; its
     oracle env is never called and its move temperature is the AbstractPlayer default of 1 (src/play.jl:37-39) */
  int player_kind;
  int minmax_depth;
  int minmax_amplify;
  double minmax_tau;
} oz_mcts_params;

/* GI.heuristic_value (games/connect-four/game.jl:172-220, games/tictactoe/game.jl:96-120, games/mancala/game.jl:212-218
   incl. its UInt8 wrap-around, games/grid-world/game.jl:118) */
double oz_heuristic_value(const oz_game* g);
/* think(::MinMax.Player, game) (src/minmax.jl:83-114): actions[n] (ascending) and pi[n]; returns n; qs (may be NULL)
   receives the root q-values */
int oz_minmax_think(const oz_game* g, int depth, int amplify_rewards, double tau, double gamma, int* actions, double* pi, double* qs);

typedef struct {
  int n_moves;                                 /* length(trace) */
  uint8_t states[OZ_MAX_PLIES + 1][OZ_STATE_BYTES];
  float pi[OZ_MAX_PLIES][OZ_MAX_ACTIONS];     /* A-wide, zero on illegal */
  uint8_t mask[OZ_MAX_PLIES][OZ_MAX_ACTIONS];
  int32_t action[OZ_MAX_PLIES];
  double rewards[OZ_MAX_PLIES];
  double z[OZ_MAX_PLIES];
  double t[OZ_MAX_PLIES];
  int64_t mem_nodes;                           /* length(env.tree) when measured (both players' trees in a duel) */
  double edepth;                               /* average_exploration_depth (both players pooled in a duel) */
  int32_t sym[OZ_MAX_PLIES];                   /* 0, or 1 + index of the symmetry applied before thinking (play.jl:305-307) */
  uint8_t think_states[OZ_MAX_PLIES][OZ_STATE_BYTES]; /* the state the player thought on (pi and mask are in its frame) */
  double pi64[OZ_MAX_PLIES][OZ_MAX_ACTIONS];   /* trace.policies[i] as the reference holds it (Float64), A-wide */
} oz_trace;

/* play_game (src/play.jl:298-315) for game index `game` on worker env `env` */
void oz_play_game(oz_env* env, const oz_mcts_params* mp, uint64_t seed, uint64_t game, oz_trace* out);
/* GI.symmetries (games/connect-four/game.jl:247-257, games/tictactoe/game.jl:149-168): number of declared symmetries
   (0: none -> flip_probability > 0 is a parameter error, src/params.jl:377-381) and the image of a state */
int oz_num_symmetries(int game_id);
void oz_apply_symmetry(int game_id, int sym, const uint8_t* state_in, uint8_t* state_out);
/* play_game with TwoPlayers(white, black) (src/play.jl:248-282) and flip_probability (src/play.jl:305-307).
   `white` thinks when white is to play, `black` otherwise (the same env twice = a single MctsPlayer).  The flip
   decision of move m is u01(stream(seed, game, m, SYMMETRY, 0)) < flip_probability and the symmetry index is
   stream(seed, game, m, SYMMETRY, 1) mod num_symmetries. */
void oz_play_game2(oz_env* white, oz_env* black, const oz_mcts_params* mp, double flip_probability, uint64_t seed,
                   uint64_t game, oz_trace* out);
/* the same with one MctsParams per player (Benchmark duels of different MctsPlayers) */
void oz_play_game2p(oz_env* white, const oz_mcts_params* mp_white, oz_env* black, const oz_mcts_params* mp_black, double flip_p,
                    uint64_t seed, uint64_t game_idx, oz_trace* tr);
/* total_reward(trace, gamma) (src/trace.jl:45-47) */
double oz_total_reward(const oz_trace* tr, double gamma);
/* one worker of simulate() (src/simulations.jl:207-244): plays games first, first+stride, ... (count games),
   measures before reset, resets every reset_every games (<=0: never) */
void oz_worker_run(int game_id, oz_oracle_fn oracle, void* octx, const oz_mcts_params* mp, uint64_t seed,
                   uint64_t first_game, uint64_t stride, int count, int reset_every, oz_trace* out);

/* synthetic random positions: k ~ U{0..max_plies} uniformly random legal plies, terminal rejected */
void oz_random_position(int game_id, uint64_t seed, uint64_t stream, int max_plies, uint8_t* state);

/* ---- batched lock-step driver (CPU baseline with a batched evaluator) ------ */
/* host threads used by oz_batch_advance / oz_batch_vectorize / oz_batch_feed (default 1) */
void oz_set_threads(int n);
int oz_get_threads(void);
typedef struct oz_batch oz_batch;
oz_batch* oz_batch_create(int game_id, int n_trees, const oz_mcts_params* mp);
void oz_batch_destroy(oz_batch*);
void oz_batch_set_roots(oz_batch*, const uint8_t* states /* n_trees * state_bytes */, const double* eta /* n*A or NULL */);
/* advance every tree until it needs an oracle answer or has finished nsims; returns number of pending leaves and
   writes their states (state_bytes each) and owning tree index */
int oz_batch_advance(oz_batch*, uint8_t* leaf_states, int32_t* leaf_tree);
/* answers for the pending leaves, in the order given by oz_batch_advance: P is A-wide (zeros on illegal) */
void oz_batch_feed(oz_batch*, const float* P, const float* V);
/* GI.vectorize_state + actions mask of the pending leaves (src/networks/network.jl:310-312), threaded over leaves */
void oz_batch_vectorize(const oz_batch*, const uint8_t* leaf_states, int n, int xdim, float* X, uint8_t* mask);
void oz_batch_root_stats(const oz_batch*, int tree, int64_t* N, double* W, float* P);
int64_t oz_batch_total_expansions(const oz_batch*);
int64_t oz_batch_total_simulations(const oz_batch*);
int oz_c4_solve(const uint8_t* state); /* test helper: exact negamax, Pons score convention */
void oz_batch_reset_trees(oz_batch*);

#ifdef __cplusplus
}
#endif
#endif
