/* Plain-C driver of the drop-in boundary (include/azb200.h): the calls a non-Python host makes for one self-play batch --
   simulate(simulator, gspec, SimParams) of src/simulations.jl:207-244 with the built-in uniform oracle
   (MCTS.RandomOracle, src/mcts.jl:62-72) -- followed by the replay-buffer side (src/memory.jl:98-130).
   Build:  gcc -std=c99 -Iinclude examples/selfplay_c_abi.c -Lalphazero.jl_b200 -lazb200 -Wl,-rpath,$PWD/alphazero.jl_b200 -o selfplay_c
   Exit codes: 0 ok, 3 no CUDA device (the library has no CPU fallback), 1 any other failure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "azb200.h"

#define CHECK(ctx, call)                                                                \
  do {                                                                                  \
    int32_t st_ = (call);                                                               \
    if (st_ != AZ_OK) {                                                                 \
      fprintf(stderr, "%s failed (%d): %s\n", #call, (int)st_, az_last_error(ctx));     \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const int num_games = argc > 1 ? atoi(argv[1]) : 64;
  const int nsims = argc > 2 ? atoi(argv[2]) : 100;
  az_ctx* ctx = NULL;
  int32_t st = az_ctx_create(0, &ctx);
  if (st != AZ_OK) {
    fprintf(stderr, "az_ctx_create failed (%d): %s\n", (int)st, az_last_error(NULL));
    return st == AZ_ECUDA ? 3 : 1;
  }
  const int32_t game = az_game_lookup("connect-four");
  const int A = az_game_num_actions(game), SB = az_game_state_bytes(game);
  az_net* oracle = NULL;
  CHECK(ctx, az_net_create_oracle(ctx, AZ_NET_UNIFORM, game, &oracle));

  az_mcts_params mp;
  memset(&mp, 0, sizeof(mp));
  mp.gamma = 1.0; mp.cpuct = 2.0; mp.num_iters_per_turn = nsims;                    /* games/connect-four/params.jl:24-30 */
  mp.dirichlet_noise_eps = 0.25; mp.dirichlet_noise_alpha = 1.0; mp.prior_temperature = 1.0;
  mp.temperature_n = 3;                                                             /* PLSchedule([0, 20, 30], [1.0, 1.0, 0.3]) */
  mp.temperature_xs[0] = 0; mp.temperature_xs[1] = 20; mp.temperature_xs[2] = 30;
  mp.temperature_ys[0] = 1.0; mp.temperature_ys[1] = 1.0; mp.temperature_ys[2] = 0.3;
  az_sim_params sp;
  memset(&sp, 0, sizeof(sp));
  sp.num_games = num_games; sp.num_workers = num_games < 32 ? num_games : 32; sp.batch_size = sp.num_workers;
  sp.fill_batches = 1; sp.reset_every = 2; sp.alternate_colors = 0; sp.flip_probability = 0.0;

  az_selfplay* run = NULL;
  CHECK(ctx, az_selfplay_create(ctx, game, oracle, &mp, &sp, /*seed=*/2024, &run));
  CHECK(ctx, az_selfplay_start(run, num_games, /*first_game_index=*/0));
  CHECK(ctx, az_selfplay_wait(run));
  int64_t nsamples = 0, ngames = 0;
  CHECK(ctx, az_selfplay_counts(run, &nsamples, &ngames));
  uint8_t* states = (uint8_t*)malloc((size_t)nsamples * SB);
  float* pi = (float*)malloc((size_t)nsamples * A * sizeof(float));
  float* z = (float*)malloc((size_t)nsamples * sizeof(float));
  if (!states || !pi || !z) return 1;
  CHECK(ctx, az_selfplay_fetch(run, states, pi, NULL, z, NULL, NULL, NULL, NULL));
  double totals[4];
  CHECK(ctx, az_selfplay_stats(run, NULL, NULL, NULL, totals));
  double zsum = 0.0;
  for (int64_t i = 0; i < nsamples; i++) zsum += z[i];
  printf("games %lld samples %lld  simulations %.0f expansions %.0f in %.3f s  mean z %.4f\n", (long long)ngames, (long long)nsamples,
         totals[1], totals[2], totals[0], nsamples ? zsum / (double)nsamples : 0.0);

  /* replay-buffer side on the device: samples of the run -> + mirror images -> one sample per distinct state */
  az_samples *smp = NULL, *aug = NULL, *mrg = NULL;
  CHECK(ctx, az_selfplay_export_samples(run, &smp));
  CHECK(ctx, az_samples_augment_with_symmetries(smp, &aug));
  CHECK(ctx, az_samples_merge_by_state(aug, &mrg));
  int64_t n_aug = 0, n_mrg = 0;
  CHECK(ctx, az_samples_count(aug, &n_aug));
  CHECK(ctx, az_samples_count(mrg, &n_mrg));
  printf("augmented %lld -> distinct states %lld\n", (long long)n_aug, (long long)n_mrg);

  free(states); free(pi); free(z);
  az_samples_destroy(mrg); az_samples_destroy(aug); az_samples_destroy(smp);
  az_selfplay_destroy(run);
  az_net_destroy(oracle);
  az_ctx_destroy(ctx);
  return 0;
}
