/*
 * az_oracle.c -- CPU restatement (plain C11) of the AlphaZero.jl self-play hot path.
 * TEST INFRASTRUCTURE ONLY -- see az_oracle.h.  "parity unpinned" for MCTS statistics.
 *
 * Compile with -ffp-contract=off: Julia never fuses a*b+c, so every floating-point
 * expression below must round after each operation.
 *
 * Reference files restated here (paths relative to the reference root):
 *   src/mcts.jl:78-89,124-151,157-271,278-296   MCTS
 *   src/play.jl:196-214,298-315                 think / play_game
 *   src/util.jl:68-110                          fix_probvec / rand_categorical / apply_temperature
 *   src/schedule.jl:64-80                       PLSchedule
 *   src/memory.jl:74-87                         push_trace!
 *   src/simulations.jl:207-244                  worker loop, reset_every
 *   games/connect-four/game.jl:40-168,226-241   rules + vectorize_state
 *   games/tictactoe/game.jl:24-92,126-143
 *   games/mancala/game.jl:43-206,224-257
 *   games/grid-world/game.jl:14-59,86-90 + src/common_rl_intf.jl:118-160
 */
#include "az_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Game table                                                                 */
/* ------------------------------------------------------------------------- */

static const char* OZ_NAMES[OZ_NUM_GAMES] = {"connect-four", "tictactoe", "mancala", "grid-world"};
static const int OZ_NACT[OZ_NUM_GAMES] = {7, 9, 6, 4};
static const int OZ_SBYTES[OZ_NUM_GAMES] = {43, 10, 15, 2};
static const int OZ_SDIM[OZ_NUM_GAMES][3] = {{7, 6, 3}, {3, 3, 3}, {14, 1, 5}, {10, 10, 1}};

int oz_game_lookup(const char* name) { /* src/examples.jl:17-21 */
  for (int i = 0; i < OZ_NUM_GAMES; i++)
    if (strcmp(name, OZ_NAMES[i]) == 0) return i;
  return -1;
}
int oz_num_actions(int g) { return OZ_NACT[g]; }
int oz_state_bytes(int g) { return OZ_SBYTES[g]; }
void oz_state_dim(int g, int d[3]) { d[0] = OZ_SDIM[g][0]; d[1] = OZ_SDIM[g][1]; d[2] = OZ_SDIM[g][2]; }

/* ---------------- Connect Four (games/connect-four/game.jl) ---------------- */
#define C4_COLS 7
#define C4_ROWS 6
#define C4(g, col, row) ((g)->cells[(col) + C4_COLS * (row)]) /* 0-based col,row */

static int c4_first_free(const oz_game* g, int col) { /* :87-93 (returns 0-based row, C4_ROWS if full) */
  int row = 0;
  while (row < C4_ROWS && C4(g, col, row) != 0) row++;
  return row;
}
static int c4_valid(int col, int row) { return col >= 0 && col < C4_COLS && row >= 0 && row < C4_ROWS; }
static int c4_connected_dir(const oz_game* g, int player, int col, int row, int dc, int dr) { /* :103-112 */
  int n = 0;
  col += dc; row += dr;
  while (c4_valid(col, row) && C4(g, col, row) == player) { n++; col += dc; row += dr; }
  return n;
}
static int c4_winning_pattern_at(const oz_game* g, int player, int col, int row) { /* :114-127 */
  static const int AX[4][2] = {{1, 1}, {1, -1}, {1, 0}, {0, 1}};
  for (int a = 0; a < 4; a++) {
    int n = 1 + c4_connected_dir(g, player, col, row, AX[a][0], AX[a][1]) +
            c4_connected_dir(g, player, col, row, -AX[a][0], -AX[a][1]);
    if (n >= 4) return 1;
  }
  return 0;
}
static int c4_any_free(const oz_game* g) {
  for (int c = 0; c < C4_COLS; c++)
    if (c4_first_free(g, c) < C4_ROWS) return 1;
  return 0;
}
static void c4_set_state(oz_game* g, const uint8_t* s) { /* :50-68 */
  memcpy(g->cells, s, 42);
  g->curplayer = s[42];
  g->finished = 0;
  g->winner = 0;
  if (!c4_any_free(g)) g->finished = 1;
  for (int col = 0; col < C4_COLS; col++) {
    int top = c4_first_free(g, col);
    if (top == 0) continue;
    int row = top - 1;
    int c = C4(g, col, row);
    if (c != 0 && c4_winning_pattern_at(g, c, col, row)) {
      g->winner = (uint8_t)c;
      g->finished = 1;
      break;
    }
  }
}
static void c4_play(oz_game* g, int col) { /* :140-146, update_status! :130-138 */
  int row = c4_first_free(g, col);
  C4(g, col, row) = g->curplayer;
  if (c4_winning_pattern_at(g, g->curplayer, col, row)) {
    g->winner = g->curplayer;
    g->finished = 1;
  } else {
    g->finished = (uint8_t)!c4_any_free(g);
  }
  g->curplayer = (uint8_t)(3 - g->curplayer);
}
static double c4_white_reward(const oz_game* g) { /* :160-168 */
  if (g->finished) {
    if (g->winner == 1) return 1.0;
    if (g->winner == 2) return -1.0;
  }
  return 0.0;
}
static void c4_vectorize(const uint8_t* s, float* x) { /* :226-241 */
  int flip = (s[42] != 1);
  for (int c = 0; c < 3; c++)
    for (int row = 0; row < C4_ROWS; row++)
      for (int col = 0; col < C4_COLS; col++) {
        int cell = s[col + 7 * row];
        if (flip && cell != 0) cell = 3 - cell;
        x[col + 7 * row + 42 * c] = (cell == c) ? 1.0f : 0.0f;
      }
}

/* ---------------- Tic-tac-toe (games/tictactoe/game.jl) -------------------- */
static const int TTT_AL[8][3] = {/* :39-49: columns (x fixed), rows (y fixed), two diagonals; pos=(y-1)*3+x (0-based here) */
                                 {0, 3, 6}, {1, 4, 7}, {2, 5, 8}, {0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {0, 4, 8}, {2, 4, 6}};
static int ttt_has_won(const oz_game* g, int player) { /* :53-59 */
  for (int a = 0; a < 8; a++)
    if (g->cells[TTT_AL[a][0]] == player && g->cells[TTT_AL[a][1]] == player && g->cells[TTT_AL[a][2]] == player)
      return 1;
  return 0;
}
static int ttt_terminal_reward(const oz_game* g, double* r) { /* :75-80 */
  if (ttt_has_won(g, 1)) { *r = 1.0; return 1; }
  if (ttt_has_won(g, 2)) { *r = -1.0; return 1; }
  int any = 0;
  for (int i = 0; i < 9; i++) any |= (g->cells[i] == 0);
  if (!any) { *r = 0.0; return 1; }
  *r = 0.0;
  return 0;
}
static void ttt_vectorize(const uint8_t* s, float* x) { /* :126-143 */
  int flip = (s[9] != 1);
  for (int c = 0; c < 3; c++)
    for (int pos = 0; pos < 9; pos++) {
      int cell = s[pos];
      if (flip && cell != 0) cell = 3 - cell;
      x[pos + 9 * c] = (cell == c) ? 1.0f : 0.0f;
    }
}

/* ---------------- Mancala (games/mancala/game.jl) -------------------------- */
/* cells[0..1] = stores (white, black); cells[2 + (player-1) + 2*(num-1)] = houses[player,num] */
#define MC_STORE(g, p) ((g)->cells[(p)-1])
#define MC_HOUSE(g, p, n) ((g)->cells[2 + ((p)-1) + 2 * ((n)-1)])
typedef struct { int is_store, player, num; } mc_pos;
static mc_pos mc_next_pos(mc_pos pos, int player) { /* :80-97 */
  mc_pos r;
  if (pos.is_store) { r.is_store = 0; r.player = 3 - player; r.num = 6; return r; }
  if (pos.num > 1) { r.is_store = 0; r.player = pos.player; r.num = pos.num - 1; return r; }
  if (pos.player == player) { r.is_store = 1; r.player = player; r.num = 0; return r; }
  r.is_store = 0; r.player = player; r.num = 6;
  return r;
}
static int mc_read(const oz_game* g, mc_pos p) { return p.is_store ? MC_STORE(g, p.player) : MC_HOUSE(g, p.player, p.num); }
static void mc_write(oz_game* g, mc_pos p, int v) {
  if (p.is_store) MC_STORE(g, p.player) = (uint8_t)v; else MC_HOUSE(g, p.player, p.num) = (uint8_t)v;
}
static int mc_sum_houses(const oz_game* g, int player) {
  int s = 0;
  for (int n = 1; n <= 6; n++) s += MC_HOUSE(g, player, n);
  return s;
}
static void mc_capture_leftovers(oz_game* g, int player) { /* :134-139 */
  MC_STORE(g, player) = (uint8_t)(MC_STORE(g, player) + mc_sum_houses(g, player));
  for (int i = 2; i < 14; i++) g->cells[i] = 0;
}
static void mc_play(oz_game* g, int a /* 1..6 */) { /* :144-177 */
  int cp = g->curplayer;
  mc_pos pos = {0, cp, a};
  int nseeds = mc_read(g, pos);
  mc_write(g, pos, 0);
  for (int i = 0; i < nseeds; i++) {
    pos = mc_next_pos(pos, cp);
    mc_write(g, pos, mc_read(g, pos) + 1);
  }
  if (mc_sum_houses(g, cp) == 0) {
    mc_capture_leftovers(g, 3 - cp);
    g->finished = 1;
  } else if (!pos.is_store) {
    if (mc_read(g, pos) == 1 && cp == pos.player) {
      mc_pos opp = {0, 3 - pos.player, 6 - pos.num + 1};
      MC_STORE(g, pos.player) = (uint8_t)(MC_STORE(g, pos.player) + mc_read(g, opp) + 1);
      mc_write(g, pos, 0);
      mc_write(g, opp, 0);
      if (mc_sum_houses(g, 3 - cp) == 0) { mc_capture_leftovers(g, cp); g->finished = 1; return; }
      if (mc_sum_houses(g, cp) == 0) { mc_capture_leftovers(g, 3 - cp); g->finished = 1; return; }
    }
    g->curplayer = (uint8_t)(3 - cp);
  }
}
static void mc_vectorize(const uint8_t* s, float* x) { /* :224-257 incl. the flip_colors quirk (returns INITIAL board) */
  uint8_t stores[2], houses[2][6];
  if (s[14] == 1) {
    stores[0] = s[0]; stores[1] = s[1];
    for (int p = 0; p < 2; p++) for (int n = 0; n < 6; n++) houses[p][n] = s[2 + p + 2 * n];
  } else {
    stores[0] = stores[1] = 0;
    for (int p = 0; p < 2; p++) for (int n = 0; n < 6; n++) houses[p][n] = 3;
  }
  /* positions: white houses 6..1, white store, black houses 6..1, black store; channels nstones,whouse,wstore,bhouse,bstore */
  for (int i = 0; i < 14; i++) {
    int is_store = (i == 6 || i == 13), player = (i < 7) ? 1 : 2;
    int num = is_store ? 0 : 6 - (i % 7);
    float nst = is_store ? (float)stores[player - 1] : (float)houses[player - 1][num - 1];
    x[i + 14 * 0] = nst;
    x[i + 14 * 1] = (!is_store && player == 1) ? 1.0f : 0.0f;
    x[i + 14 * 2] = (is_store && player == 1) ? 1.0f : 0.0f;
    x[i + 14 * 3] = (!is_store && player == 2) ? 1.0f : 0.0f;
    x[i + 14 * 4] = (is_store && player == 2) ? 1.0f : 0.0f;
  }
}

/* ---------------- Grid world (games/grid-world/game.jl) -------------------- */
static double gw_reward_at(int x, int y) { /* :24-28 */
  if (x == 9 && y == 3) return 10.0;
  if (x == 8 && y == 8) return 3.0;
  if (x == 4 && y == 3) return -10.0;
  if (x == 4 && y == 6) return -5.0;
  return 0.0;
}
static int gw_has_reward(int x, int y) { return (x == 9 && y == 3) || (x == 8 && y == 8) || (x == 4 && y == 3) || (x == 4 && y == 6); }
static const int GW_ACT[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}}; /* :36 */
static void gw_act(oz_game* g, int a /* 0..3 */, const double* u) { /* :43-51; u[0] < 0.4 -> random action floor(4*u[1]) */
  if (u && u[0] < 0.4) {
    a = (int)(u[1] * 4.0);
    if (a > 3) a = 3;
  }
  int x = g->cells[0] + GW_ACT[a][0], y = g->cells[1] + GW_ACT[a][1];
  if (x < 1) x = 1; if (x > 10) x = 10;
  if (y < 1) y = 1; if (y > 10) y = 10;
  g->cells[0] = (uint8_t)x; g->cells[1] = (uint8_t)y;
  g->time += 1;
  g->last_reward = gw_reward_at(x, y);
}

/* ---------------- generic dispatch ---------------------------------------- */
void oz_game_set_state(oz_game* g, int game_id, const uint8_t* s) {
  memset(g, 0, sizeof(*g));
  g->game_id = game_id;
  switch (game_id) {
    case OZ_CONNECT_FOUR: c4_set_state(g, s); break;
    case OZ_TICTACTOE: memcpy(g->cells, s, 9); g->curplayer = s[9]; break; /* :24,:30-33 */
    case OZ_MANCALA: /* :54-60 */
      memcpy(g->cells, s, 14); g->curplayer = s[14];
      if (mc_sum_houses(g, g->curplayer) == 0 || mc_sum_houses(g, 3 - g->curplayer) == 0) g->finished = 1;
      break;
    case OZ_GRID_WORLD: g->cells[0] = s[0]; g->cells[1] = s[1]; g->curplayer = 1; break; /* common_rl_intf.jl:128-132 */
  }
}
void oz_game_init(oz_game* g, int game_id) {
  uint8_t s[OZ_STATE_BYTES];
  memset(s, 0, sizeof(s));
  switch (game_id) {
    case OZ_CONNECT_FOUR: s[42] = 1; break;
    case OZ_TICTACTOE: s[9] = 1; break;
    case OZ_MANCALA: for (int i = 2; i < 14; i++) s[i] = 3; s[14] = 1; break;
    case OZ_GRID_WORLD: s[0] = 1; s[1] = 1; break; /* random in the reference (game.jl:32); callers set a state */
  }
  oz_game_set_state(g, game_id, s);
}
void oz_game_get_state(const oz_game* g, uint8_t* s) {
  switch (g->game_id) {
    case OZ_CONNECT_FOUR: memcpy(s, g->cells, 42); s[42] = g->curplayer; break;
    case OZ_TICTACTOE: memcpy(s, g->cells, 9); s[9] = g->curplayer; break;
    case OZ_MANCALA: memcpy(s, g->cells, 14); s[14] = g->curplayer; break;
    case OZ_GRID_WORLD: s[0] = g->cells[0]; s[1] = g->cells[1]; break; /* time is NOT in the state, game.jl:57 */
  }
}
int oz_game_terminated(const oz_game* g) {
  double r;
  switch (g->game_id) {
    case OZ_TICTACTOE: return ttt_terminal_reward(g, &r);
    case OZ_GRID_WORLD: return gw_has_reward(g->cells[0], g->cells[1]) || g->time > 200; /* :40-41 */
    default: return g->finished;
  }
}
int oz_game_white_playing(const oz_game* g) { return g->game_id == OZ_GRID_WORLD ? 1 : g->curplayer == 1; }
void oz_game_actions_mask(const oz_game* g, uint8_t* m) {
  switch (g->game_id) {
    case OZ_CONNECT_FOUR: for (int c = 0; c < 7; c++) m[c] = c4_first_free(g, c) < C4_ROWS; break;
    case OZ_TICTACTOE: for (int i = 0; i < 9; i++) m[i] = (g->cells[i] == 0); break; /* :68 */
    case OZ_MANCALA: for (int n = 1; n <= 6; n++) m[n - 1] = MC_HOUSE(g, g->curplayer, n) > 0; break; /* :121-123 */
    case OZ_GRID_WORLD: m[0] = m[1] = m[2] = m[3] = 1; break; /* :59 */
  }
}
void oz_game_play(oz_game* g, int action, const double* env_u) { /* action is 0-based */
  switch (g->game_id) {
    case OZ_CONNECT_FOUR: c4_play(g, action); break;
    case OZ_TICTACTOE: g->cells[action] = g->curplayer; g->curplayer = (uint8_t)(3 - g->curplayer); break; /* :89-92 */
    case OZ_MANCALA: mc_play(g, action + 1); break;
    case OZ_GRID_WORLD: gw_act(g, action, env_u); break;
  }
}
double oz_game_white_reward(const oz_game* g) {
  double r = 0.0;
  switch (g->game_id) {
    case OZ_CONNECT_FOUR: return c4_white_reward(g);
    case OZ_TICTACTOE: ttt_terminal_reward(g, &r); return r; /* :84-87 */
    case OZ_MANCALA: /* :199-206 */
      if (!g->finished) return 0.0;
      return MC_STORE(g, 1) > MC_STORE(g, 2) ? 1.0 : (MC_STORE(g, 1) < MC_STORE(g, 2) ? -1.0 : 0.0);
    case OZ_GRID_WORLD: return g->last_reward;
  }
  return r;
}
void oz_vectorize_state(int game_id, const uint8_t* s, float* x) {
  switch (game_id) {
    case OZ_CONNECT_FOUR: c4_vectorize(s, x); break;
    case OZ_TICTACTOE: ttt_vectorize(s, x); break;
    case OZ_MANCALA: mc_vectorize(s, x); break;
    case OZ_GRID_WORLD: /* :86-90, v[x,y] column-major */
      for (int i = 0; i < 100; i++) x[i] = 0.0f;
      x[(s[0] - 1) + 10 * (s[1] - 1)] = 1.0f;
      break;
  }
}

/* canonical 128-bit key of a state (spec shared with the device hash table / synthetic oracle) */
void oz_state_key(int game_id, const uint8_t* s, uint64_t key[2]) {
  uint64_t a = 0, b = 0;
  switch (game_id) {
    case OZ_CONNECT_FOUR:
      for (int col = 0; col < 7; col++)
        for (int row = 0; row < 6; row++) {
          int c = s[col + 7 * row];
          if (c == 1) a |= 1ull << (col * 7 + row);
          if (c == 2) b |= 1ull << (col * 7 + row);
        }
      if (s[42] == 2) b |= 1ull << 56;
      break;
    case OZ_TICTACTOE:
      for (int i = 0; i < 9; i++) {
        if (s[i] == 1) a |= 1ull << i;
        if (s[i] == 2) a |= 1ull << (16 + i);
      }
      if (s[9] == 2) a |= 1ull << 32;
      break;
    case OZ_MANCALA:
      for (int i = 0; i < 8; i++) a |= (uint64_t)s[i] << (8 * i);
      for (int i = 0; i < 6; i++) b |= (uint64_t)s[8 + i] << (8 * i);
      if (s[14] == 2) b |= 1ull << 56;
      break;
    case OZ_GRID_WORLD: a = (uint64_t)s[0] | ((uint64_t)s[1] << 8); break;
  }
  key[0] = a; key[1] = b;
}

/* ------------------------------------------------------------------------- */
/* Explicit RNG stream: Philox4x32-10 + deterministic log/exp                  */
/* ------------------------------------------------------------------------- */
void oz_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* k-th 64-bit word of stream (seed, game, move, purpose) */
static uint64_t oz_stream_u64(uint64_t seed, uint64_t game, uint32_t move, int purpose, uint32_t k) {
  uint32_t o[4];
  oz_philox(seed, k >> 1, (uint32_t)purpose | (move << 8), (uint32_t)game, (uint32_t)(game >> 32), o);
  return (k & 1) ? ((uint64_t)o[3] << 32 | o[2]) : ((uint64_t)o[1] << 32 | o[0]);
}
static double oz_u01(uint64_t x) { return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0); } /* (0,1), exact */
float oz_uniform_f32(uint64_t seed, uint64_t game, uint32_t move, int purpose, uint32_t idx) {
  uint64_t x = oz_stream_u64(seed, game, move, purpose, idx);
  return (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f); /* [0,1), 24 bits like rand(Float32) */
}

static const double OZ_LN2 = 0.6931471805599453094;
double oz_det_log(double x) { /* normal positive doubles only */
  uint64_t bits;
  memcpy(&bits, &x, 8);
  int e = (int)((bits >> 52) & 0x7FF) - 1022;
  bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull; /* m in [0.5,1) */
  double m;
  memcpy(&m, &bits, 8);
  if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
  double s = (m - 1.0) / (m + 1.0);
  double s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; k--) p = p * s2 + 1.0 / (double)(2 * k + 1);
  return (double)e * OZ_LN2 + (2.0 * s) * p;
}
double oz_det_exp(double x) {
  double kf = floor(x / OZ_LN2 + 0.5);
  if (kf < -1000.0) kf = -1000.0;
  if (kf > 1000.0) kf = 1000.0;
  double r = x - kf * OZ_LN2;
  double p = 1.0;
  for (int n = 16; n >= 1; n--) p = p * (r / (double)n) + 1.0;
  int k = (int)kf;
  uint64_t bits = (uint64_t)(k + 1023) << 52;
  double sc;
  memcpy(&sc, &bits, 8);
  return p * sc;
}
typedef struct { uint64_t seed, game; uint32_t move; int purpose; uint32_t k; } oz_stream;
static double oz_next_u01(oz_stream* st) { return oz_u01(oz_stream_u64(st->seed, st->game, st->move, st->purpose, st->k++)); }
static double oz_normal(oz_stream* st) { /* Marsaglia polar method */
  for (;;) {
    double u1 = 2.0 * oz_next_u01(st) - 1.0, u2 = 2.0 * oz_next_u01(st) - 1.0;
    double s = u1 * u1 + u2 * u2;
    if (s >= 1.0 || s == 0.0) continue;
    return u1 * sqrt((-2.0 * oz_det_log(s)) / s);
  }
}
static double oz_gamma(oz_stream* st, double alpha) { /* Marsaglia-Tsang */
  double boost = 1.0;
  if (alpha < 1.0) {
    double u = oz_next_u01(st);
    boost = oz_det_exp(oz_det_log(u) / alpha);
    alpha = alpha + 1.0;
  }
  double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = oz_normal(st);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = oz_next_u01(st);
    double x2 = x * x;
    if (u < 1.0 - (0.0331 * x2) * x2) return (d * v) * boost;
    if (oz_det_log(u) < 0.5 * x2 + d * ((1.0 - v) + oz_det_log(v))) return (d * v) * boost;
  }
}
void oz_dirichlet(uint64_t seed, uint64_t game, uint32_t move, int n, double alpha, double* eta) { /* src/mcts.jl:228-232 */
  oz_stream st = {seed, game, move, OZ_PURPOSE_DIRICHLET, 0};
  double sum = 0.0;
  for (int i = 0; i < n; i++) { eta[i] = oz_gamma(&st, alpha); sum = sum + eta[i]; }
  for (int i = 0; i < n; i++) eta[i] = eta[i] / sum;
}

/* ------------------------------------------------------------------------- */
/* Oracles                                                                     */
/* ------------------------------------------------------------------------- */
void oz_uniform_oracle(void* ctx, int game_id, const uint8_t* state, int n, float* P, float* V) { /* src/mcts.jl:62-72 */
  (void)ctx; (void)game_id; (void)state;
  for (int i = 0; i < n; i++) P[i] = (float)(1.0 / (double)n); /* ones(n) ./ n (Float64) -> ActionStats.P::Float32 */
  *V = 0.0f;
}
static uint64_t oz_splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void oz_synth_oracle(void* ctx, int game_id, const uint8_t* state, int n, float* P, float* V) {
  (void)ctx;
  uint64_t key[2];
  oz_state_key(game_id, state, key);
  uint64_t h0 = oz_splitmix(key[0] ^ oz_splitmix(key[1]));
  oz_game g;
  oz_game_set_state(&g, game_id, state);
  uint8_t mask[OZ_MAX_ACTIONS];
  oz_game_actions_mask(&g, mask);
  uint32_t raw[OZ_MAX_ACTIONS], sum = 0;
  int A = OZ_NACT[game_id], j = 0;
  for (int a = 0; a < A; a++) {
    if (!mask[a]) continue;
    uint64_t hi = oz_splitmix(h0 + (uint64_t)(a + 1) * 0x9E3779B97F4A7C15ull);
    raw[j] = 1u + (uint32_t)(hi >> 48);
    sum += raw[j];
    j++;
  }
  (void)n;
  for (int i = 0; i < j; i++) P[i] = (float)raw[i] / (float)sum;
  *V = ((float)(int)(h0 >> 48) - 32768.0f) / 32768.0f;
}

/* MCTS.RolloutOracle, src/mcts.jl:27-60.  rollout! (:42-50) is restated with its recursion; `rand(GI.available_actions(game))`
   (:43) takes the k-th available action with k = floor(u32 * n / 2^32) from the Philox stream (seed, state hash, ply). */
static double oz_rollout(oz_game* g, double gamma, uint64_t seed, uint64_t h0, int ply) {
  uint8_t mask[OZ_MAX_ACTIONS];
  oz_game_actions_mask(g, mask);
  int A = OZ_NACT[g->game_id], n = 0;
  for (int a = 0; a < A; a++) n += mask[a] ? 1 : 0;
  uint32_t o[4];
  oz_philox(seed, (uint32_t)ply, OZ_PURPOSE_ROLLOUT, (uint32_t)h0, (uint32_t)(h0 >> 32), o);
  int k = (int)(((uint64_t)o[0] * (uint64_t)n) >> 32), action = 0;
  for (int a = 0; a < A; a++)
    if (mask[a]) { if (k == 0) { action = a; break; } k--; }
  oz_game_play(g, action, NULL);
  double wr = oz_game_white_reward(g);
  if (oz_game_terminated(g)) return wr;
  return wr + gamma * oz_rollout(g, gamma, seed, h0, ply + 1);
}
void oz_rollout_oracle(void* ctx, int game_id, const uint8_t* state, int n, float* P, float* V) { /* :52-60 */
  const oz_rollout_ctx* rc = (const oz_rollout_ctx*)ctx;
  uint64_t key[2];
  oz_state_key(game_id, state, key);
  uint64_t h0 = oz_splitmix(key[0] ^ oz_splitmix(key[1]));
  oz_game g;
  oz_game_set_state(&g, game_id, state);
  int wp = oz_game_white_playing(&g);
  for (int i = 0; i < n; i++) P[i] = (float)(1.0 / (double)n);   /* ones(n) ./ n, stored as Float32 in ActionStats */
  double wr = oz_rollout(&g, rc->gamma, rc->seed, h0, 0);
  *V = (float)(wp ? wr : -wr);
}

/* ------------------------------------------------------------------------- */
/* MCTS (src/mcts.jl)                                                          */
/* ------------------------------------------------------------------------- */
typedef struct { float P; double W; int64_t N; } oz_astats;                /* :78-82 */
typedef struct { oz_state key; int used; int n; oz_astats stats[OZ_MAX_ACTIONS]; float Vest; } oz_info; /* :84-87 */

struct oz_env { /* :124-151 */
  oz_info* tab; size_t cap, count;
  oz_oracle_fn oracle; void* octx;
  double gamma, cpuct, noise_eps, noise_alpha, prior_temperature;
  int64_t total_simulations, total_nodes_traversed;
  int game_id;
  /* stochastic environments (grid-world): in-tree noise = stream (seed, game, move, ENV, (sim*256 + depth)*2 + {0,1});
     replaces Julia's global rand() inside act! (games/grid-world/game.jl:45-46) */
  uint64_t noise_seed, noise_game;
  uint32_t noise_move, cur_sim;
};

static uint64_t oz_hash_state(const oz_state* s) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < OZ_STATE_BYTES; i++) { h ^= s->b[i]; h *= 1099511628211ull; }
  return h;
}
oz_env* oz_env_create(int game_id, oz_oracle_fn oracle, void* octx, double gamma, double cpuct, double eps, double alpha,
                      double prior_temperature) {
  oz_env* e = (oz_env*)calloc(1, sizeof(oz_env));
  e->cap = 1024; e->tab = (oz_info*)calloc(e->cap, sizeof(oz_info));
  e->oracle = oracle; e->octx = octx; e->gamma = gamma; e->cpuct = cpuct; e->noise_eps = eps; e->noise_alpha = alpha;
  e->prior_temperature = prior_temperature; e->game_id = game_id;
  return e;
}
void oz_env_destroy(oz_env* e) { if (e) { free(e->tab); free(e); } }
void oz_env_reset(oz_env* e) { memset(e->tab, 0, e->cap * sizeof(oz_info)); e->count = 0; } /* :278-281, counters kept */
int64_t oz_env_num_nodes(const oz_env* e) { return (int64_t)e->count; }
int64_t oz_env_total_simulations(const oz_env* e) { return e->total_simulations; }
int64_t oz_env_total_nodes_traversed(const oz_env* e) { return e->total_nodes_traversed; }
void oz_env_set_noise(oz_env* e, uint64_t seed, uint64_t game, uint32_t move) { e->noise_seed = seed; e->noise_game = game; e->noise_move = move; }
static void oz_env_noise(uint64_t seed, uint64_t game, uint32_t move, uint32_t sim, uint32_t depth, double* u) {
  uint32_t idx = (sim * 256u + depth) * 2u;
  u[0] = oz_u01(oz_stream_u64(seed, game, move, OZ_PURPOSE_ENV, idx));
  u[1] = oz_u01(oz_stream_u64(seed, game, move, OZ_PURPOSE_ENV, idx + 1u));
}

static oz_info* oz_find(const oz_env* e, const oz_state* s) {
  size_t i = oz_hash_state(s) & (e->cap - 1);
  while (e->tab[i].used) {
    if (memcmp(&e->tab[i].key, s, sizeof(oz_state)) == 0) return &e->tab[i];
    i = (i + 1) & (e->cap - 1);
  }
  return NULL;
}
static oz_info* oz_insert(oz_env* e, const oz_state* s) {
  if ((e->count + 1) * 2 > e->cap) {
    size_t ocap = e->cap; oz_info* otab = e->tab;
    e->cap *= 2; e->tab = (oz_info*)calloc(e->cap, sizeof(oz_info));
    for (size_t j = 0; j < ocap; j++)
      if (otab[j].used) {
        size_t i = oz_hash_state(&otab[j].key) & (e->cap - 1);
        while (e->tab[i].used) i = (i + 1) & (e->cap - 1);
        e->tab[i] = otab[j];
      }
    free(otab);
  }
  size_t i = oz_hash_state(s) & (e->cap - 1);
  while (e->tab[i].used) i = (i + 1) & (e->cap - 1);
  e->tab[i].used = 1; e->tab[i].key = *s; e->count++;
  return &e->tab[i];
}

static int oz_argmax_d(const double* x, int n) { int k = 0; for (int i = 1; i < n; i++) if (x[i] > x[k]) k = i; return k; }
static int oz_argmax_f(const float* x, int n) { int k = 0; for (int i = 1; i < n; i++) if (x[i] > x[k]) k = i; return k; }

/* Util.apply_temperature on the oracle's Float32 vector (src/util.jl:98-110, eltype preserved) */
static void oz_apply_temperature_f32(float* p, int n, double tau) {
  if (tau == 1.0) return;
  if (tau == 0.0) {
    int k = oz_argmax_f(p, n);
    for (int i = 0; i < n; i++) p[i] = 0.0f;
    p[k] = 1.0f;
    return;
  }
  /* Float32 .^ Float64 promotes to Float64; the result is stored back into ActionStats.P::Float32 */
  double r[OZ_MAX_ACTIONS], s = 0.0, it = 1.0 / tau;
  for (int i = 0; i < n; i++) { r[i] = (p[i] > 0.0f) ? oz_det_exp(it * oz_det_log((double)p[i])) : 0.0; }
  for (int i = 0; i < n; i++) s = (i == 0) ? r[0] : s + r[i];
  for (int i = 0; i < n; i++) p[i] = (float)(r[i] / s);
}

static int oz_legal_actions(const oz_game* g, int* acts) {
  uint8_t m[OZ_MAX_ACTIONS];
  oz_game_actions_mask(g, m);
  int n = 0;
  for (int a = 0; a < OZ_NACT[g->game_id]; a++) if (m[a]) acts[n++] = a;
  return n;
}

/* state_info :165-174 + init_state_info :157-161 */
static oz_info* oz_state_info(oz_env* e, const oz_game* g, const oz_state* s, int n_legal, int* new_node) {
  oz_info* info = oz_find(e, s);
  if (info) { *new_node = 0; return info; }
  float P[OZ_MAX_ACTIONS], V = 0.0f;
  e->oracle(e->octx, e->game_id, s->b, n_legal, P, &V);
  (void)g;
  oz_apply_temperature_f32(P, n_legal, e->prior_temperature);
  info = oz_insert(e, s);
  info->n = n_legal;
  for (int i = 0; i < n_legal; i++) { info->stats[i].P = P[i]; info->stats[i].W = 0.0; info->stats[i].N = 0; }
  info->Vest = V;
  *new_node = 1;
  return info;
}

/* uct_scores :180-188 */
static void oz_uct_scores(const oz_info* info, double cpuct, double eps, const double* eta, double* scores) {
  int64_t ntot = 0;
  for (int i = 0; i < info->n; i++) ntot += info->stats[i].N;
  double sqrtN = sqrt((double)ntot);
  for (int i = 0; i < info->n; i++) {
    const oz_astats* a = &info->stats[i];
    double Q = a->W / (double)(a->N > 1 ? a->N : 1);
    double P = (eps == 0.0) ? (double)a->P : (1.0 - eps) * (double)a->P + eps * eta[i];
    scores[i] = Q + ((cpuct * P) * sqrtN) / (double)(a->N + 1);
  }
}

/* run_simulation! :199-226 (recursive, exactly as the reference) */
static double oz_run_simulation(oz_env* e, oz_game* g, const double* eta, int root, int depth) {
  if (oz_game_terminated(g)) return 0.0;
  oz_state s;
  memset(&s, 0, sizeof(s));
  oz_game_get_state(g, s.b);
  int acts[OZ_MAX_ACTIONS];
  int n = oz_legal_actions(g, acts);
  int new_node;
  oz_info* info = oz_state_info(e, g, &s, n, &new_node);
  if (new_node) return (double)info->Vest;
  double eps = root ? e->noise_eps : 0.0;
  double scores[OZ_MAX_ACTIONS];
  oz_uct_scores(info, e->cpuct, eps, eta, scores);
  int action_id = oz_argmax_d(scores, n);
  int wp = oz_game_white_playing(g);
  double u[2];
  const double* env_u = NULL;
  if (e->game_id == OZ_GRID_WORLD) { oz_env_noise(e->noise_seed, e->noise_game, e->noise_move, e->cur_sim, (uint32_t)depth, u); env_u = u; }
  oz_game_play(g, acts[action_id], env_u);
  double wr = oz_game_white_reward(g);
  double r = wp ? wr : -wr;
  int pswitch = (wp != oz_game_white_playing(g));
  double qnext = oz_run_simulation(e, g, eta, 0, depth + 1);
  if (pswitch) qnext = -qnext;
  double q = r + e->gamma * qnext;
  info = oz_find(e, &s); /* the table may have been rehashed by the recursive call */
  info->stats[action_id].W = info->stats[action_id].W + q; /* update_state_info! :190-194 */
  info->stats[action_id].N += 1;
  e->total_nodes_traversed += 1;
  return q;
}

void oz_explore(oz_env* e, const oz_game* root, int nsims, const double* eta) { /* :239-245 */
  for (int i = 0; i < nsims; i++) {
    e->total_simulations += 1;
    oz_game g = *root; /* GI.clone */
    e->cur_sim = (uint32_t)i;
    oz_run_simulation(e, &g, eta, 1, 0);
  }
}

int oz_root_stats(const oz_env* e, const oz_game* root, int64_t* N, double* W, float* P, float* Vest) {
  oz_state s;
  memset(&s, 0, sizeof(s));
  oz_game_get_state(root, s.b);
  const oz_info* info = oz_find(e, &s);
  int A = OZ_NACT[root->game_id];
  for (int a = 0; a < A; a++) { N[a] = 0; W[a] = 0.0; P[a] = 0.0f; }
  if (!info) return -1;
  int acts[OZ_MAX_ACTIONS];
  int n = oz_legal_actions(root, acts);
  for (int i = 0; i < n; i++) { N[acts[i]] = info->stats[i].N; W[acts[i]] = info->stats[i].W; P[acts[i]] = info->stats[i].P; }
  if (Vest) *Vest = info->Vest;
  return n;
}

int oz_policy(const oz_env* e, const oz_game* root, int* actions, double* pi) { /* :255-271 */
  oz_state s;
  memset(&s, 0, sizeof(s));
  oz_game_get_state(root, s.b);
  const oz_info* info = oz_find(e, &s);
  if (!info) return -1;
  int n = oz_legal_actions(root, actions);
  int64_t ntot = 0;
  for (int i = 0; i < n; i++) ntot += info->stats[i].N;
  double sum = 0.0;
  for (int i = 0; i < n; i++) { pi[i] = (double)info->stats[i].N / (double)ntot; sum = (i == 0) ? pi[0] : sum + pi[i]; }
  for (int i = 0; i < n; i++) pi[i] = pi[i] / sum;
  return n;
}

/* ------------------------------------------------------------------------- */
/* Schedules, temperature, sampling                                            */
/* ------------------------------------------------------------------------- */
double oz_pl_schedule(int n, const int* xs, const double* ys, int i) { /* src/schedule.jl:64-80 */
  int pt = -1;
  for (int k = 0; k < n; k++) if (xs[k] <= i) pt = k;
  if (pt < 0) return ys[0];
  if (pt == n - 1) return ys[n - 1];
  double x0 = xs[pt], y0 = ys[pt], x1 = xs[pt + 1], y1 = ys[pt + 1];
  return y0 + ((y1 - y0) / (x1 - x0)) * ((double)i - x0);
}
void oz_apply_temperature(const double* pi, int n, double tau, double* out) { /* src/util.jl:98-110 */
  if (tau == 1.0) { for (int i = 0; i < n; i++) out[i] = pi[i]; return; }
  if (tau == 0.0) {
    int k = oz_argmax_d(pi, n);
    for (int i = 0; i < n; i++) out[i] = 0.0;
    out[k] = 1.0;
    return;
  }
  double it = 1.0 / tau, s = 0.0;
  for (int i = 0; i < n; i++) out[i] = (pi[i] > 0.0) ? oz_det_exp(it * oz_det_log(pi[i])) : 0.0;
  for (int i = 0; i < n; i++) s = (i == 0) ? out[0] : s + out[i];
  for (int i = 0; i < n; i++) out[i] = out[i] / s;
}
void oz_fix_probvec(const double* pi, int n, float* out) { /* src/util.jl:68-81 */
  float s = 0.0f;
  for (int i = 0; i < n; i++) { out[i] = (float)pi[i]; s = (i == 0) ? out[0] : s + out[i]; }
  const float rtol = 3.4526698e-4f; /* sqrt(eps(Float32)) */
  float d = fabsf(s - 1.0f), m = fabsf(s) > 1.0f ? fabsf(s) : 1.0f;
  int approx = (s == 1.0f) || (isfinite(s) && d <= rtol * m);
  if (!approx) {
    if (s == 0.0f) for (int i = 0; i < n; i++) out[i] = 1.0f / (float)n;
    else for (int i = 0; i < n; i++) out[i] = out[i] / s;
  }
}
int oz_categorical(const float* p, int n, float u) { /* Distributions.jl rand(::DiscreteNonParametric): cumulative scan */
  float cp = p[0];
  int i = 0;
  while (cp <= u && i < n - 1) { i++; cp = cp + p[i]; }
  return i;
}

/* ------------------------------------------------------------------------- */
/* Self-play (src/play.jl:298-315, src/memory.jl:74-87, src/simulations.jl:221-241) */
/* ------------------------------------------------------------------------- */
/* GI.symmetries: connect-four = the column mirror (games/connect-four/game.jl:247-257); tic-tac-toe = the 7 non-trivial
   dihedral maps in the order rot, rot2, rot3, flip, flip.rot, flip.rot2, flip.rot3 where board'[p] = board[sym[p]]
   (games/tictactoe/game.jl:149-168); mancala and grid-world declare none. */
int oz_num_symmetries(int game_id) { return game_id == OZ_CONNECT_FOUR ? 1 : (game_id == OZ_TICTACTOE ? 7 : 0); }
static void ttt_sym_xy(int j, int* x, int* y) { /* 0-based: rot(x,y) = (y, 2-x); flip(x,y) = (x, 2-y) */
  int nrot = (j < 3) ? j + 1 : j - 3;
  for (int k = 0; k < nrot; k++) { int nx = *y, ny = 2 - *x; *x = nx; *y = ny; }
  if (j >= 3) *y = 2 - *y;
}
void oz_apply_symmetry(int game_id, int sym, const uint8_t* in, uint8_t* out) {
  int sb = OZ_SBYTES[game_id];
  memcpy(out, in, (size_t)sb);
  if (game_id == OZ_CONNECT_FOUR) {
    for (int col = 0; col < 7; col++)
      for (int row = 0; row < 6; row++) out[col + 7 * row] = in[(6 - col) + 7 * row];
  } else if (game_id == OZ_TICTACTOE) {
    for (int p = 0; p < 9; p++) {
      int x = p % 3, y = p / 3;
      ttt_sym_xy(sym, &x, &y);
      out[p] = in[y * 3 + x];
    }
  }
}

/* ------------------------------------------------------------------------- */
/* MinMax baseline player (src/minmax.jl)                                      */
/* ------------------------------------------------------------------------- */
static double oz_powi(double x, int k) { /* x ^ k for the small integer exponents of the heuristics */
  if (k < 0) return 1.0 / oz_powi(x, -k);
  double r = 1.0;
  for (int i = 0; i < k; i++) r = (i == 0) ? x : r * x;
  return r;
}
static double c4_alignment_value_for(const oz_game* g, int player, int col, int row, int dc, int dr) { /* game.jl:198-210 */
  int N = 0;
  for (int i = 0; i < 4; i++) {
    int cell = C4(g, col + i * dc, row + i * dr);
    if (cell == player) N++;
    else if (cell != 0) return 0.0;
  }
  return oz_powi(0.1, 4 - 1 - N);
}
static double c4_heuristic_value_for(const oz_game* g, int player) { /* :212-214 over ALIGNMENTS (:186-196) */
  static const int DIRS[4][2] = {{1, 1}, {1, -1}, {0, 1}, {1, 0}};
  double sum = 0.0;
  int first = 1;
  for (int d = 0; d < 4; d++)
    for (int x = 0; x < C4_COLS; x++)
      for (int y = 0; y < C4_ROWS; y++) {
        if (!c4_valid(x + 3 * DIRS[d][0], y + 3 * DIRS[d][1])) continue; /* alignment_from returns nothing (:176-184) */
        double v = c4_alignment_value_for(g, player, x, y, DIRS[d][0], DIRS[d][1]);
        sum = first ? v : sum + v;
        first = 0;
      }
  return sum;
}
static double ttt_heuristic_value_for(const oz_game* g, int player) { /* games/tictactoe/game.jl:98-114 */
  double sum = 0.0;
  for (int a = 0; a < 8; a++) {
    int N = 0, blocked = 0;
    for (int i = 0; i < 3 && !blocked; i++) {
      int m = g->cells[TTT_AL[a][i]];
      if (m == player) N++;
      else if (m != 0) blocked = 1;
    }
    double v = blocked ? 0.0 : oz_powi(0.3, 3 - 1 - N);
    sum = (a == 0) ? v : sum + v;
  }
  return sum;
}
double oz_heuristic_value(const oz_game* g) {
  switch (g->game_id) {
    case OZ_CONNECT_FOUR: return c4_heuristic_value_for(g, g->curplayer) - c4_heuristic_value_for(g, 3 - g->curplayer);
    case OZ_TICTACTOE: return ttt_heuristic_value_for(g, g->curplayer) - ttt_heuristic_value_for(g, 3 - g->curplayer);
    case OZ_MANCALA: { /* stores are UInt8 in the reference (game.jl:21): nw - nb and the negation wrap modulo 256 */
      uint8_t v = (uint8_t)(MC_STORE(g, 1) - MC_STORE(g, 2));
      if (g->curplayer == 2) v = (uint8_t)(0 - v);
      return (double)v;
    }
    default: return 0.0;
  }
}
typedef struct { int depth, amplify; double gamma; } oz_minmax;
static double mm_qvalue(const oz_minmax* p, const oz_game* g, int action, int depth);
static double mm_value(const oz_minmax* p, const oz_game* g, int depth) { /* src/minmax.jl:17-27 */
  if (oz_game_terminated(g)) return 0.0;
  if (depth == 0) return oz_heuristic_value(g);
  int acts[OZ_MAX_ACTIONS];
  int n = oz_legal_actions(g, acts);
  double best = 0.0;
  for (int i = 0; i < n; i++) {
    double q = mm_qvalue(p, g, acts[i], depth);
    best = (i == 0 || q > best) ? q : best;
  }
  return best;
}
static double mm_qvalue(const oz_minmax* p, const oz_game* g, int action, int depth) { /* :29-43 */
  oz_game next = *g;
  oz_game_play(&next, action, NULL);
  double wr = oz_game_white_reward(&next);
  double r = oz_game_white_playing(g) ? wr : -wr;
  if (p->amplify && r != 0.0) r = (r > 0.0) ? INFINITY : -INFINITY; /* amplify :14 */
  double nextv = mm_value(p, &next, depth - 1);
  if (oz_game_white_playing(g) != oz_game_white_playing(&next)) nextv = -nextv;
  return r + p->gamma * nextv;
}
int oz_minmax_think(const oz_game* g, int depth, int amplify, double tau, double gamma, int* acts, double* pi, double* qs_out) { /* :83-114 */
  oz_minmax p = {depth, amplify, gamma};
  double qs[OZ_MAX_ACTIONS];
  int n = oz_legal_actions(g, acts);
  int nwin = 0, nnot = 0, best = 0;
  for (int i = 0; i < n; i++) qs[i] = mm_qvalue(&p, g, acts[i], depth);
  for (int i = 0; i < n; i++) {
    nwin += (qs[i] == INFINITY);
    nnot += (qs[i] > -INFINITY);
    if (qs[i] > qs[best]) best = i;
  }
  if (nwin > 0) {
    for (int i = 0; i < n; i++) pi[i] = (qs[i] == INFINITY) ? 1.0 : 0.0;
  } else if (nnot == 0) {
    for (int i = 0; i < n; i++) pi[i] = 1.0;
  } else if (tau == 0.0) {
    for (int i = 0; i < n; i++) pi[i] = (qs[i] == qs[best]) ? 1.0 : 0.0;
  } else {
    double qmax = qs[best], Cn = 0.0;
    int hc = 0;
    for (int i = 0; i < n; i++)
      if (qs[i] > -INFINITY) { double a = fabs(qs[i]); Cn = hc ? (a > Cn ? a : Cn) : a; hc = 1; }
    Cn = Cn + 2.220446049250313e-16; /* eps() */
    double it = 1.0 / tau;
    for (int i = 0; i < n; i++) {
      double x = (qs[i] - qmax) / Cn;
      double e = (x == -INFINITY) ? 0.0 : oz_det_exp(x);
      pi[i] = (e > 0.0) ? oz_det_exp(it * oz_det_log(e)) : 0.0; /* pi .^= 1 / tau */
    }
  }
  double s = 0.0;
  for (int i = 0; i < n; i++) s = (i == 0) ? pi[0] : s + pi[i];
  for (int i = 0; i < n; i++) pi[i] = pi[i] / s;
  if (qs_out) for (int i = 0; i < n; i++) qs_out[i] = qs[i];
  return n;
}

void oz_play_game2(oz_env* white, oz_env* black, const oz_mcts_params* mp, double flip_p, uint64_t seed, uint64_t game_idx,
                   oz_trace* tr) {
  oz_play_game2p(white, mp, black, mp, flip_p, seed, game_idx, tr);
}
/* TwoPlayers of two MctsPlayers that may differ in every MctsParams field (src/play.jl:248-282, src/benchmark.jl:78-99):
   think / player_temperature dispatch on the colour to move (:258-264, :279-282).  num_iters_per_turn == 0 stands for a
   NetworkPlayer with the temperature schedule of the same parameter block (Benchmark.NetworkOnly). */
void oz_play_game2p(oz_env* white, const oz_mcts_params* mp_white, oz_env* black, const oz_mcts_params* mp_black, double flip_p,
                    uint64_t seed, uint64_t game_idx, oz_trace* tr) {
  const int gid = white->game_id;
  oz_game g;
  oz_game_init(&g, gid);
  if (gid == OZ_GRID_WORLD) { /* RL.reset!: random start cell (games/grid-world/game.jl:36) */
    uint32_t o[4];
    uint8_t st[2];
    oz_philox(seed, 0, OZ_PURPOSE_POSITION, (uint32_t)game_idx, (uint32_t)(game_idx >> 32), o);
    st[0] = (uint8_t)(1 + o[0] % 10u); st[1] = (uint8_t)(1 + o[1] % 10u);
    oz_game_set_state(&g, OZ_GRID_WORLD, st);
  }
  int A = OZ_NACT[gid];
  memset(tr, 0, sizeof(*tr));
  oz_game_get_state(&g, tr->states[0]);
  int n = 0;
  while (!oz_game_terminated(&g) && n < OZ_MAX_PLIES) {
    int acts[OZ_MAX_ACTIONS];
    double eta[OZ_MAX_ACTIONS], pi[OZ_MAX_ACTIONS], pis[OZ_MAX_ACTIONS];
    float pf[OZ_MAX_ACTIONS];
    if (flip_p != 0.0) { /* play.jl:305-307; game.jl:329-336 */
      int ns = oz_num_symmetries(gid);
      double u = oz_u01(oz_stream_u64(seed, game_idx, (uint32_t)n, OZ_PURPOSE_SYMMETRY, 0));
      if (ns > 0 && u < flip_p) {
        int j = (int)(oz_stream_u64(seed, game_idx, (uint32_t)n, OZ_PURPOSE_SYMMETRY, 1) % (uint64_t)ns);
        uint8_t cur[OZ_STATE_BYTES] = {0}, img[OZ_STATE_BYTES] = {0};
        oz_game_get_state(&g, cur);
        oz_apply_symmetry(gid, j, cur, img);
        oz_game_set_state(&g, gid, img);
        tr->sym[n] = j + 1;
      }
    }
    oz_game_get_state(&g, tr->think_states[n]);
    oz_env* env = oz_game_white_playing(&g) ? white : black;              /* think(::TwoPlayers): play.jl:258-264 */
    const oz_mcts_params* mp = oz_game_white_playing(&g) ? mp_white : mp_black;
    int nl = oz_legal_actions(&g, acts);
    if (mp->player_kind == 1) { /* MinMax.Player */
      oz_minmax_think(&g, mp->minmax_depth, mp->minmax_amplify, mp->minmax_tau, mp->gamma, acts, pi, NULL);
    } else if (mp->num_iters_per_turn == 0) {
      /* NetworkPlayer under PlayerWithTemperature = Benchmark.NetworkOnly (src/play.jl:226-235, :112-127,
         src/benchmark.jl:166-176): think returns the oracle's policy over the available actions, no search */
      float P[OZ_MAX_ACTIONS], V = 0.0f;
      env->oracle(env->octx, gid, tr->think_states[n], nl, P, &V);
      for (int i = 0; i < nl; i++) pi[i] = (double)P[i];
    } else {
      oz_dirichlet(seed, game_idx, (uint32_t)n, nl, mp->noise_alpha, eta);  /* drawn even if eps == 0 (mcts.jl:240) */
      oz_env_set_noise(env, seed, game_idx, (uint32_t)n);
      oz_explore(env, &g, mp->num_iters_per_turn, eta);                      /* think: play.jl:196-206 */
      oz_policy(env, &g, acts, pi);
    }
    double tau = mp->player_kind == 1 ? 1.0 : oz_pl_schedule(mp->sched_n, mp->sched_xs, mp->sched_ys, n); /* schedule[length(trace)] */
    oz_apply_temperature(pi, nl, tau, pis);
    oz_fix_probvec(pis, nl, pf);
    float u = oz_uniform_f32(seed, game_idx, (uint32_t)n, OZ_PURPOSE_CATEGORICAL, 0);
    int k = oz_categorical(pf, nl, u);
    for (int a = 0; a < A; a++) { tr->pi[n][a] = 0.0f; tr->pi64[n][a] = 0.0; tr->mask[n][a] = 0; }
    for (int i = 0; i < nl; i++) { tr->pi[n][acts[i]] = (float)pi[i]; tr->pi64[n][acts[i]] = pi[i]; tr->mask[n][acts[i]] = 1; }
    tr->action[n] = acts[k];
    {
      double u[2];
      const double* env_u = NULL;
      if (gid == OZ_GRID_WORLD) { oz_env_noise(seed, game_idx, (uint32_t)n, 0x7FFFFFu, 0u, u); env_u = u; }
      oz_game_play(&g, acts[k], env_u);
    }
    tr->rewards[n] = oz_game_white_reward(&g);
    n++;
    oz_game_get_state(&g, tr->states[n]);
  }
  tr->n_moves = n;
  double wr = 0.0; /* push_trace! */
  for (int i = n - 1; i >= 0; i--) {
    wr = white->gamma * wr + tr->rewards[i];
    oz_game gi;
    oz_game_set_state(&gi, gid, tr->states[i]);
    tr->z[i] = oz_game_white_playing(&gi) ? wr : -wr;
    tr->t[i] = (double)(n - i);
  }
  if (white == black) {
    tr->mem_nodes = (int64_t)white->count;
    tr->edepth = white->total_simulations == 0 ? 0.0 : (double)white->total_nodes_traversed / (double)white->total_simulations;
  } else {
    int64_t ts = white->total_simulations + black->total_simulations;
    tr->mem_nodes = (int64_t)white->count + (int64_t)black->count;
    tr->edepth = ts == 0 ? 0.0 : (double)(white->total_nodes_traversed + black->total_nodes_traversed) / (double)ts;
  }
}
void oz_play_game(oz_env* env, const oz_mcts_params* mp, uint64_t seed, uint64_t game_idx, oz_trace* tr) {
  oz_play_game2(env, env, mp, 0.0, seed, game_idx, tr);
}
double oz_total_reward(const oz_trace* tr, double gamma) { /* sum(gamma^(i-1) * r_i), left to right */
  double s = 0.0, gp = 1.0;
  for (int i = 0; i < tr->n_moves; i++) { s = (i == 0) ? gp * tr->rewards[0] : s + gp * tr->rewards[i]; gp = gp * gamma; }
  return s;
}

void oz_worker_run(int game_id, oz_oracle_fn oracle, void* octx, const oz_mcts_params* mp, uint64_t seed, uint64_t first,
                   uint64_t stride, int count, int reset_every, oz_trace* out) {
  oz_env* env = oz_env_create(game_id, oracle, octx, mp->gamma, mp->cpuct, mp->noise_eps, mp->noise_alpha, mp->prior_temperature);
  for (int i = 0; i < count; i++) {
    oz_play_game(env, mp, seed, first + (uint64_t)i * stride, &out[i]);
    if (reset_every > 0 && (i + 1) % reset_every == 0) oz_env_reset(env);
  }
  oz_env_destroy(env);
}

void oz_random_position(int game_id, uint64_t seed, uint64_t stream, int max_plies, uint8_t* state) {
  if (game_id == OZ_GRID_WORLD) {
    for (uint64_t attempt = 0;; attempt++) {
      uint64_t st = stream + (attempt << 32);
      uint32_t o[4];
      oz_philox(seed, 0, OZ_PURPOSE_POSITION, (uint32_t)st, (uint32_t)(st >> 32), o);
      int x = 1 + (int)(o[0] % 10u), y = 1 + (int)(o[1] % 10u);
      if (!gw_has_reward(x, y)) { state[0] = (uint8_t)x; state[1] = (uint8_t)y; return; }
    }
  }
  for (uint64_t attempt = 0;; attempt++) {
    uint64_t st = stream + (attempt << 32);
    oz_game g;
    oz_game_init(&g, game_id);
    uint32_t o[4];
    oz_philox(seed, 0, OZ_PURPOSE_POSITION, (uint32_t)st, (uint32_t)(st >> 32), o);
    int k = (int)(o[0] % (uint32_t)(max_plies + 1));
    int ok = 1;
    for (int ply = 0; ply < k; ply++) {
      int acts[OZ_MAX_ACTIONS];
      int n = oz_legal_actions(&g, acts);
      oz_philox(seed, (uint32_t)(ply + 1), OZ_PURPOSE_POSITION, (uint32_t)st, (uint32_t)(st >> 32), o);
      oz_game_play(&g, acts[o[0] % (uint32_t)n], NULL);
      if (oz_game_terminated(&g)) { ok = 0; break; }
    }
    if (ok) { memset(state, 0, (size_t)OZ_SBYTES[game_id]); oz_game_get_state(&g, state); return; }
  }
}

/* ------------------------------------------------------------------------- */
/* Batched lock-step driver: same algorithm with an explicit path stack so the */
/* oracle call can be deferred and answered in batches (CPU baseline only).    */
/* ------------------------------------------------------------------------- */
typedef struct { oz_state s; int action_id; double r; int pswitch; } oz_step;
typedef struct {
  oz_env* env;
  oz_game root;
  double eta[OZ_MAX_ACTIONS];
  int has_eta;
  int sims_done;
  int pending;        /* waiting for an oracle answer */
  oz_state leaf;
  int leaf_nlegal;
  int leaf_acts[OZ_MAX_ACTIONS];
  oz_step path[OZ_MAX_PLIES];
  int depth;
  int sims_local; /* simulations finished since the last oz_batch_advance (summed outside the parallel loop) */
} oz_tree;
struct oz_batch {
  int game_id, n, nsims;
  oz_mcts_params mp;
  oz_tree* t;
  int32_t* pend;
  int npend;
  int64_t expansions, sims;
};
oz_batch* oz_batch_create(int game_id, int n, const oz_mcts_params* mp) {
  oz_batch* b = (oz_batch*)calloc(1, sizeof(*b));
  b->game_id = game_id; b->n = n; b->mp = *mp; b->nsims = mp->num_iters_per_turn;
  b->t = (oz_tree*)calloc((size_t)n, sizeof(oz_tree));
  b->pend = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  for (int i = 0; i < n; i++)
    b->t[i].env = oz_env_create(game_id, NULL, NULL, mp->gamma, mp->cpuct, mp->noise_eps, mp->noise_alpha, mp->prior_temperature);
  return b;
}
void oz_batch_destroy(oz_batch* b) {
  if (!b) return;
  for (int i = 0; i < b->n; i++) oz_env_destroy(b->t[i].env);
  free(b->t); free(b->pend); free(b);
}
void oz_batch_reset_trees(oz_batch* b) { for (int i = 0; i < b->n; i++) oz_env_reset(b->t[i].env); }
void oz_batch_set_roots(oz_batch* b, const uint8_t* states, const double* eta) {
  int A = OZ_NACT[b->game_id], sb = OZ_SBYTES[b->game_id];
  for (int i = 0; i < b->n; i++) {
    oz_tree* t = &b->t[i];
    oz_game_set_state(&t->root, b->game_id, states + (size_t)i * sb);
    t->has_eta = eta != NULL;
    if (eta) for (int a = 0; a < A; a++) t->eta[a] = eta[(size_t)i * A + a];
    t->sims_done = 0; t->pending = 0; t->depth = 0;
  }
}
static void oz_tree_backup(oz_tree* t, double q) {
  oz_env* e = t->env;
  for (int d = t->depth - 1; d >= 0; d--) {
    oz_step* st = &t->path[d];
    if (st->pswitch) q = -q;
    q = st->r + e->gamma * q;
    oz_info* info = oz_find(e, &st->s);
    info->stats[st->action_id].W = info->stats[st->action_id].W + q;
    info->stats[st->action_id].N += 1;
    e->total_nodes_traversed += 1;
  }
}
/* run one tree until it needs an answer (returns 1) or has done all its simulations (returns 0) */
static int oz_tree_advance(oz_batch* b, oz_tree* t) {
  oz_env* e = t->env;
  while (t->sims_done < b->nsims) {
    e->total_simulations += 1;
    oz_game g = t->root;
    t->depth = 0;
    int root = 1;
    for (;;) {
      if (oz_game_terminated(&g)) { oz_tree_backup(t, 0.0); break; }
      oz_state s;
      memset(&s, 0, sizeof(s));
      oz_game_get_state(&g, s.b);
      int acts[OZ_MAX_ACTIONS];
      int n = oz_legal_actions(&g, acts);
      oz_info* info = oz_find(e, &s);
      if (!info) {
        t->leaf = s; t->leaf_nlegal = n; memcpy(t->leaf_acts, acts, sizeof(acts)); t->pending = 1;
        return 1;
      }
      double scores[OZ_MAX_ACTIONS];
      oz_uct_scores(info, e->cpuct, root ? e->noise_eps : 0.0, t->eta, scores);
      int k = oz_argmax_d(scores, n);
      int wp = oz_game_white_playing(&g);
      oz_game_play(&g, acts[k], NULL);
      double wr = oz_game_white_reward(&g);
      oz_step* st = &t->path[t->depth++];
      st->s = s; st->action_id = k; st->r = wp ? wr : -wr; st->pswitch = (wp != oz_game_white_playing(&g));
      root = 0;
    }
    t->sims_done++;
    t->sims_local++;
  }
  return 0;
}
/* ---- host threads of the lock-step batch driver: a plain pthread parallel-for over independent trees / leaves (the
   image has no OpenMP runtime for gcc).  Static contiguous chunks; the result never depends on the thread count. ---- */
#include <pthread.h>
static int oz_nthreads = 1;
void oz_set_threads(int n) { oz_nthreads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int oz_get_threads(void) { return oz_nthreads; }
typedef void (*oz_range_fn)(void* ctx, int lo, int hi);
typedef struct { oz_range_fn fn; void* ctx; int lo, hi; } oz_job;
static void* oz_job_run(void* p) { oz_job* j = (oz_job*)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }
static void oz_parallel_for(int n, oz_range_fn fn, void* ctx) {
  int nt = oz_nthreads < n ? oz_nthreads : n;
  if (nt <= 1) { fn(ctx, 0, n); return; }
  pthread_t th[256];
  oz_job job[256];
  for (int k = 0; k < nt; k++) {
    job[k].fn = fn; job[k].ctx = ctx;
    job[k].lo = (int)((long long)n * k / nt); job[k].hi = (int)((long long)n * (k + 1) / nt);
    if (k > 0 && pthread_create(&th[k], NULL, oz_job_run, &job[k]) != 0) { job[k].fn(ctx, job[k].lo, job[k].hi); th[k] = 0; job[k].fn = NULL; }
  }
  oz_job_run(&job[0]);
  for (int k = 1; k < nt; k++) if (job[k].fn) pthread_join(th[k], NULL);
}
/* Trees are independent (one MCTS.Env per worker, src/simulations.jl:217-218): the CPU baseline runs them on all host
   threads like the reference's Util.mapreduce over worker tasks (src/util.jl:169-200).  Leaves are collected in tree
   order afterwards, so the result does not depend on the thread count. */
static void oz_advance_range(void* ctx, int lo, int hi) {
  oz_batch* b = (oz_batch*)ctx;
  for (int i = lo; i < hi; i++) {
    oz_tree* t = &b->t[i];
    if (!t->pending) oz_tree_advance(b, t);
  }
}
int oz_batch_advance(oz_batch* b, uint8_t* leaf_states, int32_t* leaf_tree) {
  int sb = OZ_SBYTES[b->game_id];
  b->npend = 0;
  oz_parallel_for(b->n, oz_advance_range, b);
  for (int i = 0; i < b->n; i++) {
    oz_tree* t = &b->t[i];
    b->sims += t->sims_local;
    t->sims_local = 0;
    if (t->pending) {
      memcpy(leaf_states + (size_t)b->npend * sb, t->leaf.b, (size_t)sb);
      leaf_tree[b->npend] = i;
      b->pend[b->npend++] = i;
    }
  }
  return b->npend;
}
/* GI.vectorize_state + GI.actions_mask of the pending leaves (the Batchifier's `Flux.batch(vectorize_state.(...))`,
   src/networks/network.jl:310-312), on all host threads: X[npend][state_dim floats], mask[npend][A] */
typedef struct { const oz_batch* b; const uint8_t* leaf_states; int xdim; float* X; uint8_t* mask; } oz_vec_args;
static void oz_vectorize_range(void* ctx, int lo, int hi) {
  oz_vec_args* v = (oz_vec_args*)ctx;
  const oz_batch* b = v->b;
  int sb = OZ_SBYTES[b->game_id], A = OZ_NACT[b->game_id];
  for (int j = lo; j < hi; j++) {
    oz_vectorize_state(b->game_id, v->leaf_states + (size_t)j * sb, v->X + (size_t)j * v->xdim);
    const oz_tree* t = &b->t[b->pend[j]];
    for (int a = 0; a < A; a++) v->mask[(size_t)j * A + a] = 0;
    for (int i = 0; i < t->leaf_nlegal; i++) v->mask[(size_t)j * A + t->leaf_acts[i]] = 1;
  }
}
void oz_batch_vectorize(const oz_batch* b, const uint8_t* leaf_states, int n, int xdim, float* X, uint8_t* mask) {
  oz_vec_args v = {b, leaf_states, xdim, X, mask};
  oz_parallel_for(n, oz_vectorize_range, &v);
}
typedef struct { oz_batch* b; const float* P; const float* V; } oz_feed_args;
static void oz_feed_range(void* ctx, int lo, int hi) {
  oz_feed_args* f = (oz_feed_args*)ctx;
  oz_batch* b = f->b;
  const float* P = f->P; const float* V = f->V;
  int A = OZ_NACT[b->game_id];
  for (int j = lo; j < hi; j++) {
    oz_tree* t = &b->t[b->pend[j]];
    oz_env* e = t->env;
    float p[OZ_MAX_ACTIONS];
    for (int i = 0; i < t->leaf_nlegal; i++) p[i] = P[(size_t)j * A + t->leaf_acts[i]];
    oz_apply_temperature_f32(p, t->leaf_nlegal, e->prior_temperature);
    oz_info* info = oz_insert(e, &t->leaf);
    info->n = t->leaf_nlegal;
    for (int i = 0; i < t->leaf_nlegal; i++) { info->stats[i].P = p[i]; info->stats[i].W = 0.0; info->stats[i].N = 0; }
    info->Vest = V[j];
    oz_tree_backup(t, (double)info->Vest);
    t->pending = 0;
    t->sims_done++;
  }
}
void oz_batch_feed(oz_batch* b, const float* P, const float* V) {
  oz_feed_args f = {b, P, V};
  oz_parallel_for(b->npend, oz_feed_range, &f);
  b->sims += b->npend;
  b->expansions += b->npend;
  b->npend = 0;
}
void oz_batch_root_stats(const oz_batch* b, int tree, int64_t* N, double* W, float* P) {
  oz_root_stats(b->t[tree].env, &b->t[tree].root, N, W, P, NULL);
}
int64_t oz_batch_total_expansions(const oz_batch* b) { return b->expansions; }
int64_t oz_batch_total_simulations(const oz_batch* b) { return b->sims; }

/* ------------------------------------------------------------------------- */
/* Test helper: exact Connect-Four negamax in the score convention of the       */
/* known-answer files games/connect-four/benchmark/Test_L*_R* (score > 0: the  */
/* player to move wins, (43 - stones at the win)/2; 0 draw; < 0 loses).        */
/* ------------------------------------------------------------------------- */
static int c4_negamax(const oz_game* g, int nb, int alpha, int beta) {
  if (nb == 42) return 0;
  for (int col = 0; col < 7; col++)
    if (c4_first_free(g, col) < C4_ROWS) {
      oz_game h = *g;
      c4_play(&h, col);
      if (h.winner) return (43 - nb) / 2;
    }
  int max = (41 - nb) / 2;
  if (beta > max) { beta = max; if (alpha >= beta) return beta; }
  static const int ORDER[7] = {3, 2, 4, 1, 5, 0, 6};
  for (int i = 0; i < 7; i++) {
    int col = ORDER[i];
    if (c4_first_free(g, col) >= C4_ROWS) continue;
    oz_game h = *g;
    c4_play(&h, col);
    int score = -c4_negamax(&h, nb + 1, -beta, -alpha);
    if (score >= beta) return score;
    if (score > alpha) alpha = score;
  }
  return alpha;
}
int oz_c4_solve(const uint8_t* state) {
  oz_game g;
  oz_game_set_state(&g, OZ_CONNECT_FOUR, state);
  int nb = 0;
  for (int i = 0; i < 42; i++) nb += (state[i] != 0);
  return c4_negamax(&g, nb, -22, 22);
}
