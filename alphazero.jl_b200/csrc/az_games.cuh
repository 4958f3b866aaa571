// az_games.cuh -- per-game device functions (game step / legal moves / terminal reward / NN encoding).
//
// Each game is a struct of static __host__ __device__ functions over a 16-byte state key
// (a, b) plus a small `aux` word for status that the reference keeps in the mutable GameEnv
// but not in the state.  The same inline code runs inside the tree kernels and in the host
// helpers of the C ABI, so there is one implementation of the rules in the product.
//
// Reference behaviour followed (paths relative to the reference root):
//   connect-four: games/connect-four/game.jl:50-68 (set_state!), :87-99 (first_free, mask),
//                 :123-146 (win test, play!), :160-168 (white_reward), :226-241 (vectorize_state)
//   tictactoe:    games/tictactoe/game.jl:53-92, :126-143
//   mancala:      games/mancala/game.jl:54-60, :80-177, :199-206, :224-257
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#define AZ_HD __host__ __device__ __forceinline__
#include "az_rng.cuh"

struct AzEnv {
  uint64_t a, b;  // canonical state key (no table tag bits)
  uint32_t aux;   // status outside the state: bit0 finished, bits1-2 winner
};

// uniform draws consumed by a stochastic environment step (grid-world); ignored by the board games
struct AzNoise { double u0, u1; };

AZ_HD int az_popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

// ------------------------------------------------------------------------------------------
// Connect Four: a = white stones, b = black stones | (black to move) << 56; bit = col*7 + row.
// ------------------------------------------------------------------------------------------
struct GameC4 {
  static constexpr int ID = 0;
  static constexpr int A = 7;
  static constexpr int LANES = 8;          // 16-byte lanes per node line: key + 7 edges = 128 B
  static constexpr int STATE_BYTES = 43;
  static constexpr int MAX_PLIES = 42;
  static constexpr int XW = 7, XH = 6, XC = 3;
  static constexpr bool ACYCLIC = true;
  static constexpr uint64_t BOARD = 0x0000FDFBF7EFDFBFull;  // 6 rows x 7 columns, bit 6 of each column clear
  static constexpr uint64_t PLAYER = 1ull << 56;

  AZ_HD static bool aligned4(uint64_t bb) {
    uint64_t m = bb & (bb >> 7);
    if (m & (m >> 14)) return true;  // horizontal
    m = bb & (bb >> 6);
    if (m & (m >> 12)) return true;  // diagonal
    m = bb & (bb >> 8);
    if (m & (m >> 16)) return true;  // diagonal
    m = bb & (bb >> 1);
    return (m & (m >> 2)) != 0;      // vertical
  }
  AZ_HD static uint64_t occ(const AzEnv& e) { return (e.a | e.b) & BOARD; }
  AZ_HD static int height(uint64_t o, int col) { return az_popc64((o >> (7 * col)) & 0x3F); }
  AZ_HD static bool terminated(const AzEnv& e) { return e.aux & 1; }
  AZ_HD static bool white_playing(const AzEnv& e) { return (e.b & PLAYER) == 0; }
  AZ_HD static uint32_t legal_mask(const AzEnv& e) {
    uint64_t o = occ(e);
    uint32_t m = 0;
    for (int c = 0; c < 7; c++) m |= (uint32_t)(((o >> (7 * c + 5)) & 1) ^ 1) << c;
    return m;
  }
  AZ_HD static double white_reward(const AzEnv& e) {
    uint32_t w = (e.aux >> 1) & 3;
    return (e.aux & 1) ? (w == 1 ? 1.0 : (w == 2 ? -1.0 : 0.0)) : 0.0;
  }
  AZ_HD static AzEnv play(const AzEnv& e, int col) {
    AzEnv n = e;
    uint64_t o = occ(e);
    uint64_t bit = 1ull << (7 * col + height(o, col));
    bool white = white_playing(e);
    uint64_t mine;
    if (white) { n.a |= bit; mine = n.a; } else { n.b |= bit; mine = n.b & BOARD; }
    n.b ^= PLAYER;
    if (aligned4(mine)) n.aux = 1u | ((white ? 1u : 2u) << 1);
    else n.aux = (((o | bit) & BOARD) == BOARD) ? 1u : 0u;
    return n;
  }
  AZ_HD static AzEnv init() { AzEnv e = {0, 0, 0}; return e; }
  static constexpr bool STOCHASTIC = false;
  static constexpr bool HAS_PLANE = true;  // plane(e, col, row, c) gives one element of vectorize_state directly
  static constexpr long long MAX_STATES = 1ll << 40;  // no useful bound on distinct states
  AZ_HD static AzEnv play(const AzEnv& e, int a, const AzNoise&) { return play(e, a); }
  AZ_HD static AzEnv init_game(uint64_t, uint64_t) { return init(); }
  // GI.symmetries (game.jl:247-257): the column mirror.  Only applied to non-terminal states (src/play.jl:302-307), whose
  // status bits are symmetric.
  static constexpr int NSYM = 1;
  // action permutation of the symmetry (the sigma of GI.symmetries): pi'[p] = pi[sym_source(j, p)] (src/memory.jl:112-124)
  AZ_HD static int sym_source(int, int p) { return 6 - p; }
  AZ_HD static AzEnv symmetry(const AzEnv& e, int) {
    AzEnv n = {0, e.b & PLAYER, e.aux};
    for (int col = 0; col < 7; col++) {
      n.a |= ((e.a >> (7 * col)) & 0x7Full) << (7 * (6 - col));
      n.b |= ((e.b >> (7 * col)) & 0x7Full) << (7 * (6 - col));
    }
    return n;
  }
  // set_state! (game.jl:50-68): finished if no free column, or if the TOP stone of some column is part of a
  // winning pattern of its colour
  static AzEnv from_bytes(const uint8_t* s) {
    AzEnv e = {0, 0, 0};
    for (int col = 0; col < 7; col++)
      for (int row = 0; row < 6; row++) {
        int c = s[col + 7 * row];
        if (c == 1) e.a |= 1ull << (col * 7 + row);
        if (c == 2) e.b |= 1ull << (col * 7 + row);
      }
    if (s[42] == 2) e.b |= PLAYER;
    uint64_t o = occ(e);
    if (o == BOARD) e.aux = 1;
    for (int col = 0; col < 7; col++) {
      int top = 0;
      while (top < 6 && s[col + 7 * top] != 0) top++;
      if (top == 0) continue;
      int row = top - 1, c = s[col + 7 * row];
      static const int AX[4][2] = {{1, 1}, {1, -1}, {1, 0}, {0, 1}};
      bool win = false;
      for (int ax = 0; ax < 4 && !win; ax++) {
        int n = 1;
        for (int sg = -1; sg <= 1; sg += 2) {
          int cc = col + sg * AX[ax][0], rr = row + sg * AX[ax][1];
          while (cc >= 0 && cc < 7 && rr >= 0 && rr < 6 && s[cc + 7 * rr] == c) { n++; cc += sg * AX[ax][0]; rr += sg * AX[ax][1]; }
        }
        win = n >= 4;
      }
      if (win) { e.aux = 1u | ((uint32_t)c << 1); break; }
    }
    return e;
  }
  AZ_HD static void to_bytes(const AzEnv& e, uint8_t* s) {
    for (int col = 0; col < 7; col++)
      for (int row = 0; row < 6; row++) {
        int bit = col * 7 + row;
        s[col + 7 * row] = (uint8_t)(((e.a >> bit) & 1) ? 1 : (((e.b >> bit) & 1) ? 2 : 0));
      }
    s[42] = white_playing(e) ? 1 : 2;
  }
  // GI.heuristic_value (game.jl:172-220), the leaf value of the MinMax baseline: every alignment of TO_CONNECT cells that
  // holds none of the opponent's stones is worth 0.1 ^ (3 - N) with N own stones; the alignments are summed in the order of
  // the ALIGNMENTS table (:192-196: directions (1,1), (1,-1), (0,1), (1,0); start cell x outer, y inner), own minus opponent.
  // Bit (x, y) = 7 x + y, so an alignment is one 4-bit pattern shifted to its start cell.  0.1^2, 0.1^3 as products: the
  // correctly rounded powers and the products are the same doubles for this base.
  static constexpr bool HAS_HEURISTIC = true;
  AZ_HD static double heuristic_for(uint64_t mine, uint64_t theirs) {
    const double pw[5] = {0.1 * 0.1 * 0.1, 0.1 * 0.1, 0.1, 1.0, 1.0 / 0.1};
    double sum = 0.0;
    bool first = true;
    for (int d = 0; d < 4; d++) {
      const int dx = d == 2 ? 0 : 1, dy = d == 0 ? 1 : (d == 1 ? -1 : (d == 2 ? 1 : 0));
      const int step = 7 * dx + dy;
      const uint64_t pat = 1ull | (1ull << step) | (1ull << (2 * step)) | (1ull << (3 * step));
      for (int x = 0; x < 7; x++)
        for (int y = 0; y < 6; y++) {
          const int xe = x + 3 * dx, ye = y + 3 * dy;
          if (xe > 6 || ye < 0 || ye > 5) continue;
          const uint64_t m = pat << (7 * x + y);   // all four steps (8, 6, 1, 7) are positive bit distances
          const double v = (theirs & m) ? 0.0 : pw[az_popc64(mine & m)];
          sum = first ? v : sum + v;
          first = false;
        }
    }
    return sum;
  }
  AZ_HD static double heuristic_value(const AzEnv& e) {
    const uint64_t w = e.a, k = e.b & BOARD;
    return white_playing(e) ? heuristic_for(w, k) - heuristic_for(k, w) : heuristic_for(k, w) - heuristic_for(w, k);
  }
  // vectorize_state: x[col + 7*row + 42*c], c in {empty, current player, opponent}
  AZ_HD static float plane(const AzEnv& e, int col, int row, int c) {
    int bit = col * 7 + row;
    bool w = (e.a >> bit) & 1, k = (e.b >> bit) & 1;
    bool mine = white_playing(e) ? w : k, theirs = white_playing(e) ? k : w;
    return c == 0 ? (float)(!w && !k) : (c == 1 ? (float)mine : (float)theirs);
  }
  AZ_HD static void vectorize(const AzEnv& e, float* x) {
    for (int c = 0; c < 3; c++)
      for (int row = 0; row < 6; row++)
        for (int col = 0; col < 7; col++) x[col + 7 * row + 42 * c] = plane(e, col, row, c);
  }
};

// ------------------------------------------------------------------------------------------
// Tic-tac-toe: a = white bits | black bits << 16 | (black to move) << 32; pos = (y-1)*3 + (x-1).
// ------------------------------------------------------------------------------------------
struct GameTTT {
  static constexpr int ID = 1;
  static constexpr int A = 9;
  static constexpr int LANES = 16;  // key + 9 edges -> 256-byte node line
  static constexpr int STATE_BYTES = 10;
  static constexpr int MAX_PLIES = 9;
  static constexpr int XW = 3, XH = 3, XC = 3;
  static constexpr bool ACYCLIC = true;
  static constexpr uint64_t PLAYER = 1ull << 32;

  AZ_HD static bool won(uint32_t m) {
    return (m & 0x049) == 0x049 || (m & 0x092) == 0x092 || (m & 0x124) == 0x124 || (m & 0x007) == 0x007 ||
           (m & 0x038) == 0x038 || (m & 0x1C0) == 0x1C0 || (m & 0x111) == 0x111 || (m & 0x054) == 0x054;
  }
  AZ_HD static uint32_t wbits(const AzEnv& e) { return (uint32_t)e.a & 0x1FF; }
  AZ_HD static uint32_t kbits(const AzEnv& e) { return (uint32_t)(e.a >> 16) & 0x1FF; }
  AZ_HD static bool terminated(const AzEnv& e) { return won(wbits(e)) || won(kbits(e)) || ((wbits(e) | kbits(e)) == 0x1FF); }
  AZ_HD static bool white_playing(const AzEnv& e) { return (e.a & PLAYER) == 0; }
  AZ_HD static uint32_t legal_mask(const AzEnv& e) { return ~(wbits(e) | kbits(e)) & 0x1FF; }
  AZ_HD static double white_reward(const AzEnv& e) { return won(wbits(e)) ? 1.0 : (won(kbits(e)) ? -1.0 : 0.0); }
  AZ_HD static AzEnv play(const AzEnv& e, int pos) {
    AzEnv n = e;
    n.a |= white_playing(e) ? (1ull << pos) : (1ull << (16 + pos));
    n.a ^= PLAYER;
    return n;
  }
  AZ_HD static AzEnv init() { AzEnv e = {0, 0, 0}; return e; }
  static constexpr bool STOCHASTIC = false;
  static constexpr bool HAS_PLANE = false;  // plane(e, col, row, c) gives one element of vectorize_state directly
  static constexpr long long MAX_STATES = 5478;  // reachable tic-tac-toe positions: a tree never holds more nodes
  AZ_HD static AzEnv play(const AzEnv& e, int a, const AzNoise&) { return play(e, a); }
  AZ_HD static AzEnv init_game(uint64_t, uint64_t) { return init(); }
  // GI.symmetries (game.jl:149-168): rot, rot2, rot3, flip, flip.rot, flip.rot2, flip.rot3 with rot(x,y) = (y, N-x+1),
  // flip(x,y) = (x, N-y+1); the image board is board'[p] = board[sym[p]]
  static constexpr int NSYM = 7;
  AZ_HD static int sym_source(int j, int p) {  // sym[p] of SYMMETRIES[j]: board'[p] = board[sym[p]], pi'[p] = pi[sym[p]]
    const int nrot = j < 3 ? j + 1 : j - 3;
    int x = p % 3, y = p / 3;
    for (int k = 0; k < nrot; k++) { int nx = y, ny = 2 - x; x = nx; y = ny; }
    if (j >= 3) y = 2 - y;
    return y * 3 + x;
  }
  AZ_HD static AzEnv symmetry(const AzEnv& e, int j) {
    AzEnv n = {e.a & PLAYER, 0, e.aux};
    for (int p = 0; p < 9; p++) {
      const int src = sym_source(j, p);
      n.a |= ((e.a >> src) & 1ull) << p;
      n.a |= ((e.a >> (16 + src)) & 1ull) << (16 + p);
    }
    return n;
  }
  // GI.heuristic_value (game.jl:96-120): 0.3 ^ (2 - N) per alignment free of the opponent's marks, summed in the order of
  // ALIGNMENTS (:43-51: x fixed, y fixed, diagonal, anti-diagonal), own minus opponent
  static constexpr bool HAS_HEURISTIC = true;
  AZ_HD static double heuristic_for(uint32_t mine, uint32_t theirs) {
    const uint32_t al[8] = {0x049, 0x092, 0x124, 0x007, 0x038, 0x1C0, 0x111, 0x054};
    const double pw[4] = {0.3 * 0.3, 0.3, 1.0, 1.0 / 0.3};
    double sum = 0.0;
    for (int i = 0; i < 8; i++) {
      const double v = (theirs & al[i]) ? 0.0 : pw[az_popc64((uint64_t)(mine & al[i]))];
      sum = i == 0 ? v : sum + v;
    }
    return sum;
  }
  AZ_HD static double heuristic_value(const AzEnv& e) {
    const uint32_t w = wbits(e), k = kbits(e);
    return white_playing(e) ? heuristic_for(w, k) - heuristic_for(k, w) : heuristic_for(k, w) - heuristic_for(w, k);
  }
  static AzEnv from_bytes(const uint8_t* s) {
    AzEnv e = {0, 0, 0};
    for (int i = 0; i < 9; i++) {
      if (s[i] == 1) e.a |= 1ull << i;
      if (s[i] == 2) e.a |= 1ull << (16 + i);
    }
    if (s[9] == 2) e.a |= PLAYER;
    return e;
  }
  AZ_HD static void to_bytes(const AzEnv& e, uint8_t* s) {
    for (int i = 0; i < 9; i++) s[i] = (uint8_t)(((wbits(e) >> i) & 1) ? 1 : (((kbits(e) >> i) & 1) ? 2 : 0));
    s[9] = white_playing(e) ? 1 : 2;
  }
  AZ_HD static void vectorize(const AzEnv& e, float* x) {
    uint32_t mine = white_playing(e) ? wbits(e) : kbits(e), theirs = white_playing(e) ? kbits(e) : wbits(e);
    for (int p = 0; p < 9; p++) {
      x[p] = (float)(((mine | theirs) >> p & 1) ^ 1);
      x[p + 9] = (float)((mine >> p) & 1);
      x[p + 18] = (float)((theirs >> p) & 1);
    }
  }
};

// ------------------------------------------------------------------------------------------
// Mancala (Kalah(6,3)): a = bytes {store_w, store_b, h[0..5]}, b = bytes {h[6..11]} | black << 56,
// h[(player-1) + 2*(num-1)] = houses[player, num].  aux bit0 = finished.
// ------------------------------------------------------------------------------------------
struct GameMancala {
  static constexpr int ID = 2;
  static constexpr int A = 6;
  static constexpr int LANES = 8;
  static constexpr int STATE_BYTES = 15;
  static constexpr int MAX_PLIES = 400;  // safety bound on game length (reference: unbounded)
  static constexpr int XW = 14, XH = 1, XC = 5;
  static constexpr bool ACYCLIC = false;  // be conservative: sequential backup
  static constexpr uint64_t PLAYER = 1ull << 56;

  AZ_HD static void unpack(const AzEnv& e, uint8_t* c) {
    for (int i = 0; i < 8; i++) c[i] = (uint8_t)(e.a >> (8 * i));
    for (int i = 0; i < 6; i++) c[8 + i] = (uint8_t)(e.b >> (8 * i));
  }
  AZ_HD static void pack(AzEnv& e, const uint8_t* c, bool black) {
    e.a = 0; e.b = 0;
    for (int i = 0; i < 8; i++) e.a |= (uint64_t)c[i] << (8 * i);
    for (int i = 0; i < 6; i++) e.b |= (uint64_t)c[8 + i] << (8 * i);
    if (black) e.b |= PLAYER;
  }
  AZ_HD static bool terminated(const AzEnv& e) { return e.aux & 1; }
  AZ_HD static bool white_playing(const AzEnv& e) { return (e.b & PLAYER) == 0; }
  AZ_HD static int house_idx(int player, int num) { return 2 + (player - 1) + 2 * (num - 1); }
  AZ_HD static uint32_t legal_mask(const AzEnv& e) {
    uint8_t c[14];
    unpack(e, c);
    int p = white_playing(e) ? 1 : 2;
    uint32_t m = 0;
    for (int n = 1; n <= 6; n++) m |= (uint32_t)(c[house_idx(p, n)] > 0) << (n - 1);
    return m;
  }
  AZ_HD static double white_reward(const AzEnv& e) {
    if (!(e.aux & 1)) return 0.0;
    int w = (int)(e.a & 0xFF), k = (int)((e.a >> 8) & 0xFF);
    return w > k ? 1.0 : (w < k ? -1.0 : 0.0);
  }
  AZ_HD static int sum_houses(const uint8_t* c, int p) {
    int s = 0;
    for (int n = 1; n <= 6; n++) s += c[house_idx(p, n)];
    return s;
  }
  AZ_HD static void leftovers(uint8_t* c, int p) {
    c[p - 1] = (uint8_t)(c[p - 1] + sum_houses(c, p));
    for (int i = 2; i < 14; i++) c[i] = 0;
  }
  AZ_HD static AzEnv play(const AzEnv& e, int action) {
    uint8_t c[14];
    unpack(e, c);
    int cp = white_playing(e) ? 1 : 2;
    // position: store => (is_store, player); house => (player, num)
    int is_store = 0, pp = cp, num = action + 1;
    int nseeds = c[house_idx(pp, num)];
    c[house_idx(pp, num)] = 0;
    for (int i = 0; i < nseeds; i++) {
      if (is_store) { is_store = 0; pp = 3 - cp; num = 6; }
      else if (num > 1) { num--; }
      else if (pp == cp) { is_store = 1; }
      else { pp = cp; num = 6; }
      int idx = is_store ? (cp - 1) : house_idx(pp, num);
      c[idx] = (uint8_t)(c[idx] + 1);
    }
    uint32_t aux = 0;
    int np = cp;
    if (sum_houses(c, cp) == 0) {
      leftovers(c, 3 - cp);
      aux = 1;
    } else if (!is_store) {
      bool done = false;
      if (c[house_idx(pp, num)] == 1 && cp == pp) {
        int opp = house_idx(3 - pp, 6 - num + 1);
        c[pp - 1] = (uint8_t)(c[pp - 1] + c[opp] + 1);
        c[house_idx(pp, num)] = 0;
        c[opp] = 0;
        if (sum_houses(c, 3 - cp) == 0) { leftovers(c, cp); aux = 1; done = true; }
        else if (sum_houses(c, cp) == 0) { leftovers(c, 3 - cp); aux = 1; done = true; }
      }
      if (!done) np = 3 - cp;
    }
    AzEnv n;
    pack(n, c, np == 2);
    n.aux = aux;
    return n;
  }
  AZ_HD static AzEnv init() {
    uint8_t c[14] = {0, 0, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3};
    AzEnv e;
    pack(e, c, false);
    e.aux = 0;
    return e;
  }
  static constexpr bool STOCHASTIC = false;
  static constexpr bool HAS_PLANE = false;  // plane(e, col, row, c) gives one element of vectorize_state directly
  static constexpr int NSYM = 0;  // no GI.symmetries declared for this game
  AZ_HD static AzEnv symmetry(const AzEnv& e, int) { return e; }
  AZ_HD static int sym_source(int, int p) { return p; }
  static constexpr long long MAX_STATES = 1ll << 40;  // no useful bound on distinct states
  AZ_HD static AzEnv play(const AzEnv& e, int a, const AzNoise&) { return play(e, a); }
  AZ_HD static AzEnv init_game(uint64_t, uint64_t) { return init(); }
  static AzEnv from_bytes(const uint8_t* s) {
    AzEnv e;
    pack(e, s, s[14] == 2);
    int cp = s[14] == 2 ? 2 : 1;
    e.aux = (sum_houses(s, cp) == 0 || sum_houses(s, 3 - cp) == 0) ? 1 : 0;
    return e;
  }
  AZ_HD static void to_bytes(const AzEnv& e, uint8_t* s) {
    unpack(e, s);
    s[14] = white_playing(e) ? 1 : 2;
  }
  // GI.heuristic_value (game.jl:212-218): the store difference from the mover's side -- computed by the reference in UInt8
  // arithmetic (stores are UInt8, :21), so a deficit wraps to 256 - d; reproduced as is
  static constexpr bool HAS_HEURISTIC = true;
  AZ_HD static double heuristic_value(const AzEnv& e) {
    uint8_t v = (uint8_t)((uint8_t)(e.a & 0xFF) - (uint8_t)((e.a >> 8) & 0xFF));
    if (!white_playing(e)) v = (uint8_t)(0u - v);
    return (double)v;
  }
  AZ_HD static void vectorize(const AzEnv& e, float* x) {  // incl. the flip_colors quirk (game.jl:224-229)
    uint8_t c[14];
    if (white_playing(e)) unpack(e, c);
    else { c[0] = c[1] = 0; for (int i = 2; i < 14; i++) c[i] = 3; }
    for (int i = 0; i < 14; i++) {
      bool is_store = (i == 6 || i == 13);
      int player = i < 7 ? 1 : 2, num = is_store ? 0 : 6 - (i % 7);
      x[i] = is_store ? (float)c[player - 1] : (float)c[house_idx(player, num)];
      x[i + 14] = (!is_store && player == 1) ? 1.f : 0.f;
      x[i + 28] = (is_store && player == 1) ? 1.f : 0.f;
      x[i + 42] = (!is_store && player == 2) ? 1.f : 0.f;
      x[i + 56] = (is_store && player == 2) ? 1.f : 0.f;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Grid world (games/grid-world/game.jl through src/common_rl_intf.jl): 10x10 single-player MDP, 4 actions always legal,
// 40 % chance of a random action, rewards on four terminal cells, episode bound `time > 200` that is NOT part of the
// state (game.jl:57) but is kept by the cloned environment (:56).  a = x | y << 8; aux = time (bits 0-8) | acted << 16.
// ------------------------------------------------------------------------------------------------
struct GameGW {
  static constexpr int ID = 3;
  static constexpr int A = 4;
  static constexpr int LANES = 8;
  static constexpr int STATE_BYTES = 2;
  static constexpr int MAX_PLIES = 201;  // time > 200 terminates (game.jl:40-41)
  static constexpr int XW = 10, XH = 10, XC = 1;
  static constexpr bool ACYCLIC = false;  // a simulation may revisit a state (time is not in the key)
  static constexpr bool STOCHASTIC = true;
  static constexpr bool HAS_PLANE = false;  // plane(e, col, row, c) gives one element of vectorize_state directly
  static constexpr int NSYM = 0;  // no GI.symmetries declared for this game
  AZ_HD static AzEnv symmetry(const AzEnv& e, int) { return e; }
  AZ_HD static int sym_source(int, int p) { return p; }
  static constexpr long long MAX_STATES = 100;  // 10 x 10 cells: a tree never holds more nodes
  AZ_HD static int gx(const AzEnv& e) { return (int)(e.a & 0xFF); }
  AZ_HD static int gy(const AzEnv& e) { return (int)((e.a >> 8) & 0xFF); }
  AZ_HD static int time(const AzEnv& e) { return (int)(e.aux & 0x1FF); }
  AZ_HD static double reward_at(int x, int y) {  // game.jl:24-28
    if (x == 9 && y == 3) return 10.0;
    if (x == 8 && y == 8) return 3.0;
    if (x == 4 && y == 3) return -10.0;
    if (x == 4 && y == 6) return -5.0;
    return 0.0;
  }
  AZ_HD static bool has_reward(int x, int y) { return (x == 9 && y == 3) || (x == 8 && y == 8) || (x == 4 && y == 3) || (x == 4 && y == 6); }
  AZ_HD static bool terminated(const AzEnv& e) { return has_reward(gx(e), gy(e)) || time(e) > 200; }
  AZ_HD static bool white_playing(const AzEnv&) { return true; }
  AZ_HD static uint32_t legal_mask(const AzEnv&) { return 0xFu; }
  AZ_HD static double white_reward(const AzEnv& e) { return ((e.aux >> 16) & 1u) ? reward_at(gx(e), gy(e)) : 0.0; }  // last_reward
  AZ_HD static AzEnv play(const AzEnv& e, int a, const AzNoise& nz) {  // act!, game.jl:43-51
    if (nz.u0 < 0.4) { a = (int)(nz.u1 * 4.0); if (a > 3) a = 3; }
    int x = gx(e) + (a == 0 ? 1 : (a == 1 ? -1 : 0));
    int y = gy(e) + (a == 2 ? 1 : (a == 3 ? -1 : 0));
    x = x < 1 ? 1 : (x > 10 ? 10 : x);
    y = y < 1 ? 1 : (y > 10 ? 10 : y);
    AzEnv n;
    n.a = (uint64_t)x | ((uint64_t)y << 8);
    n.b = 0;
    n.aux = (uint32_t)(time(e) + 1) | (1u << 16);
    return n;
  }
  static constexpr bool HAS_HEURISTIC = false;   // GI.heuristic_value is 0 (game.jl:118); MinMax is a two-player baseline
  AZ_HD static double heuristic_value(const AzEnv&) { return 0.0; }
  AZ_HD static AzEnv play(const AzEnv& e, int a) { AzNoise nz = {1.0, 0.0}; return play(e, a, nz); }  // noise-free step
  AZ_HD static AzEnv init() { AzEnv e = {1ull | (1ull << 8), 0, 0}; return e; }
  AZ_HD static AzEnv from_xy(int x, int y) { AzEnv e = {(uint64_t)x | ((uint64_t)y << 8), 0, 0}; return e; }
  AZ_HD static AzEnv init_game(uint64_t seed, uint64_t game) { int x, y; az_gw_init_xy(seed, game, &x, &y); return from_xy(x, y); }
  static AzEnv from_bytes(const uint8_t* s) { return from_xy(s[0], s[1]); }
  AZ_HD static void to_bytes(const AzEnv& e, uint8_t* s) { s[0] = (uint8_t)gx(e); s[1] = (uint8_t)gy(e); }
  AZ_HD static void vectorize(const AzEnv& e, float* x) {  // game.jl:86-90
    for (int i = 0; i < 100; i++) x[i] = 0.0f;
    x[(gx(e) - 1) + 10 * (gy(e) - 1)] = 1.0f;
  }
};

// =====================================================================================================
// MinMax baseline player (src/minmax.jl; Benchmark.MinMaxTS, src/benchmark.jl:178-196), host + device.  qvalue (:29-43) of
// one root action by an explicit stack of value() frames (:17-27: 0 on a terminal state, the heuristic at depth 0, else the
// maximum of the q-values of the available actions).
// =====================================================================================================
constexpr int AZ_MINMAX_MAX_DEPTH = 8;
template <class G>
AZ_HD double az_minmax_qvalue(const AzEnv& root, int action, int depth, bool amplify, double gamma) {
  AzEnv st[AZ_MINMAX_MAX_DEPTH];       // st[l]: the state value() is evaluated on at level l (l plies below root + 1)
  uint32_t todo[AZ_MINMAX_MAX_DEPTH];  // actions of st[l] still to try
  double best[AZ_MINMAX_MAX_DEPTH], er[AZ_MINMAX_MAX_DEPTH];
  bool has[AZ_MINMAX_MAX_DEPTH], flip[AZ_MINMAX_MAX_DEPTH];
  const double inf = az_bits_to_double(0x7FF0000000000000ull);
  auto edge = [&](const AzEnv& from, int a, int l) {   // the part of qvalue before the recursive call
    st[l] = G::play(from, a);
    const double wr = G::white_reward(st[l]);
    double r = G::white_playing(from) ? wr : -wr;
    if (amplify && r != 0.0) r = r > 0.0 ? inf : -inf;   // amplify (:14)
    er[l] = r;
    flip[l] = G::white_playing(from) != G::white_playing(st[l]);
    todo[l] = G::legal_mask(st[l]);
    has[l] = false;
    best[l] = 0.0;
  };
  edge(root, action, 0);
  int l = 0;
  for (;;) {
    const int left = depth - 1 - l;   // the depth argument of value(st[l], .)
    double v;
    bool ret = true;
    if (G::terminated(st[l])) v = 0.0;
    else if (left == 0) v = G::heuristic_value(st[l]);
    else if (todo[l] == 0) v = best[l];
    else {
      int a = 0;
      while (!((todo[l] >> a) & 1u)) a++;   // available actions in ascending order
      todo[l] &= todo[l] - 1;
      edge(st[l], a, l + 1);
      l++;
      ret = false;
    }
    if (!ret) continue;
    const double q = er[l] + gamma * (flip[l] ? -v : v);
    if (l == 0) return q;
    l--;
    best[l] = has[l] ? (q > best[l] ? q : best[l]) : q;
    has[l] = true;
  }
}

// think(::MinMax.Player) after the q-values (src/minmax.jl:83-114): pi over the n available actions
AZ_HD void az_minmax_policy(const double* qs, int n, double tau, double* pi) {
  const double inf = az_bits_to_double(0x7FF0000000000000ull);
  int nwin = 0, nnot = 0, best = 0;
  for (int i = 0; i < n; i++) { nwin += qs[i] == inf; nnot += qs[i] > -inf; if (qs[i] > qs[best]) best = i; }
  if (nwin > 0) { for (int i = 0; i < n; i++) pi[i] = qs[i] == inf ? 1.0 : 0.0; }
  else if (nnot == 0) { for (int i = 0; i < n; i++) pi[i] = 1.0; }
  else if (tau == 0.0) { for (int i = 0; i < n; i++) pi[i] = qs[i] == qs[best] ? 1.0 : 0.0; }
  else {
    const double qmax = qs[best];
    double C = 0.0;
    bool hc = false;
    for (int i = 0; i < n; i++) if (qs[i] > -inf) { const double a = qs[i] < 0.0 ? -qs[i] : qs[i]; C = hc ? (a > C ? a : C) : a; hc = true; }
    C = C + 2.220446049250313e-16;   // eps()
    const double it = 1.0 / tau;
    for (int i = 0; i < n; i++) {
      const double x = (qs[i] - qmax) / C;
      const double e = (x == -inf) ? 0.0 : az_det_exp(x);
      pi[i] = (e > 0.0) ? az_det_exp(it * az_det_log(e)) : 0.0;   // pi .^= 1 / tau
    }
  }
  double s = 0.0;
  for (int i = 0; i < n; i++) s = (i == 0) ? pi[0] : s + pi[i];
  for (int i = 0; i < n; i++) pi[i] = pi[i] / s;
}
