// az_rng.cuh -- explicit counter-based random streams + deterministic log/exp (host + device).
//
// Julia's Xoshiro and Distributions.jl cannot be reproduced outside Julia, so every stochastic input of the
// path (Dirichlet root noise src/mcts.jl:228-232, categorical move sampling src/util.jl:87-90) is drawn from a
// Philox4x32-10 stream keyed by (seed, game index, move index, purpose, draw index).  All arithmetic is IEEE
// double with one rounding per operation (this translation unit is compiled with -fmad=false), and log/exp are
// fixed polynomial algorithms, so host and device produce identical bits.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef AZ_HD
#define AZ_HD __host__ __device__ __forceinline__
#endif

enum { AZ_PURPOSE_DIRICHLET = 0, AZ_PURPOSE_CATEGORICAL = 1, AZ_PURPOSE_SYMMETRY = 2, AZ_PURPOSE_ENV = 3, AZ_PURPOSE_POSITION = 4, AZ_PURPOSE_ROLLOUT = 5 };

AZ_HD void az_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
AZ_HD uint64_t az_stream_u64(uint64_t seed, uint64_t game, uint32_t move, int purpose, uint32_t k) {
  uint32_t o[4];
  az_philox(seed, k >> 1, (uint32_t)purpose | (move << 8), (uint32_t)game, (uint32_t)(game >> 32), o);
  return (k & 1) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
}
AZ_HD double az_u01(uint64_t x) { return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0); }
AZ_HD float az_uniform_f32(uint64_t seed, uint64_t game, uint32_t move, int purpose, uint32_t idx) {
  uint64_t x = az_stream_u64(seed, game, move, purpose, idx);
  return (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
}
AZ_HD double az_bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d; memcpy(&d, &b, 8); return d;
#endif
}
AZ_HD uint64_t az_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
#define AZ_LN2 0.6931471805599453094
AZ_HD double az_det_log(double x) {
  uint64_t bits = az_double_to_bits(x);
  int e = (int)((bits >> 52) & 0x7FF) - 1022;
  double m = az_bits_to_double((bits & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull);
  if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
  double s = (m - 1.0) / (m + 1.0);
  double s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; k--) p = p * s2 + 1.0 / (double)(2 * k + 1);
  return (double)e * AZ_LN2 + (2.0 * s) * p;
}
AZ_HD double az_det_exp(double x) {
  double kf = floor(x / AZ_LN2 + 0.5);
  if (kf < -1000.0) kf = -1000.0;
  if (kf > 1000.0) kf = 1000.0;
  double r = x - kf * AZ_LN2;
  double p = 1.0;
  for (int n = 16; n >= 1; n--) p = p * (r / (double)n) + 1.0;
  int k = (int)kf;
  return p * az_bits_to_double((uint64_t)(k + 1023) << 52);
}
struct AzStream { uint64_t seed, game; uint32_t move; int purpose; uint32_t k; };
AZ_HD double az_next_u01(AzStream& st) { return az_u01(az_stream_u64(st.seed, st.game, st.move, st.purpose, st.k++)); }
AZ_HD double az_normal(AzStream& st) {  // Marsaglia polar method
  for (;;) {
    double u1 = 2.0 * az_next_u01(st) - 1.0, u2 = 2.0 * az_next_u01(st) - 1.0;
    double s = u1 * u1 + u2 * u2;
    if (s >= 1.0 || s == 0.0) continue;
    return u1 * sqrt((-2.0 * az_det_log(s)) / s);
  }
}
AZ_HD double az_gamma(AzStream& st, double alpha) {  // Marsaglia-Tsang
  double boost = 1.0;
  if (alpha < 1.0) {
    double u = az_next_u01(st);
    boost = az_det_exp(az_det_log(u) / alpha);
    alpha = alpha + 1.0;
  }
  double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = az_normal(st);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = az_next_u01(st);
    double x2 = x * x;
    if (u < 1.0 - (0.0331 * x2) * x2) return (d * v) * boost;
    if (az_det_log(u) < 0.5 * x2 + d * ((1.0 - v) + az_det_log(v))) return (d * v) * boost;
  }
}
// eta[0..n) ~ Dirichlet(n, alpha): replaces rand(Dirichlet(n, alpha)) (src/mcts.jl:231)
AZ_HD void az_dirichlet(uint64_t seed, uint64_t game, uint32_t move, int n, double alpha, double* eta) {
  AzStream st = {seed, game, move, AZ_PURPOSE_DIRICHLET, 0};
  double sum = 0.0;
  for (int i = 0; i < n; i++) { eta[i] = az_gamma(st, alpha); sum = sum + eta[i]; }
  for (int i = 0; i < n; i++) eta[i] = eta[i] / sum;
}
AZ_HD uint64_t az_splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---- environment noise and random initial states (grid-world) -------------------------------------------------------
// In-tree noise of simulation `sim` at depth `depth` of move `move` of game `game`; the real move uses sim = AZ_REAL_MOVE.
#define AZ_REAL_MOVE 0x7FFFFFu
struct AzNoiseKey { uint64_t seed, game; uint32_t move; };
template <class Noise>
AZ_HD Noise az_env_noise(const AzNoiseKey& k, uint32_t sim, uint32_t depth) {
  const uint32_t idx = (sim * 256u + depth) * 2u;
  Noise n;
  n.u0 = az_u01(az_stream_u64(k.seed, k.game, k.move, AZ_PURPOSE_ENV, idx));
  n.u1 = az_u01(az_stream_u64(k.seed, k.game, k.move, AZ_PURPOSE_ENV, idx + 1u));
  return n;
}
// random initial cell in 1..10 x 1..10 (RL.reset!, games/grid-world/game.jl:36)
AZ_HD void az_gw_init_xy(uint64_t seed, uint64_t game, int* x, int* y) {
  uint32_t o[4];
  az_philox(seed, 0, AZ_PURPOSE_POSITION, (uint32_t)game, (uint32_t)(game >> 32), o);
  *x = 1 + (int)(o[0] % 10u);
  *y = 1 + (int)(o[1] % 10u);
}
