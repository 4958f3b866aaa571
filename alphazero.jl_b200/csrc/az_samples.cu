// az_samples.cu -- the replay-buffer side of the wire on the GPU: device-resident training samples,
// merge_by_state and augment_with_symmetries (src/memory.jl:89-130), convert_samples (src/learning.jl:17-51).
//
// A sample set is SoA in HBM: state (AzEnv, 24 B), pi (A doubles, zero on illegal actions), z, t (doubles), n (int32):
// TrainingSample (src/memory.jl:20-26).  Everything here is byte/integer/f64 streaming work bounded by HBM bandwidth
// (a Connect-Four sample is 100 B; 1.15 M samples of BASELINE config 3 are 115 MB): one thread or one warp per
// sample, coalesced SoA accesses, no tensor cores.  Sorting and prefix sums use CUB (plain library primitives).
#include <cub/cub.cuh>

#include <climits>
#include <vector>

#include "az_internal.h"

struct az_samples {
  az_ctx* ctx = nullptr;
  int game = 0, A = 0;
  int64_t n = 0;
  AzEnv* env = nullptr;
  double *pi = nullptr, *z = nullptr, *t = nullptr;
  int32_t* cnt = nullptr;
  ~az_samples() { cudaFree(env); cudaFree(pi); cudaFree(z); cudaFree(t); cudaFree(cnt); }
};

static int game_actions(int game) {
  switch (game) { case 0: return GameC4::A; case 1: return GameTTT::A; case 2: return GameMancala::A; case 3: return GameGW::A; }
  return -1;
}
#define AZS_DISPATCH(game, F, ...)                   \
  switch (game) {                                    \
    case 0: return F<GameC4>(__VA_ARGS__);           \
    case 1: return F<GameTTT>(__VA_ARGS__);          \
    case 2: return F<GameMancala>(__VA_ARGS__);      \
    case 3: return F<GameGW>(__VA_ARGS__);           \
    default: return AZ_EINVAL;                       \
  }

int az_samples_alloc(az_ctx* ctx, int game, int64_t n, az_samples** out) {
  const int A = game_actions(game);
  if (A < 0 || n < 0) { ctx->err = "az_samples: bad game or count"; return AZ_EINVAL; }
  az_samples* s = new az_samples();
  s->ctx = ctx; s->game = game; s->A = A; s->n = n;
  const size_t m = (size_t)std::max<int64_t>(n, 1);
  if (cudaMalloc((void**)&s->env, m * sizeof(AzEnv)) != cudaSuccess || cudaMalloc((void**)&s->pi, m * A * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&s->z, m * sizeof(double)) != cudaSuccess || cudaMalloc((void**)&s->t, m * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&s->cnt, m * sizeof(int32_t)) != cudaSuccess) {
    cudaGetLastError();
    delete s;
    ctx->err = "az_samples: cudaMalloc failed";
    return AZ_ENOMEM;
  }
  *out = s;
  return AZ_OK;
}
// accessors for az_engine.cu (the self-play engine exports its finished run without a host round trip)
AzEnv* az_samples_env(az_samples* s) { return s->env; }
double* az_samples_pi(az_samples* s) { return s->pi; }
double* az_samples_z(az_samples* s) { return s->z; }
double* az_samples_t(az_samples* s) { return s->t; }
int32_t* az_samples_cnt(az_samples* s) { return s->cnt; }
int az_samples_game(az_samples* s) { return s->game; }

// ---- grouping by state: open-addressing table over a 64-bit hash of the 16-byte state key --------------------------------
__device__ __forceinline__ uint64_t azs_hash(uint64_t a, uint64_t b) {
  uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15F39CC060ull) * 0xC2B2AE3D27D4EB4Full;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
  return x | 1ull;  // 0 = empty slot
}
__global__ void azs_k_insert(int64_t n, const AzEnv* __restrict__ env, unsigned long long* tab_hash, int* tab_first, uint32_t mask,
                             uint32_t* slot_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = azs_hash(env[i].a, env[i].b);
  uint32_t pos = (uint32_t)(h >> 20) & mask;
  for (;;) {
    const unsigned long long old = atomicCAS(&tab_hash[pos], 0ull, (unsigned long long)h);
    if (old == 0ull || old == h) break;
    pos = (pos + 1) & mask;
  }
  atomicMin(&tab_first[pos], (int)i);  // samples[1] of the bucket = its first occurrence (src/memory.jl:90,105-109)
  slot_of[i] = pos;
}
// flag[i] = 1 iff sample i opens its group; also verifies the 64-bit hash against the full key
__global__ void azs_k_mark(int64_t n, const AzEnv* __restrict__ env, const int* __restrict__ tab_first, const uint32_t* __restrict__ slot_of,
                           int* flag, int* first_of, int* err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = tab_first[slot_of[i]];
  if (env[f].a != env[i].a || env[f].b != env[i].b) *err = 1;  // 64-bit hash collision between different states
  first_of[i] = f;
  flag[i] = (f == (int)i) ? 1 : 0;
}
__global__ void azs_k_group(int64_t n, const int* __restrict__ first_of, const int* __restrict__ gscan, int* gid, int* gcount) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gscan[first_of[i]];  // groups are numbered in the order of their first occurrence
  gid[i] = g;
  atomicAdd(&gcount[g], 1);
}
__global__ void azs_k_iota(int64_t n, int* idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int)i;
}
// one warp per group: members in original order (stable sort by group id), lanes = the A policy entries, z, t, n.
// mean(x for x in samples) = left-to-right sum / count (src/memory.jl:89-96), one rounding per operation.
__global__ void azs_k_merge(int G, int A, const int* __restrict__ goff, const int* __restrict__ members, const AzEnv* __restrict__ env,
                            const double* __restrict__ pi, const double* __restrict__ z, const double* __restrict__ t,
                            const int32_t* __restrict__ cnt, AzEnv* oenv, double* opi, double* oz, double* ot, int32_t* ocnt) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= G) return;
  const int b = goff[g], e = goff[g + 1];
  double acc = 0.0;
  int nsum = 0;
  for (int k = b; k < e; k++) {
    const int i = members[k];
    double v = 0.0;
    if (lane < A) v = pi[(size_t)i * A + lane];
    else if (lane == A) v = z[i];
    else if (lane == A + 1) v = t[i];
    else if (lane == A + 2) nsum += cnt[i];
    acc = (k == b) ? v : acc + v;
  }
  const double m = acc / (double)(e - b);
  if (lane < A) opi[(size_t)g * A + lane] = m;
  else if (lane == A) oz[g] = m;
  else if (lane == A + 1) ot[g] = m;
  else if (lane == A + 2) ocnt[g] = nsum;
  else if (lane == A + 3) oenv[g] = env[members[b]];
}

template <class T> struct DBuf {  // scoped device scratch
  T* p = nullptr;
  int alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)) == cudaSuccess ? 0 : 1; }
  ~DBuf() { cudaFree(p); }
};

static int samples_merge(az_samples* in, az_samples** out) {
  az_ctx* ctx = in->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t n = in->n;
  if (n >= INT_MAX / 2) { ctx->err = "az_samples_merge: too many samples"; return AZ_EINVAL; }
  if (n == 0) return az_samples_alloc(ctx, in->game, 0, out);
  uint32_t cap = 1024;
  while (cap < 2 * (uint64_t)n) cap <<= 1;
  DBuf<unsigned long long> tab_hash; DBuf<int> tab_first, flag, first_of, gscan, gid, gid_s, idx, idx_s, gcount, goff, err;
  DBuf<uint32_t> slot_of; DBuf<uint8_t> tmp;
  if (tab_hash.alloc(cap) | tab_first.alloc(cap) | flag.alloc(n) | first_of.alloc(n) | gscan.alloc(n) | gid.alloc(n) | gid_s.alloc(n) |
      idx.alloc(n) | idx_s.alloc(n) | gcount.alloc(n + 1) | goff.alloc(n + 2) | err.alloc(1) | slot_of.alloc(n)) {
    cudaGetLastError(); ctx->err = "az_samples_merge: cudaMalloc failed"; return AZ_ENOMEM;
  }
  AZ_CUDA(ctx, cudaMemsetAsync(tab_hash.p, 0, (size_t)cap * 8, st));
  AZ_CUDA(ctx, cudaMemsetAsync(tab_first.p, 0x7f, (size_t)cap * 4, st));
  AZ_CUDA(ctx, cudaMemsetAsync(err.p, 0, 4, st));
  AZ_CUDA(ctx, cudaMemsetAsync(gcount.p, 0, (size_t)(n + 1) * 4, st));
  const int T = 256, B = (int)((n + T - 1) / T);
  azs_k_insert<<<B, T, 0, st>>>(n, in->env, tab_hash.p, tab_first.p, cap - 1, slot_of.p);
  azs_k_mark<<<B, T, 0, st>>>(n, in->env, tab_first.p, slot_of.p, flag.p, first_of.p, err.p);
  size_t tb = 0, tb2 = 0, tb3 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, gscan.p, (int)n, st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, gid.p, gid_s.p, idx.p, idx_s.p, (int)n, 0, 32, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, gcount.p, goff.p, (int)n + 1, st);
  tb = std::max(tb, std::max(tb2, tb3));
  if (tmp.alloc(tb)) { cudaGetLastError(); ctx->err = "az_samples_merge: cudaMalloc failed"; return AZ_ENOMEM; }
  AZ_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, gscan.p, (int)n, st));
  int last_scan = 0, last_flag = 0, herr = 0;
  AZ_CUDA(ctx, cudaMemcpyAsync(&last_scan, gscan.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  AZ_CUDA(ctx, cudaMemcpyAsync(&last_flag, flag.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  AZ_CUDA(ctx, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, st));
  AZ_CUDA(ctx, cudaStreamSynchronize(st));
  if (herr) { ctx->err = "az_samples_merge: 64-bit state hash collision"; return AZ_ESTATE; }
  const int G = last_scan + last_flag;
  azs_k_group<<<B, T, 0, st>>>(n, first_of.p, gscan.p, gid.p, gcount.p);
  azs_k_iota<<<B, T, 0, st>>>(n, idx.p);
  int bits = 1;
  while ((1ll << bits) < G) bits++;
  AZ_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, gid.p, gid_s.p, idx.p, idx_s.p, (int)n, 0, bits, st));  // stable
  AZ_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, gcount.p, goff.p, G + 1, st));
  az_samples* o = nullptr;
  int rc = az_samples_alloc(ctx, in->game, G, &o);
  if (rc != AZ_OK) return rc;
  azs_k_merge<<<(int)(((size_t)G * 32 + T - 1) / T), T, 0, st>>>(G, in->A, goff.p, idx_s.p, in->env, in->pi, in->z, in->t, in->cnt, o->env, o->pi,
                                                               o->z, o->t, o->cnt);
  ctx->launches += 4;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete o; ctx->err = std::string("az_samples_merge: ") + cudaGetErrorString(e); return AZ_ECUDA; }
  *out = o;
  return AZ_OK;
}

// ---- augment_with_symmetries: [samples ; apply_symmetry(s, sym) for s in samples for sym in symmetries(s)] ---------------
template <class G>
__global__ void azs_k_augment(int64_t n, const AzEnv* __restrict__ env, const double* __restrict__ pi, const double* __restrict__ z,
                              const double* __restrict__ t, const int32_t* __restrict__ cnt, AzEnv* oenv, double* opi, double* oz, double* ot,
                              int32_t* ocnt, int* err) {
  constexpr int A = G::A, NS = G::NSYM;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * (1 + NS)) return;
  const int64_t i = k < n ? k : (k - n) / (NS > 0 ? NS : 1);
  const int j = k < n ? -1 : (int)((k - n) % (NS > 0 ? NS : 1));
  AzEnv e = env[i];
  double p[A];
  for (int a = 0; a < A; a++) p[a] = pi[(size_t)i * A + a];
  if (j >= 0) {  // apply_symmetry (src/memory.jl:112-124)
    e = G::symmetry(e, j);
    double q[A];
    for (int a = 0; a < A; a++) q[a] = p[G::sym_source(j, a)];
    const uint32_t legal = G::legal_mask(e);
    for (int a = 0; a < A; a++) {
      if (!((legal >> a) & 1u) && q[a] != 0.0) *err = 1;  // @assert iszero(pi[.~symmask])
      p[a] = q[a];
    }
  }
  oenv[k] = e;
  for (int a = 0; a < A; a++) opi[(size_t)k * A + a] = p[a];
  oz[k] = z[i]; ot[k] = t[i]; ocnt[k] = cnt[i];
}
template <class G>
static int samples_augment(az_samples* in, az_samples** out) {
  az_ctx* ctx = in->ctx;
  const int64_t m = in->n * (1 + G::NSYM);
  az_samples* o = nullptr;
  int rc = az_samples_alloc(ctx, in->game, m, &o);
  if (rc != AZ_OK) return rc;
  DBuf<int> err;
  if (err.alloc(1)) { delete o; ctx->err = "az_samples_augment: cudaMalloc failed"; return AZ_ENOMEM; }
  cudaMemsetAsync(err.p, 0, 4, ctx->stream);
  if (m > 0) azs_k_augment<G><<<(int)((m + 255) / 256), 256, 0, ctx->stream>>>(in->n, in->env, in->pi, in->z, in->t, in->cnt, o->env, o->pi, o->z,
                                                                                 o->t, o->cnt, err.p);
  ctx->launches += 1;
  int herr = 0;
  cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { delete o; ctx->err = std::string("az_samples_augment: ") + cudaGetErrorString(e); return AZ_ECUDA; }
  if (herr) { delete o; ctx->err = "apply_symmetry: policy mass on an action that is illegal in the image state (src/memory.jl:120)"; return AZ_ESTATE; }
  *out = o;
  return AZ_OK;
}

// ---- convert_samples (src/learning.jl:17-51): (W, X, A, P, V) Float32, sample index = last (slowest) dimension -----------
template <class G>
__global__ void azs_k_convert(int64_t n, int weighing, const AzEnv* __restrict__ env, const double* __restrict__ pi,
                              const double* __restrict__ z, const int32_t* __restrict__ cnt, float* W, float* X, float* Am, float* P, float* V) {
  constexpr int A = G::A, NX = G::XW * G::XH * G::XC;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const AzEnv e = env[i];
  if (W) {
    const int c = cnt[i];
    W[i] = weighing == 0 ? 1.0f : (weighing == 1 ? (float)(log2((double)c) + 1.0) : (float)c);  // CONSTANT / LOG / LINEAR_WEIGHT
  }
  if (X) {
    float x[NX];
    G::vectorize(e, x);
    for (int k = 0; k < NX; k++) X[(size_t)i * NX + k] = x[k];
  }
  const uint32_t legal = G::legal_mask(e);
  for (int a = 0; a < A; a++) {
    const bool l = (legal >> a) & 1u;
    if (Am) Am[(size_t)i * A + a] = l ? 1.0f : 0.0f;
    if (P) P[(size_t)i * A + a] = l ? (float)pi[(size_t)i * A + a] : 0.0f;
  }
  if (V) V[i] = (float)z[i];
}
template <class G>
static int samples_convert(az_samples* s, int weighing, float* W, float* X, float* Am, float* P, float* V) {
  az_ctx* ctx = s->ctx;
  constexpr int A = G::A, NX = G::XW * G::XH * G::XC;
  const int64_t n = s->n;
  if (n == 0) return AZ_OK;
  DBuf<float> dW, dX, dA, dP, dV;
  if (dW.alloc(n) | dX.alloc((size_t)n * NX) | dA.alloc((size_t)n * A) | dP.alloc((size_t)n * A) | dV.alloc(n)) {
    cudaGetLastError(); ctx->err = "az_samples_convert: cudaMalloc failed"; return AZ_ENOMEM;
  }
  azs_k_convert<G><<<(int)((n + 127) / 128), 128, 0, ctx->stream>>>(n, weighing, s->env, s->pi, s->z, s->cnt, dW.p, dX.p, dA.p, dP.p, dV.p);
  ctx->launches += 1;
  // outputs are the caller's pageable arrays: pinned double-buffered staging (az_d2h), not one driver-staged copy per array
  int st = AZ_OK;
  if (W && st == AZ_OK) st = az_d2h(ctx, W, dW.p, (size_t)n * 4);
  if (X && st == AZ_OK) st = az_d2h(ctx, X, dX.p, (size_t)n * NX * 4);
  if (Am && st == AZ_OK) st = az_d2h(ctx, Am, dA.p, (size_t)n * A * 4);
  if (P && st == AZ_OK) st = az_d2h(ctx, P, dP.p, (size_t)n * A * 4);
  if (V && st == AZ_OK) st = az_d2h(ctx, V, dV.p, (size_t)n * 4);
  if (st != AZ_OK) return st;
  AZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AZ_OK;
}

template <class G>
static int samples_from_host(az_ctx* ctx, int64_t n, const uint8_t* states, const double* pi, const double* z, const double* t, const int32_t* cnt,
                             az_samples** out) {
  az_samples* s = nullptr;
  int rc = az_samples_alloc(ctx, G::ID, n, &s);
  if (rc != AZ_OK) return rc;
  std::vector<AzEnv> env((size_t)n);
  std::vector<int32_t> ones;
  for (int64_t i = 0; i < n; i++) env[(size_t)i] = G::from_bytes(states + (size_t)i * G::STATE_BYTES);
  if (!cnt) { ones.assign((size_t)n, 1); cnt = ones.data(); }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaSuccess;
  if (n > 0) {
    e = cudaMemcpyAsync(s->env, env.data(), (size_t)n * sizeof(AzEnv), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->pi, pi, (size_t)n * G::A * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->z, z, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->t, t, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->cnt, cnt, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  if (e != cudaSuccess) { delete s; ctx->err = std::string("az_samples_from_host: ") + cudaGetErrorString(e); return AZ_ECUDA; }
  *out = s;
  return AZ_OK;
}
template <class G>
static int samples_fetch(az_samples* s, uint8_t* states, double* pi, double* z, double* t, int32_t* cnt) {
  az_ctx* ctx = s->ctx;
  const int64_t n = s->n;
  if (n == 0) return AZ_OK;
  if (states) {
    std::vector<AzEnv> env((size_t)n);
    AZ_CUDA(ctx, cudaMemcpy(env.data(), s->env, (size_t)n * sizeof(AzEnv), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; i++) G::to_bytes(env[(size_t)i], states + (size_t)i * G::STATE_BYTES);
  }
  if (pi) AZ_CUDA(ctx, cudaMemcpy(pi, s->pi, (size_t)n * G::A * 8, cudaMemcpyDeviceToHost));
  if (z) AZ_CUDA(ctx, cudaMemcpy(z, s->z, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (t) AZ_CUDA(ctx, cudaMemcpy(t, s->t, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (cnt) AZ_CUDA(ctx, cudaMemcpy(cnt, s->cnt, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return AZ_OK;
}

#define AZS_GUARD_BEGIN try {
#define AZS_GUARD_END(ctx)                                        \
  } catch (const std::exception& ex) {                            \
    if (ctx) (ctx)->err = std::string("exception: ") + ex.what(); \
    return AZ_ESTATE;                                             \
  } catch (...) {                                                 \
    if (ctx) (ctx)->err = "unknown exception";                    \
    return AZ_ESTATE;                                             \
  }

extern "C" {

int32_t az_samples_from_host(az_ctx* ctx, int32_t game, int64_t n, const uint8_t* states, const double* pi, const double* z, const double* t,
                             const int32_t* cnt, az_samples** out) {
  if (!ctx || !out || n < 0 || (n > 0 && (!states || !pi || !z || !t))) return AZ_EINVAL;
  AZS_GUARD_BEGIN
  cudaSetDevice(ctx->device);
  AZS_DISPATCH(game, samples_from_host, ctx, n, states, pi, z, t, cnt, out)
  AZS_GUARD_END(ctx)
}
int32_t az_samples_count(az_samples* s, int64_t* n) {
  if (!s || !n) return AZ_EINVAL;
  *n = s->n;
  return AZ_OK;
}
int32_t az_samples_concat(az_samples* a, az_samples* b, az_samples** out) {
  if (!a || !b || !out) return AZ_EINVAL;
  az_ctx* ctx = a->ctx;
  AZS_GUARD_BEGIN
  if (a->game != b->game || a->ctx != b->ctx) { ctx->err = "az_samples_concat: sample sets of different games / contexts"; return AZ_EINVAL; }
  cudaSetDevice(ctx->device);
  az_samples* o = nullptr;
  int rc = az_samples_alloc(ctx, a->game, a->n + b->n, &o);
  if (rc != AZ_OK) return rc;
  const int A = a->A;
  cudaStream_t st = ctx->stream;
  az_samples* src[2] = {a, b};
  int64_t off = 0;
  for (az_samples* s : src) {
    if (s->n > 0) {
      cudaMemcpyAsync(o->env + off, s->env, (size_t)s->n * sizeof(AzEnv), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(o->pi + off * A, s->pi, (size_t)s->n * A * 8, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(o->z + off, s->z, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(o->t + off, s->t, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(o->cnt + off, s->cnt, (size_t)s->n * 4, cudaMemcpyDeviceToDevice, st);
    }
    off += s->n;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete o; ctx->err = std::string("az_samples_concat: ") + cudaGetErrorString(e); return AZ_ECUDA; }
  *out = o;
  return AZ_OK;
  AZS_GUARD_END(ctx)
}
int32_t az_samples_merge_by_state(az_samples* in, az_samples** out) {
  if (!in || !out) return AZ_EINVAL;
  AZS_GUARD_BEGIN
  cudaSetDevice(in->ctx->device);
  return samples_merge(in, out);
  AZS_GUARD_END(in->ctx)
}
int32_t az_samples_augment_with_symmetries(az_samples* in, az_samples** out) {
  if (!in || !out) return AZ_EINVAL;
  AZS_GUARD_BEGIN
  cudaSetDevice(in->ctx->device);
  AZS_DISPATCH(in->game, samples_augment, in, out)
  AZS_GUARD_END(in->ctx)
}
int32_t az_samples_convert(az_samples* s, int32_t weighing, float* W, float* X, float* A, float* P, float* V) {
  if (!s) return AZ_EINVAL;
  if (weighing < 0 || weighing > 2) { s->ctx->err = "az_samples_convert: weighing policy must be 0 (constant), 1 (log) or 2 (linear)"; return AZ_EINVAL; }
  AZS_GUARD_BEGIN
  cudaSetDevice(s->ctx->device);
  AZS_DISPATCH(s->game, samples_convert, s, weighing, W, X, A, P, V)
  AZS_GUARD_END(s->ctx)
}
int32_t az_samples_fetch(az_samples* s, uint8_t* states, double* pi, double* z, double* t, int32_t* n) {
  if (!s) return AZ_EINVAL;
  AZS_GUARD_BEGIN
  cudaSetDevice(s->ctx->device);
  AZS_DISPATCH(s->game, samples_fetch, s, states, pi, z, t, n)
  AZS_GUARD_END(s->ctx)
}
int32_t az_samples_destroy(az_samples* s) {
  if (!s) return AZ_EINVAL;
  delete s;
  return AZ_OK;
}

}  // extern "C"
