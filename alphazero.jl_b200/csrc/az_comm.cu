// az_comm.cu -- the multi-GPU side of the self-play path inside the C ABI: one rank per GPU, no collective in the
// simulation loop, and exactly the two exchanges the reference performs per iteration (src/simulations.jl:252-290):
//   * az_net_broadcast     : rank `root`'s network parameters to every rank (replaces the serialisation of the closure
//                            that captures the network when `simulate_distributed` spawns its workers, :268-281, and
//                            Network.copy(bestnn; on_gpu=true), src/training.jl:278-279)
//   * az_samples_allgather : every rank's device-resident TrainingSamples concatenated in rank order on every rank
//                            (replaces `fetch.(tasks)` + `reduce(vcat, results)`, :282-289)
// Rows never leave HBM: the local SoA set is packed into fixed-size rows by one streaming kernel, ONE ncclAllGather of
// the per-rank counts and ONE ncclAllGather of the rows padded to the largest count move them over NVLink, and one
// streaming kernel compacts the padding away into the output SoA set.  Packing and compaction are byte-copy kernels
// bounded by HBM bandwidth (8-byte words, consecutive threads on consecutive words).
//
// NCCL is bound at run time (dlopen of libnccl.so.2; a copy already loaded into the process, e.g. PyTorch's, is reused)
// so that single-GPU users of libazb200.so do not need NCCL installed at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "az_internal.h"

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};
NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.h) break; }  // reuse a loaded copy
  for (const char* n : names) { if (api.h) break; api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL); }
  if (!api.h) { api.why = std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return api; }
#define AZ_NCCL_SYM(field, name)                                                           \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name));                   \
  if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; api.h = nullptr; return api; }
  AZ_NCCL_SYM(GetVersion, "ncclGetVersion")
  AZ_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  AZ_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  AZ_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  AZ_NCCL_SYM(AllGather, "ncclAllGather")
  AZ_NCCL_SYM(Broadcast, "ncclBroadcast")
  AZ_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef AZ_NCCL_SYM
  return api;
}
}  // namespace

struct az_comm {
  az_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  int64_t* d_counts = nullptr;  // [world + 1]: [0, world) gathered counts, [world] this rank's count
  uint64_t *d_send = nullptr, *d_recv = nullptr;   // staging rows, grown on demand and kept across iterations
  size_t send_words = 0, recv_words = 0;
  double last_ms = 0;           // device time of the last collective call (CUDA events on the context's stream)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define AZ_NCCL(ctx, call)                                                                                     \
  do {                                                                                                         \
    ncclResult_t r__ = (call);                                                                                 \
    if (r__ != ncclSuccess) {                                                                                  \
      (ctx)->err = std::string(#call) + ": " + nccl().GetErrorString(r__) + " @" + __FILE__ + ":" + std::to_string(__LINE__); \
      return AZ_ECUDA;                                                                                         \
    }                                                                                                          \
  } while (0)

// ---- row packing: sample i -> ROWW 8-byte words {env.a, env.b, env.aux, pi[0..A), z, t, n} ------------------------------
__global__ void __launch_bounds__(256) azc_k_pack(int64_t n, int A, const AzEnv* __restrict__ env, const double* __restrict__ pi,
                                                  const double* __restrict__ z, const double* __restrict__ t,
                                                  const int32_t* __restrict__ cnt, uint64_t* __restrict__ rows) {
  const int ROWW = A + 6;
  const int64_t total = n * ROWW;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / ROWW;
    const int w = (int)(j - i * ROWW);
    uint64_t v;
    if (w == 0) v = env[i].a;
    else if (w == 1) v = env[i].b;
    else if (w == 2) v = env[i].aux;
    else if (w < 3 + A) v = (uint64_t)__double_as_longlong(pi[i * A + (w - 3)]);
    else if (w == 3 + A) v = (uint64_t)__double_as_longlong(z[i]);
    else if (w == 4 + A) v = (uint64_t)__double_as_longlong(t[i]);
    else v = (uint64_t)(uint32_t)cnt[i];
    rows[j] = v;
  }
}
// gathered [world][maxc][ROWW] -> compact SoA; off = exclusive prefix sums of the counts (device, world + 1 entries)
__global__ void __launch_bounds__(256) azc_k_unpack(int64_t total_n, int A, int world, int64_t maxc, const int64_t* __restrict__ counts,
                                                    const uint64_t* __restrict__ rows, AzEnv* __restrict__ env, double* __restrict__ pi,
                                                    double* __restrict__ z, double* __restrict__ t, int32_t* __restrict__ cnt) {
  const int ROWW = A + 6;
  const int64_t total = total_n * ROWW;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / ROWW;
    const int w = (int)(j - i * ROWW);
    int r = 0;
    int64_t base = 0;
    while (r + 1 < world && i >= base + counts[r]) { base += counts[r]; r++; }
    const uint64_t v = rows[((int64_t)r * maxc + (i - base)) * ROWW + w];
    if (w == 0) env[i].a = v;
    else if (w == 1) env[i].b = v;
    else if (w == 2) env[i].aux = (uint32_t)v;
    else if (w < 3 + A) pi[i * A + (w - 3)] = __longlong_as_double((long long)v);
    else if (w == 3 + A) z[i] = __longlong_as_double((long long)v);
    else if (w == 4 + A) t[i] = __longlong_as_double((long long)v);
    else cnt[i] = (int32_t)(uint32_t)v;
  }
}

// net parameters in blob order live on the host side of every network (az_net::load folds them on upload); the
// broadcast moves the fp32 blob GPU to GPU and hands it to load() on the receivers
extern "C" {

int32_t az_comm_unique_id(az_ctx* ctx, uint8_t id[AZ_COMM_ID_BYTES]) {
  if (!ctx || !id) return AZ_EINVAL;
  static_assert(sizeof(ncclUniqueId) == AZ_COMM_ID_BYTES, "AZ_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  NcclApi& api = nccl();
  if (!api.h) { ctx->err = api.why; return AZ_EUNSUPPORTED; }
  cudaSetDevice(ctx->device);
  ncclUniqueId u;
  AZ_NCCL(ctx, api.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return AZ_OK;
}

int32_t az_comm_create(az_ctx* ctx, const uint8_t id[AZ_COMM_ID_BYTES], int32_t rank, int32_t world, az_comm** out) {
  if (!ctx || !id || !out) return AZ_EINVAL;
  if (world < 1 || rank < 0 || rank >= world) { ctx->err = "az_comm_create: need 0 <= rank < world"; return AZ_EINVAL; }
  NcclApi& api = nccl();
  if (!api.h) { ctx->err = api.why; return AZ_EUNSUPPORTED; }
  cudaSetDevice(ctx->device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  az_comm* c = new az_comm();
  c->ctx = ctx; c->rank = rank; c->world = world;
  ncclResult_t r = api.CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) { ctx->err = std::string("ncclCommInitRank: ") + api.GetErrorString(r); delete c; return AZ_ECUDA; }
  if (cudaMalloc((void**)&c->d_counts, (size_t)(world + 1) * sizeof(int64_t)) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
      cudaEventCreate(&c->ev1) != cudaSuccess) {
    cudaGetLastError();
    api.CommDestroy(c->comm);
    delete c;
    ctx->err = "az_comm_create: cudaMalloc / cudaEventCreate failed";
    return AZ_ENOMEM;
  }
  // establish the NCCL channels now (the first collective on a communicator sets up its connections and proxy threads,
  // ~1-2 s): a one-element all-gather, so that the per-iteration exchanges run at link speed from their first call
  cudaMemsetAsync(c->d_counts, 0, (size_t)(world + 1) * sizeof(int64_t), ctx->stream);
  r = api.AllGather(c->d_counts + world, c->d_counts, 1, ncclInt64, c->comm, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (r != ncclSuccess || e != cudaSuccess) {
    ctx->err = std::string("az_comm_create: warm-up all-gather failed: ") + (r != ncclSuccess ? api.GetErrorString(r) : cudaGetErrorString(e));
    api.CommDestroy(c->comm); cudaFree(c->d_counts); cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    delete c;
    return AZ_ECUDA;
  }
  *out = c;
  return AZ_OK;
}

int32_t az_comm_rank(az_comm* c, int32_t* rank, int32_t* world) {
  if (!c) return AZ_EINVAL;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return AZ_OK;
}

int32_t az_comm_last_ms(az_comm* c, double* ms) {
  if (!c || !ms) return AZ_EINVAL;
  *ms = c->last_ms;
  return AZ_OK;
}

int32_t az_comm_destroy(az_comm* c) {
  if (!c) return AZ_EINVAL;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->comm) nccl().CommDestroy(c->comm);
  cudaFree(c->d_counts); cudaFree(c->d_send); cudaFree(c->d_recv);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
  return AZ_OK;
}

int32_t az_samples_allgather(az_comm* c, az_samples* local, az_samples** out, int64_t* counts_out) {
  if (!c || !local || !out) return AZ_EINVAL;
  az_ctx* ctx = c->ctx;
  try {
    NcclApi& api = nccl();
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    int64_t n = 0;
    az_samples_count(local, &n);
    const int game = az_samples_game(local);
    const int A = az_game_num_actions(game);
    const int ROWW = A + 6;
    const int world = c->world;
    // 1. counts (8 B per rank)
    AZ_CUDA(ctx, cudaMemcpyAsync(c->d_counts + world, &n, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    AZ_NCCL(ctx, api.AllGather(c->d_counts + world, c->d_counts, 1, ncclInt64, c->comm, st));
    std::vector<int64_t> counts(world);
    AZ_CUDA(ctx, cudaMemcpyAsync(counts.data(), c->d_counts, (size_t)world * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AZ_CUDA(ctx, cudaStreamSynchronize(st));
    int64_t maxc = 0, total = 0;
    for (int64_t v : counts) { maxc = std::max(maxc, v); total += v; }
    if (counts_out) for (int r = 0; r < world; r++) counts_out[r] = counts[r];
    az_samples* o = nullptr;
    int rc = az_samples_alloc(ctx, game, total, &o);
    if (rc != AZ_OK) return rc;
    if (total > 0) {
      // 2. pack -> one padded all-gather -> compact
      const size_t row_words = (size_t)maxc * ROWW;
      if (row_words > c->send_words || row_words * world > c->recv_words) {
        cudaFree(c->d_send); cudaFree(c->d_recv);
        c->d_send = c->d_recv = nullptr; c->send_words = c->recv_words = 0;
        const size_t want = row_words + row_words / 4;   // headroom: the next iteration's count differs a little
        if (cudaMalloc((void**)&c->d_send, want * 8) != cudaSuccess || cudaMalloc((void**)&c->d_recv, want * 8 * world) != cudaSuccess) {
          cudaGetLastError(); cudaFree(c->d_send); c->d_send = nullptr; az_samples_destroy(o);
          ctx->err = "az_samples_allgather: cudaMalloc of the staging rows failed";
          return AZ_ENOMEM;
        }
        c->send_words = want; c->recv_words = want * world;
      }
      uint64_t *d_send = c->d_send, *d_recv = c->d_recv;
      // last_ms = the exchange proper (pack -> all-gather -> compact); allocating the output set is host-side cudaMalloc time
      // (tens of ms once NCCL has enabled peer mappings) and is visible in the caller's wall clock
      AZ_CUDA(ctx, cudaEventRecord(c->ev0, st));
      const int grid = ctx->num_sms * 8;
      if (n > 0) azc_k_pack<<<grid, 256, 0, st>>>(n, A, az_samples_env(local), az_samples_pi(local), az_samples_z(local), az_samples_t(local),
                                                  az_samples_cnt(local), d_send);
      ncclResult_t r = api.AllGather(d_send, d_recv, row_words * 8, ncclUint8, c->comm, st);
      if (r == ncclSuccess)
        azc_k_unpack<<<grid, 256, 0, st>>>(total, A, world, maxc, c->d_counts, d_recv, az_samples_env(o), az_samples_pi(o), az_samples_z(o),
                                           az_samples_t(o), az_samples_cnt(o));
      ctx->launches += 2;
      cudaEventRecord(c->ev1, st);
      cudaError_t e = cudaStreamSynchronize(st);
      if (r != ncclSuccess) { az_samples_destroy(o); ctx->err = std::string("ncclAllGather: ") + api.GetErrorString(r); return AZ_ECUDA; }
      if (e != cudaSuccess) { az_samples_destroy(o); ctx->err = std::string("az_samples_allgather: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    } else {
      cudaEventRecord(c->ev0, st);
      cudaEventRecord(c->ev1, st);
      AZ_CUDA(ctx, cudaStreamSynchronize(st));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    *out = o;
    return AZ_OK;
  } catch (const std::exception& ex) {
    ctx->err = std::string("exception: ") + ex.what();
    return AZ_ESTATE;
  } catch (...) {
    ctx->err = "unknown exception";
    return AZ_ESTATE;
  }
}

int32_t az_net_broadcast(az_comm* c, az_net* net, const float* blob, int64_t n, int32_t root) {
  if (!c || !net) return AZ_EINVAL;
  az_ctx* ctx = c->ctx;
  if (net->ctx != ctx) { ctx->err = "az_net_broadcast: the network belongs to another context"; return AZ_EINVAL; }
  if (root < 0 || root >= c->world) { ctx->err = "az_net_broadcast: root out of range"; return AZ_EINVAL; }
  if (n != net->num_params() || n <= 0) { ctx->err = "az_net_broadcast: n must equal az_net_num_params (" + std::to_string(net->num_params()) + ")"; return AZ_EINVAL; }
  if (c->rank == root && !blob) { ctx->err = "az_net_broadcast: the root rank must pass the parameter blob"; return AZ_EINVAL; }
  try {
    NcclApi& api = nccl();
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    float* d = nullptr;
    if (cudaMalloc((void**)&d, (size_t)n * sizeof(float)) != cudaSuccess) { cudaGetLastError(); ctx->err = "az_net_broadcast: cudaMalloc failed"; return AZ_ENOMEM; }
    AZ_CUDA(ctx, cudaEventRecord(c->ev0, st));
    if (c->rank == root) cudaMemcpyAsync(d, blob, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st);
    ncclResult_t r = api.Broadcast(d, d, (size_t)n, ncclFloat32, root, c->comm, st);
    cudaEventRecord(c->ev1, st);
    std::vector<float> host;
    const float* src = blob;
    if (r == ncclSuccess && c->rank != root) {
      host.resize((size_t)n);
      cudaMemcpyAsync(host.data(), d, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st);
      src = host.data();
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (r != ncclSuccess) { ctx->err = std::string("ncclBroadcast: ") + api.GetErrorString(r); return AZ_ECUDA; }
    if (e != cudaSuccess) { ctx->err = std::string("az_net_broadcast: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    return net->load(src, n);  // fold BatchNorm + convert to the kernels' layouts (host fold of a few MB)
  } catch (const std::exception& ex) {
    ctx->err = std::string("exception: ") + ex.what();
    return AZ_ESTATE;
  } catch (...) {
    ctx->err = "unknown exception";
    return AZ_ESTATE;
  }
}

}  // extern "C"
