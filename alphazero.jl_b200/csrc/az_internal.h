// az_internal.h -- host-side types shared by the translation units of libazb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/azb200.h"
#include "az_games.cuh"

struct az_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int num_sms = 148;
  // two pinned bounce buffers for device -> caller (pageable) host copies, created on first use (az_d2h)
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
};
// Device -> caller-allocated host memory.  The ABI's output buffers are ordinary (pageable) host arrays; a plain cudaMemcpy into
// them is staged by the driver in small pieces (~2 GB/s measured for 150 MB).  Here the copy engine fills one pinned 8 MB
// buffer on the context's stream while the CPU drains the other into the caller's array.  Returns AZ_OK / AZ_ECUDA (ctx->err).
int az_d2h(az_ctx* ctx, void* dst, const void* src, size_t bytes);

#define AZ_CUDA(ctx, call)                                                                      \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ + ":" + std::to_string(__LINE__); \
      return AZ_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)

// An oracle in the sense of src/mcts.jl:6-17, evaluated for a whole leaf batch on the context's stream.
// envs/n_rows/P/V are DEVICE pointers; P is A-wide (masked, renormalised, zero on illegal actions).
struct az_net {
  az_ctx* ctx = nullptr;
  int kind = 0;
  int game = 0;
  // parity hook of az_net_forward_logits: when set, the next evaluation also writes the pre-softmax policy logits
  // [rows][A] and the pre-tanh value [rows] (device pointers; networks only)
  float* dbg_logit = nullptr;
  float* dbg_vpre = nullptr;
  virtual ~az_net() {}
  virtual int eval(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V) = 0;
  virtual int64_t num_params() { return 0; }
  virtual int load(const float*, int64_t) { return AZ_OK; }
  // parameters already on this network's device (default: stage through the host and take load(); ResNet folds on the device)
  virtual int load_device(const float* d_blob, int64_t n) {
    std::vector<float> h((size_t)std::max<int64_t>(n, 0));
    if (n > 0 && cudaMemcpy(h.data(), d_blob, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError(); ctx->err = "az_net_load_device: cudaMemcpy failed"; return AZ_ECUDA;
    }
    return load(h.data(), n);
  }
  // size every buffer an evaluation of up to max_rows leaves needs (allocations are illegal during stream capture)
  virtual int reserve(int max_rows) { (void)max_rows; return AZ_OK; }
  virtual bool capturable() { return true; }
  virtual uint64_t generation() { return 0; }  // changes whenever device pointers / weights baked into launches change  // false while per-launch CUDA events are being recorded
  virtual int set_profiling(int) { return AZ_OK; }
  virtual int get_profile(double* tower_ms, int64_t* tower_launches, double* total_ms, int64_t* evals) {
    if (tower_ms) *tower_ms = 0; if (tower_launches) *tower_launches = 0; if (total_ms) *total_ms = 0; if (evals) *evals = 0;
    return AZ_OK;
  }
  // Pinvalid (device, may be null) only for the forward_normalized ABI hook
  virtual int eval_with_pinv(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V, float* Pinv) {
    (void)Pinv;
    return eval(envs, n_rows, max_rows, P, V);
  }
};

az_net* az_make_resnet(az_ctx* ctx, int game, const az_resnet_hp* hp, int* status);
az_net* az_make_simplenet(az_ctx* ctx, int game, const az_simplenet_hp* hp, int* status);

// device-resident training samples (az_samples.cu)
struct az_samples;
int az_samples_alloc(az_ctx* ctx, int game, int64_t n, az_samples** out);
AzEnv* az_samples_env(az_samples* s);
double* az_samples_pi(az_samples* s);
double* az_samples_z(az_samples* s);
double* az_samples_t(az_samples* s);
int32_t* az_samples_cnt(az_samples* s);
int az_samples_game(az_samples* s);
extern "C" int32_t az_samples_destroy(az_samples* s);
