// az_net.cu -- policy/value ResNet forward on the leaf batch (replaces Network.forward / forward_normalized /
// evaluate_batch: src/networks/flux.jl:127-132, src/networks/network.jl:264-271,308-315, and the Flux layers of
// src/networks/architectures/resnet.jl:53-92).
//
// Data layout in HBM.  Activations are fp16 "padded NHWC" rows of F=128 channels:
//     row(board b, col x, row y) = b*BS + y*(W+1) + x,   BS = (W+1)*(H+1)      (Connect-Four: BS = 56)
// with column x = W and row y = H kept at ZERO, so that the 3x3 "same" convolution becomes nine row-shifted GEMMs
//     out[p, co] = sum_{tap} sum_{ci} act[p + off(tap), ci] * Wt[co, tap*F + ci],   off = dy*(W+1) + dx
// and each tap's A operand is ONE TMA box of 128 consecutive rows (out-of-range rows are zero-filled by TMA).
// The tower kernel is a warp-specialised tcgen05 implicit GEMM: TMA producer warp -> 6-stage smem ring (SWIZZLE_128B)
// -> single-thread tcgen05.mma (M=128, N=128, K=16, fp16 operands, fp32 accumulators in TMEM, double buffered)
// -> 4 epilogue warps (tcgen05.ld, folded-BN bias, residual, ReLU, pad-row zeroing, fp16 store).
// BatchNorm (test mode, eps = 1e-5) is folded into the conv weights/bias when the blob is loaded.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <vector>

#include "az_internal.h"

#define AZ_TRY2(expr) do { int s__ = (expr); if (s__ != AZ_OK) return s__; } while (0)

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                   // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

// ------------------------------------------------------------------------------------------------
// tower conv: implicit GEMM, one launch per conv layer
// ------------------------------------------------------------------------------------------------
namespace tc {
constexpr int BM = 128, BN = 128, BK = 64, STAGES = 6, F = 128;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int NUM_THREADS = 192;
struct Smem {
  uint8_t a[STAGES][A_BYTES];
  uint8_t b[STAGES][B_BYTES];
  uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
  float bias[BN];
};
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);  // f16 x f16 -> f32, K-major A/B
}  // namespace tc

struct ConvGeom {
  int row_stride;  // W + 1
  int board_rows;  // (W+1)*(H+1)
  int valid_rows;  // (W+1)*H
  int wcols;       // W
  int ntaps;       // 9
  int off[9];
};

__global__ void __launch_bounds__(tc::NUM_THREADS, 1)
az_k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __half* __restrict__ resid,
             __half* __restrict__ out, const float* __restrict__ bias, const int32_t* __restrict__ n_boards, ConvGeom g,
             int alloc_rows) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_used = (*n_boards) * g.board_rows;
  const int num_tiles = (rows_used + BM - 1) / BM;
  const int kblocks = g.ntaps * (F / BK);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&s.tfull[i], 1); mbar_init(&s.tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64) s.bias[threadIdx.x - 64] = bias[threadIdx.x - 64];
  if (warp == 1) {  // TMEM: 2 accumulators x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&s.empty[stage], phase ^ 1);
          mbar_expect_tx(&s.full[stage], A_BYTES + B_BYTES);
          tma_load_2d(s.a[stage], &tmA, &s.full[stage], (kb & 1) * BK, tile * BM + g.off[kb >> 1]);
          tma_load_2d(s.b[stage], &tmW, &s.full[stage], kb * BK, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&s.tempty[acc], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&s.full[stage], phase);
          tcgen05_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(s.a[stage]));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(s.b[stage]));
#pragma unroll
          for (int k = 0; k < BK / 16; k++)  // advance 32 B (= 16 fp16) inside the 128-B swizzle row
            umma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
          umma_commit(&s.empty[stage]);  // frees the smem stage when these MMAs retire
          if (kb == kblocks - 1) umma_commit(&s.tfull[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {  // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&s.tfull[acc], aphase);
      tcgen05_fence_after();
      const int p = tile * BM + quarter * 32 + lane;
      const int r = p % g.board_rows;
      const bool valid = (p < rows_used) && (r < g.valid_rows) && ((r % g.row_stride) != g.wcols);
      const bool in_alloc = p < alloc_rows;
#pragma unroll 1
      for (int c = 0; c < BN / 32; c++) {
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(quarter * 32) << 16), v);
        uint4 rs[4];
        if (resid != nullptr && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(resid + (size_t)p * F + c * 32);
#pragma unroll
          for (int j = 0; j < 4; j++) rs[j] = rp[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) rs[j] = make_uint4(0, 0, 0, 0);
        }
        const __half2* rh = reinterpret_cast<const __half2*>(rs);
        uint4 o[4];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float2 rr = __half22float2(rh[j]);
          float x0 = __uint_as_float(v[2 * j]) + s.bias[c * 32 + 2 * j] + rr.x;
          float x1 = __uint_as_float(v[2 * j + 1]) + s.bias[c * 32 + 2 * j + 1] + rr.y;
          x0 = valid ? fmaxf(x0, 0.0f) : 0.0f;
          x1 = valid ? fmaxf(x1, 0.0f) : 0.0f;
          oh[j] = __floats2half2_rn(x0, x1);
        }
        if (in_alloc) {
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)p * F + c * 32);
#pragma unroll
          for (int j = 0; j < 4; j++) op[j] = o[j];
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tempty[acc]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// stem: leaf states -> first activation (conv 3x3, C_in -> 128, folded BN, ReLU) on CUDA cores.
// Input planes come straight from the game's vectorize_state (no host round trip; replaces
// GI.vectorize_state + Flux.batch + convert_input, src/networks/network.jl:310-312).
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void __launch_bounds__(128) az_k_stem(const AzEnv* __restrict__ envs, const int32_t* __restrict__ n_boards,
                                                 const float* __restrict__ wstem /* [9][C][128] */, const float* __restrict__ bias,
                                                 __half* __restrict__ out) {
  constexpr int W = G::XW, H = G::XH, C = G::XC, RS = W + 1, BS = (W + 1) * (H + 1);
  __shared__ float x[W * H * C];
  __shared__ float xp[(H + 2) * (W + 2) * C];
  const int b = blockIdx.x;
  if (b >= *n_boards) return;
  if (threadIdx.x == 0) G::vectorize(envs[b], x);
  for (int i = threadIdx.x; i < (H + 2) * (W + 2) * C; i += blockDim.x) xp[i] = 0.0f;
  __syncthreads();
  for (int i = threadIdx.x; i < W * H * C; i += blockDim.x) {
    int c = i / (W * H), rem = i % (W * H), yy = rem / W, xx = rem % W;
    xp[((yy + 1) * (W + 2) + (xx + 1)) * C + c] = x[i];
  }
  __syncthreads();
  const int co = threadIdx.x;
  const float bs = bias[co];
  __half* ob = out + (size_t)b * BS * 128;
  for (int r = 0; r < BS; r++) {
    const int yy = r / RS, xx = r % RS;
    float acc = 0.0f;
    if (yy < H && xx < W) {
      acc = bs;
#pragma unroll
      for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          // Flux Conv is a true convolution: tap (kx,ky) reads the input at (x + 1 - kx, y + 1 - ky)
          const float* px = &xp[((yy + 1 + 1 - ky) * (W + 2) + (xx + 1 + 1 - kx)) * C];
          const float* pw = &wstem[((ky * 3 + kx) * C) * 128 + co];
#pragma unroll
          for (int c = 0; c < C; c++) acc += px[c] * pw[c * 128];
        }
      acc = fmaxf(acc, 0.0f);
    }
    ob[(size_t)r * 128 + co] = __float2half_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// heads: 1x1 convs (+BN+ReLU), flatten, dense layers, softmax / tanh, legal-action mask + renormalisation
// (resnet.jl:79-90, network.jl:264-271).  NB boards per CTA so that dense weights are reused from L1/L2.
// ------------------------------------------------------------------------------------------------
struct HeadParams {
  const float* wc;    // [128][npf+nvf] folded 1x1 conv weights (policy filters first), fp32
  const float* bc;    // [npf+nvf]
  const __half* wv1;  // [in = WH*nvf][128] (transposed: out fastest), fp16
  const float* bv1;   // [128]
  const float* wv2;   // [128]
  const float* bv2;   // [1]
  const float* wp;    // [in = WH*npf][A] (out fastest), fp32
  const float* bp;    // [A]
  int npf, nvf;
};

template <class G, int NB>
__global__ void __launch_bounds__(256) az_k_heads(const __half* __restrict__ act, const AzEnv* __restrict__ envs,
                                                  const int32_t* __restrict__ n_boards, HeadParams hp, float* __restrict__ P,
                                                  float* __restrict__ V, float* __restrict__ Pinv) {
  constexpr int W = G::XW, H = G::XH, RS = W + 1, BS = (W + 1) * (H + 1), WH = W * H, A = G::A, F = 128;
  extern __shared__ uint8_t hs_raw[];
  const int nh = hp.npf + hp.nvf;
  float* wc = reinterpret_cast<float*>(hs_raw);              // [128][nh]
  float* hfeat = wc + F * nh;                                // [NB][WH*nh]  (policy block then value block, Flux flatten order)
  float* hid = hfeat + NB * WH * nh;                         // [NB][128]
  float* logit = hid + NB * F;                               // [NB][A + 1]
  __half* xin = reinterpret_cast<__half*>(logit + NB * (A + 1));  // [NB][WH][128]
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, *n_boards - b0);
  if (nb <= 0) return;
  for (int i = threadIdx.x; i < F * nh; i += blockDim.x) wc[i] = hp.wc[i];
  for (int i = threadIdx.x; i < nb * WH * (F / 8); i += blockDim.x) {
    int bb = i / (WH * (F / 8)), rem = i % (WH * (F / 8)), pos = rem / (F / 8), ch = rem % (F / 8);
    int yy = pos / W, xx = pos % W;
    const uint4* src = reinterpret_cast<const uint4*>(act + ((size_t)(b0 + bb) * BS + yy * RS + xx) * F) + ch;
    reinterpret_cast<uint4*>(xin + ((size_t)bb * WH + pos) * F)[ch] = *src;
  }
  __syncthreads();
  // 1x1 convs: hfeat[bb][pos + WH*c'] (c' within its head), ReLU
  for (int i = threadIdx.x; i < nb * WH * nh; i += blockDim.x) {
    int bb = i / (WH * nh), rem = i % (WH * nh), c = rem / WH, pos = rem % WH;
    const __half2* xr = reinterpret_cast<const __half2*>(xin + ((size_t)bb * WH + pos) * F);
    float acc = hp.bc[c];
#pragma unroll 8
    for (int k = 0; k < F / 2; k++) {
      float2 xv = __half22float2(xr[k]);
      acc += xv.x * wc[(2 * k) * nh + c] + xv.y * wc[(2 * k + 1) * nh + c];
    }
    hfeat[(size_t)bb * WH * nh + rem] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  // value dense 1: hid[bb][o] = relu(sum_i wv1[i][o] * hv[bb][i] + b)
  {
    const int o = threadIdx.x % F, half_id = threadIdx.x / F;  // 2 halves x 128 outputs
    float acc[NB / 2];
#pragma unroll
    for (int j = 0; j < NB / 2; j++) acc[j] = hp.bv1[o];
    const int nin = WH * hp.nvf;
    for (int i = 0; i < nin; i++) {
      const float w = __half2float(hp.wv1[(size_t)i * F + o]);
#pragma unroll
      for (int j = 0; j < NB / 2; j++) acc[j] += w * hfeat[(size_t)(half_id * (NB / 2) + j) * WH * nh + WH * hp.npf + i];
    }
#pragma unroll
    for (int j = 0; j < NB / 2; j++) hid[(half_id * (NB / 2) + j) * F + o] = fmaxf(acc[j], 0.0f);
  }
  __syncthreads();
  // policy logits (A per board) and value output: one warp per output
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nin = WH * hp.npf;
    for (int t = warp; t < nb * (A + 1); t += nwarps) {
      const int bb = t / (A + 1), o = t % (A + 1);
      float acc = 0.0f;
      if (o < A) {
        for (int i = lane; i < nin; i += 32) acc += hp.wp[(size_t)i * A + o] * hfeat[(size_t)bb * WH * nh + i];
      } else {
        for (int i = lane; i < F; i += 32) acc += hp.wv2[i] * hid[bb * F + i];
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) logit[bb * (A + 1) + o] = acc + (o < A ? hp.bp[o] : hp.bv2[0]);
    }
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    const int bb = threadIdx.x, row = b0 + bb;
    const float* lg = logit + bb * (A + 1);
    float m = lg[0];
    for (int a = 1; a < A; a++) m = fmaxf(m, lg[a]);
    float e[A], se = 0.0f;
    for (int a = 0; a < A; a++) { e[a] = expf(lg[a] - m); se += e[a]; }
    const uint32_t legal = G::legal_mask(envs[row]);
    float sp = 0.0f;
    for (int a = 0; a < A; a++) { e[a] = ((legal >> a) & 1u) ? e[a] / se : 0.0f; sp += e[a]; }
    for (int a = 0; a < A; a++) P[(size_t)row * A + a] = e[a] / (sp + 1.1920929e-07f);  // eps(Float32), network.jl:268
    V[row] = tanhf(lg[A]);
    if (Pinv) Pinv[row] = 1.0f - sp;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
static int make_map_2d(az_ctx* ctx, CUtensorMap* m, void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { ctx->err = "cuTensorMapEncodeTiled not available"; return AZ_ECUDA; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return AZ_ECUDA; }
  return AZ_OK;
}

template <class G>
struct ResNetImpl : az_net {
  az_resnet_hp hp{};
  static constexpr int F = 128, W = G::XW, H = G::XH, C = G::XC, A = G::A, BS = (W + 1) * (H + 1), WH = W * H;
  static constexpr int NB = 8;
  // device weights
  float* d_wstem = nullptr; float* d_bstem = nullptr;
  std::vector<__half*> d_wconv; std::vector<float*> d_bconv;
  float *d_wc = nullptr, *d_bc = nullptr, *d_bv1 = nullptr, *d_wv2 = nullptr, *d_bv2 = nullptr, *d_wp = nullptr, *d_bp = nullptr;
  __half* d_wv1 = nullptr;
  std::vector<CUtensorMap> mapW;
  // activations (allocated for max_rows on first use)
  int act_boards = 0, alloc_rows = 0;
  __half *d_x = nullptr, *d_t = nullptr;
  CUtensorMap mapX{}, mapT{};
  ConvGeom geom{};
  bool loaded = false;
  size_t conv_smem = 0, head_smem = 0;
  // profiling: 4 events per evaluation (start, tower begin, tower end, end)
  static constexpr int PROF_SLOTS = 8192;
  bool profiling = false;
  std::vector<cudaEvent_t> pev;
  int64_t prof_evals = 0;

  int set_profiling(int enable) override {
    if (enable && pev.empty()) {
      pev.resize((size_t)PROF_SLOTS * 4);
      for (auto& e : pev) if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "cudaEventCreate failed"; return AZ_ECUDA; }
    }
    cudaStreamSynchronize(ctx->stream);
    profiling = enable != 0;
    prof_evals = 0;
    return AZ_OK;
  }
  int get_profile(double* tower_ms, int64_t* tower_launches, double* total_ms, int64_t* evals) override {
    cudaStreamSynchronize(ctx->stream);
    double tw = 0, tt = 0;
    int64_t n = std::min<int64_t>(prof_evals, PROF_SLOTS);
    for (int64_t i = 0; i < n; i++) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, pev[i * 4 + 1], pev[i * 4 + 2]);
      cudaEventElapsedTime(&b, pev[i * 4 + 0], pev[i * 4 + 3]);
      tw += a; tt += b;
    }
    if (tower_ms) *tower_ms = tw;
    if (tower_launches) *tower_launches = n * 2 * hp.num_blocks;
    if (total_ms) *total_ms = tt;
    if (evals) *evals = n;
    prof_evals = 0;
    return AZ_OK;
  }

  int init() {
    if (hp.num_filters != F || hp.conv_kernel_size[0] != 3 || hp.conv_kernel_size[1] != 3) {
      ctx->err = "ResNet: this build supports num_filters = 128 and conv_kernel_size = (3, 3)";
      return AZ_EUNSUPPORTED;
    }
    if (hp.num_blocks < 0 || hp.num_policy_head_filters < 1 || hp.num_value_head_filters < 1 ||
        hp.num_policy_head_filters + hp.num_value_head_filters > 64) {
      ctx->err = "ResNet: head filters must satisfy 1 <= npf, nvf and npf + nvf <= 64";
      return AZ_EINVAL;
    }
    geom.row_stride = W + 1; geom.board_rows = BS; geom.valid_rows = (W + 1) * H; geom.wcols = W; geom.ntaps = 9;
    for (int ky = 0; ky < 3; ky++)
      for (int kx = 0; kx < 3; kx++) geom.off[ky * 3 + kx] = (1 - ky) * (W + 1) + (1 - kx);
    conv_smem = sizeof(tc::Smem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(az_k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem);
    if (e != cudaSuccess) { ctx->err = std::string("conv smem attribute: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    const int nh = hp.num_policy_head_filters + hp.num_value_head_filters;
    head_smem = sizeof(float) * ((size_t)F * nh + (size_t)NB * WH * nh + NB * F + NB * (A + 1)) + sizeof(__half) * (size_t)NB * WH * F + 16;
    e = cudaFuncSetAttribute(az_k_heads<G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem);
    if (e != cudaSuccess) { ctx->err = std::string("heads smem attribute: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    return AZ_OK;
  }
  int64_t num_params() override {
    const int npf = hp.num_policy_head_filters, nvf = hp.num_value_head_filters;
    int64_t n = 0;
    n += 9LL * C * F + F + 4 * F;                                   // stem conv + BN
    n += (int64_t)hp.num_blocks * 2 * (9LL * F * F + F + 4 * F);    // blocks
    n += (int64_t)F * nvf + nvf + 4 * nvf + (int64_t)WH * nvf * F + F + F + 1;  // vhead
    n += (int64_t)F * npf + npf + 4 * npf + (int64_t)WH * npf * A + A;          // phead
    return n;
  }
  template <class T> int up(T** dst, const std::vector<T>& v) {
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) { ctx->err = "cudaMalloc (weights) failed"; return AZ_ENOMEM; }
    e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ctx->err = "cudaMemcpy (weights) failed"; return AZ_ECUDA; }
    return AZ_OK;
  }
  void free_weights() {
    cudaFree(d_wstem); cudaFree(d_bstem);
    for (auto p : d_wconv) cudaFree(p);
    for (auto p : d_bconv) cudaFree(p);
    d_wconv.clear(); d_bconv.clear(); mapW.clear();
    cudaFree(d_wc); cudaFree(d_bc); cudaFree(d_wv1); cudaFree(d_bv1); cudaFree(d_wv2); cudaFree(d_bv2); cudaFree(d_wp); cudaFree(d_bp);
    d_wstem = d_bstem = d_wc = d_bc = d_bv1 = d_wv2 = d_bv2 = d_wp = d_bp = nullptr; d_wv1 = nullptr;
  }
  ~ResNetImpl() override { free_weights(); cudaFree(d_x); cudaFree(d_t); for (auto e : pev) cudaEventDestroy(e); }

  // blob -> folded device weights.  Flux order: Conv W[kw,kh,cin,cout] (kw fastest), b; BatchNorm gamma, beta, mu, sigma2;
  // Dense W[out,in] (out fastest), b.  Order: common (stem, blocks), vhead, phead.
  int load(const float* blob, int64_t n) override {
    if (n != num_params()) { ctx->err = "az_net_load: blob has " + std::to_string(n) + " floats, expected " + std::to_string(num_params()); return AZ_EINVAL; }
    cudaStreamSynchronize(ctx->stream);
    free_weights();
    const float* q = blob;
    const float eps = 1e-5f;
    auto fold = [&](int cout, const float* b, const float* bn, std::vector<float>& scale, std::vector<float>& shift) {
      scale.resize(cout); shift.resize(cout);
      for (int o = 0; o < cout; o++) {
        float sc = bn[o] / std::sqrt(bn[3 * cout + o] + eps);
        scale[o] = sc;
        shift[o] = (b[o] - bn[2 * cout + o]) * sc + bn[cout + o];
      }
    };
    std::vector<float> scale, shift;
    {  // stem
      const float* w = q; q += 9 * C * F;
      const float* b = q; q += F;
      const float* bn = q; q += 4 * F;
      fold(F, b, bn, scale, shift);
      std::vector<float> ws((size_t)9 * C * F);
      for (int ky = 0; ky < 3; ky++) for (int kx = 0; kx < 3; kx++) for (int c = 0; c < C; c++) for (int o = 0; o < F; o++)
        ws[((size_t)(ky * 3 + kx) * C + c) * F + o] = w[kx + 3 * (ky + 3 * (c + (size_t)C * o))] * scale[o];
      AZ_TRY2(up(&d_wstem, ws)); AZ_TRY2(up(&d_bstem, shift));
    }
    for (int l = 0; l < 2 * hp.num_blocks; l++) {
      const float* w = q; q += 9LL * F * F;
      const float* b = q; q += F;
      const float* bn = q; q += 4 * F;
      fold(F, b, bn, scale, shift);
      std::vector<__half> wh((size_t)F * 9 * F);  // Wt[co][tap*F + ci]
      for (int o = 0; o < F; o++) for (int ky = 0; ky < 3; ky++) for (int kx = 0; kx < 3; kx++) for (int c = 0; c < F; c++)
        wh[(size_t)o * 9 * F + (size_t)(ky * 3 + kx) * F + c] = __float2half_rn(w[kx + 3 * (ky + 3 * (c + (size_t)F * o))] * scale[o]);
      __half* dw = nullptr; float* db = nullptr;
      AZ_TRY2(up(&dw, wh)); d_wconv.push_back(dw);
      AZ_TRY2(up(&db, shift)); d_bconv.push_back(db);
      CUtensorMap m;
      AZ_TRY2(make_map_2d(ctx, &m, dw, 9 * F, F, tc::BK, tc::BN));
      mapW.push_back(m);
    }
    const int npf = hp.num_policy_head_filters, nvf = hp.num_value_head_filters, nh = npf + nvf;
    std::vector<float> wc((size_t)F * nh), bc(nh);
    {  // vhead: Conv1x1 F->nvf, BN, Dense(WH*nvf -> F), Dense(F -> 1)
      const float* w = q; q += (int64_t)F * nvf;
      const float* b = q; q += nvf;
      const float* bn = q; q += 4 * nvf;
      fold(nvf, b, bn, scale, shift);
      for (int c = 0; c < F; c++) for (int o = 0; o < nvf; o++) wc[(size_t)c * nh + npf + o] = w[c + (size_t)F * o] * scale[o];
      for (int o = 0; o < nvf; o++) bc[npf + o] = shift[o];
      const float* w1 = q; q += (int64_t)WH * nvf * F;
      const float* b1 = q; q += F;
      std::vector<__half> wv1((size_t)WH * nvf * F);
      for (int i = 0; i < WH * nvf; i++) for (int o = 0; o < F; o++) wv1[(size_t)i * F + o] = __float2half_rn(w1[o + (size_t)F * i]);
      AZ_TRY2(up(&d_wv1, wv1));
      AZ_TRY2(up(&d_bv1, std::vector<float>(b1, b1 + F)));
      const float* w2 = q; q += F;
      const float* b2 = q; q += 1;
      AZ_TRY2(up(&d_wv2, std::vector<float>(w2, w2 + F)));
      AZ_TRY2(up(&d_bv2, std::vector<float>(b2, b2 + 1)));
    }
    {  // phead: Conv1x1 F->npf, BN, Dense(WH*npf -> A)
      const float* w = q; q += (int64_t)F * npf;
      const float* b = q; q += npf;
      const float* bn = q; q += 4 * npf;
      fold(npf, b, bn, scale, shift);
      for (int c = 0; c < F; c++) for (int o = 0; o < npf; o++) wc[(size_t)c * nh + o] = w[c + (size_t)F * o] * scale[o];
      for (int o = 0; o < npf; o++) bc[o] = shift[o];
      const float* w1 = q; q += (int64_t)WH * npf * A;
      const float* b1 = q; q += A;
      std::vector<float> wp((size_t)WH * npf * A);
      for (int i = 0; i < WH * npf; i++) for (int o = 0; o < A; o++) wp[(size_t)i * A + o] = w1[o + (size_t)A * i];
      AZ_TRY2(up(&d_wp, wp));
      AZ_TRY2(up(&d_bp, std::vector<float>(b1, b1 + A)));
    }
    AZ_TRY2(up(&d_wc, wc)); AZ_TRY2(up(&d_bc, bc));
    loaded = true;
    return AZ_OK;
  }
  int ensure_act(int max_boards) {
    if (max_boards <= act_boards) return AZ_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_x); cudaFree(d_t);
    d_x = d_t = nullptr;
    alloc_rows = ((max_boards * BS + 127) / 128) * 128 + 128;
    size_t bytes = (size_t)alloc_rows * F * sizeof(__half);
    if (cudaMalloc((void**)&d_x, bytes) != cudaSuccess || cudaMalloc((void**)&d_t, bytes) != cudaSuccess) {
      ctx->err = "cudaMalloc (activations) failed"; cudaGetLastError(); return AZ_ENOMEM;
    }
    cudaMemsetAsync(d_x, 0, bytes, ctx->stream);
    cudaMemsetAsync(d_t, 0, bytes, ctx->stream);
    AZ_TRY2(make_map_2d(ctx, &mapX, d_x, F, alloc_rows, tc::BK, tc::BM));
    AZ_TRY2(make_map_2d(ctx, &mapT, d_t, F, alloc_rows, tc::BK, tc::BM));
    act_boards = max_boards;
    return AZ_OK;
  }
  int eval(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V) override {
    return eval_with_pinv(envs, n_rows, max_rows, P, V, nullptr);
  }
  int eval_with_pinv(const AzEnv* envs, const int32_t* n_rows, int max_rows, float* P, float* V, float* Pinv) override {
    if (!loaded) { ctx->err = "ResNet: az_net_load must be called before the network is used"; return AZ_ESTATE; }
    AZ_TRY2(ensure_act(max_rows));
    cudaStream_t st = ctx->stream;
    const bool prof = profiling && prof_evals < PROF_SLOTS;
    cudaEvent_t* pe = prof ? &pev[(size_t)prof_evals * 4] : nullptr;
    if (prof) cudaEventRecord(pe[0], st);
    az_k_stem<G><<<max_rows, 128, 0, st>>>(envs, n_rows, d_wstem, d_bstem, d_x);
    if (prof) cudaEventRecord(pe[1], st);
    const int max_tiles = (max_rows * BS + tc::BM - 1) / tc::BM;
    const int grid = std::min(max_tiles, ctx->num_sms);
    for (int blk = 0; blk < hp.num_blocks; blk++) {
      az_k_conv_tc<<<grid, tc::NUM_THREADS, conv_smem, st>>>(mapX, mapW[2 * blk], nullptr, d_t, d_bconv[2 * blk], n_rows, geom, alloc_rows);
      az_k_conv_tc<<<grid, tc::NUM_THREADS, conv_smem, st>>>(mapT, mapW[2 * blk + 1], d_x, d_x, d_bconv[2 * blk + 1], n_rows, geom, alloc_rows);
    }
    if (prof) cudaEventRecord(pe[2], st);
    HeadParams h{d_wc, d_bc, d_wv1, d_bv1, d_wv2, d_bv2, d_wp, d_bp, hp.num_policy_head_filters, hp.num_value_head_filters};
    az_k_heads<G, NB><<<(max_rows + NB - 1) / NB, 256, head_smem, st>>>(d_x, envs, n_rows, h, P, V, Pinv);
    if (prof) { cudaEventRecord(pe[3], st); prof_evals++; }
    ctx->launches += 2 + 2 * hp.num_blocks;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("network launch: ") + cudaGetErrorString(e); return AZ_ECUDA; }
    return AZ_OK;
  }
};

az_net* az_make_resnet(az_ctx* ctx, int game, const az_resnet_hp* hp, int* status) {
  az_net* n = nullptr;
  int st = AZ_OK;
  auto mk = [&](auto* impl) {
    impl->ctx = ctx; impl->kind = AZ_NET_RESNET; impl->game = game; impl->hp = *hp;
    st = impl->init();
    if (st != AZ_OK) { delete impl; return (az_net*)nullptr; }
    return (az_net*)impl;
  };
  switch (game) {
    case 0: n = mk(new ResNetImpl<GameC4>()); break;
    case 1: n = mk(new ResNetImpl<GameTTT>()); break;
    case 2: n = mk(new ResNetImpl<GameMancala>()); break;
    default: ctx->err = "az_net_create_resnet: unknown game"; st = AZ_EINVAL;
  }
  *status = st;
  return n;
}

az_net* az_make_simplenet(az_ctx* ctx, int game, const az_simplenet_hp* hp, int* status) {
  (void)game; (void)hp;
  ctx->err = "SimpleNet forward is not built in this round (SURVEY 8f)";
  *status = AZ_EUNSUPPORTED;
  return nullptr;
}
