// Probe (not product code): is a tcgen05 K-major SWIZZLE_128B shared-memory descriptor whose start address is shifted by a
// whole number of 128-byte ROWS (not a multiple of the 1024-byte swizzle atom) read correctly, and which value of the
// descriptor's "matrix base offset" field (bits 49-51) does it need?  The tower kernel wants to serve the three dx taps of
// the 3x3 convolution from ONE shared-memory copy of an activation row by starting the A descriptor 0 / 1 / 2 rows in.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rowshift umma_rowshift.cu
//   run:   ./umma_rowshift        (prints one line per (shift, base_offset variant))
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int ROWS = 160, N = 16, K = 64;

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int shift, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                  // ROWS x 128 B, SWIZZLE_128B pattern anchored at the 1024-aligned base
  uint8_t* sb = smem + ROWS * 128;     // N x 128 B (ROWS*128 = 20480 = 20 atoms: still 1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    auto desc = [](uint32_t saddr, uint32_t bo) {
      uint64_t d = 0;
      d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(1024 >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(bo & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t a0 = smem_u32(sa) + (uint32_t)shift * 128u;
    const uint32_t bo = bo_mode == 0 ? 0u : (bo_mode == 1 ? ((a0 >> 7) & 7u) : ((8u - ((a0 >> 7) & 7u)) & 7u));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < K / 16; k++) {
      const uint64_t ad = desc(a0, bo) + (uint64_t)(k * 2), bd = desc(smem_u32(sb), 0) + (uint64_t)(k * 2);
      const uint32_t acc = k ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncwarp();
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; j++) D[tid * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  __half hA[ROWS * K], hB[N * K];
  float fA[ROWS * K], fB[N * K];
  srand(7);
  for (int i = 0; i < ROWS * K; i++) { fA[i] = (float)((rand() % 15) - 7); hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < N * K; i++) { fB[i] = (float)((rand() % 9) - 4); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const size_t smem = ROWS * 128 + N * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[3] = {"base_offset=0", "base_offset=(addr>>7)&7", "base_offset=(8-row)&7"};
  for (int shift = 0; shift <= 10; shift++)
    for (int mode = 0; mode < 3; mode++) {
      cudaMemset(dD, 0, 128 * N * 4);
      probe<<<1, 128, smem>>>(dA, dB, dD, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d %s: CUDA error %s\n", shift, names[mode], cudaGetErrorString(e)); return 1; }
      float hD[128 * N];
      cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
          float ref = 0;
          for (int k = 0; k < K; k++) ref += fA[(m + shift) * K + k] * fB[n * K + k];
          if (ref != hD[m * N + n]) bad++;
        }
      printf("shift %2d  %-26s : %s (%d mismatches of %d)\n", shift, names[mode], bad ? "WRONG" : "exact", bad, 128 * N);
    }
  return 0;
}
