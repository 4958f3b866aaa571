// Probe (not product code): how fast does one SM's TMA engine retire bulk-tensor STORES whose inner box extent is 32 / 64 /
// 128 bytes, and how long does a 16 KB A-stage LOAD (128 segments of 128 B, the tower kernel's stage) take from issue to
// mbarrier completion -- alone, and while the same CTA's other warps keep the store queue busy?  The persistent tower
// kernel's MMA warp waits 27 % of its time on `full` although its ring gives two input rows of lead: either loads are slow
// or they queue behind the epilogue's 32-byte-segment stores.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_rate tma_rate.cu -lcuda
//   run:   ./tma_rate
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// mode bit 0: epilogue warps (1..8) issue stores; bit 1: warp 0 issues loads and measures their latency
// store_cols = 16 / 32 / 64 fp16 channels per store tile (32 rows each); every warp stores `bytes_per_warp` in total
__global__ void __launch_bounds__(288, 1) probe(const __grid_constant__ CUtensorMap mS, const __grid_constant__ CUtensorMap mL,
                                               int mode, int store_cols, int stores_per_warp, int loads, int rows_total,
                                               long long* out /* [grid][4]: store clks, load clks total, load max, n */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tiles = smem;                      // 8 warps x 2 x 4 KB
  uint8_t* stage = smem + 8 * 8192;           // 4 x 16 KB load stages
  __shared__ uint64_t bar[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 8 * 8192 / 16; i += blockDim.x) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(i, i, i, i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    if (mode & 2) {
      long long tot = 0, mx = 0;
      for (int i = 0; i < loads; i++) {
        const int row0 = (int)(((long long)blockIdx.x * 977 + (long long)i * 131) % (rows_total / 128 - 1)) * 128;
        const long long a = clock64();
        if (lane == 0) { mbar_expect_tx(&bar[i & 3], 16384); tma_load_2d(stage + (i & 3) * 16384, &mL, &bar[i & 3], 0, row0); }
        __syncwarp();
        mbar_wait(&bar[i & 3], (i >> 2) & 1);
        const long long d = clock64() - a;
        tot += d; mx = d > mx ? d : mx;
        // spacing like the tower's conv stage (consumed over ~2300 clks): the next load is issued right away here,
        // so this measures back-to-back single-load latency
      }
      if (lane == 0) { out[blockIdx.x * 4 + 1] = tot; out[blockIdx.x * 4 + 2] = mx; out[blockIdx.x * 4 + 3] = loads; }
    }
  } else if (mode & 1) {
    uint8_t* my = tiles + (warp - 1) * 8192;
    const int tile_bytes = 32 * store_cols * 2;
    for (int i = 0; i < stores_per_warp; i++) {
      const int row0 = (int)(((long long)blockIdx.x * 8 + (warp - 1) + (long long)i * 1187) % (rows_total / 32)) * 32;
      const int col0 = (i * store_cols) % 128;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      if (lane == 0) { tma_store_2d(&mS, my + (i & 1) * tile_bytes, col0, row0); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x * 4 + 0] = clock64() - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(PFN_encodeTiled fn, void* base, uint64_t rows, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t dims[2] = {128, rows};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN_encodeTiled fn = (PFN_encodeTiled)p;
  const int rows = 3000 * 42;                 // one activation tensor of the tower at ~3000 leaves (32 MB: L2 resident)
  __half* buf; long long* out;
  cudaMalloc(&buf, (size_t)rows * 256 + (1 << 20));
  cudaMemset(buf, 0, (size_t)rows * 256);
  cudaMalloc(&out, 148 * 4 * sizeof(long long));
  const size_t smem = 8 * 8192 + 4 * 16384 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  CUtensorMap mL = make_map(fn, buf, rows, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  const int total_bytes_per_warp = 4096 * 128;  // 512 KB per warp per run
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int mode = 1; mode <= 3; mode++) {
    for (int cols : {16, 32, 64}) {
      if (mode == 2 && cols != 16) continue;
      CUtensorMapSwizzle sw = cols == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
      CUtensorMap mS = make_map(fn, buf, rows, cols, 32, sw);
      const int stores = total_bytes_per_warp / (32 * cols * 2);
      const int loads = 400;
      for (int rep = 0; rep < 2; rep++) {
        cudaMemset(out, 0, 148 * 4 * sizeof(long long));
        probe<<<148, 288, smem>>>(mS, mL, mode, cols, stores, loads, rows, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148 * 4];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double clk = 0, lt = 0, lm = 0, ln = 0;
      for (int i = 0; i < 148; i++) { clk += h[i * 4]; lt += h[i * 4 + 1]; lm = h[i * 4 + 2] > lm ? h[i * 4 + 2] : lm; ln += h[i * 4 + 3]; }
      clk /= 148;
      printf("mode %d (%s%s) store tile 32 rows x %3d B: ", mode, (mode & 1) ? "8 warps storing" : "", (mode & 2) ? ((mode & 1) ? " + loads" : "loads only") : "", cols * 2);
      if (mode & 1) printf("%7.0f clks for %d stores/warp x 8 warps = %.1f clks per store (%.2f B/clk/SM, %.1f clks per 32-row segment group)  ", clk, stores,
                           clk / (stores * 8.0), 8.0 * total_bytes_per_warp / clk, clk / (stores * 8.0));
      if (mode & 2) printf("16 KB load (128 x 128 B): avg %.0f clks, max %.0f", lt / ln, lm);
      printf("\n");
    }
  }
  return 0;
}
