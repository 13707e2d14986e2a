// Development micro-benchmark: how fast does TMA gather a channels-last tile whose inner extent is only 16 bytes?
// Tensor x [N][S][S][C] fp16 (C = 48 -> 96-byte pixel pitch).  One CTA per SM loads tiles {8 ch, S w, 1 h} row by row (S row boxes of
// S x 16 bytes -> one 39 KB tile with padded rows, as csrc/fft2d_mma.cu wants it) or {8, S, S} in one box, and reports clocks per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmp/mb_tma_rows tools/mb_tma_rows.cu && tools/tmp/mb_tma_rows
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int S = 48, C = 48, RS = 768;      // dense rows: TMA destinations must be 128-byte aligned (a 816-byte padded pitch is not)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (unsigned spins = 0; spins < (1u << 24); ++spins) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// mode 0: S row boxes {8, S, 1, 1} per tile issued by the 32 lanes of warp 0; mode 1: one box {8, S, S, 1} (dense 36 KB)
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm_row, const __grid_constant__ CUtensorMap tm_box, int mode,
                                            int tiles_per_cta, int cblocks, int n_images, long long* clk, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const uint32_t b0 = smem_u32(&bar[0]);
  const uint32_t t0 = (smem_u32(smem) + 127u) & ~127u;
  constexpr uint32_t kTile = S * RS + 256;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto issue = [&](int it) {
    const int ti = (blockIdx.x + it * gridDim.x) % (cblocks * n_images);
    const int n = ti / cblocks, c0 = (ti % cblocks) * 8;
    const uint32_t bar_a = b0 + 8u * (it & 1), dst = t0 + (it & 1) * kTile;
    if (warp == 0) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(S * S * 16)) : "memory");
      __syncwarp();
      if (mode == 0) {
        for (int h = lane; h < S; h += 32)
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst + (uint32_t)h * RS), "l"(&tm_row), "r"(bar_a), "r"(c0), "r"(0), "r"(h), "r"(n) : "memory");
      } else if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&tm_box), "r"(bar_a), "r"(c0), "r"(0), "r"(0), "r"(n) : "memory");
      }
    }
  };
  float acc = 0.f;
  issue(0);
  const long long c_start = clock64();
  for (int it = 0; it < tiles_per_cta; ++it) {
    if (it + 1 < tiles_per_cta) issue(it + 1);                         // double-buffered: the next tile is in flight
    mbar_wait(b0 + 8u * (it & 1), (uint32_t)((it >> 1) & 1));
    const uint8_t* tile = smem + (t0 - smem_u32(smem)) + (it & 1) * kTile;
    acc += __half2float(*reinterpret_cast<const __half*>(tile + (threadIdx.x % S) * RS + (threadIdx.x % 32) * 16));
    __syncthreads();                                                    // buffer (it & 1) is free again for tile it + 2
  }
  const long long c_end = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = c_end - c_start;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 256;
  __half* x; long long* clk; float* sink;
  cudaMalloc(&x, (size_t)N * S * S * C * 2);
  cudaMemset(x, 0, (size_t)N * S * S * C * 2);
  cudaMalloc(&clk, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 128 * sizeof(float));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap tm_row, tm_box;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)S * C * 2, (cuuint64_t)S * S * C * 2};
  cuuint32_t es[4] = {1, 1, 1, 1};
  cuuint32_t box_row[4] = {8, (cuuint32_t)S, 1, 1}, box_all[4] = {8, (cuuint32_t)S, (cuuint32_t)S, 1};
  for (int l2 = 0; l2 < 2; ++l2) {
    const CUtensorMapL2promotion prom = l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r1 = enc(&tm_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, gdim, gstr, box_row, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tm_box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, gdim, gstr, box_all, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    const int smem = 2 * (S * RS + 256) + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; ++mode) {
      const int tiles_per_cta = 24;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 128, smem>>>(tm_row, tm_box, mode, tiles_per_cta, C / 8, N, clk, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("L2 promotion %s, %s: %lld clocks per 36 KB tile per SM (max over CTAs; 148 CTAs, double-buffered)  [%s]\n", l2 ? "128B" : "none",
             mode == 0 ? "48 row boxes {8,48,1}" : "one box {8,48,48}", mx / tiles_per_cta, cudaGetErrorString(e));
    }
  }
  return 0;
}
