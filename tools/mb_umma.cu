// Development micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, SWIZZLE_128B K-major)
// as a function of N, of the A descriptor's stride-byte-offset / start alignment, and of how many K-steps share one
// 128-byte swizzle row.  One CTA per SM, one issuing thread, fixed (garbage) shared-memory operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmp/mb_umma tools/mb_umma.cu && tools/tmp/mb_umma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: same A/B descriptors every MMA; mode 1: A start walks 4 K-steps (+32 B) then next "tap" (+128 B), B walks K-steps and taps
__global__ void __launch_bounds__(128, 1) k(int n, uint32_t sbo, uint32_t a_off, int rounds, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  const uint32_t barp = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint64_t a0 = desc(base + a_off, sbo), b0 = desc(base + 64 * 1024, 1024);
    const uint32_t bstep = (uint32_t)(n * 128) >> 4;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (mode == 0) {
#pragma unroll
        for (int u = 0; u < 36; ++u) mma(tmem, a0, b0, idesc, 1u);
      } else {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma(tmem, a0 + (uint64_t)((tap / 3) * (sbo >> 4) + (tap % 3) * 8 + ks * 2), b0 + (uint64_t)(tap * bstep + ks * 2), idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barp) : "memory");
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  const int rounds = 64;
  printf("%5s %6s %6s %5s %12s\n", "N", "SBO", "a_off", "mode", "cyc/MMA");
  for (int mode = 0; mode < 2; ++mode)
    for (uint32_t sbo : {1024u, 1280u})
      for (uint32_t a_off : {0u, 128u})
        for (int n : ns) {
          if (mode == 1 && a_off) continue;
          if (mode == 0 && sbo == 1280u && a_off == 0) continue;
          for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 200 * 1024>>>(n, sbo, a_off, rounds, mode, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148; ++i) s += (double)h[i];
          printf("%5d %6u %6u %5d %12.1f\n", n, sbo, a_off, mode, s / 148 / (rounds * 36.0));
        }
  return 0;
}
