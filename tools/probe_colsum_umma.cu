// Development probe (correctness, not speed): can tcgen05.mma take the conv epilogue's STAGED fp16 tile - [128 pixel rows][64
// channels] panels, 128-byte rows, SWIZZLE_128B (16-byte chunk j of row r stored at chunk j ^ (r & 7)) - as an MN-MAJOR
// operand, so that the per-channel statistics become tensor-core work?
//   column sums      D[m][c]   = sum_r ones[m][r] * S[r][c]        A = ones (K-major), B = S (MN-major: N = channels, K = rows)
//   sums of squares  G[c1][c2] = sum_r S[r][c1] * S[r][c2]         A = S (MN-major: M = channels), B = S (MN-major); diagonal
// The probe sweeps the descriptor fields that differ from the K-major case (leading / stride byte offsets) and reports
// the max-abs error of both results against the host.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmp/probe_colsum tools/probe_colsum_umma.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);             // version 1, SWIZZLE_128B
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// tile: [2 panels][128 rows][64 halves] already swizzled on the host; ones: 16 KB of 1.0h
// out: [128 lanes][256 floats]: columns 0..127 = column-sum accumulator (N = 128), 128..255 = Gram accumulator
__global__ void __launch_bounds__(128, 1) probe(const __half* tile, uint32_t lbo_b, uint32_t sbo_b, uint32_t kstep_b,
                                                uint32_t lbo_a, uint32_t sbo_a, uint32_t kstep_a, float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* al = raw + (base - smem_u32(raw));
  __half* s_tile = reinterpret_cast<__half*>(al);                 // 32 KB: two panels
  __half* s_ones = reinterpret_cast<__half*>(al + 32 * 1024);     // 16 KB
  for (int i = threadIdx.x; i < 2 * 128 * 64; i += 128) s_tile[i] = tile[i];
  for (int i = threadIdx.x; i < 128 * 64; i += 128) s_ones[i] = __float2half(1.f);
  const uint32_t barp = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (threadIdx.x == 0) {
    // column sums: A = ones K-major (SBO 1024), B = tile MN-major, N = 128 (two panels), K = 128 rows in 8 steps of 16
    const uint32_t idesc_sum = (1u << 4) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a1 = desc(base + 32 * 1024, 0, 1024);
    for (int ks = 0; ks < 8; ++ks)
      mma(tmem, a1 + (uint64_t)(ks & 3) * 2, desc(base + ks * kstep_b, lbo_b, sbo_b), idesc_sum, ks ? 1u : 0u);
    // Gram: A = tile MN-major (M = 128 channels), B = tile MN-major (N = 128)
    const uint32_t idesc_g = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < 8; ++ks)
      mma(tmem + 128, desc(base + ks * kstep_a, lbo_a, sbo_a), desc(base + ks * kstep_b, lbo_b, sbo_b), idesc_g, ks ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
  }
  uint32_t ok = 0;
  unsigned spins = 0;
  while (!ok && ++spins < (1u << 24))
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barp) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 256; c0 += 16) {
    uint32_t v[16];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(size_t)(warp * 32 + lane) * 256 + c0 + i] = ok ? __uint_as_float(v[i]) : -12345.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  const int R = 128, C = 128;
  static float S[R][C];
  srand(1);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) S[r][c] = (float)((rand() % 9) - 4);          // small integers: exact in fp16 and in the sums
  static __half h[2 * 128 * 64];
  for (int p = 0; p < 2; ++p)
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < 64; ++c) {
        const int chunk = c >> 3, pos = chunk ^ (r & 7);
        h[(size_t)p * 128 * 64 + r * 64 + pos * 8 + (c & 7)] = __float2half(S[r][p * 64 + c]);
      }
  double colsum[C], sq[C];
  for (int c = 0; c < C; ++c) { colsum[c] = sq[c] = 0; for (int r = 0; r < R; ++r) { colsum[c] += S[r][c]; sq[c] += S[r][c] * S[r][c]; } }
  __half* dt; float* dout;
  cudaMalloc(&dt, sizeof(h)); cudaMalloc(&dout, 128 * 256 * sizeof(float));
  cudaMemcpy(dt, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  static float out[128 * 256];
  printf("%8s %8s %8s | %12s %12s %12s\n", "LBO", "SBO", "kstep", "colsum err", "gram-diag err", "gram offdiag");
  const uint32_t lbos[] = {16384, 1024, 2048};
  const uint32_t sbos[] = {1024, 2048, 16384};
  const uint32_t ksteps[] = {2048};     // 16 pixel rows of 128 bytes per K step
  for (uint32_t ks : ksteps)
    for (uint32_t lbo : lbos)
      for (uint32_t sbo : sbos) {
        probe<<<1, 128, 64 * 1024>>>(dt, lbo, sbo, ks, lbo, sbo, ks, dout);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed (%u %u %u): %s\n", lbo, sbo, ks, cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, e3 = 0;
        for (int m = 0; m < 128; m += 37)
          for (int c = 0; c < C; ++c) e1 = fmax(e1, fabs(out[m * 256 + c] - colsum[c]));
        for (int c = 0; c < C; ++c) e2 = fmax(e2, fabs(out[c * 256 + 128 + c] - sq[c]));
        for (int c = 0; c < C; c += 5) {                 // one off-diagonal entry per row as well
          const int c2 = (c * 7 + 3) % C;
          double g = 0;
          for (int r = 0; r < R; ++r) g += S[r][c] * S[r][c2];
          e3 = fmax(e3, fabs(out[c * 256 + 128 + c2] - g));
        }
        printf("%8u %8u %8u | %12.3f %12.3f %12.3f %s\n", lbo, sbo, ks, e1, e2, e3, (e1 == 0 && e2 == 0 && e3 == 0) ? "<== exact" : "");
      }
  return 0;
}
