// Development micro-benchmark: latency / throughput of mma.sync.m16n8k16 (fp16 in, fp32 accumulate) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmp/mb_hmma tools/mb_hmma.cu && tools/tmp/mb_hmma
// ILP independent accumulator chains per warp, W warps per SM sub-partition: clocks per MMA per warp and per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int ILP>
__global__ void k(float* out, long long* clk, int iters) {
  float acc[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 0x3c003c00, a3 = 0x3c003c00, b0 = 0x3c003c00, b1 = 0x38003800;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma16816(acc[i], a0, a1, a2, a3, b0, b1);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm) {
  float* out; long long* clk;
  const int blocks = 148, threads = warps_per_sm * 32, iters = 2048;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&clk, sizeof(long long) * blocks);
  k<ILP><<<blocks, threads>>>(out, clk, iters);
  k<ILP><<<blocks, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double c = (double)h[0];
  const double per_warp = c / ((double)iters * ILP);
  const double per_smsp = c / ((double)iters * ILP * (warps_per_sm / 4.0));
  printf("ILP %2d  warps/SM %2d : %6.1f clk per MMA per warp, %5.2f clk per MMA per SMSP (%s)\n", ILP, warps_per_sm, per_warp, per_smsp,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<1>(w); run<2>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); run<16>(w);
  }
  return 0;
}
