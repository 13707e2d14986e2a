// Grouped small linears: every AdaIN gamma/beta head of a net in ONE launch
// (reference: models/base_blocks.py:136-141,149-151 - 108 ADAIN instances in LNet,
// each Linear(128 -> C) x2 on its own ReLU(Linear(z)) hidden vector).
//   out[b][g.out_off + j] = bias_g[j] + sum_k hidden[b][g.in_off + k] * wt_g[k][j]
// grid (tiles, ceil(B/8)); a tile = 128 consecutive outputs of one group; thread = output j
// (coalesced transposed-weight reads), 8 batch rows per thread (hidden broadcast from smem).
#include "common.cuh"

namespace s2v {

constexpr int kLinTile = 128, kLinRows = 8, kLinMaxK = 512;

__global__ void __launch_bounds__(kLinTile) grouped_linear_kernel(const __half* __restrict__ hidden,
                                                                 long long hidden_stride, int B,
                                                                 const s2v_lin_group* __restrict__ groups,
                                                                 const int* __restrict__ tile2group,
                                                                 float* __restrict__ out, long long out_stride) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[kLinRows][kLinMaxK];
  s2v_lin_group g = groups[tile2group[2 * blockIdx.x]];
  if (g.k > kLinMaxK) g.k = kLinMaxK;        // never past the staging buffer; the host wrapper rejects k > 512 (ops.pack_lin_groups)
  const int j = tile2group[2 * blockIdx.x + 1] + threadIdx.x;
  const int b0 = blockIdx.y * kLinRows;
  for (int i = threadIdx.x; i < kLinRows * g.k; i += kLinTile) {
    const int r = i / g.k, k = i - r * g.k;
    sh[r][k] = (b0 + r < B) ? __half2float(hidden[(size_t)(b0 + r) * hidden_stride + g.in_off + k]) : 0.f;
  }
  __syncthreads();
  if (j >= g.nout) return;
  float acc[kLinRows];
  const float bj = g.bias ? g.bias[j] : 0.f;
#pragma unroll
  for (int r = 0; r < kLinRows; ++r) acc[r] = bj;
  for (int k = 0; k < g.k; ++k) {
    const float w = g.wt[(size_t)k * g.nout + j];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = fmaf(sh[r][k], w, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kLinRows; ++r)
    if (b0 + r < B) out[(size_t)(b0 + r) * out_stride + g.out_off + j] = acc[r];
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_grouped_linear(const void* hidden_f16, int64_t hidden_stride, int B, const s2v_lin_group* groups_dev,
                                  const int32_t* tile2group_dev, int n_tiles, float* out, int64_t out_stride,
                                  void* stream) {
  if (B == 0 || n_tiles == 0) return S2V_OK;
  if (!hidden_f16 || !groups_dev || !tile2group_dev || !out || B < 0 || n_tiles < 0) return S2V_EINVAL;
  launch_pdl(grouped_linear_kernel, dim3(n_tiles, ceil_div(B, kLinRows)), kLinTile, 0, (cudaStream_t)stream, 
      (const __half*)hidden_f16, hidden_stride, B, groups_dev, tile2group_dev, out, out_stride);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
