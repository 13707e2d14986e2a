// Layout converters at the module boundary: NCHW float32 (the reference's tensors,
// inference.py:260-262) <-> fp16 channels-last views used between kernels.
#include "common.cuh"

namespace s2v {

// one thread per (n, y, x): reads C strided floats (coalesced along x per channel),
// writes c_fill contiguous halves.
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, int N, int C, int H, int W,
                                                   long long src_sn, View d, int c_off, int c_fill, float scale,
                                                   float shift) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int n = (int)(idx / ((long long)W * H));
  __half* o = d.p + n * d.sn + y * d.sh + x * d.sw + c_off;
  const float* s = src + (size_t)n * src_sn + (size_t)y * W + x;
  const size_t plane = (size_t)H * W;
  for (int c = 0; c < c_fill; ++c) o[c] = __float2half_rn(c < C ? fmaf(s[c * plane], scale, shift) : 0.f);
}

__global__ void __launch_bounds__(256) unpack_kernel(View s, int c_off, int C, float* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)s.n * s.h * s.w;
  if (idx >= total) return;
  const int x = (int)(idx % s.w);
  const int y = (int)((idx / s.w) % s.h);
  const int n = (int)(idx / ((long long)s.w * s.h));
  const __half* p = s.p + n * s.sn + y * s.sh + x * s.sw + c_off;
  const size_t plane = (size_t)s.h * s.w;
  float* o = dst + ((size_t)n * C * s.h + y) * s.w + x;
  for (int c = 0; c < C; ++c) o[c * plane] = __half2float(p[c]);
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_pack_nchw_f32(const float* src, int N, int C, int H, int W, int64_t src_sn, const s2v_view* dst,
                                 int c_off, int c_fill, float scale, float shift, void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !view_ok(dst) || C <= 0 || c_fill < C || c_off < 0 || c_off + c_fill > dst->c) return S2V_EINVAL;
  if (dst->n < N || dst->h != H || dst->w != W || src_sn < (int64_t)C * H * W) return S2V_EINVAL;
  const long long total = (long long)N * H * W;
  pack_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, N, C, H, W, src_sn, mk(dst), c_off, c_fill, scale, shift);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_unpack_to_nchw_f32(const s2v_view* src, int c_off, int C, float* dst, void* stream) {
  if (!view_ok(src) || !dst || C <= 0 || c_off < 0 || c_off + C > src->c) return S2V_EINVAL;
  const long long total = (long long)src->n * src->h * src->w;
  unpack_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(src), c_off, C, dst);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
