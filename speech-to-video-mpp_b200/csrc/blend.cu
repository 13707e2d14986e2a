// Laplacian-pyramid blend of the generated face into the frame (reference: futils/inference_utils.py:181-222
// Laplacian_Pyramid_Blending_with_mask, called per frame at inference.py:312 on 512 x 512 uint8 images with a float32
// mask and 10 levels; the arithmetic is OpenCV's cv2.pyrDown / cv2.pyrUp / cv2.add on the CPU).  Batched over frames,
// images channels-last ([N,H,W,C], C <= 4) exactly as cv2 holds them:
//   pyrdown_u8   cv2.pyrDown of 8-bit images: [1 4 6 4 1]^2 / 256 on every second pixel, BORDER_REFLECT_101, integer
//                arithmetic with (sum + 128) >> 8 - bit-exact.  The Gaussian pyramids of A and B stay uint8 (the reference
//                converts each level to float32 AFTER the 8-bit pyrDown, so nothing is lost).
//   pyrdown_f32  the same kernel in float32 (the mask pyramid).
//   blend_level  one level of the collapse, fused:  out = pyrUp(coarse_out) + (A_f - pyrUp(A_c)) * m + (B_f - pyrUp(B_c)) * (1 - m)
//                i.e. both Laplacian levels, the mask blend and the reconstruction step in one pass (the Laplacian
//                pyramids and the per-level blends of the reference are never materialised); the coarsest level is
//                out = A_c * m + B_c * (1 - m).  pyrUp = zero insertion + the same kernel x 4: even outputs
//                (x[i-1] + 6 x[i] + x[i+1]) / 8, odd outputs (x[i] + x[i+1]) / 2 per axis, index -1 -> 1, index n -> n-1.
// Memory-bound integer/byte + float32 work; no tensor cores.
#include <type_traits>

#include "common.cuh"

namespace s2v {

__device__ __forceinline__ int refl101(int p, int n) {
  if (n == 1) return 0;
  p = p < 0 ? -p : p;
  return p >= n ? 2 * n - 2 - p : p;
}

// thread = one output pixel (all C channels)
template <typename T, int C>
__global__ void __launch_bounds__(256) pyrdown_kernel(const T* __restrict__ src, int N, int H, int W, T* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const int oh = (H + 1) >> 1, ow = (W + 1) >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * oh * ow) return;
  const int ox = (int)(idx % ow), oy = (int)((idx / ow) % oh), n = (int)(idx / ((long long)ow * oh));
  int xs[5], ys[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) { xs[d] = refl101(2 * ox + d - 2, W); ys[d] = refl101(2 * oy + d - 2, H); }
  const T* img = src + (size_t)n * H * W * C;
  using Acc = typename std::conditional<std::is_same<T, float>::value, float, int>::type;
  Acc row[5][C];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const T* p = img + (size_t)ys[r] * W * C;
    Acc v[5][C];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
      for (int c = 0; c < C; ++c) v[d][c] = (Acc)p[xs[d] * C + c];
#pragma unroll
    for (int c = 0; c < C; ++c) row[r][c] = v[2][c] * 6 + (v[1][c] + v[3][c]) * 4 + v[0][c] + v[4][c];
  }
  T* o = dst + ((size_t)n * oh * ow + (size_t)oy * ow + ox) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const Acc s = row[2][c] * 6 + (row[1][c] + row[3][c]) * 4 + row[0][c] + row[4][c];
    if constexpr (std::is_same<T, float>::value) o[c] = s * (1.f / 256.f);
    else o[c] = (T)((s + 128) >> 8);
  }
}

// pyrUp sample at fine position (y, x) of a coarse [h, w, C] image; U8: the coarse image is uint8
template <typename T, int C>
__device__ __forceinline__ void pyrup_at(const T* __restrict__ img, int h, int w, int y, int x, float* out) {
  const int cy = y >> 1, cx = x >> 1;
  // per axis: even -> taps (i-1, i, i+1) weights (1, 6, 1); odd -> taps (i, i+1) weights (4, 4); total weight 8
  int yi[3], xi[3];
  float wy[3], wx[3];
  if (y & 1) { yi[0] = cy; yi[1] = min(cy + 1, h - 1); yi[2] = cy; wy[0] = 4.f; wy[1] = 4.f; wy[2] = 0.f; }
  else { yi[0] = cy > 0 ? cy - 1 : min(1, h - 1); yi[1] = cy; yi[2] = min(cy + 1, h - 1); wy[0] = 1.f; wy[1] = 6.f; wy[2] = 1.f; }
  if (x & 1) { xi[0] = cx; xi[1] = min(cx + 1, w - 1); xi[2] = cx; wx[0] = 4.f; wx[1] = 4.f; wx[2] = 0.f; }
  else { xi[0] = cx > 0 ? cx - 1 : min(1, w - 1); xi[1] = cx; xi[2] = min(cx + 1, w - 1); wx[0] = 1.f; wx[1] = 6.f; wx[2] = 1.f; }
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (wy[r] == 0.f) continue;
    const T* p = img + (size_t)yi[r] * w * C;
    float rs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) rs[c] = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (wx[d] == 0.f) continue;
#pragma unroll
      for (int c = 0; c < C; ++c) rs[c] = fmaf(wx[d], (float)p[xi[d] * C + c], rs[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fmaf(wy[r], rs[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = acc[c] * (1.f / 64.f);
}

// thread = one fine pixel.  coarse == nullptr: coarsest level, out = A * m + B * (1 - m).
template <int C>
__global__ void __launch_bounds__(256) blend_level_kernel(const float* __restrict__ coarse, const uint8_t* __restrict__ a_f,
                                                          const uint8_t* __restrict__ b_f, const float* __restrict__ m_f,
                                                          const uint8_t* __restrict__ a_c, const uint8_t* __restrict__ b_c,
                                                          int N, int h, int w, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * h * w) return;
  const int x = (int)(idx % w), y = (int)((idx / w) % h), n = (int)(idx / ((long long)w * h));
  const size_t pix = (size_t)n * h * w + (size_t)y * w + x;
  const float gm = m_f[pix], gi = 1.0f - gm;
  float la[C], lb[C], up[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { la[c] = (float)a_f[pix * C + c]; lb[c] = (float)b_f[pix * C + c]; up[c] = 0.f; }
  if (coarse) {
    const int ch = h >> 1, cw = w >> 1;
    float ua[C], ub[C];
    pyrup_at<uint8_t, C>(a_c + (size_t)n * ch * cw * C, ch, cw, y, x, ua);
    pyrup_at<uint8_t, C>(b_c + (size_t)n * ch * cw * C, ch, cw, y, x, ub);
    pyrup_at<float, C>(coarse + (size_t)n * ch * cw * C, ch, cw, y, x, up);
#pragma unroll
    for (int c = 0; c < C; ++c) { la[c] -= ua[c]; lb[c] -= ub[c]; }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)       // the reference's order: ls = la*gm + lb*(1-gm) (separately rounded), then pyrUp(ls_) + ls
    out[pix * C + c] = __fadd_rn(up[c], __fadd_rn(__fmul_rn(la[c], gm), __fmul_rn(lb[c], gi)));
}

}  // namespace s2v

using namespace s2v;

template <typename T>
static int launch_pyrdown(const T* src, int N, int H, int W, int C, T* dst, cudaStream_t st) {
  const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2);
  const int grid = ceil_div(total, 256);
  switch (C) {
    case 1: launch_pdl(pyrdown_kernel<T, 1>, grid, 256, 0, st, src, N, H, W, dst); break;
    case 3: launch_pdl(pyrdown_kernel<T, 3>, grid, 256, 0, st, src, N, H, W, dst); break;
    case 4: launch_pdl(pyrdown_kernel<T, 4>, grid, 256, 0, st, src, N, H, W, dst); break;
    default: return S2V_EINVAL;
  }
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_pyrdown_u8(const uint8_t* src, int N, int H, int W, int C, uint8_t* dst, void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !dst || N < 0 || H <= 0 || W <= 0) return S2V_EINVAL;
  return launch_pyrdown<uint8_t>(src, N, H, W, C, dst, (cudaStream_t)stream);
}

extern "C" int s2v_pyrdown_f32(const float* src, int N, int H, int W, int C, float* dst, void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !dst || N < 0 || H <= 0 || W <= 0) return S2V_EINVAL;
  return launch_pyrdown<float>(src, N, H, W, C, dst, (cudaStream_t)stream);
}

extern "C" int s2v_lap_blend_level(const float* coarse_out, const uint8_t* a_fine, const uint8_t* b_fine, const float* m_fine,
                                   const uint8_t* a_coarse, const uint8_t* b_coarse, int N, int h, int w, int C, float* out,
                                   void* stream) {
  if (N == 0) return S2V_OK;
  if (!a_fine || !b_fine || !m_fine || !out || N < 0 || h <= 0 || w <= 0) return S2V_EINVAL;
  if (coarse_out && (!a_coarse || !b_coarse || (h & 1) || (w & 1))) return S2V_EINVAL;   // pyrUp doubles: fine = 2 x coarse
  const int grid = ceil_div((long long)N * h * w, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 1: launch_pdl(blend_level_kernel<1>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
    case 3: launch_pdl(blend_level_kernel<3>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
    case 4: launch_pdl(blend_level_kernel<4>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
    default: return S2V_EINVAL;
  }
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
