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
  p = p >= n ? 2 * n - 2 - p : p;
  // window taps more than one reflection away (n = 2, 3) only feed outputs outside the image; keep their address in range
  return min(max(p, 0), n - 1);
}

// thread = a 2 x 2 quad of output pixels (all C channels): the 7 x 7 input window is loaded once (49 instead of 100 pixel
// loads per quad) and the horizontal sums of a row serve both output rows that use it
template <typename T, int C>
__global__ void __launch_bounds__(256) pyrdown_kernel(const T* __restrict__ src, int N, int H, int W, T* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const int oh = (H + 1) >> 1, ow = (W + 1) >> 1;
  const int qh = (oh + 1) >> 1, qw = (ow + 1) >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * qh * qw) return;
  const int qx = (int)(idx % qw), qy = (int)((idx / qw) % qh), n = (int)(idx / ((long long)qw * qh));
  int xs[7], ys[7];
#pragma unroll
  for (int d = 0; d < 7; ++d) { xs[d] = refl101(4 * qx + d - 2, W); ys[d] = refl101(4 * qy + d - 2, H); }
  const T* img = src + (size_t)n * H * W * C;
  using Acc = typename std::conditional<std::is_same<T, float>::value, float, int>::type;
  Acc hs[7][2][C];
  const bool interior = 4 * qx - 2 >= 0 && 4 * qx + 4 < W;     // the seven columns are contiguous in memory
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const T* p = img + (size_t)ys[r] * W * C;
    Acc v[7][C];
    bool done = false;
    if constexpr (std::is_same<T, uint8_t>::value && C == 3) {
      // 21 contiguous bytes starting at an even offset: ten 16-bit loads + one byte instead of 21 byte loads
      const uint8_t* q = p + (4 * qx - 2) * 3;
      if (interior && (reinterpret_cast<uintptr_t>(q) & 1) == 0) {
        uint32_t b[21];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const uint32_t hw = *reinterpret_cast<const unsigned short*>(q + 2 * k);
          b[2 * k] = hw & 0xffu; b[2 * k + 1] = hw >> 8;
        }
        b[20] = q[20];
#pragma unroll
        for (int d = 0; d < 7; ++d)
#pragma unroll
          for (int c = 0; c < 3; ++c) v[d][c] = (Acc)b[d * 3 + c];
        done = true;
      }
    }
    if constexpr (std::is_same<T, float>::value && C == 1) {
      const float* q = p + (4 * qx - 2);
      if (interior && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {     // three 8-byte loads + one float instead of seven loads
        const float2 a = *reinterpret_cast<const float2*>(q), b2 = *reinterpret_cast<const float2*>(q + 2),
                     c2 = *reinterpret_cast<const float2*>(q + 4);
        v[0][0] = a.x; v[1][0] = a.y; v[2][0] = b2.x; v[3][0] = b2.y; v[4][0] = c2.x; v[5][0] = c2.y; v[6][0] = q[6];
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int d = 0; d < 7; ++d)
#pragma unroll
        for (int c = 0; c < C; ++c) v[d][c] = (Acc)p[xs[d] * C + c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      hs[r][0][c] = v[2][c] * 6 + (v[1][c] + v[3][c]) * 4 + v[0][c] + v[4][c];
      hs[r][1][c] = v[4][c] * 6 + (v[3][c] + v[5][c]) * 4 + v[2][c] + v[6][c];
    }
  }
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int oy = 2 * qy + dy, ox = 2 * qx + dx;
      if (oy >= oh || ox >= ow) continue;
      T* o = dst + ((size_t)n * oh * ow + (size_t)oy * ow + ox) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const Acc s = hs[2 * dy + 2][dx][c] * 6 + (hs[2 * dy + 1][dx][c] + hs[2 * dy + 3][dx][c]) * 4 + hs[2 * dy][dx][c] + hs[2 * dy + 4][dx][c];
        if constexpr (std::is_same<T, float>::value) o[c] = s * (1.f / 256.f);
        else o[c] = (T)((s + 128) >> 8);
      }
    }
}

// 3 x 3 coarse neighbourhood of coarse pixel (cy, cx) with pyrUp's border rule (index -1 -> 1, index n -> n-1), as float
template <typename T, int C>
__device__ __forceinline__ void load_nbhd(const T* __restrict__ img, int h, int w, int cy, int cx, float (&v)[3][3][C]) {
  const int ys[3] = {cy > 0 ? cy - 1 : min(1, h - 1), cy, min(cy + 1, h - 1)};
  const int xs[3] = {cx > 0 ? cx - 1 : min(1, w - 1), cx, min(cx + 1, w - 1)};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int c = 0; c < C; ++c) v[r][d][c] = (float)img[((size_t)ys[r] * w + xs[d]) * C + c];
}

// pyrUp values of the 2 x 2 fine pixels under coarse pixel (cy, cx): per axis even = (x[-1] + 6 x[0] + x[+1]) / 8,
// odd = (x[0] + x[+1]) / 2; rows first, then columns (same order for every output: deterministic)
template <int C>
__device__ __forceinline__ void pyrup_quad(const float (&v)[3][3][C], float (&q)[2][2][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float e[3], o[3];                 // horizontal pass per coarse row: even / odd fine column
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      e[r] = fmaf(6.f, v[r][1][c], v[r][0][c] + v[r][2][c]);
      o[r] = 4.f * (v[r][1][c] + v[r][2][c]);
    }
    q[0][0][c] = fmaf(6.f, e[1], e[0] + e[2]) * (1.f / 64.f);
    q[0][1][c] = fmaf(6.f, o[1], o[0] + o[2]) * (1.f / 64.f);
    q[1][0][c] = 4.f * (e[1] + e[2]) * (1.f / 64.f);
    q[1][1][c] = 4.f * (o[1] + o[2]) * (1.f / 64.f);
  }
}

// coarsest level: thread = one pixel, out = A * m + B * (1 - m)
template <int C>
__global__ void __launch_bounds__(256) blend_top_kernel(const uint8_t* __restrict__ a_f, const uint8_t* __restrict__ b_f,
                                                        const float* __restrict__ m_f, long long total, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const float gm = m_f[pix], gi = 1.0f - gm;
#pragma unroll
  for (int c = 0; c < C; ++c)
    out[pix * C + c] = __fadd_rn(__fmul_rn((float)a_f[pix * C + c], gm), __fmul_rn((float)b_f[pix * C + c], gi));
}

// thread = one COARSE pixel = a 2 x 2 quad of fine pixels: the 3 x 3 coarse neighbourhoods of A, B and the running
// reconstruction are loaded once and serve all four outputs (9 instead of 25 coarse loads per array and quad)
template <int C>
__global__ void __launch_bounds__(256) blend_level_kernel(const float* __restrict__ coarse, const uint8_t* __restrict__ a_f,
                                                          const uint8_t* __restrict__ b_f, const float* __restrict__ m_f,
                                                          const uint8_t* __restrict__ a_c, const uint8_t* __restrict__ b_c,
                                                          int N, int h, int w, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int ch = h >> 1, cw = w >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * ch * cw) return;
  const int cx = (int)(idx % cw), cy = (int)((idx / cw) % ch), n = (int)(idx / ((long long)cw * ch));
  const size_t cimg = (size_t)n * ch * cw * C;
  float v[3][3][C], ua[2][2][C], ub[2][2][C], up[2][2][C];
  load_nbhd<uint8_t, C>(a_c + cimg, ch, cw, cy, cx, v);
  pyrup_quad<C>(v, ua);
  load_nbhd<uint8_t, C>(b_c + cimg, ch, cw, cy, cx, v);
  pyrup_quad<C>(v, ub);
  load_nbhd<float, C>(coarse + cimg, ch, cw, cy, cx, v);
  pyrup_quad<C>(v, up);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const size_t pix = (size_t)n * h * w + (size_t)(2 * cy + dy) * w + (2 * cx + dx);
      const float gm = m_f[pix], gi = 1.0f - gm;
#pragma unroll
      for (int c = 0; c < C; ++c) {   // the reference's order: ls = la*gm + lb*(1-gm) (separately rounded), then pyrUp(ls_) + ls
        const float la = (float)a_f[pix * C + c] - ua[dy][dx][c], lb = (float)b_f[pix * C + c] - ub[dy][dx][c];
        up[dy][dx][c] = __fadd_rn(up[dy][dx][c], __fadd_rn(__fmul_rn(la, gm), __fmul_rn(lb, gi)));
      }
    }
  // the two pixels of a quad row are adjacent: 2 * C floats per row, 8-byte aligned when C is even or the pixel index is even
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    float* o = out + ((size_t)n * h * w + (size_t)(2 * cy + dy) * w + 2 * cx) * C;
    if ((reinterpret_cast<uintptr_t>(o) & 7) == 0 && (2 * C) % 2 == 0) {
      const float* src = &up[dy][0][0];
#pragma unroll
      for (int k = 0; k < C; ++k) reinterpret_cast<float2*>(o)[k] = make_float2(src[2 * k], src[2 * k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < 2 * C; ++k) o[k] = (&up[dy][0][0])[k];
    }
  }
}

}  // namespace s2v

using namespace s2v;

template <typename T>
static int launch_pyrdown(const T* src, int N, int H, int W, int C, T* dst, cudaStream_t st) {
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  const long long total = (long long)N * ((oh + 1) / 2) * ((ow + 1) / 2);       // one thread per 2 x 2 output quad
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
  cudaStream_t st = (cudaStream_t)stream;
  if (!coarse_out) {
    const long long total = (long long)N * h * w;
    const int grid = ceil_div(total, 256);
    switch (C) {
      case 1: launch_pdl(blend_top_kernel<1>, grid, 256, 0, st, a_fine, b_fine, m_fine, total, out); break;
      case 3: launch_pdl(blend_top_kernel<3>, grid, 256, 0, st, a_fine, b_fine, m_fine, total, out); break;
      case 4: launch_pdl(blend_top_kernel<4>, grid, 256, 0, st, a_fine, b_fine, m_fine, total, out); break;
      default: return S2V_EINVAL;
    }
  } else {
    const int grid = ceil_div((long long)N * (h / 2) * (w / 2), 256);
    switch (C) {
      case 1: launch_pdl(blend_level_kernel<1>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
      case 3: launch_pdl(blend_level_kernel<3>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
      case 4: launch_pdl(blend_level_kernel<4>, grid, 256, 0, st, coarse_out, a_fine, b_fine, m_fine, a_coarse, b_coarse, N, h, w, out); break;
      default: return S2V_EINVAL;
    }
  }
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
