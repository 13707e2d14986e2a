// ENet upsampler (reference: models/ENet.py:82-139, models/base_blocks.py:460-553) - the memory-bound pieces around the
// tcgen05 convs.  StyleGAN2's per-sample modulated conv (ModulatedConv2d, :487-508: a grouped conv with B x Cout x Cin x k x k
// weights) is run as ONE shared-weight conv_tc GEMM by moving the per-sample factors out of the weights:
//     conv(x, W * s[n,ci] * d[n,co])  ==  d[n,co] * conv(x * s[n,ci], W)
//   s = modulation Linear(style)                       (s2v_grouped_linear)
//   d = rsqrt(sum_ci s^2 * sum_k W^2 + eps)            (style_demod_kernel; sum_k W^2 is folded at load)
//   x * s                                              (fused into the bilinear x2 pass / the previous layer's epilogue)
//   d * (.) * sqrt2 + noise_w * noise + bias, LeakyReLU (style_epilogue_kernel, which also applies the NEXT layer's s)
// ToRGB (:539-553, Cout = 3, 1x1, no demodulation) is a per-pixel dot product fused with the bilinear x2 skip and the crop.
#include "common.cuh"

namespace s2v {

// out[n][co] = gain * rsqrt(sum_ci w2[co][ci] * s[n][ci]^2 + eps)
__global__ void __launch_bounds__(128) style_demod_kernel(const float* __restrict__ w2, const float* __restrict__ s, long long s_stride,
                                                         int cin, int cout, float eps, float gain, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int co = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (co >= cout) return;
  const float* sr = s + (size_t)n * s_stride;
  const float* wr = w2 + (size_t)co * cin;
  float acc = 0.f;
  for (int ci = 0; ci < cin; ++ci) acc = fmaf(wr[ci], sr[ci] * sr[ci], acc);
  out[(size_t)n * cout + co] = gain * rsqrtf(acc + eps);
}

// y = lrelu(x * a[n][c] + bias[c] + noise_w * noise[n][h][w], slope) * post[n][c]      (fp16 NHWC in / out, 8 channels per thread)
__global__ void __launch_bounds__(256) style_epilogue_kernel(View x, const float* __restrict__ a, const float* __restrict__ bias,
                                                            const float* __restrict__ noise, const float* __restrict__ noise_w, float slope,
                                                            const float* __restrict__ post, long long post_stride, View y) {
  pdl_trigger();
  pdl_wait();
  const int C8 = x.c >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)x.n * x.h * x.w * C8;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  const long long pix = idx / C8;
  const int ox = (int)(pix % x.w), oy = (int)((pix / x.w) % x.h), n = (int)(pix / ((long long)x.w * x.h));
  float v[8];
  h8_to_f(ld_h8(x.p + n * x.sn + oy * x.sh + ox * x.sw + c8 * 8), v);
  const float nz = noise ? noise_w[0] * noise[((size_t)n * x.h + oy) * x.w + ox] : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    float t = v[i] * (a ? a[(size_t)n * x.c + c] : 1.f) + (bias ? bias[c] : 0.f) + nz;
    t = t > 0.f ? t : t * slope;
    v[i] = post ? t * post[(size_t)n * post_stride + c] : t;
  }
  st_h8(y.p + n * y.sn + oy * y.sh + ox * y.sw + c8 * 8, f_to_h8(v));
}

// ToRGB: out[n][k][oy][ox] = sum_c x[n][oy+crop][ox+crop][c] * w[k][c] * s[n][c] + bias[k] + bilinear_x2(skip)[n][k][oy+crop][ox+crop]
// skip: float32 NCHW [N][3][H/2][W/2] (F.interpolate(scale_factor=2, bilinear, align_corners=False)); out: float32 NCHW
// [N][3][H-2crop][W-2crop].  A warp handles one pixel per lane; the per-sample folded weights live in shared memory.
constexpr int kRgbMaxC = 512;
__global__ void __launch_bounds__(256) to_rgb_kernel(View x, const float* __restrict__ w, const float* __restrict__ s, long long s_stride,
                                                    const float* __restrict__ bias, const float* __restrict__ skip, float* __restrict__ out,
                                                    int crop) {
  pdl_trigger();
  pdl_wait();
  __shared__ float wm[3][kRgbMaxC];
  const int n = blockIdx.y, C = x.c;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    wm[k][c] = w[i] * s[(size_t)n * s_stride + c];
  }
  __syncthreads();
  const int OH = x.h - 2 * crop, OW = x.w - 2 * crop;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= OH * OW) return;
  const int oy = pix / OW, ox = pix - oy * OW;
  const int yy = oy + crop, xx = ox + crop;
  const __half* px = x.p + n * x.sn + yy * x.sh + xx * x.sw;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  for (int c = 0; c < C; c += 8) {
    float v[8];
    h8_to_f(ld_h8(px + c), v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc0 = fmaf(v[i], wm[0][c + i], acc0);
      acc1 = fmaf(v[i], wm[1][c + i], acc1);
      acc2 = fmaf(v[i], wm[2][c + i], acc2);
    }
  }
  float r[3] = {acc0 + bias[0], acc1 + bias[1], acc2 + bias[2]};
  if (skip) {
    const int sh = x.h >> 1, sw = x.w >> 1;
    const float sy = fmaxf(0.5f * ((float)yy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * ((float)xx + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)sy, sh - 1), x0 = min((int)sx, sw - 1);
    const int y1 = y0 + (y0 < sh - 1 ? 1 : 0), x1 = x0 + (x0 < sw - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* sp = skip + ((size_t)n * 3 + k) * sh * sw;
      const float top = sp[y0 * sw + x0] + lx * (sp[y0 * sw + x1] - sp[y0 * sw + x0]);
      const float bot = sp[y1 * sw + x0] + lx * (sp[y1 * sw + x1] - sp[y1 * sw + x0]);
      r[k] += top + ly * (bot - top);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) out[(((size_t)n * 3 + k) * OH + oy) * OW + ox] = r[k];
}

// F.pad(x, (p,p,p,p), 'reflect') of a float32 NCHW tensor (models/ENet.py:118-119)
__global__ void __launch_bounds__(256) reflect_pad_nchw_kernel(const float* __restrict__ src, int planes, int H, int W, int pad,
                                                              float* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const int PH = H + 2 * pad, PW = W + 2 * pad;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)planes * PH * PW) return;
  const int x = (int)(idx % PW), y = (int)((idx / PW) % PH), p = (int)(idx / ((long long)PW * PH));
  int sy = y - pad, sx = x - pad;
  sy = sy < 0 ? -sy : (sy >= H ? 2 * H - 2 - sy : sy);
  sx = sx < 0 ? -sx : (sx >= W ? 2 * W - 2 - sx : sx);
  dst[idx] = src[((size_t)p * H + sy) * W + sx];
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_style_demod(const float* w2, const float* s, int64_t s_stride, int N, int cin, int cout, float eps, float gain,
                               float* out, void* stream) {
  if (N == 0) return S2V_OK;
  if (!w2 || !s || !out || N < 0 || N > 65535 || cin <= 0 || cout <= 0) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(style_demod_kernel, dim3(ceil_div(cout, 128), N), 128, 0, (cudaStream_t)stream, w2, s, (long long)s_stride, cin, cout,
                          eps, gain, out));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_style_epilogue(const s2v_view* x, const float* a, const float* bias, const float* noise, const float* noise_w,
                                  float slope, const float* post, int64_t post_stride, const s2v_view* y, void* stream) {
  if (!view_ok(x) || !view_ok(y) || x->n != y->n || x->h != y->h || x->w != y->w || x->c != y->c) return S2V_EINVAL;
  if (noise && !noise_w) return S2V_EINVAL;
  const long long total = (long long)x->n * x->h * x->w * (x->c >> 3);
  S2V_CUDA_TRY(launch_pdl(style_epilogue_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, mk(x), a, bias, noise, noise_w, slope, post,
                          (long long)post_stride, mk(y)));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_to_rgb(const s2v_view* x, const float* w, const float* s, int64_t s_stride, const float* bias, const float* skip,
                          float* out, int crop, void* stream) {
  if (!view_ok(x) || !w || !s || !bias || !out || x->c > kRgbMaxC || crop < 0 || 2 * crop >= x->h || 2 * crop >= x->w || x->n > 65535)
    return S2V_EINVAL;
  if (skip && ((x->h & 1) || (x->w & 1))) return S2V_EINVAL;
  const int opix = (x->h - 2 * crop) * (x->w - 2 * crop);
  S2V_CUDA_TRY(launch_pdl(to_rgb_kernel, dim3(ceil_div(opix, 256), x->n), 256, 0, (cudaStream_t)stream, mk(x), w, s, (long long)s_stride, bias,
                          skip, out, crop));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_reflect_pad_nchw_f32(const float* src, int planes, int H, int W, int pad, float* dst, void* stream) {
  if (planes == 0) return S2V_OK;
  if (!src || !dst || planes < 0 || H <= 0 || W <= 0 || pad < 0 || pad >= H || pad >= W) return S2V_EINVAL;
  const long long total = (long long)planes * (H + 2 * pad) * (W + 2 * pad);
  S2V_CUDA_TRY(launch_pdl(reflect_pad_nchw_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, src, planes, H, W, pad, dst));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
