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
// grid (pixel-chunk blocks, N): a thread keeps ONE 8-channel slice for all its pixels, so the per-(n, c) tables are read once
__global__ void __launch_bounds__(256) style_epilogue_kernel(View x, const float* __restrict__ a, const float* __restrict__ bias,
                                                            const float* __restrict__ noise, const float* __restrict__ noise_w, float slope,
                                                            const float* __restrict__ post, long long post_stride, View y, int pix_per_block) {
  pdl_trigger();
  pdl_wait();
  const int C8 = x.c >> 3, n = blockIdx.y;
  const int c8 = threadIdx.x % C8, lane_pix = threadIdx.x / C8, pix_lanes = blockDim.x / C8;
  if (lane_pix >= pix_lanes) return;
  float av[8], bv[8], pv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    av[i] = a ? a[(size_t)n * x.c + c] : 1.f;
    bv[i] = bias ? bias[c] : 0.f;
    pv[i] = post ? post[(size_t)n * post_stride + c] : 1.f;
  }
  const float nw = noise ? noise_w[0] : 0.f;
  const int hw = x.h * x.w;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, hw);
  for (int pix = p0 + lane_pix; pix < p1; pix += pix_lanes) {
    const int oy = pix / x.w, ox = pix - oy * x.w;
    float v[8];
    h8_to_f(ld_h8(x.p + n * x.sn + oy * x.sh + ox * x.sw + c8 * 8), v);
    const float nz = noise ? nw * noise[(size_t)n * hw + pix] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = fmaf(v[i], av[i], bv[i]) + nz;
      t = t > 0.f ? t : t * slope;
      v[i] = t * pv[i];
    }
    st_h8(y.p + n * y.sn + oy * y.sh + ox * y.sw + c8 * 8, f_to_h8(v));
  }
}

// ToRGB: out[n][k][oy][ox] = sum_c x[n][oy+crop][ox+crop][c] * w[k][c] * s[n][c] + bias[k] + bilinear_x2(skip)[n][k][oy+crop][ox+crop]
// skip: float32 NCHW [N][3][H/2][W/2] (F.interpolate(scale_factor=2, bilinear, align_corners=False)); out: float32 NCHW
// [N][3][H-2crop][W-2crop].  The C/8 lanes that share a pixel each own 8 channels (one coalesced row read per pixel, the folded
// per-sample weights of the lane's channels stay in registers) and finish with a butterfly; C/8 must be a power of two <= 32.
__global__ void __launch_bounds__(256) to_rgb_kernel(View x, const float* __restrict__ w, const float* __restrict__ s, long long s_stride,
                                                    const float* __restrict__ bias, const float* __restrict__ skip, float* __restrict__ out,
                                                    int crop, int pix_per_block) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y, C = x.c, L = C >> 3;                     // lanes per pixel
  const int sub = threadIdx.x % L, lane_pix = threadIdx.x / L, pix_lanes = blockDim.x / L;
  float wm[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = sub * 8 + i;
    const float sc = s[(size_t)n * s_stride + c];
#pragma unroll
    for (int k = 0; k < 3; ++k) wm[k][i] = w[k * C + c] * sc;
  }
  const int OH = x.h - 2 * crop, OW = x.w - 2 * crop;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, OH * OW);
  const int sh = x.h >> 1, sw = x.w >> 1;
  for (int base = p0; base < p1; base += pix_lanes) {                // warp-uniform trip count: every lane takes part in the shuffles
    const int pix = base + lane_pix;
    const bool ok = pix < p1;
    const int oy = ok ? pix / OW : 0, ox = ok ? pix - oy * OW : 0;
    const int yy = oy + crop, xx = ox + crop;
    float v[8];
    h8_to_f(ld_h8(x.p + n * x.sn + yy * x.sh + xx * x.sw + sub * 8), v);
    float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[0] = fmaf(v[i], wm[0][i], r[0]); r[1] = fmaf(v[i], wm[1][i], r[1]); r[2] = fmaf(v[i], wm[2][i], r[2]);
    }
    for (int o = L >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
    }
    if (!ok || sub >= 3) continue;
    const int k = sub;                                                // lanes 0..2 of the pixel's group write one plane each
    float val = (k == 0 ? r[0] : k == 1 ? r[1] : r[2]) + bias[k];
    if (skip) {
      const float sy = fmaxf(0.5f * ((float)yy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * ((float)xx + 0.5f) - 0.5f, 0.f);
      const int y0 = min((int)sy, sh - 1), x0 = min((int)sx, sw - 1);
      const int y1 = y0 + (y0 < sh - 1 ? 1 : 0), x1 = x0 + (x0 < sw - 1 ? 1 : 0);
      const float ly = sy - (float)y0, lx = sx - (float)x0;
      const float* sp = skip + ((size_t)n * 3 + k) * sh * sw;
      const float top = sp[y0 * sw + x0] + lx * (sp[y0 * sw + x1] - sp[y0 * sw + x0]);
      const float bot = sp[y1 * sw + x0] + lx * (sp[y1 * sw + x1] - sp[y1 * sw + x0]);
      val += top + ly * (bot - top);
    }
    out[(((size_t)n * 3 + k) * OH + oy) * OW + ox] = val;
  }
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
  const int C8 = x->c >> 3;
  if (C8 > 256 || x->n > 65535) return S2V_EINVAL;
  const int hw = x->h * x->w;
  const int pix_lanes = 256 / C8;                                   // pixels a block handles per iteration
  int ppb = pix_lanes * 16;                                         // ~16 pixels per thread: the per-channel tables are amortised
  if (ppb > hw) ppb = hw;
  S2V_CUDA_TRY(launch_pdl(style_epilogue_kernel, dim3(ceil_div(hw, ppb), x->n), 256, 0, (cudaStream_t)stream, mk(x), a, bias, noise, noise_w, slope,
                          post, (long long)post_stride, mk(y), ppb));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_to_rgb(const s2v_view* x, const float* w, const float* s, int64_t s_stride, const float* bias, const float* skip,
                          float* out, int crop, void* stream) {
  if (!view_ok(x) || !w || !s || !bias || !out || crop < 0 || 2 * crop >= x->h || 2 * crop >= x->w || x->n > 65535)
    return S2V_EINVAL;
  if (skip && ((x->h & 1) || (x->w & 1))) return S2V_EINVAL;
  const int L = x->c >> 3;
  if (L < 4 || L > 32 || (L & (L - 1))) return S2V_EINVAL;          // 32 <= C <= 256, C / 8 a power of two (ENet: 256, 128)
  const int opix = (x->h - 2 * crop) * (x->w - 2 * crop);
  const int ppb = (256 / L) * 16;
  S2V_CUDA_TRY(launch_pdl(to_rgb_kernel, dim3(ceil_div(opix, ppb), x->n), 256, 0, (cudaStream_t)stream, mk(x), w, s, (long long)s_stride, bias,
                          skip, out, crop, ppb));
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
