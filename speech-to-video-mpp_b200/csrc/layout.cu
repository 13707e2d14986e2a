// Layout converters at the module boundary: NCHW float32 (the reference's tensors,
// inference.py:260-262) <-> fp16 channels-last views used between kernels.
#include "common.cuh"

namespace s2v {

// one thread per (n, y, x): reads C strided floats (coalesced along x per channel),
// writes c_fill contiguous halves.
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, int N, int C, int H, int W,
                                                   long long src_sn, View d, int c_off, int c_fill, float scale,
                                                   float shift) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int n = (int)(idx / ((long long)W * H));
  __half* o = d.p + n * d.sn + y * d.sh + x * d.sw + c_off;
  const float* s = src + (size_t)n * src_sn + (size_t)y * W + x;
  const size_t plane = (size_t)H * W;
  if (c_fill == 8 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {      // the usual stem packing: one 16-byte store
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = c < C ? fmaf(s[c * plane], scale, shift) : 0.f;
    st_h8(o, f_to_h8(f));
    return;
  }
  for (int c = 0; c < c_fill; ++c) o[c] = __float2half_rn(c < C ? fmaf(s[c * plane], scale, shift) : 0.f);
}

__global__ void __launch_bounds__(256) unpack_kernel(View s, int c_off, int C, float* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
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

// DNet -> LNet glue of the synthetic full-path configuration (the real glue is
// CPU image code, inference.py:188-239,341-411):  ref = bilinear((clamp(fake,-1,1)+1)/2 -> oh x ow,
// align_corners=False), face = cat(ref with rows >= mask_row zeroed, ref).
__global__ void __launch_bounds__(256) glue_kernel(const float* __restrict__ fake, float* __restrict__ face, int B,
                                                   int C, int H, int W, int oh, int ow, int mask_row) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * oh * ow;
  if (idx >= total) return;
  const int x = (int)(idx % ow);
  const int y = (int)((idx / ow) % oh);
  const int c = (int)((idx / ((long long)ow * oh)) % C);
  const int n = (int)(idx / ((long long)ow * oh * C));
  const float sch = (float)H / (float)oh, scw = (float)W / (float)ow;
  const float sy = fmaxf(sch * ((float)y + 0.5f) - 0.5f, 0.f), sx = fmaxf(scw * ((float)x + 0.5f) - 0.5f, 0.f);
  const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f);
  const float* s = fake + ((size_t)n * C + c) * H * W;
  auto t = [](float v) { return (fminf(fmaxf(v, -1.f), 1.f) + 1.f) * 0.5f; };
  const float v = (1.f - ly) * ((1.f - lx) * t(s[y0 * W + x0]) + lx * t(s[y0 * W + x1])) +
                  ly * ((1.f - lx) * t(s[y1 * W + x0]) + lx * t(s[y1 * W + x1]));
  const size_t plane = (size_t)oh * ow;
  float* o = face + (size_t)n * 2 * C * plane + (size_t)y * ow + x;
  o[c * plane] = y >= mask_row ? 0.f : v;
  o[(C + c) * plane] = v;
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_glue_fake_to_face_f32(const float* fake, float* face, int B, int C, int H, int W, int oh, int ow,
                                         int mask_row, void* stream) {
  if (B == 0) return S2V_OK;
  if (!fake || !face || B < 0 || C <= 0 || H <= 0 || W <= 0 || oh <= 0 || ow <= 0) return S2V_EINVAL;
  const long long total = (long long)B * C * oh * ow;
  launch_pdl(glue_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, fake, face, B, C, H, W, oh, ow, mask_row);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_pack_nchw_f32(const float* src, int N, int C, int H, int W, int64_t src_sn, const s2v_view* dst,
                                 int c_off, int c_fill, float scale, float shift, void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !view_ok(dst) || C <= 0 || c_fill < C || c_off < 0 || c_off + c_fill > dst->c) return S2V_EINVAL;
  if (dst->n < N || dst->h != H || dst->w != W || src_sn < (int64_t)C * H * W) return S2V_EINVAL;
  const long long total = (long long)N * H * W;
  launch_pdl(pack_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, src, N, C, H, W, src_sn, mk(dst), c_off, c_fill, scale, shift);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_unpack_to_nchw_f32(const s2v_view* src, int c_off, int C, float* dst, void* stream) {
  if (!view_ok(src) || !dst || C <= 0 || c_off < 0 || c_off + C > src->c) return S2V_EINVAL;
  const long long total = (long long)src->n * src->h * src->w;
  launch_pdl(unpack_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, mk(src), c_off, C, dst);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

// ---- bilinear resize of fp16 channels-last tensors (align_corners = False, PyTorch's F.interpolate semantics:
// src = scale * (dst + 0.5) - 0.5 clamped at 0, neighbours clamped at the last row / column).  Building block of the
// ENet upsampler (ResBlock x0.5 / StyleConv x2 / 96 -> 256, models/base_blocks.py:42-46,500-503,
// models/ENet.py:93,104); optional per-(n, c) input scale = the StyleGAN2 modulation folded into the resize pass.
namespace s2v {

// grid (x-chunk blocks, OH, N): row taps are block-uniform, column taps per thread; 32-bit index arithmetic only
__global__ void __launch_bounds__(256) resize_bilinear_kernel(View x, View y, const float* __restrict__ chan_scale, long long scale_stride) {
  pdl_trigger();
  pdl_wait();
  const int C8 = x.c >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;         // (ox, c8) of this output row
  if (idx >= y.w * C8) return;
  const int ox = idx / C8, c8 = idx - ox * C8;
  const int oy = blockIdx.y, n = blockIdx.z;
  const float sch = (float)x.h / (float)y.h, scw = (float)x.w / (float)y.w;
  const float sy = fmaxf(sch * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(scw * ((float)ox + 0.5f) - 0.5f, 0.f);
  const int y0 = min((int)sy, x.h - 1), x0 = min((int)sx, x.w - 1);
  const int y1 = y0 + (y0 < x.h - 1 ? 1 : 0), x1 = x0 + (x0 < x.w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const __half* p = x.p + n * x.sn + c8 * 8;
  float a[8], b[8], c[8], d[8], o[8];
  h8_to_f(ld_h8(p + y0 * x.sh + x0 * x.sw), a);
  h8_to_f(ld_h8(p + y0 * x.sh + x1 * x.sw), b);
  h8_to_f(ld_h8(p + y1 * x.sh + x0 * x.sw), c);
  h8_to_f(ld_h8(p + y1 * x.sh + x1 * x.sw), d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float top = a[i] + lx * (b[i] - a[i]), bot = c[i] + lx * (d[i] - c[i]);
    o[i] = top + ly * (bot - top);
  }
  if (chan_scale) {
    const float4 s0 = *reinterpret_cast<const float4*>(chan_scale + (size_t)n * scale_stride + c8 * 8);
    const float4 s1 = *reinterpret_cast<const float4*>(chan_scale + (size_t)n * scale_stride + c8 * 8 + 4);
    o[0] *= s0.x; o[1] *= s0.y; o[2] *= s0.z; o[3] *= s0.w; o[4] *= s1.x; o[5] *= s1.y; o[6] *= s1.z; o[7] *= s1.w;
  }
  st_h8(y.p + n * y.sn + oy * y.sh + ox * y.sw + c8 * 8, f_to_h8(o));
}

// F.interpolate(mode='bilinear', align_corners=False) of float32 planes: src plane p of image n at src + n*src_sn + p*src_sp
// (a channel window of a contiguous NCHW tensor), dst contiguous [N][P][OH][OW] window with the same addressing
__global__ void __launch_bounds__(256) resize_planes_f32_kernel(const float* __restrict__ src, long long src_sn, long long src_sp, int P, int H, int W,
                                                               float* __restrict__ dst, long long dst_sn, long long dst_sp, int OH, int OW, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH);
  const long long np = idx / ((long long)OW * OH);
  const int p = (int)(np % P);
  const long long n = np / P;
  const float sch = (float)H / (float)OH, scw = (float)W / (float)OW;
  const float sy = fmaxf(sch * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(scw * ((float)ox + 0.5f) - 0.5f, 0.f);
  const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float* sp = src + n * src_sn + p * src_sp;
  // same operation order as ATen's upsample_bilinear2d: hy*(hx*a + lx*b) + ly*(hx*c + lx*d)
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v = hy * (hx * sp[(size_t)y0 * W + x0] + lx * sp[(size_t)y0 * W + x1]) + ly * (hx * sp[(size_t)y1 * W + x0] + lx * sp[(size_t)y1 * W + x1]);
  dst[n * dst_sn + p * dst_sp + (size_t)oy * OW + ox] = v;
}

}  // namespace s2v

extern "C" int s2v_resize_bilinear(const s2v_view* x, const s2v_view* y, const float* chan_scale, int64_t scale_stride, void* stream) {
  if (!view_ok(x) || !view_ok(y) || x->n != y->n || x->c != y->c) return S2V_EINVAL;
  if (y->n == 0) return S2V_OK;
  if (chan_scale && ((((uintptr_t)chan_scale) & 15) || (scale_stride & 3) || scale_stride < 0)) return S2V_EINVAL;
  if (y->h > 65535 || y->n > 65535) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(resize_bilinear_kernel, dim3(ceil_div((long long)y->w * (x->c >> 3), 256), y->h, y->n), 256, 0, (cudaStream_t)stream,
                          mk(x), mk(y), chan_scale, (long long)(scale_stride ? scale_stride : x->c)));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_resize_planes_f32(const float* src, int64_t src_sn, int64_t src_sp, int N, int P, int H, int W, float* dst, int64_t dst_sn,
                                     int64_t dst_sp, int OH, int OW, void* stream) {
  if (N == 0 || P == 0) return S2V_OK;
  if (!src || !dst || N < 0 || P < 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return S2V_EINVAL;
  const long long total = (long long)N * P * OH * OW;
  S2V_CUDA_TRY(launch_pdl(resize_planes_f32_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, src, (long long)src_sn, (long long)src_sp, P, H, W,
                          dst, (long long)dst_sn, (long long)dst_sp, OH, OW, total));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
