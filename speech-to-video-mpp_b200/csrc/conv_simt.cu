// SIMT direct convolution for the small / awkward layers of the path (Cin = 3 7x7
// stems, Cout = 3 heads, the strided audio encoder of models/LNet.py:102-120, the
// AdaIN MLP stems, MappingNet's dilated Conv1d of models/DNet.py:38-42, DNet's
// 4x4 stride-2 convs).  fp16 channels-last input, fp32 weights and accumulation,
// fused epilogue  y = act(conv*scale + bias + res1) + res2.
// Thread tile: PX=2 adjacent output pixels x CO output channels; all lanes of a warp
// share the weight address (broadcast loads), inputs are 128-bit loads.
#include "common.cuh"

namespace s2v {

struct ConvP {
  View x, y, r1, r2;
  const float* w;
  const float* scale;
  const float* bias;
  int kh, kw, sh, sw, ph, pw, dh, dw, pad_mode, up2, act;
  float ap;
  int out_mode;
  float* yf;
  int cout, cout_pad;
};

__device__ __forceinline__ int reflect_idx(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

template <int PX, int CO>
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvP p) {
  pdl_trigger();
  pdl_wait();
  const int OH = p.y.h, OW = p.y.w;
  const int OWG = (OW + PX - 1) / PX;
  const long long groups = (long long)p.y.n * OH * OWG;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cob = (int)(idx / groups);
  if (cob >= p.cout_pad / CO) return;
  long long r = idx - (long long)cob * groups;
  const int oxg = (int)(r % OWG); r /= OWG;
  const int oy = (int)(r % OH);
  const int n = (int)(r / OH);
  const int ox0 = oxg * PX;
  const int IH = p.x.h << p.up2, IW = p.x.w << p.up2;
  const int C8 = p.x.c >> 3;

  float acc[PX][CO];
#pragma unroll
  for (int a = 0; a < PX; ++a)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[a][c] = 0.f;

  const __half* xb = p.x.p + n * p.x.sn;
  for (int ky = 0; ky < p.kh; ++ky) {
    int iy = oy * p.sh - p.ph + ky * p.dh;
    bool vy = true;
    if (p.pad_mode == S2V_PAD_REFLECT) iy = reflect_idx(iy, IH);
    else vy = (iy >= 0) && (iy < IH);
    const int sy = iy >> p.up2;
    for (int kx = 0; kx < p.kw; ++kx) {
      const __half* px[PX];
      bool ok[PX];
#pragma unroll
      for (int a = 0; a < PX; ++a) {
        int ix = (ox0 + a) * p.sw - p.pw + kx * p.dw;
        bool vx = true;
        if (p.pad_mode == S2V_PAD_REFLECT) ix = reflect_idx(ix, IW);
        else vx = (ix >= 0) && (ix < IW);
        ok[a] = vy && vx && (ox0 + a < OW);
        px[a] = xb + sy * p.x.sh + (ix >> p.up2) * p.x.sw;
      }
      bool any = false;
#pragma unroll
      for (int a = 0; a < PX; ++a) any |= ok[a];
      if (!any) continue;
      const float* wp = p.w + ((size_t)(ky * p.kw + kx) * p.x.c) * p.cout_pad + cob * CO;
      for (int c8 = 0; c8 < C8; ++c8) {
        float xin[PX][8];
#pragma unroll
        for (int a = 0; a < PX; ++a) {
          if (ok[a]) h8_to_f(ld_h8(px[a] + c8 * 8), xin[a]);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) xin[a][i] = 0.f;
          }
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          float wv[CO];
          const float4* w4 = reinterpret_cast<const float4*>(wp + (size_t)(c8 * 8 + ci) * p.cout_pad);
#pragma unroll
          for (int q = 0; q < CO / 4; ++q) {
            const float4 t = __ldg(w4 + q);
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int a = 0; a < PX; ++a)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[a][c] = fmaf(xin[a][ci], wv[c], acc[a][c]);
        }
      }
    }
  }

  // epilogue
#pragma unroll
  for (int a = 0; a < PX; ++a) {
    const int ox = ox0 + a;
    if (ox >= OW) continue;
    float v[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const int ch = cob * CO + c;
      float t = acc[a][c];
      if (ch < p.cout) {
        if (p.scale) t *= p.scale[ch];
        if (p.bias) t += p.bias[ch];
        if (p.r1.p) t += __half2float(p.r1.p[n * p.r1.sn + oy * p.r1.sh + ox * p.r1.sw + ch]);
        t = act_apply(t, p.act, p.ap);
        if (p.r2.p) t += __half2float(p.r2.p[n * p.r2.sn + oy * p.r2.sh + ox * p.r2.sw + ch]);
      } else {
        t = 0.f;
      }
      v[c] = t;
    }
    if (p.out_mode == S2V_OUT_F32_NCHW) {
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        const int ch = cob * CO + c;
        if (ch < p.cout) p.yf[(((size_t)n * p.cout + ch) * OH + oy) * OW + ox] = v[c];
      }
    } else {
      __half* yo = p.y.p + n * p.y.sn + oy * p.y.sh + ox * p.y.sw + cob * CO;
      if (CO % 8 == 0 && cob * CO + CO <= p.y.c) {
#pragma unroll
        for (int q = 0; q < CO / 8; ++q) st_h8(yo + 8 * q, f_to_h8(v + 8 * q));
      } else {
#pragma unroll
        for (int c = 0; c < CO; ++c)
          if (cob * CO + c < p.y.c) yo[c] = __float2half_rn(v[c]);
      }
    }
  }
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_conv_simt(const s2v_conv* d, void* stream) {
  if (!d || !view_ok(&d->x) || !d->w) return S2V_EINVAL;
  if (d->out_mode == S2V_OUT_F16_NHWC && !view_ok(&d->y)) return S2V_EINVAL;
  if (d->out_mode == S2V_OUT_F32_NCHW && !d->y_f32) return S2V_EINVAL;
  if (d->kh <= 0 || d->kw <= 0 || d->stride_h <= 0 || d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0) return S2V_EINVAL;
  if (d->up2 != 0 && d->up2 != 1) return S2V_EINVAL;
  const int IH = d->x.h << d->up2, IW = d->x.w << d->up2;
  const int OH = (IH + 2 * d->pad_h - d->dil_h * (d->kh - 1) - 1) / d->stride_h + 1;
  const int OW = (IW + 2 * d->pad_w - d->dil_w * (d->kw - 1) - 1) / d->stride_w + 1;
  if (OH != d->y.h || OW != d->y.w || d->y.n != d->x.n) return S2V_EINVAL;
  if (d->pad_mode == S2V_PAD_REFLECT && (d->pad_h >= IH || d->pad_w >= IW)) return S2V_EINVAL;
  ConvP p;
  p.x = mk(d->x); p.y = mk(d->y);
  p.r1 = mk(d->res1.ptr ? &d->res1 : nullptr);
  p.r2 = mk(d->res2.ptr ? &d->res2 : nullptr);
  p.w = (const float*)d->w; p.scale = d->scale; p.bias = d->bias;
  p.kh = d->kh; p.kw = d->kw; p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w;
  p.dh = d->dil_h; p.dw = d->dil_w; p.pad_mode = d->pad_mode; p.up2 = d->up2; p.act = d->act; p.ap = d->act_param;
  p.out_mode = d->out_mode; p.yf = d->y_f32;
  p.cout = d->y.c;
  const int cout = d->y.c;
  const bool wide = cout >= 16;
  p.cout_pad = wide ? ((cout + 15) / 16) * 16 : ((cout + 3) / 4) * 4;
  const long long groups = (long long)d->y.n * OH * ((OW + 1) / 2);
  if (wide) {
    const long long threads = groups * (p.cout_pad / 16);
    launch_pdl(conv_simt_kernel<2, 16>, ceil_div(threads, 128), 128, 0, (cudaStream_t)stream, p);
  } else {
    const long long threads = groups * (p.cout_pad / 4);
    launch_pdl(conv_simt_kernel<2, 4>, ceil_div(threads, 128), 128, 0, (cudaStream_t)stream, p);
  }
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
