// Shared helpers for libs2v kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/s2v.h"

#define S2V_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return S2V_ECUDA;                \
  } while (0)

namespace s2v {

struct View {           // device-side copy of s2v_view with typed pointer
  __half* p;
  int n, h, w, c;
  long long sn, sh, sw;
};

static inline View mk(const s2v_view* v) {
  View r;
  r.p = v ? (__half*)v->ptr : nullptr;
  if (v) { r.n = v->n; r.h = v->h; r.w = v->w; r.c = v->c; r.sn = v->sn; r.sh = v->sh; r.sw = v->sw; }
  else { r.n = r.h = r.w = r.c = 0; r.sn = r.sh = r.sw = 0; }
  return r;
}
static inline View mk(const s2v_view& v) { return mk(&v); }

static inline bool view_ok(const s2v_view* v) {
  if (!v || !v->ptr) return false;
  if (v->n <= 0 || v->h <= 0 || v->w <= 0 || v->c <= 0) return false;
  if ((v->c & 7) || (v->sw & 7) || (v->sh & 7) || (v->sn & 7)) return false;
  if (((uintptr_t)v->ptr) & 15) return false;
  return true;
}

__device__ __forceinline__ float act_apply(float v, int act, float p) {
  switch (act) {
    case S2V_ACT_RELU: return fmaxf(v, 0.f);
    case S2V_ACT_LRELU: return v > 0.f ? v : v * p;
    case S2V_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case S2V_ACT_TANH: return tanhf(v);
    case S2V_ACT_GELU: {
      float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      return 0.5f * v * (1.f + tanhf(u));
    }
    default: return v;
  }
}

struct __align__(16) H8 { __half2 v[4]; };

__device__ __forceinline__ void h8_to_f(const H8& h, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ H8 f_to_h8(const float* f) {
  H8 h;
#pragma unroll
  for (int i = 0; i < 4; ++i) h.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return h;
}
__device__ __forceinline__ H8 ld_h8(const __half* p) { return *reinterpret_cast<const H8*>(p); }
__device__ __forceinline__ void st_h8(__half* p, const H8& v) { *reinterpret_cast<H8*>(p) = v; }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace s2v
