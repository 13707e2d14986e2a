// Cross-attention core of models/transformer.py:77-86: per (frame n, head h)
//   out = softmax(q k^T * scale) v,  q,k from the masked-face tokens, v from the reference.
// 144 tokens x 64 dims: K and V of one head live in shared memory (fp16, 36 KB), one thread
// per query row keeps q and the output row in registers and runs an online softmax over keys
// in chunks of 8 (exp2 with the scale folded into q).  0.04 GFLOP/frame - latency-bound.
#include "common.cuh"

namespace s2v {

constexpr int kDh = 64, kMaxT = 256, kChunk = 8;

__global__ void __launch_bounds__(kMaxT) attention_kernel(View q, View k, View v, View o, int heads, float scale_log2e) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __half sm[];        // K [T][64], V [T][64]
  const int T = q.w;
  const int n = blockIdx.x / heads, h = blockIdx.x - n * heads;
  __half* sk = sm;
  __half* sv = sm + T * kDh;
  for (int i = threadIdx.x; i < T * (kDh / 8); i += blockDim.x) {
    const int tok = i / (kDh / 8), c8 = i - tok * (kDh / 8);
    st_h8(sk + tok * kDh + c8 * 8, ld_h8(k.p + n * k.sn + tok * k.sw + h * kDh + c8 * 8));
    st_h8(sv + tok * kDh + c8 * 8, ld_h8(v.p + n * v.sn + tok * v.sw + h * kDh + c8 * 8));
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= T) return;
  float qr[kDh], acc[kDh];
#pragma unroll
  for (int c8 = 0; c8 < kDh / 8; ++c8) {
    float f[8];
    h8_to_f(ld_h8(q.p + n * q.sn + i * q.sw + h * kDh + c8 * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { qr[c8 * 8 + j] = f[j] * scale_log2e; acc[c8 * 8 + j] = 0.f; }
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < T; j0 += kChunk) {
    float s[kChunk];
    float cm = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kChunk; ++jj) {
      const int j = j0 + jj;
      float d = -INFINITY;
      if (j < T) {
        d = 0.f;
        const __half2* kr = reinterpret_cast<const __half2*>(sk + j * kDh);
#pragma unroll
        for (int c = 0; c < kDh / 2; ++c) {
          const float2 kk = __half22float2(kr[c]);
          d = fmaf(qr[2 * c], kk.x, fmaf(qr[2 * c + 1], kk.y, d));
        }
      }
      s[jj] = d;
      cm = fmaxf(cm, d);
    }
    if (cm > m) {
      const float corr = exp2f(m - cm);     // m = -inf on the first chunk -> 0
      l *= corr;
#pragma unroll
      for (int c = 0; c < kDh; ++c) acc[c] *= corr;
      m = cm;
    }
#pragma unroll
    for (int jj = 0; jj < kChunk; ++jj) {
      const int j = j0 + jj;
      if (j < T) {
        const float p = exp2f(s[jj] - m);
        l += p;
        const __half2* vr = reinterpret_cast<const __half2*>(sv + j * kDh);
#pragma unroll
        for (int c = 0; c < kDh / 2; ++c) {
          const float2 vv = __half22float2(vr[c]);
          acc[2 * c] = fmaf(p, vv.x, acc[2 * c]);
          acc[2 * c + 1] = fmaf(p, vv.y, acc[2 * c + 1]);
        }
      }
    }
  }
  const float inv = 1.f / l;
  __half* op = o.p + n * o.sn + i * o.sw + h * kDh;
#pragma unroll
  for (int c8 = 0; c8 < kDh / 8; ++c8) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = acc[c8 * 8 + j] * inv;
    st_h8(op + c8 * 8, f_to_h8(f));
  }
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_attention(const s2v_view* q, const s2v_view* k, const s2v_view* v, const s2v_view* o, int heads,
                             float scale, void* stream) {
  if (!view_ok(q) || !view_ok(k) || !view_ok(v) || !view_ok(o) || heads <= 0) return S2V_EINVAL;
  const int T = q->w;
  if (q->h != 1 || k->h != 1 || v->h != 1 || o->h != 1 || T > kMaxT || k->w != T || v->w != T || o->w != T) return S2V_EINVAL;
  if (q->c < heads * kDh || k->c < heads * kDh || v->c < heads * kDh || o->c < heads * kDh) return S2V_EINVAL;
  if (k->n != q->n || v->n != q->n || o->n != q->n) return S2V_EINVAL;
  const int threads = ((T + 31) / 32) * 32;
  const size_t smem = (size_t)2 * T * kDh * sizeof(__half);
  if (smem > 48 * 1024) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMaxT * kDh * 2); attr = true; }
  }
  launch_pdl(attention_kernel, q->n * heads, threads, smem, (cudaStream_t)stream, mk(q), mk(k), mk(v), mk(o), heads,
                                                                         scale * 1.4426950408889634f);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
