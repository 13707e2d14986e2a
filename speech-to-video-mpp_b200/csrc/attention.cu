// Cross-attention core of models/transformer.py:77-86: per (frame n, head h)
//   out = softmax(q k^T * scale) v,  q,k from the masked-face tokens, v from the reference.
// 144 tokens x 64 dims: K and V of one head live in shared memory (fp16, 36 KB), one thread
// per query row keeps q and the output row in registers and runs an online softmax over keys
// in chunks of 8 (exp2 with the scale folded into q).  This SIMT kernel is the general-shape path; the 144-token
// LNet geometry runs on the tensor-core kernel below.
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


// ---- tensor-core path ---------------------------------------------------------------------------------------------
// One CTA per (frame, head); warp w owns query rows [16w, 16w+16).  S = Q K^T and O = P V run on mma.sync.m16n8k16
// (fp16 operands, fp32 accumulate): a head is 144 x 144 x 64 - far below the 128-row tcgen05 tile, and the whole score
// row block (16 x T fp32) fits in registers, so softmax is exact (no online rescaling) and P never leaves registers:
// the accumulator layout of two adjacent 8-key score tiles IS the A-fragment layout of the next 16-key P tile.
// Q, K, V of the head are staged in shared memory with a 16-byte row pad (conflict-free ldmatrix).
constexpr int kTcMaxT = 144, kRowH = kDh + 8;      // padded row: 72 halves = 144 B

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int T>     // T = tokens, a multiple of 16
__global__ void __launch_bounds__(T * 2) attention_mma_kernel(View q, View k, View v, View o, int heads, float scale_log2e) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) __half smh[];
  __half *sq = smh, *sk = smh + T * kRowH, *sv = smh + 2 * T * kRowH;
  const int n = blockIdx.x / heads, h = blockIdx.x - n * heads;
  for (int i = threadIdx.x; i < T * (kDh / 8); i += blockDim.x) {
    const int tok = i >> 3, c8 = i & 7;
    st_h8(sq + tok * kRowH + c8 * 8, ld_h8(q.p + n * q.sn + tok * q.sw + h * kDh + c8 * 8));
    st_h8(sk + tok * kRowH + c8 * 8, ld_h8(k.p + n * k.sn + tok * k.sw + h * kDh + c8 * 8));
    st_h8(sv + tok * kRowH + c8 * 8, ld_h8(v.p + n * v.sn + tok * v.sw + h * kDh + c8 * 8));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * 16;
  const uint32_t sq_u = (uint32_t)__cvta_generic_to_shared(sq), sk_u = (uint32_t)__cvta_generic_to_shared(sk),
                 sv_u = (uint32_t)__cvta_generic_to_shared(sv);
  // Q fragments of this warp's 16 rows: 4 k16 steps (ldmatrix.x4: lanes 0-15 -> rows, lanes 16-31 -> +8 columns)
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(sq_u + (uint32_t)(((row0 + (lane & 15)) * kRowH + ks * 16 + (lane >> 4) * 8) * 2), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
  // S = Q K^T: T/8 score tiles of 16 x 8
  float sc[T / 8][4];
#pragma unroll
  for (int nt = 0; nt < T / 8; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sc[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b0, b1;     // B[k = dim][n = key]: rows of K are keys -> non-transposed ldmatrix; lanes 0-7: dims ks*16.., 8-15: +8
      ldsm_x2(sk_u + (uint32_t)(((nt * 8 + (lane & 7)) * kRowH + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0, b1);
      mma16816(sc[nt], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
    }
  }
  // exact softmax over the T keys: thread holds rows r = lane/4 (c0,c1) and r+8 (c2,c3); a row lives in 4 lanes
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < T / 8; ++nt) { m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1])); m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3])); }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float mb0 = m0 * scale_log2e, mb1 = m1 * scale_log2e;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < T / 8; ++nt) {
    sc[nt][0] = exp2f(fmaf(sc[nt][0], scale_log2e, -mb0)); sc[nt][1] = exp2f(fmaf(sc[nt][1], scale_log2e, -mb0));
    sc[nt][2] = exp2f(fmaf(sc[nt][2], scale_log2e, -mb1)); sc[nt][3] = exp2f(fmaf(sc[nt][3], scale_log2e, -mb1));
    l0 += sc[nt][0] + sc[nt][1]; l1 += sc[nt][2] + sc[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // O = P V: 8 output tiles of 16 x 8 dims, T/16 key steps
  float oc[kDh / 8][4];
#pragma unroll
  for (int dt = 0; dt < kDh / 8; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) oc[dt][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < T / 16; ++kt) {
    const uint32_t a0 = pack_h2(sc[2 * kt][0], sc[2 * kt][1]), a1 = pack_h2(sc[2 * kt][2], sc[2 * kt][3]);
    const uint32_t a2 = pack_h2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]), a3 = pack_h2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
    for (int dt = 0; dt < kDh / 8; ++dt) {
      uint32_t b0, b1;     // B[k = key][n = dim]: V rows are keys -> transposed ldmatrix; lanes 0-15 give the 16 key rows
      ldsm_x2_trans(sv_u + (uint32_t)(((kt * 16 + (lane & 15)) * kRowH + dt * 8) * 2), b0, b1);
      mma16816(oc[dt], a0, a1, a2, a3, b0, b1);
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r = row0 + (lane >> 2), cc = (lane & 3) * 2;
  __half* op0 = o.p + n * o.sn + r * o.sw + h * kDh + cc;
  __half* op1 = op0 + 8 * o.sw;
#pragma unroll
  for (int dt = 0; dt < kDh / 8; ++dt) {
    *reinterpret_cast<__half2*>(op0 + dt * 8) = __floats2half2_rn(oc[dt][0] * i0, oc[dt][1] * i0);
    *reinterpret_cast<__half2*>(op1 + dt * 8) = __floats2half2_rn(oc[dt][2] * i1, oc[dt][3] * i1);
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
  if (T == 144 && !getenv("S2V_ATTN_SIMT")) {      // the LNet geometry (12 x 12 tokens): tensor-core path
    const size_t smem_tc = (size_t)3 * 144 * kRowH * sizeof(__half);
    static DeviceOnce attr_tc;       // per device, idempotent
    const int dev = current_device();
    if (dev < 0) return S2V_ECUDA;
    if (attr_tc.needed(dev)) {
      S2V_CUDA_TRY(cudaFuncSetAttribute(attention_mma_kernel<144>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
      attr_tc.mark(dev);
    }
    launch_pdl(attention_mma_kernel<144>, q->n * heads, 288, smem_tc, (cudaStream_t)stream, mk(q), mk(k), mk(v), mk(o), heads,
               scale * 1.4426950408889634f);
    S2V_CHECK_LAUNCH();
    return S2V_OK;
  }
  const int threads = ((T + 31) / 32) * 32;
  const size_t smem = (size_t)2 * T * kDh * sizeof(__half);
  if (smem > 48 * 1024) {
    static DeviceOnce attr;
    const int dev = current_device();
    if (dev < 0) return S2V_ECUDA;
    if (attr.needed(dev)) {
      S2V_CUDA_TRY(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMaxT * kDh * 2));
      attr.mark(dev);
    }
  }
  launch_pdl(attention_kernel, q->n * heads, threads, smem, (cudaStream_t)stream, mk(q), mk(k), mk(v), mk(o), heads,
                                                                         scale * 1.4426950408889634f);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
