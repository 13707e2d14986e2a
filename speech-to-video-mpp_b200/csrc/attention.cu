// Cross-attention core of models/transformer.py:77-86: per (frame n, head h)
//   out = softmax(q k^T * scale) v,  q,k from the masked-face tokens, v from the reference.
// 144 tokens x 64 dims: K and V of one head live in shared memory (fp16, 36 KB), one thread
// per query row keeps q and the output row in registers and runs an online softmax over keys
// in chunks of 8 (exp2 with the scale folded into q).  This SIMT kernel is the general-shape path; the 144-token
// LNet geometry runs on the tensor-core kernel below.
#include <cuda.h>

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


// ---- tcgen05 / TMEM / TMA path (the default for the 144-token LNet geometry) ---------------------------------------------
// One CTA (128 threads) per (frame, head).  Q, K, V of the head ([144 tokens][64 dims] fp16 = channel window h*64 .. of the
// channels-last token tensors) arrive by three tiled TMA loads in the K-major SWIZZLE_128B layout tcgen05.mma reads.
// For each of the two 128-row query tiles (rows 0-127, 128-143 + 112 don't-care rows):
//   S[128 x 144]  = Q_tile K^T      4 x tcgen05.mma (M 128, N 144, K 16), accumulator in TMEM columns [0, 144)
//   softmax       thread = query row: two passes over its TMEM row (max, then exp2 / sum); P (fp16) is written to shared
//                 memory as the K-major A operand of the second GEMM (three 64-column panels, SWIZZLE_128B)
//   O[128 x 64]   = P V             9 x tcgen05.mma (M 128, N 64, K 16): V stays as loaded ([token][dim]) and is read as an
//                 MN-MAJOR B operand (N = dims contiguous, K = tokens: one 16-row step = 2 048 bytes; verified layout of
//                 tools/probe_colsum_umma.cu), accumulator in TMEM columns [160, 224)
//   epilogue      tcgen05.ld of the row, * 1 / sum, fp16, 128-byte row store
// 102 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM.
namespace attn_tc {
constexpr int kT = 144, kRowsBytes = kT * 128;                  // 18 432 B per operand tile
constexpr int kOffQ = 0, kOffK = kRowsBytes, kOffV = 2 * kRowsBytes, kOffP = 3 * kRowsBytes, kPBytes = 3 * 128 * 128;
constexpr int kSmem = kOffP + kPBytes + 64 + 1024;              // + barriers / tmem pointer + 1024-byte alignment slack
constexpr unsigned kSpin = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (++spins > kSpin) __trap();
  }
}
// shared-memory matrix descriptor, SWIZZLE_128B: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | swizzle 2 << 61
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(128, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, View o, int heads, float scale_log2e) {
  extern __shared__ __align__(1024) uint8_t raw[];
  pdl_trigger();
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* al = raw + (base - smem_u32(raw));
  const uint32_t sQ = base + kOffQ, sK = base + kOffK, sV = base + kOffV, sP = base + kOffP;
  const uint32_t bar_tma = sP + kPBytes, bar_mma = bar_tma + 8, tptr = bar_tma + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(al + kOffP + kPBytes + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / heads, h = blockIdx.x - n * heads;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_tma) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tptr) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr_gen;
  pdl_wait();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_tma), "r"(3u * kRowsBytes) : "memory");
    const int c0 = h * kDh;
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sQ), "l"(&tmQ), "r"(bar_tma), "r"(c0), "r"(0), "r"(0), "r"(n) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sK), "l"(&tmK), "r"(bar_tma), "r"(c0), "r"(0), "r"(0), "r"(n) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sV), "l"(&tmV), "r"(bar_tma), "r"(c0), "r"(0), "r"(0), "r"(n) : "memory");
  }
  mbar_wait(bar_tma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // instruction descriptors: D fp32 (bit 4), A / B fp16, N >> 3 at bit 17, M >> 4 at bit 24; bit 16 = B is MN-major
  const uint32_t idesc_s = (1u << 4) | ((uint32_t)(kT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(kDh >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t tS = tmem, tO = tmem + 160;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const int rl = warp * 32 + lane;                         // row of the 128-row tile this thread owns
  uint32_t phase = 0;
  for (int mt = 0; mt < 2; ++mt) {
    // ---- S = Q_tile K^T: the second tile starts 128 rows (16 KB) into Q; its rows past token 143 read whatever follows
    //      (K's bytes - finite fp16 values) and are never used
    if (threadIdx.x == 0) {
      const uint64_t da = desc(sQ + (uint32_t)mt * 16384u, 16, 1024), db = desc(sK, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma(tS, da + 2u * k, db + 2u * k, idesc_s, k ? 1u : 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int tok = mt * 128 + rl;
    const bool valid = tok < kT;
    // tcgen05.ld is warp-collective: every lane walks its TMEM row (the don't-care rows of the second tile too); only the
    // shared-memory / global stores are predicated
    float l = 0.f;
    {
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kT; c += 16) {
        float v[16];
        tmem_ld16(tS + lane_sel + (uint32_t)c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
      }
      const float mb = mx * scale_log2e;
#pragma unroll 1
      for (int c = 0; c < kT; c += 16) {
        float v[16];
        tmem_ld16(tS + lane_sel + (uint32_t)c, v);
        H8 p8[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __half2 pp = __floats2half2_rn(exp2f(fmaf(v[2 * i], scale_log2e, -mb)), exp2f(fmaf(v[2 * i + 1], scale_log2e, -mb)));
          const float2 pf = __half22float2(pp);             // the sum uses the rounded values: rows are exactly normalised
          l += pf.x + pf.y;
          p8[i >> 2].v[i & 3] = pp;
        }
        // K-major SWIZZLE_128B A tile: panel = 64 key columns, row = 128 B, 16-byte chunk j stored at j ^ (row & 7)
        const int panel = c >> 6, chunk = (c & 63) >> 3;
        uint8_t* rowp = al + kOffP + panel * 16384 + rl * 128;
        if (valid) {
          *reinterpret_cast<H8*>(rowp + (((chunk) ^ (rl & 7)) << 4)) = p8[0];
          *reinterpret_cast<H8*>(rowp + (((chunk + 1) ^ (rl & 7)) << 4)) = p8[1];
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes of P -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- O = P V
    if (threadIdx.x == 0) {
#pragma unroll
      for (int ks = 0; ks < kT / 16; ++ks) {
        const uint64_t da = desc(sP + (uint32_t)(ks >> 2) * 16384u, 16, 1024) + 2u * (ks & 3);
        const uint64_t db = desc(sV + (uint32_t)ks * 2048u, 16384, 1024);   // MN-major: 16 token rows per K step
        mma(tO, da, db, idesc_o, ks ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const float inv = 1.f / l;
      __half* op = o.p + n * o.sn + (valid ? tok : 0) * o.sw + h * kDh;
#pragma unroll
      for (int c = 0; c < kDh; c += 16) {
        float v[16];
        tmem_ld16(tO + lane_sel + (uint32_t)c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= inv;
        if (valid) {
          st_h8(op + c, f_to_h8(v));
          st_h8(op + c + 8, f_to_h8(v + 8));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // TMEM reads done before the next tile's MMAs overwrite S / O
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;     // resolved once; immutable afterwards
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}
// the head's [144 tokens][64 dims] window of a [N,1,144,C] token tensor as a 4-D tiled map {c, token, 1, n}
static bool make_map(EncodeTiledFn enc, const s2v_view* t, CUtensorMap* tm) {
  cuuint64_t gdim[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, 1, (cuuint64_t)t->n};
  cuuint64_t gstr[3] = {(cuuint64_t)t->sw * 2, (cuuint64_t)t->sh * 2, (cuuint64_t)t->sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)kDh, (cuuint32_t)kT, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t->ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace attn_tc

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_attention(const s2v_view* q, const s2v_view* k, const s2v_view* v, const s2v_view* o, int heads,
                             float scale, void* stream) {
  if (!view_ok(q) || !view_ok(k) || !view_ok(v) || !view_ok(o) || heads <= 0) return S2V_EINVAL;
  const int T = q->w;
  if (q->h != 1 || k->h != 1 || v->h != 1 || o->h != 1 || T > kMaxT || k->w != T || v->w != T || o->w != T) return S2V_EINVAL;
  if (q->c < heads * kDh || k->c < heads * kDh || v->c < heads * kDh || o->c < heads * kDh) return S2V_EINVAL;
  if (k->n != q->n || v->n != q->n || o->n != q->n) return S2V_EINVAL;
  static const int path = [] { const char* e = getenv("S2V_ATTN"); return e ? atoi(e) : 0; }();   // development knob: 0 tcgen05, 1 mma.sync, 2 SIMT
  if (T == 144 && path == 0 && (q->sw & 7) == 0 && (k->sw & 7) == 0 && (v->sw & 7) == 0) {
    // the LNet geometry (12 x 12 tokens): tcgen05 / TMEM / TMA path
    attn_tc::EncodeTiledFn enc = attn_tc::get_encode();
    if (!enc) return S2V_EUNSUPPORTED;
    CUtensorMap tq, tk, tv;
    if (!attn_tc::make_map(enc, q, &tq) || !attn_tc::make_map(enc, k, &tk) || !attn_tc::make_map(enc, v, &tv)) return S2V_ECUDA;
    static DeviceOnce attr_t5;
    const int dev = current_device();
    if (dev < 0) return S2V_ECUDA;
    if (attr_t5.needed(dev)) {
      S2V_CUDA_TRY(cudaFuncSetAttribute(attn_tc::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::kSmem));
      attr_t5.mark(dev);
    }
    S2V_CUDA_TRY(launch_pdl(attn_tc::attention_tc_kernel, q->n * heads, 128, (size_t)attn_tc::kSmem, (cudaStream_t)stream, tq, tk, tv, mk(o), heads,
                            scale * 1.4426950408889634f));
    S2V_CHECK_LAUNCH();
    return S2V_OK;
  }
  if (T == 144 && path <= 1) {      // mma.sync path (kept as a cross-check of the tcgen05 kernel)
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
