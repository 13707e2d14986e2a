// Normalisation kernels (memory-bound, 128-bit vectorised, deterministic):
//   chan_stats      per-(n, chunk, c) partial sum / sum of squares (fixed-order reduction)
//   ln2d_finalize   LayerNorm2d over (C,H,W)   (models/base_blocks.py:52-69)  -> per-(n,c) affine
//   adain_finalize  InstanceNorm2d + AdaIN     (models/base_blocks.py:127-157) -> per-(n,c) affine
//   affine_act      y = act(x*a + b) [2x2 avg-pool] [+res] [reflect border]  (one read, one write)
//   token_layernorm nn.LayerNorm(C) per token   (models/transformer.py:27,35-36)
//   add, reflect_border, mean_over_w
// No float atomics anywhere: sharded and unsharded runs give bit-identical frames.
#include "common.cuh"

namespace s2v {

// grid (chunks, N); block = PG * C8 threads (C8 = C/8); thread (pg, c8) walks pixels pg, pg+PG, ...
__global__ void chan_stats_kernel(View x, int chunks, int PG, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];   // [PG][C][2]
  const int C8 = x.c >> 3;
  const int c8 = threadIdx.x % C8, pg = threadIdx.x / C8;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int HW = x.h * x.w;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (pg < PG) {
    // 4 independent 16-byte loads in flight per thread (addresses computed first, then loads, then math)
    for (int p = p0 + pg; p < p1; p += 4 * PG) {
      H8 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pu = p + u * PG;
        ok[u] = pu < p1;
        const int pc = ok[u] ? pu : p;
        const int yy = pc / x.w, xx = pc - yy * x.w;
        v[u] = ld_h8(x.p + n * x.sn + yy * x.sh + xx * x.sw + c8 * 8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        float f[8];
        h8_to_f(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    float* o = sm + ((size_t)pg * x.c + c8 * 8) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[2 * i] = s[i]; o[2 * i + 1] = q[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < x.c; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < PG; ++g) { a += sm[((size_t)g * x.c + c) * 2]; b += sm[((size_t)g * x.c + c) * 2 + 1]; }
    float* o = partial + (((size_t)n * chunks + chunk) * x.c + c) * 2;
    o[0] = a; o[1] = b;
  }
}

// grid N, block T (256, or 1024 for long partial lists): fixed-order reduction over chunks and channels (double
// accumulators).  T depends only on chunks * C (per-image geometry), never on the batch, so results are batch-independent.
template <int T>
__global__ void __launch_bounds__(T) ln2d_finalize_kernel(const float* __restrict__ partial, int chunks, int PC, int C,
                                                          double inv_count, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          float* __restrict__ a, float* __restrict__ b) {
  pdl_trigger();
  pdl_wait();
  __shared__ double ss[T], sq[T];
  const int n = blockIdx.x;
  double s = 0.0, q = 0.0;
  // thread = (channel c, chunk lane): all T threads stream the [chunks][C][2] partials of this image with
  // 8 independent float2 loads in flight; fixed order per thread + fixed tree below => deterministic
  // PC = entries per chunk of the partial list: C (per-channel partials) or 4 (LayerNorm2d totals written by s2v_conv_tc)
  const int lanes = PC >= T ? 1 : T / PC;                   // chunk lanes when PC < T (PC divides T or lanes = 1)
  const int c_of = threadIdx.x % (lanes > 1 ? PC : T), lane = lanes > 1 ? threadIdx.x / PC : 0;
  if (lane < lanes) {
    for (int c = c_of; c < PC; c += (lanes > 1 ? PC : T)) {
      const float2* pp = reinterpret_cast<const float2*>(partial) + (size_t)n * chunks * PC + c;
#pragma unroll 8
      for (int k = lane; k < chunks; k += lanes) {
        const float2 v = pp[(size_t)k * PC];
        s += (double)v.x; q += (double)v.y;
      }
    }
  }
  ss[threadIdx.x] = s; sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = T / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  const double mean = ss[0] * inv_count;
  double var = sq[0] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float fmean = (float)mean;
  for (int c = threadIdx.x; c < C; c += T) {
    const float av = rstd * gamma[c];
    a[(size_t)n * C + c] = av;
    b[(size_t)n * C + c] = beta[c] - fmean * av;
  }
}

// block = 32 channels x 8 chunk slices: slice w sums the chunks w, w + 8, ... of its (n, c) - independent loads, all in flight
// at once (one thread per (n, c) walking the chunk list was a chain of dependent-latency batches: 5.3 us per launch, 54 launches
// per LNet forward) - and slice 0 adds the eight slice sums in a fixed order: deterministic, independent of the batch.
constexpr int kFinSlices = 8;
__global__ void __launch_bounds__(32 * kFinSlices) adain_finalize_kernel(const float* __restrict__ partial, int N, int chunks,
                                                             int C, float inv_count, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long gb_stride,
                                                             float eps, float* __restrict__ a, float* __restrict__ b) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh_s[kFinSlices][32], sh_q[kFinSlices][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  const bool ok = idx < N * C;
  const int n = ok ? idx / C : 0, c = ok ? idx - n * C : 0;
  double s = 0.0, q = 0.0;
  if (ok) {
    const float2* pp = reinterpret_cast<const float2*>(partial) + (size_t)n * chunks * C + c;
#pragma unroll 4
    for (int k = slice; k < chunks; k += kFinSlices) {
      const float2 v = __ldg(pp + (size_t)k * C);
      s += (double)v.x; q += (double)v.y;
    }
  }
  sh_s[slice][lane] = s; sh_q[slice][lane] = q;
  __syncthreads();
  if (slice != 0 || !ok) return;
#pragma unroll
  for (int w = 1; w < kFinSlices; ++w) { s += sh_s[w][lane]; q += sh_q[w][lane]; }
  const double mean = s * (double)inv_count;
  double var = q * (double)inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[(size_t)n * gb_stride + c] : 0.f;
  const float be = beta ? beta[(size_t)n * gb_stride + c] : 0.f;
  const float av = rstd * (1.f + g);
  a[idx] = av;
  b[idx] = be - (float)mean * av;
}

// one thread per TWO output vectors (n, oy, ox, c8): items idx and idx + half are loaded together so two
// independent 16-byte loads (four with a residual) are in flight per thread
__device__ __forceinline__ void affine_store(const View& y, int n, int oy, int ox, int c8, const float* o, int reflect1) {
  const H8 hv = f_to_h8(o);
  __half* yp = y.p + n * y.sn + c8 * 8;
  st_h8(yp + oy * y.sh + ox * y.sw, hv);
  if (reflect1) {
    // padded(-1) = in(1), padded(H) = in(H-2); h,w >= 4 is enforced on the host
    const int my = (oy == 1) ? -1 : (oy == y.h - 2 ? y.h : -2);
    const int mx = (ox == 1) ? -1 : (ox == y.w - 2 ? y.w : -2);
    if (my != -2) st_h8(yp + my * y.sh + ox * y.sw, hv);
    if (mx != -2) st_h8(yp + oy * y.sh + mx * y.sw, hv);
    if (my != -2 && mx != -2) st_h8(yp + my * y.sh + mx * y.sw, hv);
  }
}

__device__ __forceinline__ void load_ab(const float* __restrict__ a, const float* __restrict__ b, int n, int C, int c8,
                                        float* av, float* bv) {
  const float4* pa = reinterpret_cast<const float4*>(a + (size_t)n * C + c8 * 8);
  const float4* pb = reinterpret_cast<const float4*>(b + (size_t)n * C + c8 * 8);
  const float4 t0 = pa[0], t1 = pa[1], u0 = pb[0], u1 = pb[1];
  av[0] = t0.x; av[1] = t0.y; av[2] = t0.z; av[3] = t0.w; av[4] = t1.x; av[5] = t1.y; av[6] = t1.z; av[7] = t1.w;
  bv[0] = u0.x; bv[1] = u0.y; bv[2] = u0.z; bv[3] = u0.w; bv[4] = u1.x; bv[5] = u1.y; bv[6] = u1.z; bv[7] = u1.w;
}

template <int ACT>
__device__ __forceinline__ float act_c(float v, float ap) {
  if (ACT == S2V_ACT_LRELU) return v > 0.f ? v : v * ap;
  if (ACT == S2V_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == S2V_ACT_NONE) return v;
  return v;
}

// grid (ceil(W*C8/256), ceil(H/2), N): thread = (ox, c8) of output rows oy and oy + ceil(H/2) of image n.
// Both rows share the per-(n,c) affine; all loads of both rows are issued before any math.
template <int POOL, int ACT, int R2>
__global__ void __launch_bounds__(256) affine_act_kernel(View x, const float* __restrict__ a,
                                                         const float* __restrict__ b, float ap, View res, View y,
                                                         int reflect1, int rev, const float* __restrict__ ra,
                                                         const float* __restrict__ rb) {
  pdl_trigger();
  pdl_wait();
  constexpr int U = 2;
  const int C8 = x.c >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= y.w * C8) return;
  const int ox = idx / C8, c8 = idx - ox * C8;
  // rev: walk the images (and rows) from the END of the tensor - the producer (a conv walking tiles in ascending order)
  // wrote those last, so they are the part of x still resident in L2; this kernel's own first writes (high images) age
  // out while its last writes (low images) are what the next ascending conv reads first.  Pure scheduling, same results.
  const int n = rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int by = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int hh = (y.h + 1) >> 1;
  int oy[U];
  bool ok[U];
  oy[0] = by; oy[1] = by + hh;
  ok[0] = true; ok[1] = oy[1] < y.h;
  H8 xin[U][POOL ? 4 : 1], rin[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!ok[u]) continue;
    if (POOL) {
#pragma unroll
      for (int d = 0; d < 4; ++d)
        xin[u][d] = ld_h8(x.p + n * x.sn + (2 * oy[u] + (d >> 1)) * x.sh + (2 * ox + (d & 1)) * x.sw + c8 * 8);
    } else {
      xin[u][0] = ld_h8(x.p + n * x.sn + oy[u] * x.sh + ox * x.sw + c8 * 8);
    }
    if (res.p) rin[u] = ld_h8(res.p + n * res.sn + oy[u] * res.sh + ox * res.sw + c8 * 8);
  }
  float av[8], bv[8], rav[R2 ? 8 : 1], rbv[R2 ? 8 : 1];
  load_ab(a, b, n, x.c, c8, av, bv);
  if (R2) load_ab(ra, rb, n, x.c, c8, rav, rbv);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!ok[u]) continue;
    float o[8];
    if (POOL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float f[8];
        h8_to_f(xin[u][d], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += act_c<ACT>(fmaf(f[i], av[i], bv[i]), ap);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= 0.25f;
    } else {
      float f[8];
      h8_to_f(xin[u][0], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_c<ACT>(fmaf(f[i], av[i], bv[i]), ap);
    }
    if (res.p) {
      float f[8];
      h8_to_f(rin[u], f);
      if (R2) {               // the residual is itself a raw conv output: its own per-(n,c) affine + the same activation
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += act_c<ACT>(fmaf(f[i], rav[R2 ? i : 0], rbv[R2 ? i : 0]), ap);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += f[i];
      }
    }
    affine_store(y, n, oy[u], ox, c8, o, reflect1);
  }
}

// Flat streaming form (the default): the OUTPUT of one image is a 1-D array of 16-byte vectors (pixel-major, 8 channels each); a
// block owns a contiguous span of them inside ONE image and every thread streams kFlatU vectors that are blockDim apart - fully
// coalesced, and all loads of a thread (4-16 independent 16-byte loads) are in flight before the first use.  C/8 divides the block
// size, so a thread's channel slice never changes and the per-(n, c) affine is loaded once.  Views may be strided (channel slices,
// interiors of reflect-padded buffers); pooling reads the 2 x 2 input window of every output pixel.  Same arithmetic as
// affine_act_kernel (the row-pair form above, kept for channel counts whose C/8 does not divide 256).  Measured on B200, DNet's
// 256 x 256 x 64 tensors: 4.9 -> 6.3 TB/s (x -> y), 4.7 -> 6.0 TB/s (two raw inputs), 6.1 -> 6.9 TB/s (x + res -> y).
constexpr int kFlatU = 4, kFlatT = 256;
template <int POOL, int ACT, int R2, bool DENSE>
__global__ void __launch_bounds__(kFlatT) affine_act_flat_kernel(View x, const float* __restrict__ a, const float* __restrict__ b, float ap,
                                                                View res, View y, int reflect1, int spans_per_img,
                                                                const float* __restrict__ ra, const float* __restrict__ rb) {
  pdl_trigger();
  pdl_wait();
  const int C8 = y.c >> 3;
  const int n = blockIdx.x / spans_per_img, span = blockIdx.x - n * spans_per_img;
  const int vec_per_img = y.h * y.w * C8;
  const int v0 = span * (kFlatT * kFlatU) + threadIdx.x;            // vector index inside the image
  const int c8 = v0 % C8;                                           // kFlatT % C8 == 0: constant for this thread
  const int pstep = kFlatT / C8;                                    // pixels between a thread's consecutive vectors
  const int pix0 = v0 / C8;
  int oy[kFlatU], ox[kFlatU];
  bool ok[kFlatU];
  H8 xin[kFlatU][POOL ? 4 : 1], rin[kFlatU];
  // DENSE (x, y, res contiguous, no pooling / border): plain pointer + vector index, no per-vector coordinate arithmetic - the
  // leaner address path is worth 5.1 -> 6.3 TB/s on the 256 x 256 x 64 tensors
  const __half* xd = x.p + (size_t)n * x.sn;
  const __half* rd = res.p ? res.p + (size_t)n * res.sn : nullptr;
  __half* yd = y.p + (size_t)n * y.sn;
#pragma unroll
  for (int u = 0; u < kFlatU; ++u) {
    const int pix = pix0 + u * pstep;
    ok[u] = v0 + u * kFlatT < vec_per_img;
    if (DENSE) {
      if (ok[u]) {
        xin[u][0] = ld_h8(xd + (size_t)(v0 + u * kFlatT) * 8);
        if (rd) rin[u] = ld_h8(rd + (size_t)(v0 + u * kFlatT) * 8);
      }
      continue;
    }
    // strided views: one division for the thread's first pixel, increments afterwards; 32-bit offsets inside the image
    if (u == 0) { oy[0] = pix / y.w; ox[0] = pix - oy[0] * y.w; }
    else {
      oy[u] = oy[u - 1]; ox[u] = ox[u - 1] + pstep;
      while (ox[u] >= y.w) { ox[u] -= y.w; ++oy[u]; }
    }
    if (!ok[u]) continue;
    const int xsh = (int)x.sh, xsw = (int)x.sw;
    if (POOL) {
      const __half* p00 = xd + (2 * oy[u]) * xsh + (2 * ox[u]) * xsw + c8 * 8;
      xin[u][0] = ld_h8(p00); xin[u][1] = ld_h8(p00 + xsw); xin[u][2] = ld_h8(p00 + xsh); xin[u][3] = ld_h8(p00 + xsh + xsw);
    } else {
      xin[u][0] = ld_h8(xd + oy[u] * xsh + ox[u] * xsw + c8 * 8);
    }
    if (rd) rin[u] = ld_h8(rd + oy[u] * (int)res.sh + ox[u] * (int)res.sw + c8 * 8);
  }
  float av[8], bv[8], rav[R2 ? 8 : 1], rbv[R2 ? 8 : 1];
  load_ab(a, b, n, x.c, c8, av, bv);
  if (R2) load_ab(ra, rb, n, x.c, c8, rav, rbv);
#pragma unroll
  for (int u = 0; u < kFlatU; ++u) {
    if (!ok[u]) continue;
    float f[8], o[8];
    if (POOL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        h8_to_f(xin[u][d], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += act_c<ACT>(fmaf(f[i], av[i], bv[i]), ap);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= 0.25f;
    } else {
      h8_to_f(xin[u][0], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_c<ACT>(fmaf(f[i], av[i], bv[i]), ap);
    }
    if (res.p) {
      h8_to_f(rin[u], f);
      if (R2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += act_c<ACT>(fmaf(f[i], rav[R2 ? i : 0], rbv[R2 ? i : 0]), ap);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += f[i];
      }
    }
    if (DENSE) st_h8(yd + (size_t)(v0 + u * kFlatT) * 8, f_to_h8(o));
    else affine_store(y, n, oy[u], ox[u], c8, o, reflect1);
  }
}

// Single-pass InstanceNorm + AdaIN + activation (+res, +reflect border) for feature maps whose (image, channel
// group) slab fits in shared memory: one block = image n x CG channels.  The slab is read ONCE from HBM/L2 into
// smem while per-thread partial sums are accumulated; the per-channel reduction runs in a fixed order
// (deterministic); the normalised result is written straight from smem.  Replaces chan_stats + adain_finalize +
// affine_act (three launches, two reads of x) for the FFC levels of LNet.
constexpr int kFusedThreads = 256;

template <int CG, int ACT>
__global__ void __launch_bounds__(kFusedThreads) adain_fused_kernel(View x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, long long gb_stride,
                                                                    float eps, float ap, View res, View y, int reflect1) {
  pdl_trigger();
  pdl_wait();
  constexpr int G8 = CG / 8;              // 16-byte vectors per pixel in this channel group
  constexpr int PL = kFusedThreads / G8;  // pixel lanes
  constexpr int NW = kFusedThreads / 32;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int HW = x.h * x.w;
  H8* slab = reinterpret_cast<H8*>(smraw);                          // [HW][G8]
  float* red = reinterpret_cast<float*>(smraw + (size_t)HW * G8 * 16);   // [NW][CG][2]
  float* ab = red + NW * CG * 2;                                    // [CG][2]
  const int n = blockIdx.y, c0 = blockIdx.x * CG;
  const int v = threadIdx.x % G8, pl = threadIdx.x / G8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const __half* xb = x.p + n * x.sn + c0 + v * 8;
  for (int p = pl; p < HW; p += 2 * PL) {
    H8 t[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int pu = p + u * PL;
      ok[u] = pu < HW;
      const int pc = ok[u] ? pu : p;
      const int yy = pc / x.w, xx = pc - yy * x.w;
      t[u] = ld_h8(xb + yy * x.sh + xx * x.sw);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      slab[(size_t)(p + u * PL) * G8 + v] = t[u];
      float f[8];
      h8_to_f(t[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
  // fixed-order reduction: across the pixel lanes of a warp (xor shuffles), then across warps through smem
#pragma unroll
  for (int off = G8; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
      q[i] += __shfl_xor_sync(0xffffffffu, q[i], off);
    }
  }
  if (lane < G8) {
    float* o = red + ((size_t)warp * CG + v * 8) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[2 * i] = s[i]; o[2 * i + 1] = q[i]; }
  }
  __syncthreads();
  if (threadIdx.x < CG) {
    const int c = threadIdx.x;
    double sd = 0.0, qd = 0.0;
    for (int g = 0; g < NW; ++g) { sd += (double)red[((size_t)g * CG + c) * 2]; qd += (double)red[((size_t)g * CG + c) * 2 + 1]; }
    const double mean = sd / (double)HW;
    double var = qd / (double)HW - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[(size_t)n * gb_stride + c0 + c] : 0.f;
    const float be = beta ? beta[(size_t)n * gb_stride + c0 + c] : 0.f;
    const float av = rstd * (1.f + g);
    ab[2 * c] = av;
    ab[2 * c + 1] = be - (float)mean * av;
  }
  __syncthreads();
  float av[8], bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { av[i] = ab[2 * (v * 8 + i)]; bv[i] = ab[2 * (v * 8 + i) + 1]; }
  const int c8g = (c0 >> 3) + v;           // vector index within the full channel dim
  for (int p = pl; p < HW; p += 2 * PL) {
    H8 r[2];
    bool ok[2];
    int yy[2], xx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int pu = p + u * PL;
      ok[u] = pu < HW;
      const int pc = ok[u] ? pu : p;
      yy[u] = pc / x.w; xx[u] = pc - yy[u] * x.w;
      if (res.p && ok[u]) r[u] = ld_h8(res.p + n * res.sn + yy[u] * res.sh + xx[u] * res.sw + c0 + v * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      float f[8], o[8];
      h8_to_f(slab[(size_t)(p + u * PL) * G8 + v], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_c<ACT>(fmaf(f[i], av[i], bv[i]), ap);
      if (res.p) {
        float g[8];
        h8_to_f(r[u], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += g[i];
      }
      affine_store(y, n, yy[u], xx[u], c8g, o, reflect1);
    }
  }
}

__global__ void __launch_bounds__(256) reflect_border_kernel(View v) {
  pdl_trigger();
  pdl_wait();
  // v = interior view; fills rows -1 / H and cols -1 / W (pad 1, reflect)
  const int C8 = v.c >> 3;
  const int PW = v.w + 2, PH = v.h + 2;
  const int border = 2 * PW + 2 * v.h;
  const long long total = (long long)v.n * border * C8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  long long r = idx / C8;
  const int e = (int)(r % border);
  const int n = (int)(r / border);
  int py, px;   // padded coords in [-1, H] x [-1, W]
  if (e < PW) { py = -1; px = e - 1; }
  else if (e < 2 * PW) { py = v.h; px = e - PW - 1; }
  else { const int k = e - 2 * PW; py = k >> 1; px = (k & 1) ? v.w : -1; }
  (void)PH;
  const int sy = py < 0 ? 1 : (py >= v.h ? v.h - 2 : py);
  const int sx = px < 0 ? 1 : (px >= v.w ? v.w - 2 : px);
  __half* base = v.p + n * v.sn + c8 * 8;
  st_h8(base + py * v.sh + px * v.sw, ld_h8(base + sy * v.sh + sx * v.sw));
}

// one warp per token; C <= 1024 (C/8 vectors spread over lanes)
__global__ void __launch_bounds__(256) token_ln_kernel(View x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, View y) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)x.n * x.h * x.w;
  if (tok >= total) return;
  const int xx = (int)(tok % x.w);
  const int yy = (int)((tok / x.w) % x.h);
  const int n = (int)(tok / ((long long)x.w * x.h));
  const __half* px = x.p + n * x.sn + yy * x.sh + xx * x.sw;
  __half* py = y.p + n * y.sn + yy * y.sh + xx * y.sw;
  const int C8 = x.c >> 3;
  float f[4][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c8 = lane + 32 * k;
    if (c8 < C8) {
      h8_to_f(ld_h8(px + c8 * 8), f[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[k][i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)x.c;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c8 = lane + 32 * k;
    if (c8 < C8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = f[k][i] - mean; q = fmaf(d, d, q); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)x.c + eps);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c8 = lane + 32 * k;
    if (c8 < C8) {
      float o[8], g[8], be[8];
      *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(gamma + c8 * 8);        // 16-byte parameter loads (32 scalar
      *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(gamma + c8 * 8 + 4);  // loads per lane made the kernel
      *reinterpret_cast<float4*>(be) = *reinterpret_cast<const float4*>(beta + c8 * 8);         // instruction bound: 0.22 of HBM)
      *reinterpret_cast<float4*>(be + 4) = *reinterpret_cast<const float4*>(beta + c8 * 8 + 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (f[k][i] - mean) * rstd * g[i] + be[i];
      st_h8(py + c8 * 8, f_to_h8(o));
    }
  }
}

__global__ void __launch_bounds__(256) add_kernel(View a, View b, View y) {
  pdl_trigger();
  pdl_wait();
  const int C8 = y.c >> 3;
  const long long total = (long long)y.n * y.h * y.w * C8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  long long r = idx / C8;
  const int xx = (int)(r % y.w); r /= y.w;
  const int yy = (int)(r % y.h);
  const int n = (int)(r / y.h);
  float fa[8], fb[8];
  h8_to_f(ld_h8(a.p + n * a.sn + yy * a.sh + xx * a.sw + c8 * 8), fa);
  h8_to_f(ld_h8(b.p + n * b.sn + yy * b.sh + xx * b.sw + c8 * 8), fb);
#pragma unroll
  for (int i = 0; i < 8; ++i) fa[i] += fb[i];
  st_h8(y.p + n * y.sn + yy * y.sh + xx * y.sw + c8 * 8, f_to_h8(fa));
}

__global__ void __launch_bounds__(256) mean_over_w_kernel(View x, View y) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= x.n * x.c) return;
  const int n = idx / x.c, c = idx - n * x.c;
  float s = 0.f;
  for (int i = 0; i < x.w; ++i) s += __half2float(x.p[n * x.sn + i * x.sw + c]);
  y.p[n * y.sn + c] = __float2half_rn(s / (float)x.w);
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_chan_stats(const s2v_view* x, int chunks, float* partial, void* stream) {
  if (!view_ok(x) || !partial || chunks <= 0 || chunks > 65535 || x->n > 65535) return S2V_EINVAL;
  const int C8 = x->c >> 3;
  if (C8 > 1024) return S2V_EINVAL;
  int PG = 256 / C8;
  if (PG < 1) PG = 1;
  if (PG > 32) PG = 32;
  const int threads = PG * C8;
  const size_t smem = (size_t)PG * x->c * 2 * sizeof(float);
  if (smem > 48 * 1024) return S2V_EINVAL;
  launch_pdl(chan_stats_kernel, dim3(chunks, x->n), threads, smem, (cudaStream_t)stream, mk(x), chunks, PG, partial);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_ln2d_finalize(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                                 const float* gamma, const float* beta, float eps, float* a, float* b, void* stream) {
  if (!partial || !gamma || !beta || !a || !b || N <= 0 || chunks <= 0 || C <= 0 || count_per_channel <= 0) return S2V_EINVAL;
  const double inv = 1.0 / ((double)count_per_channel * C);
  if ((long long)chunks * C >= 8192)
    launch_pdl(ln2d_finalize_kernel<1024>, N, 1024, 0, (cudaStream_t)stream, partial, chunks, C, C, inv, gamma, beta, eps, a, b);
  else
    launch_pdl(ln2d_finalize_kernel<256>, N, 256, 0, (cudaStream_t)stream, partial, chunks, C, C, inv, gamma, beta, eps, a, b);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_ln2d_finalize_totals(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                                        const float* gamma, const float* beta, float eps, float* a, float* b, void* stream) {
  if (!partial || !gamma || !beta || !a || !b || N <= 0 || chunks <= 0 || C <= 0 || count_per_channel <= 0) return S2V_EINVAL;
  const double inv = 1.0 / ((double)count_per_channel * C);
  launch_pdl(ln2d_finalize_kernel<256>, N, 256, 0, (cudaStream_t)stream, partial, chunks, 4, C, inv, gamma, beta, eps, a, b);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_adain_finalize(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                                  const float* gamma, const float* beta, int64_t gb_stride, float eps, float* a,
                                  float* b, void* stream) {
  if (!partial || !a || !b || N <= 0 || chunks <= 0 || C <= 0 || count_per_channel <= 0) return S2V_EINVAL;
  launch_pdl(adain_finalize_kernel, ceil_div((long long)N * C, 32), 32 * kFinSlices, 0, (cudaStream_t)stream, 
      partial, N, chunks, C, 1.f / (float)count_per_channel, gamma, beta, gb_stride, eps, a, b);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

static int affine_act_impl(const s2v_view* x, const float* a, const float* b, int act, float act_param, int pool2,
                           const s2v_view* res, const float* ra, const float* rb, const s2v_view* y, int reflect1, void* stream);

extern "C" int s2v_affine_act(const s2v_view* x, const float* a, const float* b, int act, float act_param, int pool2,
                              const s2v_view* res, const s2v_view* y, int reflect1, void* stream) {
  return affine_act_impl(x, a, b, act, act_param, pool2, res, nullptr, nullptr, y, reflect1, stream);
}

extern "C" int s2v_affine_act2(const s2v_view* x, const float* a, const float* b, int act, float act_param,
                               const s2v_view* res, const float* ra, const float* rb, const s2v_view* y, int reflect1,
                               void* stream) {
  if (!res || !res->ptr || !ra || !rb) return S2V_EINVAL;
  return affine_act_impl(x, a, b, act, act_param, 0, res, ra, rb, y, reflect1, stream);
}

static int affine_act_impl(const s2v_view* x, const float* a, const float* b, int act, float act_param, int pool2,
                           const s2v_view* res, const float* ra, const float* rb, const s2v_view* y, int reflect1, void* stream) {
  if (!view_ok(x) || !view_ok(y) || !a || !b) return S2V_EINVAL;
  if (y->c != x->c || y->n != x->n) return S2V_EINVAL;
  if (pool2 ? (x->h != 2 * y->h || x->w != 2 * y->w) : (x->h != y->h || x->w != y->w)) return S2V_EINVAL;
  if (res && res->ptr && (!view_ok(res) || res->c != y->c || res->h != y->h || res->w != y->w || res->n != y->n)) return S2V_EINVAL;
  if (reflect1 && (y->h < 4 || y->w < 4)) return S2V_EINVAL;
  if (act != S2V_ACT_NONE && act != S2V_ACT_LRELU && act != S2V_ACT_RELU) return S2V_EINVAL;
  if (y->n > 65535 || (y->h + 1) / 2 > 65535) return S2V_EINVAL;
  const dim3 grid(ceil_div((long long)y->w * (y->c >> 3), 256), (y->h + 1) / 2, y->n);
  const View vx = mk(x), vr = mk(res && res->ptr ? res : nullptr), vy = mk(y);
  cudaStream_t st = (cudaStream_t)stream;
  {
    static const int flat_env = [] { const char* e = getenv("S2V_AFFINE_FLAT"); return e ? atoi(e) : 1; }();      // development knob
    const int C8 = y->c >> 3;
    const long long vec_per_img = (long long)y->h * y->w * C8;
    if (flat_env && kFlatT % C8 == 0 && vec_per_img < (1ll << 30)) {
      const int spans = ceil_div(vec_per_img, kFlatT * kFlatU);
      const long long blocks = (long long)spans * y->n;
      if (blocks <= 0x7fffffffLL) {
        auto is_dense = [](const s2v_view* v) { return v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c; };
        const bool dense = !pool2 && !reflect1 && is_dense(x) && is_dense(y) && (!(res && res->ptr) || is_dense(res));
#define S2V_FLAT(P, A, R)                                                                                                                        \
  do {                                                                                                                                           \
    if (dense) launch_pdl(affine_act_flat_kernel<P, A, R, true>, (int)blocks, kFlatT, 0, st, vx, a, b, act_param, vr, vy, reflect1, spans, ra, rb); \
    else launch_pdl(affine_act_flat_kernel<P, A, R, false>, (int)blocks, kFlatT, 0, st, vx, a, b, act_param, vr, vy, reflect1, spans, ra, rb);   \
  } while (0)
        if (ra) {
          if (act == S2V_ACT_LRELU) S2V_FLAT(0, S2V_ACT_LRELU, 1);
          else if (act == S2V_ACT_NONE) S2V_FLAT(0, S2V_ACT_NONE, 1);
          else return S2V_EINVAL;
        } else if (pool2) {
          if (act == S2V_ACT_LRELU) S2V_FLAT(1, S2V_ACT_LRELU, 0);
          else if (act == S2V_ACT_RELU) S2V_FLAT(1, S2V_ACT_RELU, 0);
          else S2V_FLAT(1, S2V_ACT_NONE, 0);
        } else {
          if (act == S2V_ACT_LRELU) S2V_FLAT(0, S2V_ACT_LRELU, 0);
          else if (act == S2V_ACT_RELU) S2V_FLAT(0, S2V_ACT_RELU, 0);
          else S2V_FLAT(0, S2V_ACT_NONE, 0);
        }
#undef S2V_FLAT
        S2V_CHECK_LAUNCH();
        return S2V_OK;
      }
    }
  }
  static const int rev = [] { const char* e = getenv("S2V_AFFINE_REV"); return e ? atoi(e) : 0; }();
#define S2V_AFFINE(P, A, R) launch_pdl(affine_act_kernel<P, A, R>, grid, 256, 0, st, vx, a, b, act_param, vr, vy, reflect1, rev, ra, rb)
  if (ra) {                     // double affine (s2v_affine_act2): LeakyReLU / none only, no pooling
    if (act == S2V_ACT_LRELU) S2V_AFFINE(0, S2V_ACT_LRELU, 1);
    else if (act == S2V_ACT_NONE) S2V_AFFINE(0, S2V_ACT_NONE, 1);
    else return S2V_EINVAL;
  } else if (pool2) {
    if (act == S2V_ACT_LRELU) S2V_AFFINE(1, S2V_ACT_LRELU, 0);
    else if (act == S2V_ACT_RELU) S2V_AFFINE(1, S2V_ACT_RELU, 0);
    else S2V_AFFINE(1, S2V_ACT_NONE, 0);
  } else {
    if (act == S2V_ACT_LRELU) S2V_AFFINE(0, S2V_ACT_LRELU, 0);
    else if (act == S2V_ACT_RELU) S2V_AFFINE(0, S2V_ACT_RELU, 0);
    else S2V_AFFINE(0, S2V_ACT_NONE, 0);
  }
#undef S2V_AFFINE
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

static size_t adain_fused_smem(int hw, int cg) { return (size_t)hw * (cg / 8) * 16 + (size_t)(kFusedThreads / 32) * cg * 2 * sizeof(float) + cg * 2 * sizeof(float); }

extern "C" int s2v_adain_fused_fits(int h, int w, int c) {
  const int hw = h * w;
  if (c % 64 == 0 && adain_fused_smem(hw, 64) <= 100 * 1024) return 64;
  if (c % 16 == 0 && adain_fused_smem(hw, 16) <= 100 * 1024) return 16;
  return 0;
}

extern "C" int s2v_adain_fused(const s2v_view* x, const float* gamma, const float* beta, int64_t gb_stride, float eps,
                               int act, float act_param, const s2v_view* res, const s2v_view* y, int reflect1,
                               void* stream) {
  if (!view_ok(x) || !view_ok(y)) return S2V_EINVAL;
  if (y->c != x->c || y->n != x->n || y->h != x->h || y->w != x->w || x->n > 65535) return S2V_EINVAL;
  if (res && res->ptr && (!view_ok(res) || res->c != y->c || res->h != y->h || res->w != y->w || res->n != y->n)) return S2V_EINVAL;
  if (reflect1 && (y->h < 4 || y->w < 4)) return S2V_EINVAL;
  if (act != S2V_ACT_NONE && act != S2V_ACT_LRELU && act != S2V_ACT_RELU) return S2V_EINVAL;
  const int cg = s2v_adain_fused_fits(x->h, x->w, x->c);
  if (!cg) return S2V_EINVAL;
  const size_t smem = adain_fused_smem(x->h * x->w, cg);
  const View vx = mk(x), vr = mk(res && res->ptr ? res : nullptr), vy = mk(y);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(x->c / cg, x->n);
#define S2V_FUSED(CGV, A)                                                                                          \
  do {                                                                                                             \
    static DeviceOnce attr;                                                                                        \
    const int dev = current_device();                                                                              \
    if (dev < 0) return S2V_ECUDA;                                                                                 \
    if (attr.needed(dev)) {                                                                                        \
      S2V_CUDA_TRY(cudaFuncSetAttribute(adain_fused_kernel<CGV, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
      attr.mark(dev);                                                                                              \
    }                                                                                                              \
    launch_pdl(adain_fused_kernel<CGV, A>, grid, kFusedThreads, smem, st, vx, gamma, beta, gb_stride, eps, act_param, vr, vy, reflect1); \
  } while (0)
  if (cg == 64) {
    if (act == S2V_ACT_LRELU) S2V_FUSED(64, S2V_ACT_LRELU);
    else if (act == S2V_ACT_RELU) S2V_FUSED(64, S2V_ACT_RELU);
    else S2V_FUSED(64, S2V_ACT_NONE);
  } else {
    if (act == S2V_ACT_LRELU) S2V_FUSED(16, S2V_ACT_LRELU);
    else if (act == S2V_ACT_RELU) S2V_FUSED(16, S2V_ACT_RELU);
    else S2V_FUSED(16, S2V_ACT_NONE);
  }
#undef S2V_FUSED
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_reflect_border(const s2v_view* interior, void* stream) {
  if (!view_ok(interior) || interior->h < 2 || interior->w < 2) return S2V_EINVAL;
  const long long total = (long long)interior->n * (2 * (interior->w + 2) + 2 * interior->h) * (interior->c >> 3);
  launch_pdl(reflect_border_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, mk(interior));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_token_layernorm(const s2v_view* x, const float* gamma, const float* beta, float eps,
                                   const s2v_view* y, void* stream) {
  if (!view_ok(x) || !view_ok(y) || !gamma || !beta || x->c > 1024) return S2V_EINVAL;
  if ((((uintptr_t)gamma) | ((uintptr_t)beta)) & 15) return S2V_EINVAL;        // 16-byte parameter loads
  if (x->n != y->n || x->h != y->h || x->w != y->w || x->c != y->c) return S2V_EINVAL;
  const long long total = (long long)x->n * x->h * x->w;
  launch_pdl(token_ln_kernel, ceil_div(total, 8), 256, 0, (cudaStream_t)stream, mk(x), gamma, beta, eps, mk(y));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_add(const s2v_view* a, const s2v_view* b, const s2v_view* y, void* stream) {
  if (!view_ok(a) || !view_ok(b) || !view_ok(y)) return S2V_EINVAL;
  if (a->n != y->n || a->h != y->h || a->w != y->w || a->c != y->c) return S2V_EINVAL;
  if (b->n != y->n || b->h != y->h || b->w != y->w || b->c != y->c) return S2V_EINVAL;
  const long long total = (long long)y->n * y->h * y->w * (y->c >> 3);
  launch_pdl(add_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, mk(a), mk(b), mk(y));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_mean_over_w(const s2v_view* x, const s2v_view* y, void* stream) {
  if (!view_ok(x) || !view_ok(y) || x->h != 1 || y->h != 1 || y->w != 1 || x->c != y->c || x->n != y->n) return S2V_EINVAL;
  launch_pdl(mean_over_w_kernel, ceil_div((long long)x->n * x->c, 256), 256, 0, (cudaStream_t)stream, mk(x), mk(y));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
