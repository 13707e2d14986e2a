// 7x7 zero-padded "head" convolutions with a handful of output channels (Cout <= 8): FinalBlock2d of LNet's decoder and
// of DNet's editing net (models/base_blocks.py:444-457, 64 -> 3 + sigmoid / tanh) and WarpingNet.flow_out's conv
// (models/DNet.py:72-76, 256 -> 2), fp16 channels-last in, fp32 NCHW out.
//
// As a plain implicit GEMM these layers are N = 16 wide and need 49 taps x 4 K-steps = 196 tcgen05.mma per 128-pixel tile
// and 64-channel chunk - and one MMA costs ~60 cycles whatever N <= 128 is (tools/mb_umma.cu), i.e. 46-76 TFLOP/s.
// Here the kx taps are FOLDED INTO N:
//     P[pos, (kx, co)] = sum over ky, ci of  X[pos + ky * PW, ci] * W[co, ky, kx, ci]          (N = 7 * CP, CP = Cout padded to 2 / 4 / 8)
//     out[y, x, co]    = bias[co] + sum over kx of  P[y * PW + x + kx, (kx, co)]
// where pos runs over the positions of a PW = 32 pixel wide input patch (row-major).  The M tile is 128 CONSECUTIVE patch
// positions (4 patch rows), so the A operand of row tap ky is simply the K-major tile that starts ky * PW rows further
// down the same TMA-loaded patch (10 rows x 32 pixels x 64 channels, zero fill = zero padding): 7 x 4 = 28 MMAs per tile
// and chunk instead of 196.  A tile yields 4 rows x 26 output pixels (104 of 128 accumulator rows are useful).  The
// epilogue parks the 128 x (7 CP) fp32 accumulator in shared memory (column-major, conflict-free both ways) and every thread
// gathers the 7 shifted partial sums of its output pixel.
//
// Roles: warp 0 = TMA producer (patch slots + weight tiles), warp 1 = TMEM allocation + MMA issue, warps 2-9 = two
// epilogue groups (group g drains accumulator buffer g = the tiles j = g, g+2, ... of this CTA); persistent over tiles.
// The weight tile of one (chunk, ky) is [N = 7 CP -> 16 / 32 / 64 rows][64 channels]: 2 / 4 / 8 KB.  With the channel pad matched to
// Cout the weights of every head of the nets stay RESIDENT in smem (FinalBlock2d, Cin 64, Cout 3: 28 KB; flow_out, Cin 256, Cout 2:
// 56 KB - at a fixed pad of 8 flow_out's 224 KB had to stream through a ring, 896 KB of L2 -> smem traffic per tile, and that
// stream, not the MMAs, bounded the launch); a layer whose weights do not fit still streams them through an 8-deep ring.
#include <cuda.h>

#include "common.cuh"

namespace s2v {
namespace head {

constexpr int kK = 7, kPW = 32, kRows = 4, kOW = kPW - (kK - 1);   // 26 output columns, 4 output rows per 128-position M tile
// SUB vertically stacked M tiles share ONE patch of SUB * 4 + 6 rows (SUB = 2: 14 rows = 56 KB for 8 x 26 outputs instead of two
// 10-row patches = 80 KB): 30 % less patch traffic and twice the MMA work behind every patch load.  SUB = 2 needs the weights
// resident (Cin = 64: the FinalBlock2d heads); the streaming-weights head (flow_out, Cin = 256) keeps SUB = 1.
__host__ __device__ constexpr int patch_rows(int sub) { return sub * kRows + kK - 1; }
__host__ __device__ constexpr int patch_bytes(int sub) { return kPW * patch_rows(sub) * 128; }
__host__ __device__ constexpr int n_rows(int cp) { return cp == 8 ? 64 : cp == 4 ? 32 : 16; }   // MMA N: 7 * cp rounded up to 16
__host__ __device__ constexpr int b_tile(int cp) { return n_rows(cp) * 128; }                  // (kx, co) x 64 channels of one (chunk, ky)
__host__ __device__ constexpr int stage_bytes(int cp) { return 7 * cp * 128 * 4; }             // fp32 accumulator parked column-major, per epilogue group
constexpr int kASlots = 2, kBRing = 8, kThreads = 320;   // producer, MMA, 2 x 4 epilogue warps (the epilogue is the longer stage)
constexpr unsigned kSpin = 1u << 26;

struct Params {
  int N, H, W, cout, chunks, tiles_x, tiles_y, total_tiles, b_resident, act, rows_per_tile;
  float ap;
  const float* bias;
  float* out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B descriptor: start >> 4 | LBO 1 << 16 | SBO (1024 B) >> 4 << 32 | version 1 << 46 | SW128 << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& n, int& y0, int& x0) {
  const int per_img = p.tiles_x * p.tiles_y;
  n = tile / per_img;
  const int t = tile - n * per_img;
  const int ty = t / p.tiles_x;
  y0 = ty * p.rows_per_tile;
  x0 = (t - ty * p.tiles_x) * kOW;
}

template <int SUB, int CP>
__global__ void __launch_bounds__(kThreads, 1)
conv_head_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  constexpr int kPatchBytes = patch_bytes(SUB);
  constexpr int kNB = n_rows(CP), kBTile = b_tile(CP), kStageBytes = stage_bytes(CP), kCols = 7 * CP;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nb = p.b_resident ? p.chunks * kK : kBRing;               // weight tiles held in smem
  const uint32_t a0 = base, b0 = a0 + kASlots * kPatchBytes, s0 = b0 + (uint32_t)nb * kBTile;      // patches | weights | fp32 staging
  const uint32_t bar0 = s0 + 2 * kStageBytes;
  const uint32_t afull = bar0, aempty = afull + 8u * kASlots, bfull = aempty + 8u * kASlots, bempty = bfull + 8u * kBRing,
                 tfull = bempty + 8u * kBRing, tempty = tfull + 16u, ball = tempty + 16u, tptr = ball + 8u;
  float* stage0 = reinterpret_cast<float*>(smem_raw + (s0 - smem_u32(smem_raw)));
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < kASlots; ++i) { mbar_init(afull + 8u * i, 1); mbar_init(aempty + 8u * i, 1); }
    for (int i = 0; i < kBRing; ++i) { mbar_init(bfull + 8u * i, 1); mbar_init(bempty + 8u * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + 8u * i, 1); mbar_init(tempty + 8u * i, 4 * SUB); }
    mbar_init(ball, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(128u * SUB) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (p.b_resident) {                       // weights are never written by a kernel: fetch them before the PDL wait
        mbar_expect_tx(ball, (uint32_t)nb * kBTile);
        for (int i = 0; i < nb; ++i) tma_load_2d(b0 + (uint32_t)i * kBTile, &tmB, ball, i * 64, 0);
      }
      pdl_wait();
      uint32_t a = 0, aph = 0, s = 0, sph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int n, y0, x0;
        tile_coords(p, tile, n, y0, x0);
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(aempty + 8u * a, aph ^ 1u);
          mbar_expect_tx(afull + 8u * a, kPatchBytes);
          tma_load_4d(a0 + a * kPatchBytes, &tmA, afull + 8u * a, c * 64, x0 - kK / 2, y0 - kK / 2, n);
          if (++a == kASlots) { a = 0; aph ^= 1u; }
          if (!p.b_resident) {
            for (int ky = 0; ky < kK; ++ky) {
              mbar_wait(bempty + 8u * s, sph ^ 1u);
              mbar_expect_tx(bfull + 8u * s, kBTile);
              tma_load_2d(b0 + s * kBTile, &tmB, bfull + 8u * s, (c * kK + ky) * 64, 0);
              if (++s == kBRing) { s = 0; sph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issue: D[128 positions, kNB] (fp32, TMEM) += A(ky)[128, 64] * W(chunk, ky)[kNB, 64] =====
      const uint32_t idesc = (1u << 4) | ((uint32_t)(kNB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      if (p.b_resident) { mbar_wait(ball, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      uint32_t a = 0, aph = 0, s = 0, sph = 0;
      int j = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++j) {
        const uint32_t buf = (uint32_t)j & 1u;
        mbar_wait(tempty + 8u * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem + buf * (64u * SUB);                              // SUB accumulators of 64 columns per buffer
        uint32_t accum = 0;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(afull + 8u * a, aph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t adesc0 = umma_desc(a0 + a * kPatchBytes);
#pragma unroll 1
          for (int ky = 0; ky < kK; ++ky) {
            uint64_t bdesc;
            if (p.b_resident) {
              bdesc = umma_desc(b0 + (uint32_t)(c * kK + ky) * kBTile);
            } else {
              mbar_wait(bfull + 8u * s, sph);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              bdesc = umma_desc(b0 + s * kBTile);
            }
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {                                       // the stacked M tiles read the same weight tile
              const uint64_t adesc = adesc0 + (uint64_t)(((ky + sub * kRows) * kPW * 128) >> 4);   // (ky + 4 sub) patch rows further down
#pragma unroll
              for (int k = 0; k < 4; ++k)                                               // +32 B = 16 channels along K
                umma_f16(d_tmem + 64u * sub, adesc + 2u * k, bdesc + 2u * k, idesc, accum | (uint32_t)(k > 0));
            }
            accum = 1u;
            if (!p.b_resident) {
              umma_commit(bempty + 8u * s);
              if (++s == kBRing) { s = 0; sph ^= 1u; }
            }
          }
          umma_commit(aempty + 8u * a);
          if (++a == kASlots) { a = 0; aph ^= 1u; }
        }
        umma_commit(tfull + 8u * buf);
      }
    }
  } else {
    // ===== epilogue (2 groups x 4 warps): TMEM -> smem (column-major fp32) -> 7 shifted partial sums per output pixel =====
    // bias in registers for the kernel's lifetime: read per tile from global memory (a dependent load in front of every output
    // channel's sum) it was the epilogue's top stall - ncu of the DNet head, B = 192: long_scoreboard 5.5 cycles per issued
    // instruction, tensor pipe 20 % active, 1.56 TB/s of DRAM reads - the kernel is bound by this stage, not by its patch loads
    float bias_r[CP];
#pragma unroll
    for (int co = 0; co < CP; ++co) bias_r[co] = (p.bias && co < p.cout) ? __ldg(p.bias + co) : 0.f;     // weights: no PDL wait needed
    pdl_wait();
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;                                   // TMEM lane quarter of this warp
    const int pos = q * 32 + lane;                            // accumulator row = patch position
    const int et = ((warp - 2) & 3) * 32 + lane;              // 0..127: output slot (row et / 32, column et % 32)
    const int oy = et >> 5, ox = et & 31;
    float* stage = stage0 + grp * (kStageBytes / 4);
    const uint32_t bar_id = 1u + (uint32_t)grp;
    // SUB = 1: tile j -> group j & 1 -> accumulator buffer j & 1 (a group drains every other tile).
    // SUB = 2: group g drains stacked tile g of EVERY tile (accumulator g of buffer j & 1).
    int jj = 0;                                               // tiles this group has drained
    for (int tile = blockIdx.x + (SUB == 1 ? grp : 0) * gridDim.x; tile < p.total_tiles; tile += (SUB == 1 ? 2 : 1) * gridDim.x, ++jj) {
      int n, y0, x0;
      tile_coords(p, tile, n, y0, x0);
      const uint32_t buf = SUB == 1 ? (uint32_t)grp : ((uint32_t)jj & 1u);
      if (SUB == 2) y0 += grp * kRows;
      mbar_wait(tfull + 8u * buf, SUB == 1 ? ((uint32_t)jj & 1u) : (((uint32_t)jj >> 1) & 1u));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + buf * (64u * SUB) + (SUB == 2 ? 64u * grp : 0u);
      float v[kNB];
#pragma unroll
      for (int i = 0; i < kNB; i += 16) tmem_ld16(trow + (uint32_t)i, v + i);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty + 8u * buf) : "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");     // the previous tile's gathers are done with the staging buffer
#pragma unroll
      for (int i = 0; i < kCols; ++i) stage[i * 128 + pos] = v[i];
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const int Y = y0 + oy, X = x0 + ox;
      if (ox < kOW && Y < p.H && X < p.W) {
        float acc[CP];
#pragma unroll
        for (int co = 0; co < CP; ++co) {                    // all gathers of the pixel are independent: issued back to back
          acc[co] = bias_r[co];
          if (co < p.cout) {
#pragma unroll
            for (int kx = 0; kx < kK; ++kx) acc[co] += stage[(kx * CP + co) * 128 + et + kx];
          }
        }
        const size_t plane = (size_t)p.H * p.W;
        float* op = p.out + (size_t)n * p.cout * plane + (size_t)Y * p.W + X;
#pragma unroll
        for (int co = 0; co < CP; ++co) {
          if (co < p.cout) {
            float v = acc[co];
            if (p.act == S2V_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
            else if (p.act == S2V_ACT_TANH) v = tanhf(v);
            else if (p.act == S2V_ACT_RELU) v = fmaxf(v, 0.f);
            else if (p.act == S2V_ACT_LRELU) v = v > 0.f ? v : v * p.ap;
            op[co * plane] = v;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u * SUB) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;     // resolved once; immutable afterwards
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace head
}  // namespace s2v

using namespace s2v;
using namespace s2v::head;

// d: x fp16 NHWC (C a multiple of 64), kh = kw = 7, stride 1, pad 3, out_mode F32_NCHW with y_f32 [N][Cout][H][W], Cout <= 8,
// bias optional, act NONE / RELU / LRELU / SIGMOID / TANH.  d->w: fp16 [NB][chunks * 7 * 64] with CP = 2 / 4 / 8 the smallest pad
// >= Cout and NB = 16 / 32 / 64: row kx * CP + co, column (chunk * 7 + ky) * 64 + ci  (ops.pack_w_head); rows >= 7 CP and
// co >= Cout are zero.
extern "C" int s2v_conv_head(const s2v_conv* d, void* stream) {
  if (!d || !view_ok(&d->x) || !d->w || !d->y_f32) return S2V_EINVAL;
  if (d->kh != kK || d->kw != kK || d->stride_h != 1 || d->stride_w != 1 || d->dil_h != 1 || d->dil_w != 1 || d->pad_h != kK / 2 ||
      d->pad_w != kK / 2 || d->out_mode != S2V_OUT_F32_NCHW || d->res1.ptr || d->res2.ptr || d->x2.ptr || d->scale || d->stats_partial)
    return S2V_EINVAL;
  const int N = d->x.n, H = d->x.h, W = d->x.w, cout = d->y.c;
  if (cout <= 0 || cout > 8 || (d->x.c % 64) || d->y.n != N || d->y.h != H || d->y.w != W) return S2V_EINVAL;
  if (d->act != S2V_ACT_NONE && d->act != S2V_ACT_RELU && d->act != S2V_ACT_LRELU && d->act != S2V_ACT_SIGMOID && d->act != S2V_ACT_TANH)
    return S2V_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return S2V_EUNSUPPORTED;
  Params p;
  p.N = N; p.H = H; p.W = W; p.cout = cout; p.chunks = d->x.c / 64;
  // two stacked M tiles per patch when the weights of the single chunk stay resident next to two 14-row patches (Cin = 64)
  static const int sub_env = [] { const char* e = getenv("S2V_HEAD_SUB"); return e ? atoi(e) : 2; }();      // development knob
  const int cp = cout <= 2 ? 2 : cout <= 4 ? 4 : 8;
  const int kBTile = b_tile(cp);
  const int fixed = 2 * stage_bytes(cp) + 256 + 1024;
  const int sub = (sub_env == 2 && (long long)p.chunks * kK * kBTile + kASlots * patch_bytes(2) + fixed <= 227 * 1024) ? 2 : 1;
  p.rows_per_tile = sub * kRows;
  p.tiles_x = ceil_div(W, kOW); p.tiles_y = ceil_div(H, p.rows_per_tile);
  const long long tiles = (long long)p.tiles_x * p.tiles_y * N;
  if (tiles <= 0 || tiles > 0x7fffffff) return S2V_EINVAL;
  p.total_tiles = (int)tiles;
  p.act = d->act; p.ap = d->act_param; p.bias = d->bias; p.out = d->y_f32;
  const int kPatchBytes = patch_bytes(sub);
  p.b_resident = (sub == 2 || (long long)p.chunks * kK * kBTile + kASlots * kPatchBytes + fixed <= 224 * 1024) ? 1 : 0;
  const size_t smem = (size_t)kASlots * kPatchBytes + (size_t)(p.b_resident ? p.chunks * kK : kBRing) * kBTile + fixed;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)d->x.c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)d->x.sw * 2, (cuuint64_t)d->x.sh * 2, (cuuint64_t)d->x.sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kPW, (cuuint32_t)patch_rows(sub), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->x.ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)p.chunks * kK * 64;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)n_rows(cp)};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)n_rows(cp)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  static DeviceOnce attr;     // per device, idempotent
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const Params);
  static const KernelFn kernels[2][3] = {{conv_head_kernel<1, 2>, conv_head_kernel<1, 4>, conv_head_kernel<1, 8>},
                                         {conv_head_kernel<2, 2>, conv_head_kernel<2, 4>, conv_head_kernel<2, 8>}};
  if (attr.needed(dev)) {
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < 3; ++k) S2V_CUDA_TRY(cudaFuncSetAttribute(kernels[i][k], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr.mark(dev);
  }
  const int n_sm = sm_count(dev);
  if (n_sm <= 0) return S2V_ECUDA;
  const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
  if (smem > 227 * 1024) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(kernels[sub - 1][cp == 2 ? 0 : cp == 4 ? 1 : 2], grid, kThreads, smem, (cudaStream_t)stream, tmA, tmB, p));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
