// FourierUnit transforms as small DENSE DFT matrix products on the tensor cores (reference: models/ffc.py:99-102 rfftn +
// re/im channel interleave, :116-121 de-interleave + irfftn, norm='ortho'; same entry points and tensors as fft2d.cu).
//
// Why: the register FFT of fft2d.cu executes ~77 k thread-instructions per image-channel at 48 x 48 and is bound by its
// instruction issue (ncu: 14.8 M warp instructions per launch, 55 % issue-active, one 400-thread block per SM, 0.37-0.41 of the
// HBM peak).  A 48-point DFT is a 48 x 48 matrix: with the channels as the GEMM's N dimension (fp16 channels-last tensors are
// already [.., K index, channel]) both passes of the 2-D transform are [M <= 64] x [K <= 64] constant matrices times the tile,
// 0.76 MFLOP per image-channel on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate) = 5 x fewer issued instructions, and the
// tile needs 39 KB of shared memory instead of 154 KB, so four blocks per SM overlap their load / compute / store phases.
// STATUS (profiles/r2c_summary.md): parity-tested (tests/test_gpu_kernels.py::test_fft2_mma_path, CPU emulation in
// tests/test_fft_mma_emulation.py).  With cp.async tile loads the kernels only MATCH the register FFT: an 8-channel tile is a 16-byte
// slice of every pixel, one LDGSTS.128 of a warp touches 24 cache lines, and those L1TEX line accesses - not the MMAs - set the time
// (19 of 46 us at 48 x 48, B = 256).  With the tiles moved by TMA (rfft2_mma_tma_kernel / irfft2_mma_tma_kernel below: one 16-byte row
// per clock per SM, off the LSU) the 48 x 48 transforms run 1.2 - 1.6 x faster than the register FFT and are the default for that size
// (S2V_FFT_MMA, fft2d.cu); at 24 and 12 px the register FFT stays.  The contraction sizes are far below a 128-row tcgen05 tile and the operands change role between the two passes (the
// accumulator of pass 1 is the B operand of pass 2), which is what warp-level mma + ldmatrix.trans is for.
//
// One block (4 warps) = one image n x 8 channels.  The tile lives in shared memory as rows of RS = 32 (S/2+1) + 16 bytes:
//   forward:  X[h][w][8 c]  --pass 1 per h, in place-->  Y[h][(k, re|im)][8 c]  --pass 2 per k-->  spec[kh][k][2 c + ri]  (global)
//   inverse:  spec[kh][(k, re|im)][8 c] (de-interleaved on load)  --pass A per k, in place-->  T[h][(k, re|im)][8 c]
//             --pass B per h-->  y[h][w][c] (+ residual)
// A "row" of 16 bytes = the 8 channels of one K index, so ldmatrix.trans delivers B fragments directly (K = w, (k, ri) along a
// tile row; K = h / kh across tile rows, conflict-free because eight consecutive rows start in eight different 16-byte bank groups: Cfg).  Every pass is in place: a warp
// owns whole tile rows (pass 1 / B) or whole (k, ri) column pairs (pass 2 / A), reads them into fragments, and only then writes.
// The DFT matrices (1/sqrt(S) folded into each pass = ortho; Hermitian weights 1, 2, .., 2, 1 and the ignored imaginary parts of
// the DC / Nyquist columns folded into the c2r matrix) are built on the host in fp16, already in A-fragment order.
// Accuracy: fp16 twiddles add one rounding (2^-12 relative) per product to the fp16 rounding of the stored result; measured
// against torch.fft in tests/test_gpu_kernels.py::test_fft2 (same 2e-3 of max bound as the register FFT).
#include <cuda.h>
#include <math.h>

#ifndef S2V_FFT_EXP
#define S2V_FFT_EXP 0      // development: bit 0 = rfft2 without its global loads, bit 1 = without its global stores (profiles/r2c_summary.md)
#endif

#include "common.cuh"

namespace s2v {
namespace fftmma {

constexpr int kThreads = 128, kWarps = 4, kFragsTotal = 90;

template <int S>
struct Cfg {
  static constexpr int K1 = S / 2 + 1;                 // stored half-spectrum columns
  static constexpr int KP = (S + 15) / 16 * 16;        // K extent of the w / h / kh contractions
  static constexpr int KT = KP / 16;
  static constexpr int MT1 = (K1 + 7) / 8;             // pass 1 m-tiles: 8 k's each, rows 0-7 = re, rows 8-15 = im
  static constexpr int MT2 = (S + 15) / 16;            // m-tiles over kh / h / w
  static constexpr int K2P = (2 * K1 + 15) / 16 * 16;  // K extent of the c2r contraction over (k, ri)
  static constexpr int KT2 = K2P / 16;
  static constexpr int RS = K1 * 32 + 16;              // tile row stride in bytes: 816 / 432 / 240 = 48, 48, 112 mod 128 -> eight
                                                       // consecutive rows start in eight different 16-byte bank groups
  static constexpr int ROWS = KP;                      // rows >= S stay zero (K padding of the contractions over h / kh)
  static constexpr int TILE = ROWS * RS + 256;         // + tail: the K padding of pass 1 / B reads past the last row
  static constexpr int N_F1 = MT1 * KT, N_G = MT2 * KT, N_A2 = MT2 * KT2;
  static constexpr int COUNT = N_F1 + 4 * N_G + N_A2;  // 60 / 24 / 6 fragments
  static constexpr int BASE = S == 48 ? 0 : S == 24 ? 60 : 84;
  static constexpr int O_F1 = BASE, O_GR = O_F1 + N_F1, O_GI = O_GR + N_G, O_WR = O_GI + N_G, O_WI = O_WR + N_G, O_A2 = O_WI + N_G;
};
static_assert(Cfg<48>::COUNT == 60 && Cfg<24>::COUNT == 24 && Cfg<12>::COUNT == 6, "fragment table layout");

// A fragments of every DFT matrix: [fragment][lane] -> 4 registers of mma.m16n8k16 (filled by fft_mma_init, per device)
__device__ uint4 g_frag[kFragsTotal * 32];

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// B fragments of KTN k-tiles whose K rows are `stride` bytes apart, starting at shared address `a` (16-byte rows)
template <int KTN>
__device__ __forceinline__ void load_b(uint32_t a, uint32_t stride, int lane, uint32_t (&b)[KTN][2]) {
#pragma unroll
  for (int kt = 0; kt + 1 < KTN; kt += 2) ldsm_x4_t(a + (uint32_t)(kt * 16 + lane) * stride, b[kt][0], b[kt][1], b[kt + 1][0], b[kt + 1][1]);
  if (KTN & 1) ldsm_x2_t(a + (uint32_t)((KTN - 1) * 16 + (lane & 15)) * stride, b[KTN - 1][0], b[KTN - 1][1]);
}

// The complex pass over the tile rows (forward pass 2 along h, inverse pass A along kh) for column pair k:
//   zr = Ar * Bre - Ai * Bim,  zi = Ar * Bim + Ai * Bre     (rows = MT2 m-tiles, columns = 8 channels)
template <int S>
__device__ __forceinline__ void complex_pass(uint32_t tile, int k, int lane, const uint4 (&ar)[Cfg<S>::MT2][Cfg<S>::KT],
                                             const uint4 (&ai)[Cfg<S>::MT2][Cfg<S>::KT], float (&zr)[Cfg<S>::MT2][4],
                                             float (&zi)[Cfg<S>::MT2][4], uint32_t pitch = Cfg<S>::RS) {
  using C = Cfg<S>;
  uint32_t bre[C::KT][2], bim[C::KT][2], bin[C::KT][2];
  load_b<C::KT>(tile + (uint32_t)(2 * k) * 16u, pitch, lane, bre);
  load_b<C::KT>(tile + (uint32_t)(2 * k + 1) * 16u, pitch, lane, bim);
#pragma unroll
  for (int kt = 0; kt < C::KT; ++kt) { bin[kt][0] = bim[kt][0] ^ 0x80008000u; bin[kt][1] = bim[kt][1] ^ 0x80008000u; }   // -Bim
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) { zr[mt][i] = 0.f; zi[mt][i] = 0.f; }
  // issue order: 2 * MT2 independent accumulators round-robin, so that an accumulator is touched again only every 2 * MT2 MMAs
  // (mma.sync issues in program order; back-to-back MMAs on one accumulator wait out the whole pipeline latency)
#pragma unroll
  for (int kt = 0; kt < C::KT; ++kt) {
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt) mma16816(zr[mt], ar[mt][kt], bre[kt][0], bre[kt][1]);
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt) mma16816(zi[mt], ar[mt][kt], bim[kt][0], bim[kt][1]);
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt) mma16816(zr[mt], ai[mt][kt], bin[kt][0], bin[kt][1]);
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt) mma16816(zi[mt], ai[mt][kt], bre[kt][0], bre[kt][1]);
  }
}

// Blocks are PERSISTENT (BPS blocks per SM, block b takes tiles b, b + grid, ...) and register-light: the A fragments of the DFT
// matrices are copied to shared memory once per block and re-read from there in front of each pass (12 - 18 LDS.128 per thread and
// tile), so a thread needs ~120 registers and FOUR 48 x 48 blocks (16 warps) share an SM, each with one 39 KB tile + 15 KB of tables.
// Versions measured on B200 (us for rfft2 / irfft2, 48 x 48 x 48 channels; register FFT of fft2d.cu: 23.6 / 37.1 at B = 128, 53.2 /
// 75.9 at B = 256; profiles/r2c_summary.md):
//   v1  one block per tile, tables from L2 in front of each pass, 3 - 4 blocks per SM          21.4 / 32.2 (B = 128)
//       ncu: tensor pipe 44 % / 29 % active while resident, long_scoreboard the top stall (tile DRAM latency + two L2 round trips
//       for the tables + the rolled cp.async loop, whose address registers every LDGSTS has to release first: 15 % of the samples)
//   v2  persistent, double-buffered tile, all tables resident in 250 registers, 2 blocks per SM  24.4 / 36.0 (B = 128), 45.0 / 67.5 (B = 256)
//       ncu: 7.2 stall cycles per issued instruction with two warps per scheduler, tensor pipe 41 % / 27 %
//   v3  this one
template <int S> constexpr int blocks_per_sm() { return S == 48 ? 4 : S == 24 ? 6 : 8; }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// block prologue: zero the tile (pads, K-padding rows and the tail stay zero for the block's lifetime) and copy NF fragments
// [fragment][lane] of the table, starting at fragment F0, behind it
template <int S>
__device__ __forceinline__ void block_prologue(uint32_t tile, int f0, int nf) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < Cfg<S>::TILE / 16; i += kThreads) st_shared_u4(tile + (uint32_t)i * 16u, z);
  for (int i = threadIdx.x; i < nf * 32; i += kThreads) st_shared_u4(tile + Cfg<S>::TILE + (uint32_t)i * 16u, g_frag[f0 * 32 + i]);
}
template <int MT, int KTN>
__device__ __forceinline__ void load_a(uint32_t tab, int frag0, int lane, uint4 (&a)[MT][KTN]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int kt = 0; kt < KTN; ++kt) a[mt][kt] = ld_shared_u4(tab + (uint32_t)((frag0 + mt * KTN + kt) * 32 + lane) * 16u);
}

// ---- forward passes (shared by the cp.async kernel and the TMA kernel) ------------------------------------------------------
// pass 1: real-input DFT along w.  X rows at xb + h * xpitch (K rows = pixels, 16 bytes apart), Y rows at yb + h * RS; in place when
// xb == yb and xpitch == RS (a warp reads its rows into fragments before it writes them).  Two rows per step = 2 * MT1 independent
// accumulators in flight (mma.sync issues in program order; back-to-back MMAs on one accumulator wait out the pipeline latency).
template <int S>
__device__ __forceinline__ void rfft_pass1(uint32_t xb, uint32_t xpitch, uint32_t yb, const uint4 (&a1)[Cfg<S>::MT1][Cfg<S>::KT], int warp, int lane) {
  using C = Cfg<S>;
  const int g = lane >> 2, t = lane & 3;
  for (int h = warp; h < S; h += 2 * kWarps) {
    const bool two = h + kWarps < S;                             // (warp-uniform; an odd last step repeats its row, stores once)
    const int h1 = two ? h + kWarps : h;
    uint32_t b0[C::KT][2], b1[C::KT][2];
    load_b<C::KT>(xb + (uint32_t)h * xpitch, 16u, lane, b0);
    load_b<C::KT>(xb + (uint32_t)h1 * xpitch, 16u, lane, b1);
    float acc0[C::MT1][4], acc1[C::MT1][4];
#pragma unroll
    for (int mt = 0; mt < C::MT1; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc0[mt][i] = 0.f; acc1[mt][i] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < C::KT; ++kt) {
#pragma unroll
      for (int mt = 0; mt < C::MT1; ++mt) mma16816(acc0[mt], a1[mt][kt], b0[kt][0], b0[kt][1]);
#pragma unroll
      for (int mt = 0; mt < C::MT1; ++mt) mma16816(acc1[mt], a1[mt][kt], b1[kt][0], b1[kt][1]);
    }
    __syncwarp();
    const uint32_t row0 = yb + (uint32_t)h * C::RS, row1 = yb + (uint32_t)h1 * C::RS;
#pragma unroll
    for (int mt = 0; mt < C::MT1; ++mt) {
      const int k = mt * 8 + g;
      if (k < C::K1) {
        const uint32_t o = (uint32_t)(2 * k) * 16u + (uint32_t)t * 4u;
        st_shared_u32(row0 + o, pack_h2(acc0[mt][0], acc0[mt][1]));            // re, channels 2t, 2t+1
        st_shared_u32(row0 + o + 16u, pack_h2(acc0[mt][2], acc0[mt][3]));      // im
        if (two) {
          st_shared_u32(row1 + o, pack_h2(acc1[mt][0], acc1[mt][1]));
          st_shared_u32(row1 + o + 16u, pack_h2(acc1[mt][2], acc1[mt][3]));
        }
      }
    }
  }
}
// pass 2: complex DFT along h for column k of the Y tile, straight to global (op = the tile's channel 0 of spec, image n)
template <int S>
__device__ __forceinline__ void rfft_pass2(uint32_t yb, const uint4 (&ar)[Cfg<S>::MT2][Cfg<S>::KT], const uint4 (&ai)[Cfg<S>::MT2][Cfg<S>::KT],
                                           const View& sp, __half* op0, int warp, int lane) {
  using C = Cfg<S>;
  const int g = lane >> 2, t = lane & 3;
  __half* op = op0 + 4 * t;
  for (int k = warp; k < C::K1; k += kWarps) {
    float zr[C::MT2][4], zi[C::MT2][4];
    complex_pass<S>(yb, k, lane, ar, ai, zr, zi);
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt) {
      const int kh0 = mt * 16 + g, kh1 = kh0 + 8;
#if (S2V_FFT_EXP & 2)      // experiment: keep the arithmetic alive, store (practically) nothing
      if (kh0 < S && zr[mt][0] == 1234.5f) *reinterpret_cast<uint2*>(op + kh0 * sp.sh + k * sp.sw) = make_uint2(pack_h2(zr[mt][0], zi[mt][0]), pack_h2(zr[mt][1], zi[mt][1]));
      if (kh1 < S && zr[mt][2] == 1234.5f) *reinterpret_cast<uint2*>(op + kh1 * sp.sh + k * sp.sw) = make_uint2(pack_h2(zr[mt][2], zi[mt][2]), pack_h2(zr[mt][3], zi[mt][3]));
#else
      if (kh0 < S) *reinterpret_cast<uint2*>(op + kh0 * sp.sh + k * sp.sw) = make_uint2(pack_h2(zr[mt][0], zi[mt][0]), pack_h2(zr[mt][1], zi[mt][1]));
      if (kh1 < S) *reinterpret_cast<uint2*>(op + kh1 * sp.sh + k * sp.sw) = make_uint2(pack_h2(zr[mt][2], zi[mt][2]), pack_h2(zr[mt][3], zi[mt][3]));
#endif
    }
  }
}

// x [N,S,S,C] -> spec [N,S,S/2+1,2C]; tiles = N * C / 8 (image-major), grid = min(tiles, BPS * #SMs)
template <int S>
__global__ void __launch_bounds__(kThreads, blocks_per_sm<S>()) rfft2_mma_kernel(View x, View sp, int cblocks, int tiles) {
  using C = Cfg<S>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t tile = (uint32_t)__cvta_generic_to_shared(smem_raw), tab = tile + C::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  block_prologue<S>(tile, C::O_F1, C::N_F1 + 2 * C::N_G);       // constant tables: copied ahead of the PDL wait
  __syncthreads();
  pdl_wait();
  for (int ti = blockIdx.x; ti < tiles; ti += gridDim.x) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    {
      const __half* xp = x.p + n * x.sn + ch0;
      // fully unrolled: a rolled loop re-uses the address registers, and every LDGSTS then waits for the previous one to have read them
#pragma unroll
      for (int j = 0; j < (S * S + kThreads - 1) / kThreads; ++j) {
        const int i = threadIdx.x + j * kThreads;
        if (i < S * S) {
          const int h = i / S, w = i - h * S;
#if !(S2V_FFT_EXP & 1)
          cp_async16(tile + (uint32_t)(h * C::RS + w * 16), xp + h * x.sh + w * x.sw);
#else
          if (h < 0) cp_async16(tile + (uint32_t)(h * C::RS + w * 16), xp + h * x.sh + w * x.sw);
#endif
        }
      }
      cp_async_commit();
    }
    {
      uint4 a1[C::MT1][C::KT];
      load_a<C::MT1, C::KT>(tab, 0, lane, a1);
      cp_async_wait_all();
      __syncthreads();
      rfft_pass1<S>(tile, C::RS, tile, a1, warp, lane);       // in place
    }
    uint4 ar[C::MT2][C::KT], ai[C::MT2][C::KT];
    load_a<C::MT2, C::KT>(tab, C::N_F1, lane, ar);
    load_a<C::MT2, C::KT>(tab, C::N_F1 + C::N_G, lane, ai);
    __syncthreads();
    rfft_pass2<S>(tile, ar, ai, sp, sp.p + n * sp.sn + 2 * ch0, warp, lane);
    __syncthreads();                           // every warp is done reading the tile before the next tile's loads land in it
  }
}

// ---- the forward transform with the tile gathered by TMA ---------------------------------------------------------------------
// The 16-byte channel slices make the cp.async loads of the kernel above the largest part of its time (19 of 46 us at 48 x 48,
// B = 256: 24 cache lines per warp request).  One bulk-tensor load {8 ch, S, S, 1} per tile moves the same slices at one 16-byte row
// per clock per SM (tools/mb_tma_rows.cu: 2 326 clocks per 36 KB tile) WITHOUT occupying the LSU or an issue slot, into a dense
// X buffer (TMA destinations are 128-byte aligned, so the padded in-place pitch is not an option); pass 1 reads X and writes the
// padded Y tile, the next tile's load is issued as soon as pass 1 is done and lands during pass 2.  76 KB -> 2 blocks per SM, so all
// A fragments stay in registers.
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  for (unsigned spins = 0; spins < (1u << 26); ++spins) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
template <int S> constexpr int x_bytes() { return S * S * 16 + 256; }          // dense tile + tail (K padding of pass 1 reads past the last row)

template <int S> constexpr int blocks_per_sm_tma() { return S == 48 ? 2 : S == 24 ? 4 : 8; }
template <int S>
__global__ void __launch_bounds__(kThreads, blocks_per_sm_tma<S>()) rfft2_mma_tma_kernel(const __grid_constant__ CUtensorMap tmx, View sp, int cblocks, int tiles) {
  using C = Cfg<S>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t xb = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u, yb = xb + x_bytes<S>(), bar = yb + C::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < (x_bytes<S>() + C::TILE) / 16; i += kThreads) st_shared_u4(xb + (uint32_t)i * 16u, z);
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint4 a1[C::MT1][C::KT], ar[C::MT2][C::KT], ai[C::MT2][C::KT];     // constant tables: fetched ahead of the PDL wait
#pragma unroll
  for (int mt = 0; mt < C::MT1; ++mt)
#pragma unroll
    for (int kt = 0; kt < C::KT; ++kt) a1[mt][kt] = g_frag[(C::O_F1 + mt * C::KT + kt) * 32 + lane];
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int kt = 0; kt < C::KT; ++kt) {
      ar[mt][kt] = g_frag[(C::O_GR + mt * C::KT + kt) * 32 + lane];
      ai[mt][kt] = g_frag[(C::O_GI + mt * C::KT + kt) * 32 + lane];
    }
  __syncthreads();
  pdl_wait();
  auto issue_load = [&](int ti) {              // one thread; every generic access to X is behind a block barrier at this point
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(S * S * 16)) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(xb), "l"(&tmx), "r"(bar), "r"(ch0), "r"(0), "r"(0), "r"(n) : "memory");
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < tiles) issue_load(blockIdx.x);
  int it = 0;
  for (int ti = blockIdx.x; ti < tiles; ti += gridDim.x, ++it) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    mbar_wait_parity(bar, (uint32_t)(it & 1));
    rfft_pass1<S>(xb, (uint32_t)(S * 16), yb, a1, warp, lane);
    __syncthreads();                           // X is consumed, Y is complete
    if (threadIdx.x == 0 && ti + (int)gridDim.x < tiles) issue_load(ti + gridDim.x);
    rfft_pass2<S>(yb, ar, ai, sp, sp.p + n * sp.sn + 2 * ch0, warp, lane);
    __syncthreads();                           // every warp is done reading Y before the next tile's pass 1 writes it
  }
}

// spec [N,S,S/2+1,2C] -> y [N,S,S,C] (+ add); tiles = N * C / 8 (image-major), grid = min(tiles, BPS * #SMs)
template <int S>
__global__ void __launch_bounds__(kThreads, blocks_per_sm<S>()) irfft2_mma_kernel(View sp, View add, View y, int cblocks, int tiles) {
  using C = Cfg<S>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t tile = (uint32_t)__cvta_generic_to_shared(smem_raw), tab = tile + C::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  block_prologue<S>(tile, C::O_WR, 2 * C::N_G + C::N_A2);
  __syncthreads();
  pdl_wait();
  constexpr int kItems = S * C::K1;          // one item = the 8 channels of one (kh, k): 32 bytes, (re, im) interleaved in global
  const bool has_add = add.p != nullptr;
  for (int ti = blockIdx.x; ti < tiles; ti += gridDim.x) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    {
      const __half* spp = sp.p + n * sp.sn + 2 * ch0;
#pragma unroll
      for (int j = 0; j < (kItems + kThreads - 1) / kThreads; ++j) {      // unrolled: see rfft2_mma_kernel
        const int i = threadIdx.x + j * kThreads;
        if (i < kItems) {
          const int kh = i / C::K1, k = i - kh * C::K1;
          const __half* src = spp + kh * sp.sh + k * sp.sw;
          const uint32_t d = tile + (uint32_t)(kh * C::RS + k * 32);
          cp_async16(d, src);
          cp_async16(d + 16u, src + 8);
        }
      }
      cp_async_commit();
    }
    {
      uint4 ar[C::MT2][C::KT], ai[C::MT2][C::KT];
      load_a<C::MT2, C::KT>(tab, 0, lane, ar);
      load_a<C::MT2, C::KT>(tab, C::N_G, lane, ai);
      cp_async_wait_all();
      // de-interleave this thread's own items in place: [c0r c0i .. c7r c7i] -> [c0r .. c7r][c0i .. c7i]
#pragma unroll 2
      for (int i = threadIdx.x; i < kItems; i += kThreads) {
        const int kh = i / C::K1, k = i - kh * C::K1;
        const uint32_t d = tile + (uint32_t)(kh * C::RS + k * 32);
        const uint4 u0 = ld_shared_u4(d), u1 = ld_shared_u4(d + 16u);
        st_shared_u4(d, make_uint4(__byte_perm(u0.x, u0.y, 0x5410), __byte_perm(u0.z, u0.w, 0x5410),
                                   __byte_perm(u1.x, u1.y, 0x5410), __byte_perm(u1.z, u1.w, 0x5410)));
        st_shared_u4(d + 16u, make_uint4(__byte_perm(u0.x, u0.y, 0x7632), __byte_perm(u0.z, u0.w, 0x7632),
                                         __byte_perm(u1.x, u1.y, 0x7632), __byte_perm(u1.z, u1.w, 0x7632)));
      }
      __syncthreads();
      // ---- pass A: inverse complex DFT along kh for column k, in place -------------------------------------------------
      for (int k = warp; k < C::K1; k += kWarps) {
        float zr[C::MT2][4], zi[C::MT2][4];
        complex_pass<S>(tile, k, lane, ar, ai, zr, zi);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) {
          const int h0 = mt * 16 + g, h1 = h0 + 8;
          const uint32_t c0 = tile + (uint32_t)(2 * k) * 16u + (uint32_t)t * 4u;
          if (h0 < S) {
            st_shared_u32(c0 + (uint32_t)h0 * C::RS, pack_h2(zr[mt][0], zr[mt][1]));
            st_shared_u32(c0 + (uint32_t)h0 * C::RS + 16u, pack_h2(zi[mt][0], zi[mt][1]));
          }
          if (h1 < S) {
            st_shared_u32(c0 + (uint32_t)h1 * C::RS, pack_h2(zr[mt][2], zr[mt][3]));
            st_shared_u32(c0 + (uint32_t)h1 * C::RS + 16u, pack_h2(zi[mt][2], zi[mt][3]));
          }
        }
      }
    }
    uint4 a2[C::MT2][C::KT2];
    load_a<C::MT2, C::KT2>(tab, 2 * C::N_G, lane, a2);
    __syncthreads();
    // ---- pass B: complex-to-real along w for tile row h (K = (k, re|im)), + residual, to global ---------------------------
    const __half* ap = has_add ? add.p + n * add.sn + ch0 + 2 * t : nullptr;
    __half* yp = y.p + n * y.sn + ch0 + 2 * t;
    uint32_t rn[2][C::MT2][2];
    auto load_res = [&](int r, int h) {
#pragma unroll
      for (int mt = 0; mt < C::MT2; ++mt) {
        const int w0 = mt * 16 + g, w1 = w0 + 8;
        rn[r][mt][0] = (has_add && h < S && w0 < S) ? *reinterpret_cast<const uint32_t*>(ap + h * add.sh + w0 * add.sw) : 0u;
        rn[r][mt][1] = (has_add && h < S && w1 < S) ? *reinterpret_cast<const uint32_t*>(ap + h * add.sh + w1 * add.sw) : 0u;
      }
    };
    load_res(0, warp);
    load_res(1, warp + kWarps);
    for (int h = warp; h < S; h += 2 * kWarps) {                 // two rows per step = 2 * MT2 independent accumulators in flight
      const bool two = h + kWarps < S;                             // (warp-uniform; an odd last step repeats its row, stores once)
      const int hh[2] = {h, two ? h + kWarps : h};
      uint32_t rc[2][C::MT2][2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) { rc[r][mt][0] = rn[r][mt][0]; rc[r][mt][1] = rn[r][mt][1]; }
      load_res(0, h + 2 * kWarps);                                 // the next step's residual is in flight behind this step's MMAs
      load_res(1, h + 3 * kWarps);
      uint32_t b[2][C::KT2][2];
      load_b<C::KT2>(tile + (uint32_t)hh[0] * C::RS, 16u, lane, b[0]);
      load_b<C::KT2>(tile + (uint32_t)hh[1] * C::RS, 16u, lane, b[1]);
      float acc[2][C::MT2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[r][mt][i] = 0.f;
#pragma unroll
      for (int kt = 0; kt < C::KT2; ++kt)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int mt = 0; mt < C::MT2; ++mt) mma16816(acc[r][mt], a2[mt][kt], b[r][kt][0], b[r][kt][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) {
          const int w0 = mt * 16 + g, w1 = w0 + 8;
          const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&rc[r][mt][0]));
          const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&rc[r][mt][1]));
          if (w0 < S) *reinterpret_cast<uint32_t*>(yp + hh[r] * y.sh + w0 * y.sw) = pack_h2(acc[r][mt][0] + r0.x, acc[r][mt][1] + r0.y);
          if (w1 < S) *reinterpret_cast<uint32_t*>(yp + hh[r] * y.sh + w1 * y.sw) = pack_h2(acc[r][mt][2] + r1.x, acc[r][mt][3] + r1.y);
        }
      }
    }
    __syncthreads();                           // every warp is done reading the tile before the next tile's loads land in it
  }
}

// ---- the inverse transform with all three global streams on TMA --------------------------------------------------------------
// spec tile {2 x 8 ch, S/2+1, S} (32-byte inner rows) -> SP (dense 32 (S/2+1)-byte rows: two-way bank conflicts on pass A's fragment
// loads, accepted), residual tile {8 ch, S, S} -> R; pass A in place in SP, pass B adds its result into R in place (4-byte shared
// accesses instead of 4-byte global ones: 8 cache lines per warp request), one bulk-tensor store R -> y.  75 KB -> 2 blocks per SM.
template <int S> constexpr int sp_pitch() { return Cfg<S>::K1 * 32; }
template <int S> constexpr int sp_bytes() { return (Cfg<S>::KP * sp_pitch<S>() + 256 + 127) / 128 * 128; }      // rows >= S (K padding over kh) and the tail stay zero
template <int S> constexpr int r_bytes() { return S * S * 16; }

// STORE_TMA = false: y leaves through 4-byte global stores from the registers instead (the TMA unit moves one 16 / 32-byte row per
// clock, and with the store on it as well it is the unit that bounds this kernel; the LSU is idle otherwise)
template <int S, bool STORE_TMA>
__global__ void __launch_bounds__(kThreads, blocks_per_sm_tma<S>()) irfft2_mma_tma_kernel(const __grid_constant__ CUtensorMap tmsp,
                                                                                         const __grid_constant__ CUtensorMap tmadd,
                                                                                         const __grid_constant__ CUtensorMap tmy, View y, int has_add,
                                                                                         int cblocks, int tiles) {
  using C = Cfg<S>;
  constexpr uint32_t kPitch = sp_pitch<S>();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t rb = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u, spb = rb + r_bytes<S>(), bar = spb + sp_bytes<S>();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < (r_bytes<S>() + sp_bytes<S>()) / 16; i += kThreads) st_shared_u4(rb + (uint32_t)i * 16u, z);
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmsp) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
    if (has_add) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmadd) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint4 ar[C::MT2][C::KT], ai[C::MT2][C::KT], a2[C::MT2][C::KT2];     // constant tables: fetched ahead of the PDL wait
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt) {
#pragma unroll
    for (int kt = 0; kt < C::KT; ++kt) {
      ar[mt][kt] = g_frag[(C::O_WR + mt * C::KT + kt) * 32 + lane];
      ai[mt][kt] = g_frag[(C::O_WI + mt * C::KT + kt) * 32 + lane];
    }
#pragma unroll
    for (int kt = 0; kt < C::KT2; ++kt) a2[mt][kt] = g_frag[(C::O_A2 + mt * C::KT2 + kt) * 32 + lane];
  }
  __syncthreads();
  pdl_wait();
  constexpr int kItems = S * C::K1;
  int it = 0;
  for (int ti = blockIdx.x; ti < tiles; ti += gridDim.x, ++it) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    if (threadIdx.x == 0) {                    // every generic access to SP / R of the previous tile is behind the loop's last barrier
      if (STORE_TMA) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the previous tile's store has read R
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kItems * 32 + (has_add ? r_bytes<S>() : 0))) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(spb), "l"(&tmsp), "r"(bar), "r"(2 * ch0), "r"(0), "r"(0), "r"(n) : "memory");
      if (has_add)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(rb), "l"(&tmadd), "r"(bar), "r"(ch0), "r"(0), "r"(0), "r"(n) : "memory");
    }
    mbar_wait_parity(bar, (uint32_t)(it & 1));
    // de-interleave in place: [c0r c0i .. c7r c7i] -> [c0r .. c7r][c0i .. c7i]
#pragma unroll 2
    for (int i = threadIdx.x; i < kItems; i += kThreads) {
      const int kh = i / C::K1, k = i - kh * C::K1;
      const uint32_t d = spb + (uint32_t)kh * kPitch + (uint32_t)k * 32u;
      const uint4 u0 = ld_shared_u4(d), u1 = ld_shared_u4(d + 16u);
      st_shared_u4(d, make_uint4(__byte_perm(u0.x, u0.y, 0x5410), __byte_perm(u0.z, u0.w, 0x5410),
                                 __byte_perm(u1.x, u1.y, 0x5410), __byte_perm(u1.z, u1.w, 0x5410)));
      st_shared_u4(d + 16u, make_uint4(__byte_perm(u0.x, u0.y, 0x7632), __byte_perm(u0.z, u0.w, 0x7632),
                                       __byte_perm(u1.x, u1.y, 0x7632), __byte_perm(u1.z, u1.w, 0x7632)));
    }
    __syncthreads();
    // ---- pass A: inverse complex DFT along kh for column k, in place ---------------------------------------------------
    for (int k = warp; k < C::K1; k += kWarps) {
      float zr[C::MT2][4], zi[C::MT2][4];
      complex_pass<S>(spb, k, lane, ar, ai, zr, zi, kPitch);
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < C::MT2; ++mt) {
        const int h0 = mt * 16 + g, h1 = h0 + 8;
        const uint32_t c0 = spb + (uint32_t)(2 * k) * 16u + (uint32_t)t * 4u;
        if (h0 < S) {
          st_shared_u32(c0 + (uint32_t)h0 * kPitch, pack_h2(zr[mt][0], zr[mt][1]));
          st_shared_u32(c0 + (uint32_t)h0 * kPitch + 16u, pack_h2(zi[mt][0], zi[mt][1]));
        }
        if (h1 < S) {
          st_shared_u32(c0 + (uint32_t)h1 * kPitch, pack_h2(zr[mt][2], zr[mt][3]));
          st_shared_u32(c0 + (uint32_t)h1 * kPitch + 16u, pack_h2(zi[mt][2], zi[mt][3]));
        }
      }
    }
    __syncthreads();
    // ---- pass B: complex-to-real along w for tile row h (K = (k, re|im)), added into R in place --------------------------
    for (int h = warp; h < S; h += 2 * kWarps) {                 // two rows per step = 2 * MT2 independent accumulators in flight
      const bool two = h + kWarps < S;
      const int hh[2] = {h, two ? h + kWarps : h};
      uint32_t b[2][C::KT2][2];
      load_b<C::KT2>(spb + (uint32_t)hh[0] * kPitch, 16u, lane, b[0]);
      load_b<C::KT2>(spb + (uint32_t)hh[1] * kPitch, 16u, lane, b[1]);
      float acc[2][C::MT2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[r][mt][i] = 0.f;
#pragma unroll
      for (int kt = 0; kt < C::KT2; ++kt)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int mt = 0; mt < C::MT2; ++mt) mma16816(acc[r][mt], a2[mt][kt], b[r][kt][0], b[r][kt][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) {
          const int w0 = mt * 16 + g, w1 = w0 + 8;
          const uint32_t p0 = rb + (uint32_t)((hh[r] * S + w0) * 16 + t * 4), p1 = rb + (uint32_t)((hh[r] * S + w1) * 16 + t * 4);
          float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
          if (has_add) {
            uint32_t u0 = 0u, u1 = 0u;
            if (w0 < S) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u0) : "r"(p0) : "memory");
            if (w1 < S) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u1) : "r"(p1) : "memory");
            r0 = __half22float2(*reinterpret_cast<const __half2*>(&u0));
            r1 = __half22float2(*reinterpret_cast<const __half2*>(&u1));
          }
          if (STORE_TMA) {
            if (w0 < S) st_shared_u32(p0, pack_h2(acc[r][mt][0] + r0.x, acc[r][mt][1] + r0.y));
            if (w1 < S) st_shared_u32(p1, pack_h2(acc[r][mt][2] + r1.x, acc[r][mt][3] + r1.y));
          } else {
            __half* yp = y.p + n * y.sn + hh[r] * y.sh + ch0 + 2 * t;
            if (w0 < S) *reinterpret_cast<uint32_t*>(yp + w0 * y.sw) = pack_h2(acc[r][mt][0] + r0.x, acc[r][mt][1] + r0.y);
            if (w1 < S) *reinterpret_cast<uint32_t*>(yp + w1 * y.sw) = pack_h2(acc[r][mt][2] + r1.x, acc[r][mt][3] + r1.y);
          }
        }
      }
    }
    __syncthreads();                           // R holds the finished tile; SP is free
    if (STORE_TMA && threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   ::"l"(&tmy), "r"(rb), "r"(ch0), "r"(0), "r"(0), "r"(n) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (STORE_TMA && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the last store is done with R before the block exits
}

// Pipelined form of the inverse kernel: TWO spec buffers (the next tile's spec load is issued at the top of the current tile), the
// residual tile loaded one phase ahead (issued as soon as pass B is done with the buffer), y through 4-byte stores from the registers
// (fire and forget).  112 KB -> still 2 blocks per SM.  The default (S2V_FFT_TMA_STORE=0): 27.7 / 46.7 us at 48 x 48, B = 128 / 256, against
// 33.7 / 60.1 us for the single-buffered kernel above and 37.1 / 75.9 us for the register FFT.
template <int S>
__global__ void __launch_bounds__(kThreads, blocks_per_sm_tma<S>()) irfft2_mma_tma2_kernel(const __grid_constant__ CUtensorMap tmsp,
                                                                                          const __grid_constant__ CUtensorMap tmadd, View y, int has_add,
                                                                                          int cblocks, int tiles) {
  using C = Cfg<S>;
  constexpr uint32_t kPitch = sp_pitch<S>();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t rb = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u, sp0 = rb + r_bytes<S>(),
                 bar_sp = sp0 + 2 * sp_bytes<S>(), bar_r = bar_sp + 16u;
  static_assert(sp_bytes<S>() % 128 == 0, "second spec buffer must be 128-byte aligned");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < (r_bytes<S>() + 2 * sp_bytes<S>()) / 16; i += kThreads) st_shared_u4(rb + (uint32_t)i * 16u, z);
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmsp) : "memory");
    if (has_add) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmadd) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_sp) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_sp + 8u) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_r) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint4 ar[C::MT2][C::KT], ai[C::MT2][C::KT], a2[C::MT2][C::KT2];     // constant tables: fetched ahead of the PDL wait
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt) {
#pragma unroll
    for (int kt = 0; kt < C::KT; ++kt) {
      ar[mt][kt] = g_frag[(C::O_WR + mt * C::KT + kt) * 32 + lane];
      ai[mt][kt] = g_frag[(C::O_WI + mt * C::KT + kt) * 32 + lane];
    }
#pragma unroll
    for (int kt = 0; kt < C::KT2; ++kt) a2[mt][kt] = g_frag[(C::O_A2 + mt * C::KT2 + kt) * 32 + lane];
  }
  __syncthreads();
  pdl_wait();
  constexpr int kItems = S * C::K1;
  auto issue_spec = [&](int ti, int buf) {     // one thread; every generic access to that buffer is behind a block barrier
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_sp + 8u * buf), "r"((uint32_t)(kItems * 32)) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sp0 + (uint32_t)buf * sp_bytes<S>()), "l"(&tmsp), "r"(bar_sp + 8u * buf), "r"(2 * ch0), "r"(0), "r"(0), "r"(n) : "memory");
  };
  auto issue_res = [&](int ti) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_r), "r"((uint32_t)r_bytes<S>()) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(rb), "l"(&tmadd), "r"(bar_r), "r"(ch0), "r"(0), "r"(0), "r"(n) : "memory");
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < tiles) {
    issue_spec(blockIdx.x, 0);
    if (has_add) issue_res(blockIdx.x);
  }
  int it = 0;
  for (int ti = blockIdx.x; ti < tiles; ti += gridDim.x, ++it) {
    const int n = ti / cblocks, ch0 = (ti - n * cblocks) * 8, buf = it & 1;
    const uint32_t spb = sp0 + (uint32_t)buf * sp_bytes<S>();
    // the other spec buffer was last touched by the previous tile, which ended behind a block barrier: prefetch the next tile into it
    if (threadIdx.x == 0 && ti + (int)gridDim.x < tiles) issue_spec(ti + gridDim.x, buf ^ 1);
    mbar_wait_parity(bar_sp + 8u * buf, (uint32_t)((it >> 1) & 1));
    // de-interleave in place: [c0r c0i .. c7r c7i] -> [c0r .. c7r][c0i .. c7i]
#pragma unroll 2
    for (int i = threadIdx.x; i < kItems; i += kThreads) {
      const int kh = i / C::K1, k = i - kh * C::K1;
      const uint32_t d = spb + (uint32_t)kh * kPitch + (uint32_t)k * 32u;
      const uint4 u0 = ld_shared_u4(d), u1 = ld_shared_u4(d + 16u);
      st_shared_u4(d, make_uint4(__byte_perm(u0.x, u0.y, 0x5410), __byte_perm(u0.z, u0.w, 0x5410),
                                 __byte_perm(u1.x, u1.y, 0x5410), __byte_perm(u1.z, u1.w, 0x5410)));
      st_shared_u4(d + 16u, make_uint4(__byte_perm(u0.x, u0.y, 0x7632), __byte_perm(u0.z, u0.w, 0x7632),
                                       __byte_perm(u1.x, u1.y, 0x7632), __byte_perm(u1.z, u1.w, 0x7632)));
    }
    __syncthreads();
    // ---- pass A: inverse complex DFT along kh for column k, in place ---------------------------------------------------
    for (int k = warp; k < C::K1; k += kWarps) {
      float zr[C::MT2][4], zi[C::MT2][4];
      complex_pass<S>(spb, k, lane, ar, ai, zr, zi, kPitch);
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < C::MT2; ++mt) {
        const int h0 = mt * 16 + g, h1 = h0 + 8;
        const uint32_t c0 = spb + (uint32_t)(2 * k) * 16u + (uint32_t)t * 4u;
        if (h0 < S) {
          st_shared_u32(c0 + (uint32_t)h0 * kPitch, pack_h2(zr[mt][0], zr[mt][1]));
          st_shared_u32(c0 + (uint32_t)h0 * kPitch + 16u, pack_h2(zi[mt][0], zi[mt][1]));
        }
        if (h1 < S) {
          st_shared_u32(c0 + (uint32_t)h1 * kPitch, pack_h2(zr[mt][2], zr[mt][3]));
          st_shared_u32(c0 + (uint32_t)h1 * kPitch + 16u, pack_h2(zi[mt][2], zi[mt][3]));
        }
      }
    }
    __syncthreads();
    if (has_add) mbar_wait_parity(bar_r, (uint32_t)(it & 1));
    // ---- pass B: complex-to-real along w for tile row h (K = (k, re|im)) + residual (shared memory) -> y (global) ---------
    for (int h = warp; h < S; h += 2 * kWarps) {                 // two rows per step = 2 * MT2 independent accumulators in flight
      const bool two = h + kWarps < S;
      const int hh[2] = {h, two ? h + kWarps : h};
      uint32_t b[2][C::KT2][2];
      load_b<C::KT2>(spb + (uint32_t)hh[0] * kPitch, 16u, lane, b[0]);
      load_b<C::KT2>(spb + (uint32_t)hh[1] * kPitch, 16u, lane, b[1]);
      float acc[2][C::MT2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[r][mt][i] = 0.f;
#pragma unroll
      for (int kt = 0; kt < C::KT2; ++kt)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int mt = 0; mt < C::MT2; ++mt) mma16816(acc[r][mt], a2[mt][kt], b[r][kt][0], b[r][kt][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        __half* yp = y.p + n * y.sn + hh[r] * y.sh + ch0 + 2 * t;
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) {
          const int w0 = mt * 16 + g, w1 = w0 + 8;
          float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
          if (has_add) {
            uint32_t u0 = 0u, u1 = 0u;
            if (w0 < S) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u0) : "r"(rb + (uint32_t)((hh[r] * S + w0) * 16 + t * 4)) : "memory");
            if (w1 < S) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u1) : "r"(rb + (uint32_t)((hh[r] * S + w1) * 16 + t * 4)) : "memory");
            r0 = __half22float2(*reinterpret_cast<const __half2*>(&u0));
            r1 = __half22float2(*reinterpret_cast<const __half2*>(&u1));
          }
          if (w0 < S) *reinterpret_cast<uint32_t*>(yp + w0 * y.sw) = pack_h2(acc[r][mt][0] + r0.x, acc[r][mt][1] + r0.y);
          if (w1 < S) *reinterpret_cast<uint32_t*>(yp + w1 * y.sw) = pack_h2(acc[r][mt][2] + r1.x, acc[r][mt][3] + r1.y);
        }
      }
    }
    __syncthreads();                           // every warp is done with the residual tile and with this spec buffer
    if (threadIdx.x == 0 && has_add && ti + (int)gridDim.x < tiles) issue_res(ti + gridDim.x);
  }
}

// ---- host: DFT matrices in A-fragment order ---------------------------------------------------------------------------
template <typename F>
static void fill_frags(uint4* dst, int mts, int kts, F f) {
  for (int mt = 0; mt < mts; ++mt)
    for (int kt = 0; kt < kts; ++kt)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        uint32_t r[4];
        for (int i = 0; i < 4; ++i) {       // a0,a1: row g | a2,a3: row g+8 | a4,a5: row g, col +8 | a6,a7: row g+8, col +8
          const int row = mt * 16 + g + 8 * (i & 1), col = kt * 16 + 2 * t + 8 * (i >> 1);
          const __half_raw lo = static_cast<__half_raw>(__float2half_rn((float)f(row, col))), hi = static_cast<__half_raw>(__float2half_rn((float)f(row, col + 1)));
          r[i] = (uint32_t)lo.x | ((uint32_t)hi.x << 16);
        }
        dst[(mt * kts + kt) * 32 + lane] = make_uint4(r[0], r[1], r[2], r[3]);
      }
}

template <int S>
static void build_tables(uint4* tab) {
  using C = Cfg<S>;
  const double s = 1.0 / sqrt((double)S), w0 = 2.0 * 3.14159265358979323846 / S;
  auto cs = [&](int a, int b) { return cos(w0 * ((a * b) % S)) * s; };
  auto sn = [&](int a, int b) { return sin(w0 * ((a * b) % S)) * s; };
  // pass 1: row = m-tile of 8 k's, rows 0-7 re, 8-15 im; col = w
  fill_frags(tab + C::O_F1 * 32, C::MT1, C::KT, [&](int row, int w) {
    const int k = (row / 16) * 8 + (row % 8), ri = (row % 16) / 8;
    if (k >= C::K1 || w >= S) return 0.0;
    return ri ? -sn(k, w) : cs(k, w);
  });
  auto in = [&](int a, int b) { return a < S && b < S; };
  fill_frags(tab + C::O_GR * 32, C::MT2, C::KT, [&](int kh, int h) { return in(kh, h) ? cs(kh, h) : 0.0; });
  fill_frags(tab + C::O_GI * 32, C::MT2, C::KT, [&](int kh, int h) { return in(kh, h) ? -sn(kh, h) : 0.0; });
  fill_frags(tab + C::O_WR * 32, C::MT2, C::KT, [&](int h, int kh) { return in(h, kh) ? cs(h, kh) : 0.0; });
  fill_frags(tab + C::O_WI * 32, C::MT2, C::KT, [&](int h, int kh) { return in(h, kh) ? sn(h, kh) : 0.0; });
  // c2r: col = (k, ri); Hermitian weights 1 (k = 0, S/2) / 2; sin(0) = sin(pi w) = 0 drops the imaginary parts of DC / Nyquist
  fill_frags(tab + C::O_A2 * 32, C::MT2, C::KT2, [&](int w, int col) {
    const int k = col / 2, ri = col % 2;
    if (w >= S || k >= C::K1) return 0.0;
    const double ck = (k == 0 || k == S / 2) ? 1.0 : 2.0;
    return ri ? -ck * sn(k, w) : ck * cs(k, w);
  });
}

template <int S>
static int grid_for(int tiles, int dev) {
  const int n_sm = sm_count(dev);
  if (n_sm <= 0) return -1;
  const int cap = n_sm * blocks_per_sm<S>();
  return tiles < cap ? tiles : cap;
}
template <int S>
static int launch_r(const s2v_view* x, const s2v_view* sp, cudaStream_t st) {
  static DeviceOnce attr;
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  constexpr int smem = Cfg<S>::TILE + (Cfg<S>::N_F1 + 2 * Cfg<S>::N_G) * 512;
  if (attr.needed(dev)) {
    S2V_CUDA_TRY(cudaFuncSetAttribute(rfft2_mma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark(dev);
  }
  const int cblocks = x->c / 8, tiles = cblocks * x->n, grid = grid_for<S>(tiles, dev);
  if (grid <= 0) return S2V_ECUDA;
  S2V_CUDA_TRY(launch_pdl(rfft2_mma_kernel<S>, dim3(grid), kThreads, (size_t)smem, st, mk(x), mk(sp), cblocks, tiles));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
template <int S>
static int launch_r_tma(const s2v_view* x, const s2v_view* sp, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return S2V_EUNSUPPORTED;
  static DeviceOnce attr;
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  constexpr int smem = x_bytes<S>() + Cfg<S>::TILE + 16 + 128;
  if (attr.needed(dev)) {
    S2V_CUDA_TRY(cudaFuncSetAttribute(rfft2_mma_tma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark(dev);
  }
  CUtensorMap tmx;
  cuuint64_t gdim[4] = {(cuuint64_t)x->c, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
  cuuint64_t gstr[3] = {(cuuint64_t)x->sw * 2, (cuuint64_t)x->sh * 2, (cuuint64_t)x->sn * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)S, (cuuint32_t)S, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x->ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return launch_r<S>(x, sp, st);            // a view the tensor map cannot describe: the cp.async kernel takes any view_ok view
  const int n_sm = sm_count(dev);
  if (n_sm <= 0) return S2V_ECUDA;
  const int cblocks = x->c / 8, tiles = cblocks * x->n, cap = blocks_per_sm_tma<S>() * n_sm, grid = tiles < cap ? tiles : cap;
  typedef void (*KernelFn)(const CUtensorMap, View, int, int);
  const KernelFn kfn = rfft2_mma_tma_kernel<S>;
  S2V_CUDA_TRY(launch_pdl(kfn, dim3(grid), kThreads, (size_t)smem, st, tmx, mk(sp), cblocks, tiles));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

template <int S>
static int launch_i(const s2v_view* sp, const s2v_view* add, const s2v_view* y, cudaStream_t st) {
  static DeviceOnce attr;
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  constexpr int smem = Cfg<S>::TILE + (2 * Cfg<S>::N_G + Cfg<S>::N_A2) * 512;
  if (attr.needed(dev)) {
    S2V_CUDA_TRY(cudaFuncSetAttribute(irfft2_mma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark(dev);
  }
  const int cblocks = y->c / 8, tiles = cblocks * y->n, grid = grid_for<S>(tiles, dev);
  if (grid <= 0) return S2V_ECUDA;
  S2V_CUDA_TRY(launch_pdl(irfft2_mma_kernel<S>, dim3(grid), kThreads, (size_t)smem, st, mk(sp), mk(add), mk(y), cblocks, tiles));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

static bool make_map4(EncodeTiledFn enc, CUtensorMap* tm, const s2v_view* v, int c_elems, int w_elems, int box_c, int box_w, int box_h) {
  cuuint64_t gdim[4] = {(cuuint64_t)c_elems, (cuuint64_t)w_elems, (cuuint64_t)v->h, (cuuint64_t)v->n};
  cuuint64_t gstr[3] = {(cuuint64_t)v->sw * 2, (cuuint64_t)v->sh * 2, (cuuint64_t)v->sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, v->ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <int S>
static int launch_i_tma(const s2v_view* sp, const s2v_view* add, const s2v_view* y, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return S2V_EUNSUPPORTED;
  static DeviceOnce attr;
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  constexpr int smem = r_bytes<S>() + sp_bytes<S>() + 16 + 128;
  if (attr.needed(dev)) {
    S2V_CUDA_TRY(cudaFuncSetAttribute(irfft2_mma_tma_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2V_CUDA_TRY(cudaFuncSetAttribute(irfft2_mma_tma_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark(dev);
  }
  static const int store_tma = [] { const char* e = getenv("S2V_FFT_TMA_STORE"); return e ? atoi(e) : 0; }();      // development knob: 0 = pipelined kernel (default), 1 / 2 = single-buffered with TMA / LSU store
  CUtensorMap tmsp, tmadd, tmy;
  if (!make_map4(enc, &tmsp, sp, sp->c, sp->w, 16, Cfg<S>::K1, S) || !make_map4(enc, &tmy, y, y->c, y->w, 8, S, S) ||
      !make_map4(enc, &tmadd, add ? add : y, y->c, y->w, 8, S, S))
    return launch_i<S>(sp, add, y, st);       // a view the tensor maps cannot describe: the cp.async kernel takes any view_ok view
  const int n_sm = sm_count(dev);
  if (n_sm <= 0) return S2V_ECUDA;
  const int cblocks = y->c / 8, tiles = cblocks * y->n, cap = blocks_per_sm_tma<S>() * n_sm, grid = tiles < cap ? tiles : cap;
  if (store_tma == 0) {                       // pipelined form: two spec buffers, LSU stores
    constexpr int smem2 = r_bytes<S>() + 2 * sp_bytes<S>() + 32 + 128;
    static DeviceOnce attr2;
    if (attr2.needed(dev)) {
      S2V_CUDA_TRY(cudaFuncSetAttribute(irfft2_mma_tma2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr2.mark(dev);
    }
    typedef void (*KernelFn2)(const CUtensorMap, const CUtensorMap, View, int, int, int);
    const KernelFn2 kfn2 = irfft2_mma_tma2_kernel<S>;
    S2V_CUDA_TRY(launch_pdl(kfn2, dim3(grid), kThreads, (size_t)smem2, st, tmsp, tmadd, mk(y), add ? 1 : 0, cblocks, tiles));
    S2V_CHECK_LAUNCH();
    return S2V_OK;
  }
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, View, int, int, int);
  const KernelFn kfn = store_tma == 1 ? irfft2_mma_tma_kernel<S, true> : irfft2_mma_tma_kernel<S, false>;
  S2V_CUDA_TRY(launch_pdl(kfn, dim3(grid), kThreads, (size_t)smem, st, tmsp, tmadd, tmy, mk(y), add ? 1 : 0, cblocks, tiles));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

}  // namespace fftmma

// ---- entry points used by fft2d.cu ------------------------------------------------------------------------------------
int fft_mma_init() {
  static uint4 host[fftmma::kFragsTotal * 32];     // rebuilt per call (per device): same values every time
  fftmma::build_tables<48>(host);
  fftmma::build_tables<24>(host);
  fftmma::build_tables<12>(host);
  if (cudaMemcpyToSymbol(fftmma::g_frag, host, sizeof(host)) != cudaSuccess) return S2V_ECUDA;
  return S2V_OK;
}
int rfft2_mma(const s2v_view* x, const s2v_view* sp, cudaStream_t st) {
  static const int tma = [] { const char* e = getenv("S2V_FFT_TMA"); return e ? atoi(e) : 1; }();       // development knob: 0 = cp.async tile loads
  switch (x->h) {
    case 48: return tma ? fftmma::launch_r_tma<48>(x, sp, st) : fftmma::launch_r<48>(x, sp, st);
    case 24: return tma ? fftmma::launch_r_tma<24>(x, sp, st) : fftmma::launch_r<24>(x, sp, st);
    case 12: return tma ? fftmma::launch_r_tma<12>(x, sp, st) : fftmma::launch_r<12>(x, sp, st);
  }
  return S2V_EINVAL;
}
int irfft2_mma(const s2v_view* sp, const s2v_view* add, const s2v_view* y, cudaStream_t st) {
  static const int tma = [] { const char* e = getenv("S2V_FFT_TMA"); return e ? atoi(e) : 1; }();       // development knob: 0 = cp.async tile loads
  switch (y->h) {
    case 48: return tma ? fftmma::launch_i_tma<48>(sp, add, y, st) : fftmma::launch_i<48>(sp, add, y, st);
    case 24: return tma ? fftmma::launch_i_tma<24>(sp, add, y, st) : fftmma::launch_i<24>(sp, add, y, st);
    case 12: return tma ? fftmma::launch_i_tma<12>(sp, add, y, st) : fftmma::launch_i<12>(sp, add, y, st);
  }
  return S2V_EINVAL;
}

}  // namespace s2v
