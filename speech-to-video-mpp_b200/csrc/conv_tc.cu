// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d (3x3 / 1x1, stride 1) and nn.Linear on
// the hot path (models/base_blocks.py:99,116,433; models/ffc.py:139-153,197-204;
// models/transformer.py:45-48,65-71).
//
//   D[M = pixels, N = Cout] = sum over taps (ky,kx) and 64-channel chunks of
//                             A_tap[M, 64] * W_tap[64, N]          (fp16 x fp16 -> fp32)
//
// * The M tile is a BOX of 128 output pixels (box_n x box_h x box_w) of the channels-last
//   activation.  For tap (ky,kx) the A operand is the same box shifted by the tap offset, so
//   one tiled 4-D TMA load {64 ch, box_w, box_h, box_n} lands it in shared memory already in
//   the K-major SWIZZLE_128B layout tcgen05.mma wants (row = pixel, 128 B = 64 fp16 channels).
//   TMA out-of-bounds zero fill IS the zero padding; channel counts that are not a multiple
//   of 64 are zero-filled the same way.  Reflect padding is pre-materialised by the producer
//   (s2v_affine_act reflect1) and arrives here as pad = 0 on a padded view.
// * Weights are packed [Cout][tap][Cin64] (K-major) and fetched by a 2-D TMA box {64, BN}.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2..5 = epilogue (tcgen05.ld -> scale/bias/residual/activation -> fp16 stores).
//   A multi-stage smem ring is handed over with mbarriers (TMA complete_tx -> MMA,
//   tcgen05.commit -> producer); the accumulator (128 lanes x BN fp32 columns) lives in TMEM.
#include <cuda.h>
#include <cstdio>

#include "common.cuh"

namespace s2v {

constexpr int kTileM = 128, kChunkK = 64, kUmmaK = 16;
constexpr int kABytes = kTileM * kChunkK * 2;        // 16 KB
constexpr int kThreads = 320;                        // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kMaxASlots = 8;                        // A-patch slots in flight (small 1x1 patches need many to cover the load latency)
constexpr unsigned kSpinLimit = 1u << 26;            // bounded waits: trap instead of hanging the GPU

#ifdef S2V_EPI_PROF
// development build only (tools/build_variant.py): per-phase clock sums of ONE epilogue thread (CTA 0, group 0, thread 0)
__device__ unsigned long long g_epi_prof[16];
#define EPI_T(i) if (prof_on) tprof[i] = clock64()
#else
#define EPI_T(i)
#endif

struct TcParams {
  int N, OH, OW;
  int box_w, box_h, box_n;
  int tiles_w, tiles_h;
  int kh, kw, pad_h, pad_w, dil_h, dil_w, str_h, str_w;
  int cin_chunks, cout, bn, stages, tmem_cols, ring_bytes;
  int m_tiles, total_tiles, tmem_buf_cols, stage_out_bytes, use_tma_store, pass_cols, stage_bufs, n_tiles_n, epi_groups;
  int tab;                                                // per-group scale / bias table length (bn rounded up to 32)
  int nf_chunk, bn_narrow;                                // segment-0 chunks >= nf_chunk run as N = bn_narrow MMAs (0 = off)
  int m_pairs, total_pair_tiles;                          // CTA-pair mode: the pair (leader, peer) owns M tiles (2*mp, 2*mp+1)
  int ki0, k2w, pad2_h, pad2_w, cin2_chunks, ki_total;   // second K segment (x2)
  // halo mode: one A patch (box + (k-1) halo) per 64-channel chunk serves every tap; B tiles stream per tap
  int halo, taps0, taps2, pw0, prows0, pw2, prows2, a_slots, a_slot_bytes, b_resident, pf_dist;
  int a_slots2, a_slot2_bytes;                            // second A ring (segment-2 patches) behind the first; 0 = one shared ring
  float* stats;                                           // fused per-(image, tile, channel) sum / sumsq
  int st_c_off, st_c_total, st_chunk_off, st_chunks_total, st_groups, st_gmax;
  View y, r1, r2;
  const float* scale;
  const float* bias;
  int act;
  float ap;
  int out_mode;
  float* yf;
  int exp;                                                // development experiment bits (S2V_EXP): 1 = epilogue skips its work, 2 = no MMAs issued, 4 = no stats walk
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a patch a few tiles ahead: the A-patch slots are few (2 when the weights are resident next to a 128-column
// staging tile), so a load's latency is only partly hidden by the previous chunk's MMAs; pulling the NEXT tiles' patches into
// L2 while the slots are busy turns the slot loads into L2 hits
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(8 rows * 128 B = 1024 B)>>4 <<32 | version 1 <<46 | SW128 (2) <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes = 1024) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// same instruction with the 64-bit descriptors assembled from (lo, hi) words inside the asm block: the issue loop
// only does 32-bit adds on the address field (the single issuing thread is latency-bound, every instruction counts)
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }

// ---- CTA-pair (cta_group::2) variants: one MMA instruction spans two SMs (M = 256), each CTA stages its own 128
// A rows and HALF of the B tile; TMA completions of both CTAs land on the leader's mbarrier; commits are multicast.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_lh_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {     // arrives on `bar` (same offset) in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// one elected lane of a fully converged warp (the whole MMA warp runs the loop so that descriptors stay warp-uniform)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int ACT>
__device__ __forceinline__ float act_t(float v, float ap) {
  if (ACT == S2V_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == S2V_ACT_LRELU) return v > 0.f ? v : v * ap;
  if (ACT == S2V_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  if (ACT == S2V_ACT_TANH) return tanhf(v);
  if (ACT == S2V_ACT_GELU) {
    const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    return 0.5f * v * (1.f + tanhf(u));
  }
  return v;
}

// Issue of every tap (kh x KW) x 4 K-steps of ONE 64-channel chunk against smem-resident weights, by the single elected
// lane.  One patch row (KW taps = 4*KW MMAs) is straight-line code whose descriptors are two base registers plus
// compile-time offsets - a couple of uniform adds between two tcgen05.mma (the generic loop spends ~50 instructions per
// 4 MMAs, which bounds every N <= 128 layer); the rows are a rolled loop so that the body stays a few hundred bytes: the
// MMA warp shares its scheduler's instruction cache with two epilogue warps (ncu: stall_no_inst was its top stall with
// the fully unrolled 3x3 / 7x7 bodies).
template <int KW, bool C2>
__device__ __forceinline__ void issue_chunk_resident(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t row_units,
                                                     uint32_t b_lo, uint32_t b_units, uint32_t b_hi, uint32_t idesc, uint32_t accum,
                                                     int kh) {
#pragma unroll 1
  for (int ky = 0; ky < kh; ++ky, a_lo += row_units) {
#pragma unroll
    for (int kx = 0; kx < KW; ++kx, b_lo += b_units) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (C2) umma_f16_lh_2sm(d_tmem, a_lo + (uint32_t)(kx * 8 + 2 * k), a_hi, b_lo + 2u * k, b_hi, idesc, accum);
        else umma_f16_lh(d_tmem, a_lo + (uint32_t)(kx * 8 + 2 * k), a_hi, b_lo + 2u * k, b_hi, idesc, accum);
        accum = 1u;
      }
    }
  }
}
// row widths with a specialised instantiation (0 = generic loop)
__device__ __forceinline__ int issue_shape(int kh, int kw) { return (kw == 1 || kw == 2 || kw == 3 || kw == 7) ? kw : 0; }

struct Smem {            // offsets (shared-space addresses) of the carved regions
  uint32_t ring, stage_out, full0, empty0, tfull0, tempty0, tptr, afull0, aempty0, ball, bring;
  float* s_scale;
  float* s_bias;
  uint8_t* s_valid;
  __half* stage;
};

// work item -> (N tile, box origin).  single-CTA mode: item = tile index; pair mode: item = pair-tile index, the CTA of
// rank `crank` takes M tile 2*mp + crank (an odd tail gives the peer a tile fully outside the tensor: TMA zero-fills its
// loads and clips its stores).
template <bool C2>
__device__ __forceinline__ void tile_coords(const TcParams& p, int tile, int crank, int& ntile, int& n0, int& y0, int& x0, int& tile_sp) {
  int mt;
  if (C2) {
    ntile = tile / p.m_pairs;
    mt = 2 * (tile - ntile * p.m_pairs) + crank;
  } else {
    ntile = tile / p.m_tiles;
    mt = tile - ntile * p.m_tiles;
  }
  const int per_img = p.tiles_w * p.tiles_h;
  const int tn = mt / per_img;
  tile_sp = mt - tn * per_img;
  const int th = tile_sp / p.tiles_w, tw = tile_sp - th * p.tiles_w;
  x0 = tw * p.box_w; y0 = th * p.box_h; n0 = tn * p.box_n;
}

// NC (32 or 16) accumulator columns of this thread's row: TMEM -> scale / bias -> activation -> [LayerNorm2d totals] -> fp16 ->
// NC / 8 128-bit stores into the swizzled staging tile.  Straight-line code on NC independent values (the per-launch conditions
// are warp-uniform branches around whole blocks): the generic column loop of epilogue_loop is a chain of small basic blocks and
// ran at ~600 clocks per 32 columns with two epilogue warps per scheduler, this one at ~250.
template <int ACT, int NC>
__device__ __forceinline__ void epi_cols(uint32_t taddr, const float* s_scale, const float* s_bias, bool affine, bool st_scalar, float ap,
                                         float& ln_s, float& ln_q, __half* row_base, int col, int r7) {
  float v[NC];
  if (NC == 32) tmem_ld32_nowait(taddr, v);
  else tmem_ld16_nowait(taddr, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (affine) {
    const float4* s4 = reinterpret_cast<const float4*>(s_scale);
    const float4* b4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
    for (int g = 0; g < NC / 4; ++g) {
      const float4 sc = s4[g], bi = b4[g];
      v[4 * g] = fmaf(v[4 * g], sc.x, bi.x); v[4 * g + 1] = fmaf(v[4 * g + 1], sc.y, bi.y);
      v[4 * g + 2] = fmaf(v[4 * g + 2], sc.z, bi.z); v[4 * g + 3] = fmaf(v[4 * g + 3], sc.w, bi.w);
    }
  }
  if (ACT != S2V_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = act_t<ACT>(v[i], ap);
  }
  if (st_scalar) {                                // LayerNorm2d totals from the fp32 values (pairwise trees)
    float ts[NC / 4], tq[NC / 4];
#pragma unroll
    for (int g = 0; g < NC / 4; ++g) {
      ts[g] = (v[4 * g] + v[4 * g + 1]) + (v[4 * g + 2] + v[4 * g + 3]);
      tq[g] = fmaf(v[4 * g + 1], v[4 * g + 1], v[4 * g] * v[4 * g]) + fmaf(v[4 * g + 3], v[4 * g + 3], v[4 * g + 2] * v[4 * g + 2]);
    }
#pragma unroll
    for (int w = NC / 8; w >= 1; w >>= 1)
#pragma unroll
      for (int g = 0; g < w; ++g) { ts[g] += ts[g + w]; tq[g] += tq[g + w]; }
    ln_s += ts[0];
    ln_q += tq[0];
  }
  __half* const panel = row_base + (size_t)(col >> 6) * (kTileM * 64);
  const int c0 = (col >> 3) & 7;                  // multiple of 2 (NC = 16) or 4 (NC = 32): chunks c0 .. c0 + NC / 8 - 1 never wrap
#pragma unroll
  for (int g = 0; g < NC / 8; ++g) st_h8(panel + (((c0 + g) ^ r7) << 3), f_to_h8(v + 8 * g));
}

// Epilogue warps (4): for every tile of this CTA, TMEM -> registers -> scale/bias/activation ->
// (fp16 tile staged in smem -> optional column statistics -> coalesced 16-byte stores [+ residual]) or direct
// stores (fp32 NCHW heads, pre-activation residual).  Runs concurrently with the producer / MMA warps
// working on the NEXT tile (the accumulator is double-buffered in TMEM).
template <int ACT, bool C2>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, const Smem& sm, const CUtensorMap* tmY, uint32_t tmem, int warp,
                                              int lane, int crank) {
  const int w_first = C2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_step = C2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_total = C2 ? p.total_pair_tiles : p.total_tiles;
  const uint32_t tempty_remote0 = C2 ? mapa_rank0(sm.tempty0) : 0u;
  const int q = warp & 3;                          // TMEM lane quarter this warp may access
  // two epilogue groups of 4 warps: group g drains accumulator buffer g, i.e. the tiles j = g, g+2, ... of this CTA,
  // with its own staging buffer, tables and named barrier - two tiles' epilogues are in flight at once
  const int grp = (warp - 2) >> 2;
  if (grp >= p.epi_groups) return;
  const int et = threadIdx.x - 64 - 128 * grp;     // 0..127 within the group
  const uint32_t bar_id = 1u + (uint32_t)grp;
  float* const s_scale = sm.s_scale + 2 * p.tab * grp;
  float* const s_bias = s_scale + p.tab;
  uint8_t* const s_valid = sm.s_valid + 128 * grp;
  const int m = q * 32 + lane;                     // tile row owned in phase 1
  const int ww = m % p.box_w, hh = (m / p.box_w) % p.box_h, nn = m / (p.box_w * p.box_h);
  const bool direct = (p.out_mode == S2V_OUT_F32_NCHW) || (p.r1.p != nullptr);
  const bool affine = p.scale != nullptr || p.bias != nullptr;      // identity tables are skipped altogether
  const bool st_scalar = p.stats != nullptr && p.st_gmax == 0;      // LayerNorm2d totals instead of per-channel partials
  const int half_n = p.bn > p.pass_cols ? p.pass_cols : p.bn;      // columns staged per pass (1 or 2 panels of 64 channels)
  // staging layout = what a SWIZZLE_128B TMA store expects: panels of 64 channels, [128 rows][128 B] each,
  // 16-byte chunk j of row r stored at chunk position j ^ (r & 7)  (also makes the smem stores conflict-free)
  const bool tma_store = !direct && (p.r2.p == nullptr) && p.use_tma_store;
  uint32_t sb = p.epi_groups == 2 ? (uint32_t)grp : 0u;   // staging buffer: one per group, or alternating per pass
  auto stage_ptr = [&](int row, int col) -> __half* {      // col multiple of 8 within the pass
    const int panel = col >> 6, chunk = (col >> 3) & 7;
    return sm.stage + (size_t)sb * (p.stage_out_bytes >> 2) + (size_t)panel * (kTileM * 64) + (size_t)row * 64 + ((chunk ^ (row & 7)) << 3);
  };
  auto load_tables = [&](int ntile) {
    if (!affine) return;                             // identity tables are not even allocated (p.tab = 0)
    for (int i = et; i < p.bn; i += 128) {
      const int c = ntile * p.bn + i;
      s_scale[i] = (p.scale && c < p.cout) ? p.scale[c] : 1.f;
      s_bias[i] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    }
  };
  if (p.n_tiles_n == 1) { load_tables(0); asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); }
#ifdef S2V_EPI_PROF
  const bool prof_on = blockIdx.x == 0 && grp == 0 && et == 0;
  long long tprof[8];
#endif
  const int lg_w = __ffs(p.box_w) - 1, lg_wh = lg_w + __ffs(p.box_h) - 1;    // box_w * box_h * box_n = 128: powers of two
  const int jstep = p.epi_groups;
  int j = grp;
  for (int tile = w_first + grp * w_step; tile < w_total; tile += jstep * w_step, j += jstep) {
    int ntile, n0, y0, x0, tile_sp;
    tile_coords<C2>(p, tile, crank, ntile, n0, y0, x0, tile_sp);
    const int n = n0 + nn, oy = y0 + hh, ox = x0 + ww;
    const bool valid = (n < p.N) && (oy < p.OH) && (ox < p.OW);
    const uint32_t buf = (uint32_t)j & 1u;
    if (p.n_tiles_n > 1) {                          // per-channel tables change with the N tile
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      load_tables(ntile);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
    EPI_T(0);
    mbar_wait(sm.tfull0 + 8u * buf, ((uint32_t)j >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    EPI_T(1);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.tmem_buf_cols;
    float ln_s = 0.f, ln_q = 0.f;
    if (p.exp & 1) {                                // experiment: drain nothing, just hand the accumulator back
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (C2) mbar_arrive_cluster(tempty_remote0 + 8u * buf);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm.tempty0 + 8u * buf) : "memory");
      }
      continue;
    }
    for (int pass0 = 0; pass0 < p.bn; pass0 += half_n) {
      const int pass_n = min(half_n, p.bn - pass0);
      if (!direct) {
        // the staging buffer of this pass is free once the bulk store issued stage_bufs passes ago has read it
        if (tma_store && et == 0) {
          if (p.stage_bufs == 2 && p.epi_groups == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        if (p.stats && pass0 == 0) s_valid[m] = valid ? 1 : 0;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      EPI_T(2);
      // Fast path (fp16 staged output, whole 32-column groups, every channel valid): one 32-column TMEM load per iteration
      // and straight-line code with the per-launch conditions hoisted, so that the 32 values are independent instruction
      // streams - the generic loop below is a chain of small basic blocks (latency-bound with two epilogue warps per scheduler:
      // measured ~600 clocks per 32 columns)
      const bool fast_cols = !direct && (pass_n & 15) == 0 && ntile * p.bn + pass0 + pass_n <= p.cout && !(p.exp & 8);
      if (fast_cols) {
        __half* const row_base = sm.stage + (size_t)sb * (p.stage_out_bytes >> 2) + (size_t)m * 64;
        const int r7 = m & 7;
#pragma unroll 1
        for (int cb = 0; cb + 32 <= pass_n; cb += 32)
          epi_cols<ACT, 32>(trow + (uint32_t)(pass0 + cb), s_scale + pass0 + cb, s_bias + pass0 + cb, affine, st_scalar, p.ap, ln_s, ln_q,
                            row_base, cb, r7);
        if (pass_n & 16)                            // 48 / 80 / 112-column passes: one 16-column tail group
          epi_cols<ACT, 16>(trow + (uint32_t)(pass0 + (pass_n & ~31)), s_scale + pass0 + (pass_n & ~31), s_bias + pass0 + (pass_n & ~31), affine,
                            st_scalar, p.ap, ln_s, ln_q, row_base, pass_n & ~31, r7);
      } else
      for (int cb = 0; cb < pass_n; cb += 32) {
        float v[32];
        tmem_ld16_nowait(trow + (uint32_t)(pass0 + cb), v);
        if (cb + 16 < pass_n) tmem_ld16_nowait(trow + (uint32_t)(pass0 + cb + 16), v + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cl = cb + 8 * g;                // column within the pass
          if (cl >= pass_n) break;
          const int ct = pass0 + cl;                // column within the N tile
          const int c = ntile * p.bn + ct;
          float* o = v + 8 * g;
          if (affine) {
            const float4 s0 = *reinterpret_cast<const float4*>(s_scale + ct), s1 = *reinterpret_cast<const float4*>(s_scale + ct + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + ct), b1 = *reinterpret_cast<const float4*>(s_bias + ct + 4);
            o[0] = fmaf(o[0], s0.x, b0.x); o[1] = fmaf(o[1], s0.y, b0.y); o[2] = fmaf(o[2], s0.z, b0.z); o[3] = fmaf(o[3], s0.w, b0.w);
            o[4] = fmaf(o[4], s1.x, b1.x); o[5] = fmaf(o[5], s1.y, b1.y); o[6] = fmaf(o[6], s1.z, b1.z); o[7] = fmaf(o[7], s1.w, b1.w);
          }
          if (!direct) {
            if (ACT != S2V_ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = act_t<ACT>(o[i], p.ap);
            }
            if (st_scalar && c < p.cout) {          // LayerNorm2d totals: this thread's row, all channels, straight from fp32
              // (pairwise tree: a serial chain of 16 dependent adds per 8 columns stalls the two epilogue warps of a scheduler)
              const float s0 = o[0] + o[1], s1 = o[2] + o[3], s2 = o[4] + o[5], s3 = o[6] + o[7];
              const float q0 = fmaf(o[1], o[1], o[0] * o[0]), q1 = fmaf(o[3], o[3], o[2] * o[2]);
              const float q2 = fmaf(o[5], o[5], o[4] * o[4]), q3 = fmaf(o[7], o[7], o[6] * o[6]);
              ln_s += (s0 + s1) + (s2 + s3);
              ln_q += (q0 + q1) + (q2 + q3);
            }
            st_h8(stage_ptr(m, cl), f_to_h8(o));
            continue;
          }
          if (!valid || c >= p.cout) continue;
          if (p.r1.p) {
            float f[8];
            h8_to_f(ld_h8(p.r1.p + n * p.r1.sn + oy * p.r1.sh + ox * p.r1.sw + c), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += f[i];
          }
          if (ACT != S2V_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = act_t<ACT>(o[i], p.ap);
          }
          if (p.r2.p) {
            float f[8];
            h8_to_f(ld_h8(p.r2.p + n * p.r2.sn + oy * p.r2.sh + ox * p.r2.sw + c), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += f[i];
          }
          if (p.out_mode == S2V_OUT_F32_NCHW) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c + i < p.cout) p.yf[(((size_t)n * p.cout + c + i) * p.OH + oy) * p.OW + ox] = o[i];
          } else {
            st_h8(p.y.p + n * p.y.sn + oy * p.y.sh + ox * p.y.sw + c, f_to_h8(o));
          }
        }
      }
      EPI_T(3);
      const bool last_pass = pass0 + half_n >= p.bn;
      if (last_pass) {
        // all TMEM reads of this accumulator buffer are complete: hand it back to the MMA warp
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (C2) mbar_arrive_cluster(tempty_remote0 + 8u * buf);       // the leader's MMA thread owns both accumulators' schedule
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm.tempty0 + 8u * buf) : "memory");
        }
      }
      if (direct) continue;
      if (tma_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes -> visible to the TMA engine
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      EPI_T(4);
      if (tma_store && et == 0) {
        // one bulk tensor store per 64-channel panel; TMA clips rows/channels outside the output view
        for (int pc = 0; pc < pass_n; pc += 64) {
          const uint32_t src = sm.stage_out + sb * (uint32_t)(p.stage_out_bytes >> 1) + (uint32_t)(pc >> 6) * (kTileM * 128);
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(tmY), "r"(src), "r"(ntile * p.bn + pass0 + pc), "r"(x0), "r"(y0), "r"(n0)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#ifdef S2V_EPI_PROF
        if ((p.exp & 16) && prof_on) {              // how long until the bulk store has READ the staging buffer?
          const long long w0 = clock64();
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          atomicAdd(&g_epi_prof[13], (unsigned long long)(clock64() - w0));
          atomicAdd(&g_epi_prof[14], 1ull);
        }
#endif
      }
      EPI_T(5);
      if (st_scalar && last_pass) {
        // LayerNorm2d consumer: only the totals over (C, H, W) are needed.  Every thread summed its own output row over
        // all channels in registers (fp32, before the fp16 rounding); a fixed xor-butterfly over the lanes that share an
        // image (32, or rows-per-image when a box holds several images) leaves ONE partial per (image, tile, row group of
        // 32): stats[n][chunk][(m % rows_per_img) / 32][2].  No shared-memory pass, no per-channel work.
        const int rows_per_img = p.box_w * p.box_h;
        const int wdt = rows_per_img < 32 ? rows_per_img : 32;
        float ss = valid ? ln_s : 0.f, qq = valid ? ln_q : 0.f;
        for (int o = wdt >> 1; o > 0; o >>= 1) {
          ss += __shfl_xor_sync(0xffffffffu, ss, o);
          qq += __shfl_xor_sync(0xffffffffu, qq, o);
        }
        if ((lane & (wdt - 1)) == 0 && n < p.N) {
          const int slot = (m % rows_per_img) >> 5;
          float2* o2 = reinterpret_cast<float2*>(p.stats) + ((size_t)n * p.st_chunks_total + p.st_chunk_off + tile_sp) * 4 + slot;
          *o2 = make_float2(ss, qq);
        }
      }
      if (p.stats && !st_scalar && !(p.exp & 4)) {
        // Column sums (sum, sum of squares) of the staged fp16 tile, ONE partial per (image, spatial tile, channel):
        // the G = 128 / (pass_n / 8) consecutive lanes that share a 16-byte chunk (8 channels) each walk rows g, g+G, ...
        // of one image of the box with conflict-free 128-bit smem loads, then a fixed butterfly over those lanes
        // (deterministic, independent of the batch) leaves the total in lane g == 0.
        const int rows_per_img = p.box_w * p.box_h;
        const int cks = pass_n >> 3;                // 4, 8 or 16 (validated on the host)
        const int G = 128 / cks;
        // multi-image boxes (12x12 level: 8 images x 16 rows): the G lanes are spread over min(G, box_n) images at a time,
        // L = G / that lanes per image - with box_n >= G every lane owns whole images and no shuffle is needed
        const int ipg = G < p.box_n ? G : p.box_n;  // images in flight per chunk group (box_n and G are powers of two)
        const int L = G / ipg;
        // L > 1: the lanes sharing a chunk are consecutive (butterfly inside the warp; their rows differ, so the swizzle
        // spreads them over the banks).  L == 1: chunk-fastest mapping - a quarter warp reads one row's 8 chunks
        // (lanes that own different images would otherwise hit the same banks).
        const int ck = L > 1 ? et / G : et % cks, g = L > 1 ? et - ck * G : et / cks;
        const int c = ntile * p.bn + pass0 + ck * 8;
        const int chunk = p.st_chunk_off + tile_sp;
        const bool full_tile = (n0 + p.box_n <= p.N) && (y0 + p.box_h <= p.OH) && (x0 + p.box_w <= p.OW);
        const int sub = g / L, r0 = g - sub * L;
        for (int im = sub; im < p.box_n; im += ipg) {           // box_n % ipg == 0: the trip count is warp-uniform
          const bool img_ok = n0 + im < p.N;
          float sa[8], qa[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sa[i] = qa[i] = 0.f;
          const uint8_t* vp = s_valid + im * rows_per_img;
          if (full_tile) {                          // common case: no row mask, loads issued back to back
#pragma unroll 4
            for (int r = r0; r < rows_per_img; r += L) {
              float f[8];
              h8_to_f(ld_h8(stage_ptr(im * rows_per_img + r, ck * 8)), f);
#pragma unroll
              for (int i = 0; i < 8; ++i) { sa[i] += f[i]; qa[i] = fmaf(f[i], f[i], qa[i]); }
            }
          } else if (img_ok) {
#pragma unroll 2
            for (int r = r0; r < rows_per_img; r += L) {
              if (!vp[r]) continue;
              float f[8];
              h8_to_f(ld_h8(stage_ptr(im * rows_per_img + r, ck * 8)), f);
#pragma unroll
              for (int i = 0; i < 8; ++i) { sa[i] += f[i]; qa[i] = fmaf(f[i], f[i], qa[i]); }
            }
          }
          // Reduce-scatter over the L lanes of the group (recursive halving): at every stage a lane keeps one half of its
          // values and hands the other half to its xor partner, so the 16 values (sum, sumsq of 8 channels) cost
          // 8+4+2+1 = 15 shuffles instead of 16 per stage, and the result ends up spread over the lanes - lane r0 holds
          // the 16 / L consecutive values starting at r0 * 16 / L (one coalesced store per group).  Fixed order:
          // deterministic and independent of the batch.
          float val[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) { val[2 * i] = sa[i]; val[2 * i + 1] = qa[i]; }
#pragma unroll
          for (int st = 0; st < 5; ++st) {                // offsets L/2, L/4, ... 1; 16 >> st values held at stage st
            const int o = L >> (st + 1);
            if (o == 0) break;
            if (st < 4) {
              const bool up = (r0 & o) != 0;              // this lane keeps the upper half
#pragma unroll
              for (int i = 0; i < (8 >> st); ++i) {
                const float lo = val[i], hi = val[i + (8 >> st)];
                const float recv = __shfl_xor_sync(0xffffffffu, up ? lo : hi, o);
                val[i] = (up ? hi : lo) + recv;
              }
            } else {
              val[0] += __shfl_xor_sync(0xffffffffu, val[0], o);     // L = 32: the last stage has a single value left
            }
          }
          if (img_ok && c < p.cout) {
            float* o = p.stats + (((size_t)(n0 + im) * p.st_chunks_total + chunk) * p.st_c_total + p.st_c_off + c) * 2;
            if (L == 32) { if ((r0 & 1) == 0) o[r0 >> 1] = val[0]; }
            else if (L == 16) o[r0] = val[0];
            else if (L == 8) *reinterpret_cast<float2*>(o + 2 * r0) = make_float2(val[0], val[1]);
            else if (L == 4) *reinterpret_cast<float4*>(o + 4 * r0) = make_float4(val[0], val[1], val[2], val[3]);
            else {                                        // L = 2 (8 values per lane) or L = 1 (all 16)
              float4* o4 = reinterpret_cast<float4*>(o + (L == 2 ? 8 * r0 : 0));
              o4[0] = make_float4(val[0], val[1], val[2], val[3]);
              o4[1] = make_float4(val[4], val[5], val[6], val[7]);
              if (L == 1) {
                o4[2] = make_float4(val[8], val[9], val[10], val[11]);
                o4[3] = make_float4(val[12], val[13], val[14], val[15]);
              }
            }
          }
        }
      }
      const int cpr = pass_n >> 3;                  // 16-byte chunks per staged row
      const int cpr_sh = (cpr & (cpr - 1)) == 0 ? __ffs(cpr) - 1 : -1;
      const int total = tma_store ? 0 : kTileM * cpr;
      // 4 independent (row, 16 B chunk) items per iteration: the residual loads are issued together so
      // their L2 latency overlaps (y and res2 may be the same buffer, so the compiler cannot hoist them)
      for (int base = et; base < total; base += 4 * 128) {
        H8 hv[4], rv[4];
        __half* dst[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * 128;
          ok[u] = idx < total;
          // (the box dimensions are powers of two - their product is 128 - and so is cpr for every pass width but 48 / 24 / 40 ...:
          //  shifts and masks instead of five runtime divisions per item; this loop was ~5 000 clocks per pass)
          const int row = ok[u] ? (cpr_sh >= 0 ? idx >> cpr_sh : idx / cpr) : 0, ch = ok[u] ? idx - row * cpr : 0;
          const int c = ntile * p.bn + pass0 + ch * 8;
          const int rw = row & (p.box_w - 1), rh = (row >> lg_w) & (p.box_h - 1), rn = row >> lg_wh;
          const int pn = n0 + rn, py = y0 + rh, px = x0 + rw;
          ok[u] = ok[u] && pn < p.N && py < p.OH && px < p.OW && c < p.cout;
          dst[u] = p.y.p + pn * p.y.sn + py * p.y.sh + px * p.y.sw + c;
          if (ok[u]) {
            hv[u] = ld_h8(stage_ptr(row, ch * 8));
            if (p.r2.p) rv[u] = ld_h8(p.r2.p + pn * p.r2.sn + py * p.r2.sh + px * p.r2.sw + c);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          if (p.r2.p) {
            float a[8], f[8];
            h8_to_f(hv[u], a);
            h8_to_f(rv[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += f[i];
            hv[u] = f_to_h8(a);
          }
          st_h8(dst[u], hv[u]);
        }
      }
      if (p.stage_bufs == 2 && p.epi_groups == 1) sb ^= 1u;
#ifdef S2V_EPI_PROF
      if (prof_on) {                                 // single-pass layers only are meaningful
        tprof[6] = clock64();
        for (int i = 0; i < 6; ++i) atomicAdd(&g_epi_prof[i], (unsigned long long)(tprof[i + 1] - tprof[i]));
        atomicAdd(&g_epi_prof[8], 1ull);
      }
#endif
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); CTA b processes tiles b, b+grid, ...  The smem ring and its
// mbarrier phases run continuously across tiles; the fp32 accumulator is double-buffered in TMEM so the
// epilogue of tile j overlaps the TMA/MMA main loop of tile j+1.
template <bool C2, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmBn, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger_conv();     // the next kernel may start its prologue while this one runs (it waits for our completion itself)
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const int crank = C2 ? (int)cluster_ctarank() : 0;
  const bool leader = crank == 0;
  const uint32_t b_bytes = (uint32_t)(C2 ? (p.bn >> 1) : p.bn) * 128u;      // pair mode: each CTA stages half of the B tile
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int w_first = C2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_step = C2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_total = C2 ? p.total_pair_tiles : p.total_tiles;
  Smem sm;
  sm.ring = smem_base;
  sm.stage_out = smem_base + (uint32_t)p.ring_bytes;
  const uint32_t bar_base = sm.stage_out + (uint32_t)p.stage_out_bytes;
  sm.full0 = bar_base;
  sm.empty0 = bar_base + 8u * p.stages;
  sm.tfull0 = bar_base + 16u * p.stages;
  sm.tempty0 = sm.tfull0 + 16u;
  sm.afull0 = sm.tempty0 + 16u;             // up to kMaxASlots A-patch slots (halo mode)
  sm.aempty0 = sm.afull0 + 8u * kMaxASlots;
  sm.ball = sm.aempty0 + 8u * kMaxASlots;               // resident-weights barrier (halo mode)
  sm.tptr = sm.ball + 16u;
  sm.bring = sm.ring + (uint32_t)(p.a_slots * p.a_slot_bytes + p.a_slots2 * p.a_slot2_bytes);
  sm.s_scale = reinterpret_cast<float*>(smem_raw + (sm.tptr + 16u - raw_u32));
  sm.s_bias = sm.s_scale + p.tab;
  sm.s_valid = reinterpret_cast<uint8_t*>(sm.s_scale + 4 * p.tab);     // two groups x (scale[tab] | bias[tab])
  sm.stage = reinterpret_cast<__half*>(smem_raw + (sm.stage_out - raw_u32));
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (sm.tptr - raw_u32));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KI = p.ki_total;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (p.bn_narrow) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBn) : "memory");
    if (p.ki_total > p.ki0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    if (p.use_tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(sm.full0 + 8u * s, 1); mbar_init(sm.empty0 + 8u * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(sm.tfull0 + 8u * b, 1); mbar_init(sm.tempty0 + 8u * b, C2 ? 8 : 4); }
    for (int a = 0; a < kMaxASlots; ++a) { mbar_init(sm.afull0 + 8u * a, 1); mbar_init(sm.aempty0 + 8u * a, 1); }
    mbar_init(sm.ball, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (C2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm.tptr), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm.tptr), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (C2) cluster_sync_all(); else __syncthreads();     // pair mode: the peer's barriers must be initialised before remote arrivals
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tptr_gen, 0);
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel; nothing
  // below may touch activations before it has completed.  The producer thread waits later: weights are never written
  // by a kernel, so a resident weight matrix is fetched first.
  if (threadIdx.x != 0) pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====  (ring state is kept incrementally: no divisions in the single-thread loops)
      // pair mode: both CTAs load (own A rows, own half of B); every completion lands on the LEADER's barrier, which the
      // leader arms with the bytes of both CTAs
      uint32_t s = 0, ph = 0;                       // B (or A+B) ring slot / phase
      const uint32_t full_r = C2 ? mapa_rank0(sm.full0) : sm.full0;
      const uint32_t afull_r = C2 ? mapa_rank0(sm.afull0) : sm.afull0;
      const uint32_t ball_r = C2 ? mapa_rank0(sm.ball) : sm.ball;
      const uint32_t txm = C2 ? 2u : 1u;
      const int brow0 = C2 ? crank * (p.bn >> 1) : 0;
#define load_a(dst, tm, bar, c0, c1, c2, c3)                        \
  do {                                                               \
    if (C2) tma_load_4d_2sm(dst, tm, bar, c0, c1, c2, c3);           \
    else tma_load_4d(dst, tm, bar, c0, c1, c2, c3);                  \
  } while (0)
#define load_b(dst, bar, kcol, row)                                 \
  do {                                                               \
    if (C2) tma_load_2d_2sm(dst, &tmB, bar, kcol, row);              \
    else tma_load_2d(dst, &tmB, bar, kcol, row);                     \
  } while (0)
      if (p.halo) {
        if (p.b_resident) {                         // the whole weight matrix of this (single) N tile stays in smem
          if (p.bn_narrow) {
            // segment-0 chunks >= nf_chunk only feed the first bn_narrow output channels: their taps keep just those
            // weight rows (box of the narrow map), packed right behind the full-width taps
            const uint32_t nb_bytes = (uint32_t)(C2 ? (p.bn_narrow >> 1) : p.bn_narrow) * 128u;
            const int it_n0 = p.nf_chunk * p.taps0, it_n1 = p.ki0;            // narrow chunk-taps [it_n0, it_n1)
            const uint32_t n_narrow = (uint32_t)(it_n1 - it_n0);
            if (leader) mbar_expect_tx(sm.ball, txm * (((uint32_t)KI - n_narrow) * b_bytes + n_narrow * nb_bytes));
            const int brow0n = C2 ? crank * (p.bn_narrow >> 1) : 0;
            uint32_t dst = sm.bring;
            for (int it = 0; it < KI; ++it) {
              if (it >= it_n0 && it < it_n1) {
                if (C2) tma_load_2d_2sm(dst, &tmBn, ball_r, it * kChunkK, brow0n);
                else tma_load_2d(dst, &tmBn, ball_r, it * kChunkK, brow0n);
                dst += nb_bytes;
              } else {
                load_b(dst, ball_r, it * kChunkK, brow0);
                dst += b_bytes;
              }
            }
          } else {
            if (leader) mbar_expect_tx(sm.ball, txm * (uint32_t)KI * b_bytes);
            for (int it = 0; it < KI; ++it) load_b(sm.bring + (uint32_t)it * b_bytes, ball_r, it * kChunkK, brow0);
          }
        }
        pdl_wait();
        // A-patch rings: ring 0 (slots 0 .. a_slots-1) and, when a_slots2 > 0, a ring of its own for the segment-2 patches
        // (slots a_slots .. a_slots+a_slots2-1): with few slots a short segment-2 chunk otherwise leaves the next big patch
        // load uncovered (its slot frees only when the chunk before it retires)
        uint32_t a = 0, aph = 0, a2 = 0, aph2 = 0;
        const bool ring2 = p.a_slots2 > 0;
        const uint32_t ring1_base = sm.ring + (uint32_t)(p.a_slots * p.a_slot_bytes);
        const int chunks2 = p.ki_total > p.ki0 ? p.cin2_chunks : 0;
        for (int tile = w_first; tile < w_total; tile += w_step) {
          int ntile, n0, y0, x0, tile_sp;
          if (p.pf_dist) {                            // this CTA's tile pf_dist rounds ahead -> L2
            const int pt = tile + p.pf_dist * w_step;
            if (pt < w_total) {
              tile_coords<C2>(p, pt, crank, ntile, n0, y0, x0, tile_sp);
              for (int c = 0; c < p.cin_chunks; ++c) tma_prefetch_4d(&tmA, c * kChunkK, x0 - p.pad_w, y0 - p.pad_h, n0);
              for (int c = 0; c < chunks2; ++c) tma_prefetch_4d(&tmA2, c * kChunkK, x0 - p.pad2_w, y0 - p.pad2_h, n0);
            }
          }
          tile_coords<C2>(p, tile, crank, ntile, n0, y0, x0, tile_sp);
          int kcol = 0;
          for (int seg = 0; seg < 2; ++seg) {
            const int chunks = seg == 0 ? p.cin_chunks : chunks2;
            const int taps = seg == 0 ? p.taps0 : p.taps2;
            const uint32_t abytes = (uint32_t)(seg == 0 ? p.prows0 : p.prows2) * 128u;
            const CUtensorMap* tm = seg == 0 ? &tmA : &tmA2;
            const int ax = x0 - (seg == 0 ? p.pad_w : p.pad2_w), ay = y0 - (seg == 0 ? p.pad_h : p.pad2_h);
            const bool r1 = ring2 && seg == 1;
            for (int c = 0; c < chunks; ++c) {
              const uint32_t idx = r1 ? (uint32_t)p.a_slots + a2 : a;
              const uint32_t dst = r1 ? ring1_base + a2 * (uint32_t)p.a_slot2_bytes : sm.ring + a * (uint32_t)p.a_slot_bytes;
              mbar_wait(sm.aempty0 + 8u * idx, (r1 ? aph2 : aph) ^ 1u);
              if (leader) mbar_expect_tx(sm.afull0 + 8u * idx, txm * abytes);
              load_a(dst, tm, afull_r + 8u * idx, c * kChunkK, ax, ay, n0);
              if (r1) { if (++a2 == (uint32_t)p.a_slots2) { a2 = 0; aph2 ^= 1u; } }
              else if (++a == (uint32_t)p.a_slots) { a = 0; aph ^= 1u; }
              if (p.b_resident) continue;
              for (int t = 0; t < taps; ++t, kcol += kChunkK) {
                mbar_wait(sm.empty0 + 8u * s, ph ^ 1u);
                if (leader) mbar_expect_tx(sm.full0 + 8u * s, txm * b_bytes);
                load_b(sm.bring + s * b_bytes, full_r + 8u * s, kcol, ntile * p.bn + brow0);
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
              }
            }
          }
        }
      } else {
        pdl_wait();
        for (int tile = w_first; tile < w_total; tile += w_step) {
          int ntile, n0, y0, x0, tile_sp;
          tile_coords<C2>(p, tile, crank, ntile, n0, y0, x0, tile_sp);
          const int bx = x0 * p.str_w - p.pad_w, by = y0 * p.str_h - p.pad_h;
          int kcol = 0;
          for (int seg = 0; seg < 2; ++seg) {
            const int chunks = seg == 0 ? p.cin_chunks : (p.ki_total > p.ki0 ? p.cin2_chunks : 0);
            const int kh = seg == 0 ? p.kh : p.taps2 / p.k2w, kw = seg == 0 ? p.kw : p.k2w;
            for (int c = 0; c < chunks; ++c) {
              for (int ky = 0; ky < kh; ++ky) {
                for (int kx = 0; kx < kw; ++kx, kcol += kChunkK) {     // K order is chunk-major: (chunk, tap)
                  mbar_wait(sm.empty0 + 8u * s, ph ^ 1u);
                  const uint32_t a_dst = sm.ring + s * stage_bytes;
                  if (leader) mbar_expect_tx(sm.full0 + 8u * s, txm * stage_bytes);
                  if (seg == 0) load_a(a_dst, &tmA, full_r + 8u * s, c * kChunkK, bx + kx * p.dil_w, by + ky * p.dil_h, n0);
                  else load_a(a_dst, &tmA2, full_r + 8u * s, c * kChunkK, x0 - p.pad2_w + kx, y0 - p.pad2_h + ky, n0);
                  load_b(a_dst + kABytes, full_r + 8u * s, kcol, ntile * p.bn + brow0);
                  if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
                }
              }
            }
          }
        }
      }
#if S2V_PDL_TRIG_CONV == 2
      // every load of this CTA is issued: let the next kernel's launch / prologue overlap our last tiles
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    }
#undef load_a
#undef load_b
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer: the whole warp walks the loop (warp-uniform descriptors live in uniform registers, no
      // per-instruction uniformisation), ONE elected lane issues; in pair mode the leader drives both SMs =====
      // instruction descriptor: D=f32 (1<<4), A=B=f16 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)((C2 ? 2 * kTileM : kTileM) >> 4) << 24);
      const uint32_t hi1024 = desc_hi(1024);
#define mma4(d, alo, ahi, blo, bhi, acc)                                        \
  do {                                                                           \
    if (elect_one()) {                                                           \
      if (C2) {                                                                  \
        umma_f16_lh_2sm(d, alo, ahi, blo, bhi, idesc, acc);                      \
        umma_f16_lh_2sm(d, (alo) + 2u, ahi, (blo) + 2u, bhi, idesc, 1u);         \
        umma_f16_lh_2sm(d, (alo) + 4u, ahi, (blo) + 4u, bhi, idesc, 1u);         \
        umma_f16_lh_2sm(d, (alo) + 6u, ahi, (blo) + 6u, bhi, idesc, 1u);         \
      } else {                                                                   \
        umma_f16_lh(d, alo, ahi, blo, bhi, idesc, acc);                          \
        umma_f16_lh(d, (alo) + 2u, ahi, (blo) + 2u, bhi, idesc, 1u);             \
        umma_f16_lh(d, (alo) + 4u, ahi, (blo) + 4u, bhi, idesc, 1u);             \
        umma_f16_lh(d, (alo) + 6u, ahi, (blo) + 6u, bhi, idesc, 1u);             \
      }                                                                          \
    }                                                                            \
    __syncwarp();                                                                \
  } while (0)
#define commit(bar)                     \
  do {                                  \
    if (elect_one()) {                  \
      if (C2) umma_commit_2sm(bar);     \
      else umma_commit(bar);            \
    }                                   \
    __syncwarp();                       \
  } while (0)
      uint32_t s = 0, ph = 0, a = 0, aph = 0, a2 = 0, aph2 = 0;
      const int chunks2 = p.ki_total > p.ki0 ? p.cin2_chunks : 0;
      int j = 0;
      if (p.halo) {
        // per-segment invariants, hoisted: this single-warp loop is the critical path of every N <= 128 layer
        int sg_chunks[2], sg_kh[2], sg_kw[2], sg_shape[2];
        uint32_t sg_ahi[2], sg_rowstep[2], sg_ru[2], sg_bstep[2];
#pragma unroll
        for (int seg = 0; seg < 2; ++seg) {
          const int taps = seg == 0 ? p.taps0 : p.taps2;
          const int kw = seg == 0 ? p.kw : p.k2w;
          const int pw = seg == 0 ? p.pw0 : p.pw2;
          sg_chunks[seg] = seg == 0 ? p.cin_chunks : chunks2;
          sg_kw[seg] = kw;
          sg_kh[seg] = taps / kw;
          // 8-row core groups of the M tile: contiguous rows for 1x1 (SBO 1024 B); for k > 1 the box is 8 pixels
          // wide, one group per image row of the patch -> SBO = patch width * 128 B
          sg_ahi[seg] = desc_hi(taps > 1 ? (uint32_t)pw * 128u : 1024u);
          sg_rowstep[seg] = (uint32_t)(pw - kw) * 8u;        // (addr >> 4) units: next patch row after kw taps
          sg_ru[seg] = (uint32_t)pw * 8u;
          sg_bstep[seg] = (uint32_t)taps * (b_bytes >> 4);
          sg_shape[seg] = p.b_resident ? issue_shape(sg_kh[seg], kw) : 0;
        }
        const uint32_t bu = b_bytes >> 4;
        const uint32_t bu_n = ((uint32_t)(C2 ? (p.bn_narrow >> 1) : p.bn_narrow) * 128u) >> 4;
        const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.bn_narrow >> 3) << 17) | ((uint32_t)((C2 ? 2 * kTileM : kTileM) >> 4) << 24);
        const uint32_t a_slot_units = (uint32_t)p.a_slot_bytes >> 4, a_slots = (uint32_t)p.a_slots;
        const uint32_t a_slot2_units = (uint32_t)p.a_slot2_bytes >> 4, a_slots2 = (uint32_t)p.a_slots2;
        const bool ring2 = a_slots2 > 0;
        const uint32_t ring_lo = desc_lo(sm.ring), bring_lo = desc_lo(sm.bring);
        const uint32_t ring1_lo = ring_lo + a_slots * a_slot_units;
        const bool b_res = p.b_resident != 0;
        if (b_res) { mbar_wait(sm.ball, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        for (int tile = w_first; tile < w_total; tile += w_step, ++j) {
          const uint32_t buf = (uint32_t)j & 1u;
#ifdef S2V_EPI_PROF
          const bool mprof = blockIdx.x == 0 && lane == 0;
          long long mt0 = clock64(), mt1, mw_a = 0, mw_i = 0;
#endif
          mbar_wait(sm.tempty0 + 8u * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef S2V_EPI_PROF
          mt1 = clock64();
#endif
          const uint32_t d_tmem = tmem + buf * (uint32_t)p.tmem_buf_cols;
          uint32_t accum = 0;
          uint32_t b_res_lo = bring_lo;                                  // resident weights: walks the whole matrix per tile
#pragma unroll
          for (int seg = 0; seg < 2; ++seg) {
            const int kh = sg_kh[seg], kw = sg_kw[seg], shape = sg_shape[seg];
            const uint32_t a_hi = sg_ahi[seg], row_step = sg_rowstep[seg], ru = sg_ru[seg];
            for (int c = 0; c < sg_chunks[seg]; ++c) {
#ifdef S2V_EPI_PROF
              const long long ca0 = clock64();
#endif
              const bool r1 = ring2 && seg == 1;
              const uint32_t aidx = r1 ? a_slots + a2 : a;
              mbar_wait(sm.afull0 + 8u * aidx, r1 ? aph2 : aph);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef S2V_EPI_PROF
              const long long ca1 = clock64();
              mw_a += ca1 - ca0;
#endif
              uint32_t a_lo = r1 ? ring1_lo + a2 * a_slot2_units : ring_lo + a * a_slot_units;
              if (shape) {
                // narrow chunks (structural zero block of the weights): same A patch, N = bn_narrow columns of the accumulator
                const bool nrw = seg == 0 && p.bn_narrow != 0 && c >= p.nf_chunk;
                const uint32_t idc = nrw ? idesc_n : idesc, buc = nrw ? bu_n : bu;
                if (!(p.exp & 2) && elect_one()) {
                  switch (shape) {
                    case 1: issue_chunk_resident<1, C2>(d_tmem, a_lo, a_hi, ru, b_res_lo, buc, hi1024, idc, accum, kh); break;
                    case 2: issue_chunk_resident<2, C2>(d_tmem, a_lo, a_hi, ru, b_res_lo, buc, hi1024, idc, accum, kh); break;
                    case 3: issue_chunk_resident<3, C2>(d_tmem, a_lo, a_hi, ru, b_res_lo, buc, hi1024, idc, accum, kh); break;
                    default: issue_chunk_resident<7, C2>(d_tmem, a_lo, a_hi, ru, b_res_lo, buc, hi1024, idc, accum, kh); break;
                  }
                }
                __syncwarp();
                b_res_lo += (uint32_t)(kh * kw) * buc;
                accum = 1u;
              } else {
                for (int ky = 0; ky < kh; ++ky, a_lo += row_step) {
                  for (int kx = 0; kx < kw; ++kx, a_lo += 8u) {          // +128 B = next pixel of the patch row
                    uint32_t b_lo;
                    if (b_res) {
                      b_lo = b_res_lo;
                      b_res_lo += bu;
                    } else {
                      mbar_wait(sm.full0 + 8u * s, ph);
                      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                      b_lo = bring_lo + s * bu;
                    }
                    mma4(d_tmem, a_lo, a_hi, b_lo, hi1024, accum);
                    accum = 1u;
                    if (!b_res) {
                      commit(sm.empty0 + 8u * s);
                      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
                    }
                  }
                }
              }
              commit(sm.aempty0 + 8u * aidx);       // patch slot free once these MMAs retire
              if (r1) { if (++a2 == a_slots2) { a2 = 0; aph2 ^= 1u; } }
              else if (++a == a_slots) { a = 0; aph ^= 1u; }
#ifdef S2V_EPI_PROF
              mw_i += clock64() - ca1;
#endif
            }
          }
          commit(sm.tfull0 + 8u * buf);               // accumulator of this tile complete (signalled to both CTAs)
#ifdef S2V_EPI_PROF
          if (mprof) {
            atomicAdd(&g_epi_prof[9], (unsigned long long)(mt1 - mt0));     // wait for the accumulator buffer
            atomicAdd(&g_epi_prof[10], (unsigned long long)mw_a);           // wait for A patches
            atomicAdd(&g_epi_prof[11], (unsigned long long)mw_i);           // issue
            atomicAdd(&g_epi_prof[12], 1ull);
          }
#endif
        }
      } else {
        const uint32_t stages = (uint32_t)p.stages;
        for (int tile = w_first; tile < w_total; tile += w_step, ++j) {
          const uint32_t buf = (uint32_t)j & 1u;
          mbar_wait(sm.tempty0 + 8u * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem + buf * (uint32_t)p.tmem_buf_cols;
          uint32_t accum = 0;
          for (int it = 0; it < KI; ++it) {
            mbar_wait(sm.full0 + 8u * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = sm.ring + s * stage_bytes;
            const uint32_t a_lo = desc_lo(a_addr), b_lo = desc_lo(a_addr + kABytes);
            // advance 16 fp16 = 32 B along K inside the 128 B swizzle span: +2 in the (addr>>4) field
            mma4(d_tmem, a_lo, hi1024, b_lo, hi1024, accum);
            accum = 1u;
            commit(sm.empty0 + 8u * s);             // frees the smem slot (of both CTAs) when these MMAs retire
            if (++s == stages) { s = 0; ph ^= 1u; }
          }
          commit(sm.tfull0 + 8u * buf);               // accumulator of this tile complete (signalled to both CTAs)
        }
      }
    }
#undef mma4
#undef commit
  } else {
    epilogue_loop<ACT, C2>(p, sm, &tmY, tmem, warp, lane, crank);
    if (threadIdx.x == 64 || threadIdx.x == 192) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores landed
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (C2) cluster_sync_all(); else __syncthreads();     // pair mode: nobody leaves while the peer may still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (C2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
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

}  // namespace s2v

using namespace s2v;

#ifdef S2V_EPI_PROF
extern "C" int s2v_dbg_epi_prof(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, g_epi_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return S2V_ECUDA;
  if (reset) { unsigned long long z[16] = {}; if (cudaMemcpyToSymbol(g_epi_prof, z, sizeof(z)) != cudaSuccess) return S2V_ECUDA; }
  return S2V_OK;
}
#endif

extern "C" int s2v_conv_tc_tile_n(int cout) {
  static int bn_max = 0;                   // development knob, resolved once
  if (!bn_max) { const char* e = getenv("S2V_BN_MAX"); bn_max = e ? atoi(e) : 256; if (bn_max != 128) bn_max = 256; }
  if (bn_max == 128 && cout > 128 && cout % 128 == 0) return 128;
  if (cout <= 256) return ((cout + 15) / 16) * 16;
  // pick the N tile that wastes the least: 256, 192 or 128
  const int cands[3] = {256, 192, 128};
  int best = 256, best_waste = 1 << 30;
  for (int i = 0; i < 3; ++i) {
    const int w = ceil_div(cout, cands[i]) * cands[i] - cout;
    if (w < best_waste) { best_waste = w; best = cands[i]; }
  }
  return best;
}

extern "C" int s2v_conv_tc(const s2v_conv* d, int box_w, int box_h, int box_n, void* stream) {
  if (!d || !view_ok(&d->x) || !d->w) return S2V_EINVAL;
  if (d->out_mode == S2V_OUT_F16_NHWC && !view_ok(&d->y)) return S2V_EINVAL;
  if (d->out_mode == S2V_OUT_F32_NCHW && !d->y_f32) return S2V_EINVAL;
  if (d->stride_h <= 0 || d->stride_w <= 0 || d->up2 || d->pad_mode != S2V_PAD_ZERO) return S2V_EINVAL;
  if (box_w * d->stride_w > 256 || box_h * d->stride_h > 256) return S2V_EINVAL;
  if (d->kh <= 0 || d->kw <= 0 || d->dil_h <= 0 || d->dil_w <= 0) return S2V_EINVAL;
  if (box_w <= 0 || box_h <= 0 || box_n <= 0 || box_w * box_h * box_n != kTileM || box_w > 256 || box_h > 256 || box_n > 256) return S2V_EINVAL;
  if ((box_w * box_h * box_n) % 8) return S2V_EINVAL;
  const int N = d->x.n, OH = d->y.h, OW = d->y.w, cout = d->y.c;
  if (d->y.n != N) return S2V_EINVAL;
  if (OH <= 0 || OW <= 0) return S2V_EINVAL;   // rows/cols beyond the input are TMA zero fill (bottom/right padding)
  if ((cout % 8) && (d->out_mode != S2V_OUT_F32_NCHW || d->res1.ptr || d->res2.ptr)) return S2V_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return S2V_EUNSUPPORTED;

  TcParams p;
  p.N = N; p.OH = OH; p.OW = OW;
  p.box_w = box_w; p.box_h = box_h; p.box_n = box_n;
  p.tiles_w = ceil_div(OW, box_w); p.tiles_h = ceil_div(OH, box_h);
  const int tiles_n = ceil_div(N, box_n);
  p.kh = d->kh; p.kw = d->kw; p.pad_h = d->pad_h; p.pad_w = d->pad_w; p.dil_h = d->dil_h; p.dil_w = d->dil_w;
  p.str_h = d->stride_h; p.str_w = d->stride_w;
  p.cin_chunks = ceil_div(d->x.c, kChunkK);
  p.cout = cout;
  int bn = s2v_conv_tc_tile_n(cout);
  {
    // Tiny feature maps (<= 64 output pixels per image: the 8 x 8 level of DNet's hourglass, MappingNet, the tail of the audio
    // encoder, the AdaIN / style MLPs): even at batch 64 there are only a few M tiles, and with one wide N tile only that many
    // CTAs run and EACH streams the whole weight matrix through its own L2 port (1.2 MB for a 3x3 256 -> 256 layer: ~20 us for
    // ~1 us of MMAs).  N tiles of 64 spread the weight stream over 2-4 x more SMs (decoder4 phase convs 24 -> 12 us, MappingNet
    // 25 -> 13 us).  The rule depends on the LAYER geometry only, never on the batch: the tile shape a frame is computed with must
    // not change with the batch it sits in (sharded == unsharded bit for bit).  Not with LayerNorm2d totals (one thread sums a
    // tile's channels) or the narrow hint (both need a single N tile).
    static const int split_env = [] { const char* e = getenv("S2V_SPLIT_N"); return e ? atoi(e) : 1; }();      // development knob
    const bool single_only = (d->stats_partial && d->stats_gmax == 0) || d->narrow_cout > 0;
    if (split_env && !single_only && OH * OW <= 64 && cout >= 128 && cout % 64 == 0) bn = 64;
  }
  p.bn = bn;
  // two accumulator buffers (epilogue of tile j overlaps the main loop of tile j+1)
  p.tmem_buf_cols = bn <= 16 ? 16 : bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256;
  p.tmem_cols = 2 * p.tmem_buf_cols < 32 ? 32 : 2 * p.tmem_buf_cols;
  const bool seg2 = d->x2.ptr != nullptr;
  if (seg2 && (!view_ok(&d->x2) || d->x2.n != N || d->k2h <= 0 || d->k2w <= 0 || d->stride_h != 1 || d->stride_w != 1)) return S2V_EINVAL;
  p.taps0 = d->kh * d->kw;
  p.taps2 = seg2 ? d->k2h * d->k2w : 1;
  p.ki0 = p.taps0 * p.cin_chunks;
  p.cin2_chunks = seg2 ? ceil_div(d->x2.c, kChunkK) : 1;
  p.k2w = seg2 ? d->k2w : 1; p.pad2_h = d->pad2_h; p.pad2_w = d->pad2_w;
  p.ki_total = p.ki0 + (seg2 ? p.taps2 * p.cin2_chunks : 0);
  p.n_tiles_n = ceil_div(cout, bn);
  p.tab = (d->scale || d->bias) ? (bn + 31) / 32 * 32 : 0;     // identity epilogues carry no tables
  // halo mode: stride 1, no dilation, and (1x1, any box) or (k > 1 with an 8-pixel-wide single-image box)
  const bool unit = d->stride_h == 1 && d->stride_w == 1 && d->dil_h == 1 && d->dil_w == 1;
  const bool box816 = box_w == 8 && box_n == 1;
  p.pw0 = box_w + (p.taps0 > 1 ? d->kw - 1 : 0);
  const int ph0 = box_h + (p.taps0 > 1 ? d->kh - 1 : 0);
  p.prows0 = p.pw0 * ph0 * box_n;
  p.pw2 = box_w + (seg2 && p.taps2 > 1 ? d->k2w - 1 : 0);
  const int ph2 = box_h + (seg2 && p.taps2 > 1 ? d->k2h - 1 : 0);
  p.prows2 = seg2 ? p.pw2 * ph2 * box_n : 0;
  const bool halo_ok = unit && (p.taps0 == 1 || box816) && (!seg2 || p.taps2 == 1 || box816) && p.pw0 <= 256 && ph0 <= 256 &&
                       p.pw2 <= 256 && ph2 <= 256;
  const int prmax = p.prows0 > p.prows2 ? p.prows0 : p.prows2;
  const int a_slot_bytes = (prmax * 128 + 1023) / 1024 * 1024;

  // ---- shared-memory plan.  In order of preference:
  //  * weights RESIDENT (single N tile, halo-capable geometry): loaded once per CTA, every tile only streams its A patches;
  //  * TWO output staging buffers = two epilogue groups (tiles j and j+1 drain concurrently) - with 64-column passes if
  //    128-column ones do not fit;  one staging buffer / one group otherwise;
  //  * streaming weights: halo mode (one A patch per chunk serves all taps) when taps share a patch, else the plain
  //    (A box + B tile) per tap ring.
  // structural zero block hint (see s2v.h): honoured only with resident weights and a specialised issue shape
  int narrow_n = 0, narrow_taps = 0, nf_chunk = 0;
  if (d->narrow_cout > 0 && d->narrow_cin_from > 0 && d->narrow_cin_from % kChunkK == 0 && d->narrow_cout % 32 == 0 && d->narrow_cout < bn &&
      d->narrow_cin_from < d->x.c && p.n_tiles_n == 1 && (d->kw == 1 || d->kw == 2 || d->kw == 3 || d->kw == 7)) {
    narrow_n = d->narrow_cout;
    nf_chunk = d->narrow_cin_from / kChunkK;
    narrow_taps = (p.cin_chunks - nf_chunk) * p.taps0;
  }
  struct SmemPlan { bool ok, resident, halo; int pass_cols, bufs, a_slots, stages, ring_bytes, stage_out_bytes, b_bytes, a_slots2, s0, s1; };
  const int cap = 227 * 1024 - (16 * 8 + 240 + 4 * p.tab * (int)sizeof(float) + 512 + 1024);
  auto plan_smem = [&](bool pair) -> SmemPlan {
    SmemPlan sp = {};
    const int bb = (pair ? bn / 2 : bn) * 128;
    sp.b_bytes = bb;
    const int nbb = (pair ? narrow_n / 2 : narrow_n) * 128;
    const long long wb = (long long)(p.ki_total - narrow_taps) * bb + (long long)narrow_taps * nbb;   // narrow taps only matter when resident
    auto staging = [&](int pc, int bufs) { return (((bn > pc ? pc : bn) + 63) / 64) * kTileM * 128 * bufs; };
    if (halo_ok && p.n_tiles_n == 1) {
      static const int pc_max = [] { const char* e = getenv("S2V_PC_MAX"); return e && atoi(e) == 64 ? 64 : 128; }();   // development knob
      for (int bufs = 2; bufs >= 1; --bufs)
        for (int pc = (bn > 64 ? pc_max : 128); pc >= 64; pc -= 64) {
          if (pc == 64 && bn <= 64) continue;
          const int so = staging(pc, bufs);
          if (wb + 2LL * a_slot_bytes + so > cap) continue;
          sp.ok = sp.resident = sp.halo = true;
          sp.pass_cols = pc; sp.bufs = bufs; sp.stage_out_bytes = so; sp.stages = 1;
          sp.a_slots = (int)((cap - so - wb) / a_slot_bytes);
          if (sp.a_slots > kMaxASlots) sp.a_slots = kMaxASlots;
          // ~64 KB of patches in flight covers the load latency; beyond that extra slots only cost shared memory
          while (sp.a_slots > 4 && (long long)(sp.a_slots - 1) * a_slot_bytes >= 96 * 1024) --sp.a_slots;
          sp.s0 = a_slot_bytes; sp.s1 = 0; sp.a_slots2 = 0;
          // Few slots and a second K segment: give the segment-2 patches a ring of their own, each ring with slots of its own
          // size (48 x 48 merged FFC GEMM: 2 x 23 KB + 1 x 16 KB instead of 2 x 23 KB; 24 x 24 l2g: 1 x 23 KB + 2 x 16 KB).
          // Measured on B200: the MMA warp of the 48 x 48 GEMM spent 26 % of its time waiting for patches with the shared ring.
          static const int ring2_env = [] { const char* e = getenv("S2V_RING2"); return e ? atoi(e) : 1; }();      // development knob
          if (ring2_env && seg2 && sp.a_slots < 4) {
            const int s0 = (p.prows0 * 128 + 1023) / 1024 * 1024, s1 = (p.prows2 * 128 + 1023) / 1024 * 1024;
            const long long avail = cap - so - wb;
            for (int n1 = p.cin2_chunks < 2 ? p.cin2_chunks : 2; n1 >= 1; --n1) {
              int n0 = (int)((avail - (long long)n1 * s1) / s0);
              if (n0 > kMaxASlots - n1) n0 = kMaxASlots - n1;
              if (n0 >= 1 && n0 + n1 > sp.a_slots) {
                sp.a_slots = n0; sp.a_slots2 = n1; sp.s0 = s0; sp.s1 = s1;
                break;
              }
            }
          }
          sp.ring_bytes = sp.a_slots * sp.s0 + sp.a_slots2 * sp.s1 + (int)wb;
          return sp;
        }
    }
    sp.halo = halo_ok && (p.taps0 > 1 || p.taps2 > 1);
    sp.pass_cols = 128;
    for (int bufs = 2; bufs >= 1; --bufs) {
      const int so = staging(128, bufs), avail = cap - so;
      sp.bufs = bufs; sp.stage_out_bytes = so;
      if (sp.halo) {
        sp.a_slots = p.taps0 > 1 ? 2 : 3;
        if (sp.a_slots * a_slot_bytes + 2 * bb > avail) sp.a_slots = 2;
        sp.stages = (avail - sp.a_slots * a_slot_bytes) / bb;
        if (sp.stages > 8) sp.stages = 8;
        if (sp.stages < (bufs == 2 ? 4 : 2)) continue;
        sp.ring_bytes = sp.a_slots * a_slot_bytes + sp.stages * bb;
      } else {
        sp.a_slots = 0;
        sp.stages = avail / (kABytes + bb);
        if (sp.stages > 8) sp.stages = 8;
        if (sp.stages < (bufs == 2 ? 3 : 2)) continue;
        sp.ring_bytes = sp.stages * (kABytes + bb);
      }
      sp.ok = true;
      return sp;
    }
    return sp;
  };
  // CTA-pair mode (cta_group::2): M = 256 per MMA instruction, each CTA stages half of the B tile -> half the weight
  // traffic / footprint.  Used when that is what makes the weights resident, or for long streaming K loops (measured on
  // B200: 24x24 level 82 -> 56 us, 12x12 3x3 63 -> 60 us); short streaming layers lose (pair synchronisation).
  const int m_tiles_all = p.tiles_w * p.tiles_h * tiles_n;
  SmemPlan sp = plan_smem(false);
  bool cta2 = false;
  if (m_tiles_all >= 2) {
    const char* env2 = getenv("S2V_CTA2");
    if (env2) cta2 = atoi(env2) != 0;
    else if (!sp.resident) {
      const SmemPlan s2 = plan_smem(true);
      cta2 = s2.ok && (s2.resident || p.ki_total >= 16);
    }
    if (cta2) sp = plan_smem(true);
  }
  if (!sp.ok) return S2V_EINVAL;
  const int b_bytes = sp.b_bytes;
  const int stage_bytes = kABytes + b_bytes;
  p.pass_cols = sp.pass_cols;
  p.stage_out_bytes = sp.stage_out_bytes;
  p.stage_bufs = sp.bufs;
  p.epi_groups = sp.bufs == 2 ? 2 : 1;
  p.halo = sp.halo ? 1 : 0;
  p.b_resident = sp.resident ? 1 : 0;
  p.bn_narrow = sp.resident ? narrow_n : 0;
  p.nf_chunk = nf_chunk;
  // L2 prefetch distance in tiles.  Measured on B200 (LNet B=128 / DNet B=64 plans): 0 / 1 / 2 / 4 -> 13.69 / 13.88 / 13.82 / 13.88 ms
  // and 9.29 / 9.52 / 9.41 / 9.39 ms - the patch loads already hit L2 (the producer kernel just wrote the tensor), so it stays off.
  static const int pf_env = [] { const char* e = getenv("S2V_PF"); return e ? atoi(e) : 0; }();      // development knob
  p.pf_dist = (sp.halo && p.n_tiles_n == 1) ? pf_env : 0;
  p.a_slots = sp.halo ? sp.a_slots : 0;
  p.a_slot_bytes = sp.halo ? (sp.a_slots2 ? sp.s0 : a_slot_bytes) : 0;
  p.a_slots2 = sp.halo ? sp.a_slots2 : 0;
  p.a_slot2_bytes = sp.halo ? sp.s1 : 0;
  p.ring_bytes = sp.ring_bytes;
  int stages = sp.stages;
  p.stages = stages;
  p.y = mk(d->y);
  p.r1 = mk(d->res1.ptr ? &d->res1 : nullptr);
  p.r2 = mk(d->res2.ptr ? &d->res2 : nullptr);
  p.scale = d->scale; p.bias = d->bias; p.act = d->act; p.ap = d->act_param;
  p.out_mode = d->out_mode; p.yf = d->y_f32;
  { static const int exp_env = [] { const char* e = getenv("S2V_EXP"); return e ? atoi(e) : 0; }(); p.exp = exp_env; }
  p.stats = d->stats_partial;
  p.st_c_off = d->stats_c_off; p.st_c_total = d->stats_c_total;
  p.st_chunk_off = d->stats_chunk_off; p.st_chunks_total = d->stats_chunks_total;
  p.st_groups = d->stats_groups; p.st_gmax = d->stats_gmax;
  // fused statistics: one partial per (image, spatial tile, channel); every <=128-column epilogue pass must be 32, 64 or
  // 128 columns wide (the lanes sharing a 16-byte chunk form a power-of-two group inside a warp)
  const bool st_bn_ok = bn == 32 || bn == 64 || bn == 128 || bn == 192 || bn == 256;
  if (p.stats && p.st_gmax == 0) {
    // LayerNorm2d totals: [N][chunks][4][2], one N tile only (a tile's channels are summed by one thread)
    if (d->out_mode != S2V_OUT_F16_NHWC || d->res1.ptr || d->res2.ptr || p.st_groups != 4 || p.st_c_total != 4 || p.st_c_off != 0 ||
        p.n_tiles_n != 1 || p.st_chunks_total <= 0 || p.st_chunk_off + p.tiles_w * p.tiles_h > p.st_chunks_total)
      return S2V_EINVAL;
  } else
  if (p.stats && (d->out_mode != S2V_OUT_F16_NHWC || d->res1.ptr || d->res2.ptr || p.st_c_total <= 0 || p.st_chunks_total <= 0 ||
                  p.st_groups != 1 || p.st_gmax != 1 || !st_bn_ok || (p.st_c_off & 7) || (p.st_c_total & 7) ||
                  p.st_chunk_off + p.tiles_w * p.tiles_h > p.st_chunks_total || p.st_c_off + cout > p.st_c_total))
    return S2V_EINVAL;

  CUtensorMap tmA, tmB, tmA2;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)d->x.c, (cuuint64_t)d->x.w, (cuuint64_t)d->x.h, (cuuint64_t)d->x.n};
    cuuint64_t gstr[3] = {(cuuint64_t)d->x.sw * 2, (cuuint64_t)d->x.sh * 2, (cuuint64_t)d->x.sn * 2};
    // strided convs: the box spans box*stride input pixels and TMA's element stride picks every stride-th one
    cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(box_w * d->stride_w), (cuuint32_t)(box_h * d->stride_h), (cuuint32_t)box_n};
    if (p.halo) { box[1] = (cuuint32_t)p.pw0; box[2] = (cuuint32_t)ph0; }     // the patch: box + (k-1) halo
    cuuint32_t es[4] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->x.ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)p.ki_total * kChunkK;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)((cout + 7) / 8 * 8)};   // weight rows are padded to 8 by the packer
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)(cta2 ? bn / 2 : bn)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  CUtensorMap tmBn = tmB;
  if (p.bn_narrow) {
    const cuuint64_t ktot = (cuuint64_t)p.ki_total * kChunkK;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)((cout + 7) / 8 * 8)};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)(cta2 ? p.bn_narrow / 2 : p.bn_narrow)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tmBn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  tmA2 = tmA;
  if (seg2) {
    cuuint64_t gdim[4] = {(cuuint64_t)d->x2.c, (cuuint64_t)d->x2.w, (cuuint64_t)d->x2.h, (cuuint64_t)d->x2.n};
    cuuint64_t gstr[3] = {(cuuint64_t)d->x2.sw * 2, (cuuint64_t)d->x2.sh * 2, (cuuint64_t)d->x2.sn * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(p.halo ? p.pw2 : box_w), (cuuint32_t)(p.halo ? ph2 : box_h), (cuuint32_t)box_n};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tmA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->x2.ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return S2V_ECUDA;
  }
  CUtensorMap tmY = tmA;
  p.use_tma_store = (d->out_mode == S2V_OUT_F16_NHWC && !d->res1.ptr && !d->res2.ptr) ? 1 : 0;
  {
    // development knob: 0 = coalesced manual stores from the staging tile everywhere, 2 = manual stores when a tile takes several
    // passes through ONE staging buffer per group (the next pass otherwise waits for the bulk store to have read the buffer)
    static const int ts_env = [] { const char* e = getenv("S2V_TMA_STORE"); return e ? atoi(e) : 1; }();
    if (ts_env == 0 || (ts_env == 2 && bn > sp.pass_cols)) p.use_tma_store = 0;
  }
  if (p.use_tma_store) {
    cuuint64_t gdim[4] = {(cuuint64_t)d->y.c, (cuuint64_t)d->y.w, (cuuint64_t)d->y.h, (cuuint64_t)d->y.n};
    cuuint64_t gstr[3] = {(cuuint64_t)d->y.sw * 2, (cuuint64_t)d->y.sh * 2, (cuuint64_t)d->y.sn * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->y.ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      p.use_tma_store = 0;                  // e.g. a stride the encoder rejects: fall back to the manual coalesced stores
  }
  // ring | output staging tile(s) | barriers + tmem ptr | scale/bias tables | row-valid mask
  const size_t smem = (size_t)p.ring_bytes + p.stage_out_bytes + 16 * stages + 240 + 4 * p.tab * sizeof(float) + 512 + 1024;
  if (smem > 227 * 1024) return S2V_EINVAL;
  p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  p.total_tiles = p.m_tiles * ceil_div(cout, bn);
  // one kernel per (pair mode, activation): keeps each launch's instruction footprint small (the MMA warp's issue
  // loop is instruction-fetch sensitive)
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
  static const KernelFn kernels[2][6] = {
      {conv_tc_kernel<false, S2V_ACT_NONE>, conv_tc_kernel<false, S2V_ACT_RELU>, conv_tc_kernel<false, S2V_ACT_LRELU>,
       conv_tc_kernel<false, S2V_ACT_SIGMOID>, conv_tc_kernel<false, S2V_ACT_TANH>, conv_tc_kernel<false, S2V_ACT_GELU>},
      {conv_tc_kernel<true, S2V_ACT_NONE>, conv_tc_kernel<true, S2V_ACT_RELU>, conv_tc_kernel<true, S2V_ACT_LRELU>,
       conv_tc_kernel<true, S2V_ACT_SIGMOID>, conv_tc_kernel<true, S2V_ACT_TANH>, conv_tc_kernel<true, S2V_ACT_GELU>}};
  int act_idx;
  switch (d->act) {
    case S2V_ACT_NONE: act_idx = 0; break;
    case S2V_ACT_RELU: act_idx = 1; break;
    case S2V_ACT_LRELU: act_idx = 2; break;
    case S2V_ACT_SIGMOID: act_idx = 3; break;
    case S2V_ACT_TANH: act_idx = 4; break;
    case S2V_ACT_GELU: act_idx = 5; break;
    default: return S2V_EINVAL;
  }
  static DeviceOnce attr;     // function attributes are per device; idempotent
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  if (attr.needed(dev)) {
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < 6; ++k)
        S2V_CUDA_TRY(cudaFuncSetAttribute(kernels[i][k], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr.mark(dev);
  }
  const int n_sm = sm_count(dev);
  if (n_sm <= 0) return S2V_ECUDA;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("S2V_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  if (dbg)
    fprintf(stderr, "conv_tc: N=%d %dx%d cin=%d cout=%d k=%dx%d seg2=%d | bn=%d halo=%d resident=%d a_slots=%d stages=%d pass_cols=%d stage_bufs=%d "
            "epi_groups=%d cta2=%d m_tiles=%d smem=%zu narrow=%d a_slots2=%d\n", N, OH, OW, d->x.c, cout, d->kh, d->kw, (int)seg2, bn, p.halo, p.b_resident, p.a_slots,
            stages, p.pass_cols, p.stage_bufs, p.epi_groups, (int)cta2, p.m_tiles, smem, p.bn_narrow, p.a_slots2);
  if (cta2) {
    p.m_pairs = (p.m_tiles + 1) / 2;
    p.total_pair_tiles = p.m_pairs * p.n_tiles_n;
    int pairs = n_sm / 2;
    if (pairs > p.total_pair_tiles) pairs = p.total_pair_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    S2V_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernels[1][act_idx], tmA, tmB, tmA2, tmY, tmBn, p));
  } else {
    p.m_pairs = 0; p.total_pair_tiles = 0;
    const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;      // persistent: one CTA per SM
    S2V_CUDA_TRY(launch_pdl(kernels[0][act_idx], grid, kThreads, smem, (cudaStream_t)stream, tmA, tmB, tmA2, tmY, tmBn, p));
  }
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
