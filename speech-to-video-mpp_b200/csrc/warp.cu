// Fused flow warp: convert_flow_to_deformation + bilinear grid resize + grid_sample
// (reference: futils/flow_util.py:3-15, :41-56; closed form SURVEY A.4) as ONE
// memory-bound kernel.  fp32 NCHW in/out (the module boundary), optional fp16 NHWC
// side output feeding DNet's editing net.
//
// Algorithmic bytes per frame (C=3, 256^2, flow 64^2): src 786432 + flow 32768 +
// out 786432 = 1605632 B.  One thread per output pixel, consecutive threads along X
// (coalesced stores; gathers are spatially local and hit L1/L2).
#include "common.cuh"

namespace s2v {

// deformation value at flow cell (i,j): grid + 2*flow/(size-1)   (flow_util.py:12-14,30-31)
__device__ __forceinline__ float2 deform_at(const float* __restrict__ fx, const float* __restrict__ fy, int i, int j,
                                            int w, float inv_wm1, float inv_hm1) {
  float gx = __fadd_rn(__fmul_rn(2.f, __fmul_rn((float)j, inv_wm1)), -1.f);
  float gy = __fadd_rn(__fmul_rn(2.f, __fmul_rn((float)i, inv_hm1)), -1.f);
  float dx = __fadd_rn(gx, __fmul_rn(2.f, __fmul_rn(fx[i * w + j], inv_wm1)));
  float dy = __fadd_rn(gy, __fmul_rn(2.f, __fmul_rn(fy[i * w + j], inv_hm1)));
  return make_float2(dx, dy);
}

// grid_sample, bilinear, zeros padding, align_corners=False.  Branch-free: neighbour coordinates are clamped and
// the weights of out-of-range neighbours zeroed, so the 4*C gathers are issued back to back (ILP) before any FMA.
template <int CMAX>
__device__ __forceinline__ void sample_store(const float* __restrict__ src, float* __restrict__ out, int C, int H,
                                             int W, int Y, int X, float gx, float gy, __half* out16) {
  float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 0.5f);
  float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 0.5f);
  // keep the float -> int conversion defined for huge / NaN coordinates (they sample nothing)
  ix = (fabsf(ix) < 1e9f) ? ix : -10.f;
  iy = (fabsf(iy) < 1e9f) ? iy : -10.f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
  const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
  const float w00 = (vy0 & vx0) ? wx0 * wy0 : 0.f, w01 = (vy0 & vx1) ? wx1 * wy0 : 0.f;
  const float w10 = (vy1 & vx0) ? wx0 * wy1 : 0.f, w11 = (vy1 & vx1) ? wx1 * wy1 : 0.f;
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  const int o00 = cy0 * W + cx0, o01 = cy0 * W + cx1, o10 = cy1 * W + cx0, o11 = cy1 * W + cx1;
  const size_t plane = (size_t)H * W;
  const size_t po = (size_t)Y * W + X;
  if (C <= CMAX) {
    float v[CMAX][4];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        const float* sp = src + c * plane;
        v[c][0] = __ldg(sp + o00); v[c][1] = __ldg(sp + o01); v[c][2] = __ldg(sp + o10); v[c][3] = __ldg(sp + o11);
      }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        // same accumulation order as the reference kernel: nw, ne, sw, se
        float r = v[c][0] * w00;
        r += v[c][1] * w01;
        r += v[c][2] * w10;
        r += v[c][3] * w11;
        out[c * plane + po] = r;
        if (out16) out16[c] = __float2half_rn(r);
      }
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const float* sp = src + c * plane;
      float r = __ldg(sp + o00) * w00;
      r += __ldg(sp + o01) * w01;
      r += __ldg(sp + o10) * w10;
      r += __ldg(sp + o11) * w11;
      out[c * plane + po] = r;
      if (out16) out16[c] = __float2half_rn(r);
    }
  }
}

constexpr int kWarpBX = 32, kWarpBY = 8;

// block = 32 x 8 output pixels; the deformation values of the flow cells under the tile are computed once into
// shared memory (every 4x4 pixels share their 4 corner cells when the flow is 4x coarser than the image).
__global__ void __launch_bounds__(kWarpBX * kWarpBY, 4) flow_warp_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                                       float* __restrict__ out, int C, int H, int W, int h, int w,
                                                                       View o16, int c_off) {
  pdl_trigger();
  pdl_wait();
  constexpr int kMaxCells = 36 * 12;            // smem cell tile capacity (falls back to direct loads beyond it)
  __shared__ float2 s_def[kMaxCells];
  const int X = blockIdx.x * kWarpBX + threadIdx.x;
  const int Y = blockIdx.y * kWarpBY + threadIdx.y;
  const int b = blockIdx.z;
  const float* fx = flow + (size_t)b * 2 * h * w;
  const float* fy = fx + (size_t)h * w;
  const float inv_wm1 = 1.f / (float)(w - 1), inv_hm1 = 1.f / (float)(h - 1);
  const bool same = (h == H && w == W);
  // flow-cell rectangle touched by this tile (bilinear resize of the grid reads cells floor(s) and floor(s)+1)
  const float sch = (float)h / (float)H, scw = (float)w / (float)W;
  int cy_lo = 0, cx_lo = 0, ncx = 0, ncy = 0;
  bool tiled = false;
  if (!same) {
    const int ty0 = blockIdx.y * kWarpBY, tx0 = blockIdx.x * kWarpBX;
    cy_lo = min((int)fmaxf(sch * ((float)ty0 + 0.5f) - 0.5f, 0.f), h - 1);
    cx_lo = min((int)fmaxf(scw * ((float)tx0 + 0.5f) - 0.5f, 0.f), w - 1);
    const int cy_hi = min((int)fmaxf(sch * ((float)min(ty0 + kWarpBY, H) - 0.5f) - 0.5f, 0.f) + 1, h - 1);
    const int cx_hi = min((int)fmaxf(scw * ((float)min(tx0 + kWarpBX, W) - 0.5f) - 0.5f, 0.f) + 1, w - 1);
    ncx = cx_hi - cx_lo + 1; ncy = cy_hi - cy_lo + 1;
    tiled = ncx * ncy <= kMaxCells;
    if (tiled) {
      for (int i = threadIdx.y * kWarpBX + threadIdx.x; i < ncx * ncy; i += kWarpBX * kWarpBY) {
        const int ci = i / ncx, cj = i - ci * ncx;
        s_def[i] = deform_at(fx, fy, cy_lo + ci, cx_lo + cj, w, inv_wm1, inv_hm1);
      }
    }
    __syncthreads();
  }
  if (X >= W || Y >= H) return;
  float gx, gy;
  if (same) {
    float2 d = deform_at(fx, fy, Y, X, w, inv_wm1, inv_hm1);
    gx = d.x; gy = d.y;
  } else {
    // F.interpolate(mode='bilinear', align_corners=False) of the deformation grid
    float sy = fmaxf(sch * ((float)Y + 0.5f) - 0.5f, 0.f);
    float sx = fmaxf(scw * ((float)X + 0.5f) - 0.5f, 0.f);
    int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f);
    float hy = 1.f - ly, hx = 1.f - lx;
    float2 d00, d01, d10, d11;
    if (tiled) {
      d00 = s_def[(y0 - cy_lo) * ncx + (x0 - cx_lo)]; d01 = s_def[(y0 - cy_lo) * ncx + (x1 - cx_lo)];
      d10 = s_def[(y1 - cy_lo) * ncx + (x0 - cx_lo)]; d11 = s_def[(y1 - cy_lo) * ncx + (x1 - cx_lo)];
    } else {
      d00 = deform_at(fx, fy, y0, x0, w, inv_wm1, inv_hm1); d01 = deform_at(fx, fy, y0, x1, w, inv_wm1, inv_hm1);
      d10 = deform_at(fx, fy, y1, x0, w, inv_wm1, inv_hm1); d11 = deform_at(fx, fy, y1, x1, w, inv_wm1, inv_hm1);
    }
    gx = hy * (hx * d00.x + lx * d01.x) + ly * (hx * d10.x + lx * d11.x);
    gy = hy * (hx * d00.y + lx * d01.y) + ly * (hx * d10.y + lx * d11.y);
  }
  __half* p16 = o16.p ? o16.p + (size_t)b * o16.sn + (size_t)Y * o16.sh + (size_t)X * o16.sw + c_off : nullptr;
  sample_store<4>(src + (size_t)b * C * H * W, out + (size_t)b * C * H * W, C, H, W, Y, X, gx, gy, p16);
}

// ---- 4 pixels per thread --------------------------------------------------------------------------------------------
// The one-pixel-per-thread kernel above executes ~400 instructions per pixel (tile bounding box, scale divisions, 64-bit
// addressing - all per thread) and is issue-bound at 1.2-1.4 TB/s whatever the flow looks like.  Here a thread owns a
// COLUMN of 4 pixels (lanes = consecutive x, so every gather / store instruction of a warp touches one or two 128-byte
// lines when the flow is smooth): the per-thread setup and the column interpolation are shared, the resize scales come
// precomputed from the host (same fp32 divisions), offsets are 32-bit and the 48 gathers of the 4 pixels are issued
// before any arithmetic.  Arithmetic (operation order, rounding) is identical to the kernel above.
constexpr int kW4BX = 32, kW4BY = 4, kW4Cells = 16 * 12;      // block = 32 x 16 pixels

// PACK (C <= 4, c_off == C): the fp16 side output is the whole 8-channel texel [src | warp | 0] in one 16-byte store (the source
// pixel is read once more, coalesced) - replaces a separate pack launch and three 2-byte partial-sector stores per pixel.
template <bool PACK>
__global__ void __launch_bounds__(kW4BX * kW4BY) flow_warp4_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                                  float* __restrict__ out, int C, int H, int W, int h, int w,
                                                                  View o16, int c_off, float sch, float scw, float inv_wm1,
                                                                  float inv_hm1) {
  pdl_trigger();
  pdl_wait();
  __shared__ float2 s_def[kW4Cells];
  const int b = blockIdx.z;
  const float* fx = flow + (size_t)b * 2 * h * w;
  const float* fy = fx + (size_t)h * w;
  const int ty0 = blockIdx.y * (kW4BY * 4), tx0 = blockIdx.x * kW4BX;
  // flow-cell rectangle under this tile (block-uniform)
  const int cy_lo = min((int)fmaxf(sch * ((float)ty0 + 0.5f) - 0.5f, 0.f), h - 1);
  const int cx_lo = min((int)fmaxf(scw * ((float)tx0 + 0.5f) - 0.5f, 0.f), w - 1);
  const int cy_hi = min((int)fmaxf(sch * ((float)min(ty0 + kW4BY * 4, H) - 0.5f) - 0.5f, 0.f) + 1, h - 1);
  const int cx_hi = min((int)fmaxf(scw * ((float)min(tx0 + kW4BX, W) - 0.5f) - 0.5f, 0.f) + 1, w - 1);
  const int ncx = cx_hi - cx_lo + 1, ncy = cy_hi - cy_lo + 1;          // <= kW4Cells (checked on the host)
  for (int i = threadIdx.y * kW4BX + threadIdx.x; i < ncx * ncy; i += kW4BX * kW4BY) {
    const int ci = i / ncx, cj = i - ci * ncx;
    s_def[i] = deform_at(fx, fy, cy_lo + ci, cx_lo + cj, w, inv_wm1, inv_hm1);
  }
  __syncthreads();
  const int X = tx0 + threadIdx.x, Y0 = ty0 + threadIdx.y * 4;
  if (X >= W || Y0 >= H) return;
  // column interpolation of the deformation grid (shared by the 4 pixels)
  const float sx = fmaxf(scw * ((float)X + 0.5f) - 0.5f, 0.f);
  const int x0 = min((int)sx, w - 1);
  const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f), hx = 1.f - lx;
  const int plane = H * W;
  const float* sb = src + (size_t)b * C * plane;
  int off[4][4];
  float wt[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int Y = Y0 + p;                                   // rows beyond H (H % 4 != 0 never reaches here) are not stored
    const float sy = fmaxf(sch * ((float)Y + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)sy, h - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
    const float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), hy = 1.f - ly;
    const float2* r0 = s_def + (y0 - cy_lo) * ncx - cx_lo;
    const float2* r1 = s_def + (y1 - cy_lo) * ncx - cx_lo;
    const float2 d00 = r0[x0], d01 = r0[x1], d10 = r1[x0], d11 = r1[x1];
    const float gx = hy * (hx * d00.x + lx * d01.x) + ly * (hx * d10.x + lx * d11.x);
    const float gy = hy * (hx * d00.y + lx * d01.y) + ly * (hx * d10.y + lx * d11.y);
    // grid_sample, bilinear, zeros padding, align_corners=False (same arithmetic as sample_store)
    float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 0.5f);
    float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 0.5f);
    ix = (fabsf(ix) < 1e9f) ? ix : -10.f;
    iy = (fabsf(iy) < 1e9f) ? iy : -10.f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int ax0 = (int)fx0, ay0 = (int)fy0, ax1 = ax0 + 1, ay1 = ay0 + 1;
    const bool vx0 = (ax0 >= 0) & (ax0 < W), vx1 = (ax1 >= 0) & (ax1 < W);
    const bool vy0 = (ay0 >= 0) & (ay0 < H), vy1 = (ay1 >= 0) & (ay1 < H);
    wt[p][0] = (vy0 & vx0) ? wx0 * wy0 : 0.f; wt[p][1] = (vy0 & vx1) ? wx1 * wy0 : 0.f;
    wt[p][2] = (vy1 & vx0) ? wx0 * wy1 : 0.f; wt[p][3] = (vy1 & vx1) ? wx1 * wy1 : 0.f;
    const int cx0 = min(max(ax0, 0), W - 1), cx1 = min(max(ax1, 0), W - 1);
    const int cy0 = min(max(ay0, 0), H - 1) * W, cy1 = min(max(ay1, 0), H - 1) * W;
    off[p][0] = cy0 + cx0; off[p][1] = cy0 + cx1; off[p][2] = cy1 + cx0; off[p][3] = cy1 + cx1;
  }
  const int po = Y0 * W + X;
  float* ob = out + (size_t)b * C * plane + po;
  __half* p16 = o16.p ? o16.p + (size_t)b * o16.sn + (size_t)Y0 * o16.sh + (size_t)X * o16.sw + c_off : nullptr;
  if (PACK) {                                      // C <= 4: every plane's gathers in flight at once, one texel store per pixel
    float v[4][4][4], sv[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) {
        const float* sp = sb + c * plane;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          sv[c][p] = __ldg(sp + po + p * W);
#pragma unroll
          for (int t = 0; t < 4; ++t) v[c][p][t] = __ldg(sp + off[p][t]);
        }
      }
    float aw[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {               // same accumulation order as the reference kernel: nw, ne, sw, se
          float a = v[c][p][0] * wt[p][0];
          a += v[c][p][1] * wt[p][1];
          a += v[c][p][2] * wt[p][2];
          a += v[c][p][3] * wt[p][3];
          ob[c * plane + p * W] = a;
          aw[c][p] = a;
        }
      }
    __half* tp = p16 - c_off;                       // channel 0 of the texel (16-byte aligned: checked on the host)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {                 // slot i = src channel i (i < C), warp channel i - C (C <= i < 2 C), else 0
        float val = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < C && i == c) val = sv[c][p];
          if (c < C && i == C + c) val = aw[c][p];
        }
        f[i] = val;
      }
      st_h8(tp + (size_t)p * o16.sh, f_to_h8(f));
    }
  } else {
  for (int c0 = 0; c0 < C; c0 += 3) {              // 3 planes at a time: 48 gathers in flight
    float v[3][4][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) {
        const float* sp = sb + (c0 + c) * plane;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int t = 0; t < 4; ++t) v[c][p][t] = __ldg(sp + off[p][t]);
      }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {               // same accumulation order as the reference kernel: nw, ne, sw, se
          float a = v[c][p][0] * wt[p][0];
          a += v[c][p][1] * wt[p][1];
          a += v[c][p][2] * wt[p][2];
          a += v[c][p][3] * wt[p][3];
          ob[(c0 + c) * plane + p * W] = a;
          if (p16) p16[(size_t)p * o16.sh + c0 + c] = __float2half_rn(a);
        }
      }
  }
  }
}

__global__ void flow_to_deformation_kernel(const float* __restrict__ flow, float* __restrict__ def, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
  if (j >= w) return;
  const float* fx = flow + (size_t)b * 2 * h * w;
  float2 d = deform_at(fx, fx + (size_t)h * w, i, j, w, 1.f / (float)(w - 1), 1.f / (float)(h - 1));
  reinterpret_cast<float2*>(def)[((size_t)b * h + i) * w + j] = d;
}

__global__ void __launch_bounds__(256) warp_deformation_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ def, float* __restrict__ out,
                                                               int C, int H, int W, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, b = blockIdx.z;
  if (X >= W) return;
  const float2* d = reinterpret_cast<const float2*>(def) + (size_t)b * h * w;
  float gx, gy;
  if (h == H && w == W) {
    float2 v = d[Y * w + X];
    gx = v.x; gy = v.y;
  } else {
    const float sch = (float)h / (float)H, scw = (float)w / (float)W;
    float sy = fmaxf(sch * ((float)Y + 0.5f) - 0.5f, 0.f);
    float sx = fmaxf(scw * ((float)X + 0.5f) - 0.5f, 0.f);
    int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f);
    float hy = 1.f - ly, hx = 1.f - lx;
    float2 d00 = d[y0 * w + x0], d01 = d[y0 * w + x1], d10 = d[y1 * w + x0], d11 = d[y1 * w + x1];
    gx = hy * (hx * d00.x + lx * d01.x) + ly * (hx * d10.x + lx * d11.x);
    gy = hy * (hx * d00.y + lx * d01.y) + ly * (hx * d10.y + lx * d11.y);
  }
  sample_store<4>(src + (size_t)b * C * H * W, out + (size_t)b * C * H * W, C, H, W, Y, X, gx, gy, nullptr);
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_flow_warp_f32(const float* src, const float* flow, float* out, int B, int C, int H, int W, int h,
                                 int w, const s2v_view* out16, int c_off, void* stream) {
  if (B == 0) return S2V_OK;
  if (!src || !flow || !out || B < 0 || C <= 0 || H <= 0 || W <= 0 || h < 2 || w < 2) return S2V_EINVAL;
  if (B > 65535 || H > 65535) return S2V_EINVAL;
  const bool pack = (c_off & S2V_WARP_PACK_SRC) != 0;
  c_off &= ~S2V_WARP_PACK_SRC;
  View o16 = mk(out16);
  if (out16 && out16->ptr && (out16->n < B || out16->h != H || out16->w != W || c_off + C > out16->c)) return S2V_EINVAL;
  if (pack && (!out16 || !out16->ptr || !view_ok(out16) || c_off != C || 2 * C > 8 || out16->c < 8)) return S2V_EINVAL;
  const float sch = (float)h / (float)H, scw = (float)w / (float)W;
  const bool fits4 = ((int)(scw * kW4BX) + 3) * ((int)(sch * (kW4BY * 4)) + 3) <= kW4Cells;
  if (!(h == H && w == W) && H % 4 == 0 && fits4 && (long long)C * H * W < (1ll << 31) && !getenv("S2V_WARP1")) {
    dim3 grid(ceil_div(W, kW4BX), ceil_div(H, kW4BY * 4), B);
    if (pack)
      S2V_CUDA_TRY(launch_pdl(flow_warp4_kernel<true>, grid, dim3(kW4BX, kW4BY), 0, (cudaStream_t)stream, src, flow, out, C, H, W, h, w, o16,
                              c_off, sch, scw, 1.f / (float)(w - 1), 1.f / (float)(h - 1)));
    else
      S2V_CUDA_TRY(launch_pdl(flow_warp4_kernel<false>, grid, dim3(kW4BX, kW4BY), 0, (cudaStream_t)stream, src, flow, out, C, H, W, h, w, o16,
                              c_off, sch, scw, 1.f / (float)(w - 1), 1.f / (float)(h - 1)));
    S2V_CHECK_LAUNCH();
    return S2V_OK;
  }
  if (pack) {   // geometries the 4-pixel kernel does not take: the texel is written by the pack launch + the partial side output
    const int rc = s2v_pack_nchw_f32(src, B, C, H, W, (int64_t)C * H * W, out16, 0, 8, 1.f, 0.f, stream);
    if (rc != S2V_OK) return rc;
  }
  dim3 grid(ceil_div(W, kWarpBX), ceil_div(H, kWarpBY), B);
  launch_pdl(flow_warp_kernel, grid, dim3(kWarpBX, kWarpBY), 0, (cudaStream_t)stream, src, flow, out, C, H, W, h, w, o16, c_off);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_flow_to_deformation_f32(const float* flow, float* deformation, int B, int h, int w, void* stream) {
  if (B == 0) return S2V_OK;
  if (!flow || !deformation || B < 0 || h < 2 || w < 2 || B > 65535 || h > 65535) return S2V_EINVAL;
  dim3 grid(ceil_div(w, 128), h, B);
  launch_pdl(flow_to_deformation_kernel, grid, 128, 0, (cudaStream_t)stream, flow, deformation, h, w);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_warp_deformation_f32(const float* src, const float* deformation, float* out, int B, int C, int H,
                                        int W, int h, int w, void* stream) {
  if (B == 0) return S2V_OK;
  if (!src || !deformation || !out || B < 0 || C <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return S2V_EINVAL;
  if (B > 65535 || H > 65535) return S2V_EINVAL;
  dim3 grid(ceil_div(W, 256), H, B);
  launch_pdl(warp_deformation_kernel, grid, 256, 0, (cudaStream_t)stream, src, deformation, out, C, H, W, h, w);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
