// Per-frame image glue around the networks, batched over frames (the reference runs these lines per frame on the CPU with
// numpy / OpenCV):
//   resize_linear_u8   cv2.resize(INTER_LINEAR) of 8-bit images, bit-exact: OpenCV's fixed-point bilinear (11-bit weights,
//                      cvRound(w * 2048); vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; exact 2x2 down-scale =
//                      INTER_AREA (a+b+c+d+2)>>2) - inference.py:292 (face -> box), :308 (frames -> 512 x 512), :392-393 (crops ->
//                      img_size).  Every frame may have its own destination box (y1,y2,x1,x2) inside a larger frame: the resize
//                      writes straight into that window, which IS the paste of inference.py:295-297.
//   resize_linear_f32  the float32 form (mask -> 512 x 512 at :308, blended image back to the frame size at :313, with the
//                      np.clip(0,255) before and the np.uint8 truncation after fused in)
//   fake_to_bgr_u8     preprocessing/facing.py:190-192: clamp(-1,1), (x+1)/2*255, uint8, RGB -> BGR
//   face_batch         inference.py:394-399, :260-262: lower-half mask, concat, / 255, HWC -> CHW
//   compose_pred_u8    inference.py:267, :282-288, :290: clamp(pred,0,1), mask mix with img_original, *255, uint8
// HBM-bound byte work: one thread per output pixel (all channels), coalesced along x.
#include "common.cuh"

namespace s2v {

struct Tap { int i0, i1; float f; };

// source taps of destination index d (OpenCV: fx = (d + 0.5) * scale - 0.5 in float; x taps clamp the FRACTION at the borders,
// y taps only clamp the row index)
__device__ __forceinline__ Tap tap_f32(int d, double scale, int n, bool is_x) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  Tap t;
  if (is_x) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
    t.i0 = s; t.i1 = min(s + 1, n - 1);
  } else {
    t.i0 = min(max(s, 0), n - 1); t.i1 = min(max(s + 1, 0), n - 1);
  }
  t.f = f;
  return t;
}

template <int C>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, long long src_sn, int H, int W,
                                                       uint8_t* __restrict__ dst, long long dst_sn, long long dst_sh,
                                                       const int* __restrict__ boxes, int OH, int OW, int max_pix) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  int oh = OH, ow = OW, oy0 = 0, ox0 = 0;
  if (boxes) { oy0 = boxes[4 * n]; oh = boxes[4 * n + 1] - oy0; ox0 = boxes[4 * n + 2]; ow = boxes[4 * n + 3] - ox0; }
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= max_pix || oh <= 0 || ow <= 0 || pix >= oh * ow) return;
  const int oy = pix / ow, ox = pix - oy * ow;
  const uint8_t* s = src + (size_t)n * src_sn;
  uint8_t* o = dst + (size_t)n * dst_sn + (size_t)(oy0 + oy) * dst_sh + (size_t)(ox0 + ox) * C;
  if (H == 2 * oh && W == 2 * ow) {                 // INTER_LINEAR of an exact 2x down-scale is INTER_AREA in OpenCV
    const uint8_t* p = s + ((size_t)(2 * oy) * W + 2 * ox) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)(((int)p[c] + p[C + c] + p[(size_t)W * C + c] + p[(size_t)W * C + C + c] + 2) >> 2);
    return;
  }
  const Tap tx = tap_f32(ox, (double)W / (double)ow, W, true), ty = tap_f32(oy, (double)H / (double)oh, H, false);
  const int a0 = __float2int_rn((1.f - tx.f) * 2048.f), a1 = __float2int_rn(tx.f * 2048.f);
  const int b0 = __float2int_rn((1.f - ty.f) * 2048.f), b1 = __float2int_rn(ty.f * 2048.f);
  const uint8_t* r0 = s + (size_t)ty.i0 * W * C;
  const uint8_t* r1 = s + (size_t)ty.i1 * W * C;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int s0 = r0[tx.i0 * C + c] * a0 + r0[tx.i1 * C + c] * a1;
    const int s1 = r1[tx.i0 * C + c] * a0 + r1[tx.i1 * C + c] * a1;
    const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

// float32 [N,H,W,C] -> [N,OH,OW,C]; taps in double (cv2's default IPP path; OpenCV's own code rounds the coordinate to float).
// clip_in: np.clip(x, 0, 255) on load; out_u8: np.uint8() truncation on store (inference.py:313)
template <int C>
__global__ void __launch_bounds__(256) resize_f32_kernel(const float* __restrict__ src, int H, int W, void* __restrict__ dst, int OH, int OW,
                                                        int clip_in, int out_u8) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= OH * OW) return;
  const int oy = pix / OW, ox = pix - oy * OW;
  const float* s = src + (size_t)n * H * W * C;
  float r[C];
  auto ld = [&](const float* p) { const float v = *p; return clip_in ? fminf(fmaxf(v, 0.f), 255.f) : v; };
  if (H == 2 * OH && W == 2 * OW) {
    const float* p = s + ((size_t)(2 * oy) * W + 2 * ox) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = (ld(p + c) + ld(p + C + c) + ld(p + (size_t)W * C + c) + ld(p + (size_t)W * C + C + c)) * 0.25f;
  } else {
    double fx = ((double)ox + 0.5) * ((double)W / (double)OW) - 0.5, fy = ((double)oy + 0.5) * ((double)H / (double)OH) - 0.5;
    int sx = (int)floor(fx), sy = (int)floor(fy);
    fx -= (double)sx; fy -= (double)sy;
    if (sx < 0) { fx = 0.0; sx = 0; }
    if (sx >= W - 1) { fx = 0.0; sx = W - 1; }
    const int x1 = min(sx + 1, W - 1), y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
    const float* r0 = s + (size_t)y0 * W * C;
    const float* r1 = s + (size_t)y1 * W * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double h0 = (double)ld(r0 + sx * C + c) * (1.0 - fx) + (double)ld(r0 + x1 * C + c) * fx;
      const double h1 = (double)ld(r1 + sx * C + c) * (1.0 - fx) + (double)ld(r1 + x1 * C + c) * fx;
      r[c] = (float)(h0 * (1.0 - fy) + h1 * fy);
    }
  }
  const size_t oo = ((size_t)n * OH * OW + pix) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (out_u8) reinterpret_cast<uint8_t*>(dst)[oo + c] = (uint8_t)(int)fminf(fmaxf(r[c], 0.f), 255.f);
    else reinterpret_cast<float*>(dst)[oo + c] = r[c];
  }
}

__global__ void __launch_bounds__(256) fake_to_bgr_kernel(const float* __restrict__ fake, int H, int W, uint8_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const float* p = fake + (size_t)n * 3 * H * W + pix;
  uint8_t* o = out + ((size_t)n * H * W + pix) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = fminf(fmaxf(p[(size_t)c * H * W], -1.f), 1.f);
    o[2 - c] = (uint8_t)(int)__fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), 255.f);       // np.uint8((x + 1) / 2. * 255), RGB -> BGR
  }
}

__global__ void __launch_bounds__(256) face_batch_kernel(const uint8_t* __restrict__ oface, const uint8_t* __restrict__ face, int S,
                                                        float* __restrict__ img_batch, float* __restrict__ img_original) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= S * S) return;
  const int y = pix / S;
  const uint8_t* po = oface + ((size_t)n * S * S + pix) * 3;
  const uint8_t* pf = face + ((size_t)n * S * S + pix) * 3;
  float* ib = img_batch + (size_t)n * 6 * S * S + pix;
  float* io = img_original + (size_t)n * 3 * S * S + pix;
  const bool masked = y >= S / 2;                   // img_masked[:, img_size//2:] = 0
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ib[(size_t)c * S * S] = masked ? 0.f : (float)((double)po[c] / 255.0);       // numpy: uint8 / 255. is float64, then FloatTensor
    ib[(size_t)(3 + c) * S * S] = (float)((double)pf[c] / 255.0);
    io[(size_t)c * S * S] = __fdiv_rn((float)po[c], 255.f);                      // FloatTensor(...) / 255.
  }
}

__global__ void __launch_bounds__(256) compose_pred_kernel(const float* __restrict__ pred, const float* __restrict__ img_batch,
                                                          const float* __restrict__ img_original, int S, int compose, uint8_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= S * S) return;
  uint8_t* o = out + ((size_t)n * S * S + pix) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p = fminf(fmaxf(pred[((size_t)n * 3 + c) * S * S + pix], 0.f), 1.f);
    if (compose) {
      const float m = img_batch[((size_t)n * 6 + c) * S * S + pix] == 0.f ? 1.f : 0.f;
      p = __fadd_rn(__fmul_rn(p, m), __fmul_rn(img_original[((size_t)n * 3 + c) * S * S + pix], 1.f - m));
    }
    o[c] = (uint8_t)(int)__fmul_rn(p, 255.f);
  }
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_resize_linear_u8(const uint8_t* src, int N, int H, int W, int C, uint8_t* dst, int64_t dst_sn, int64_t dst_sh,
                                    const int32_t* boxes_dev, int OH, int OW, int max_box_pixels, void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !dst || N < 0 || N > 65535 || H <= 0 || W <= 0 || (C != 1 && C != 3)) return S2V_EINVAL;
  if (!boxes_dev && (OH <= 0 || OW <= 0)) return S2V_EINVAL;
  const int max_pix = boxes_dev ? max_box_pixels : OH * OW;
  if (max_pix <= 0) return S2V_EINVAL;
  const dim3 grid(ceil_div(max_pix, 256), N);
  const long long ssn = (long long)H * W * C;
  if (C == 3) S2V_CUDA_TRY(launch_pdl(resize_u8_kernel<3>, grid, 256, 0, (cudaStream_t)stream, src, ssn, H, W, dst, (long long)dst_sn, (long long)dst_sh, boxes_dev, OH, OW, max_pix));
  else S2V_CUDA_TRY(launch_pdl(resize_u8_kernel<1>, grid, 256, 0, (cudaStream_t)stream, src, ssn, H, W, dst, (long long)dst_sn, (long long)dst_sh, boxes_dev, OH, OW, max_pix));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_resize_linear_f32(const float* src, int N, int H, int W, int C, void* dst, int OH, int OW, int clip_in, int out_u8,
                                     void* stream) {
  if (N == 0) return S2V_OK;
  if (!src || !dst || N < 0 || N > 65535 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || (C != 1 && C != 3)) return S2V_EINVAL;
  const dim3 grid(ceil_div((long long)OH * OW, 256), N);
  if (C == 3) S2V_CUDA_TRY(launch_pdl(resize_f32_kernel<3>, grid, 256, 0, (cudaStream_t)stream, src, H, W, dst, OH, OW, clip_in, out_u8));
  else S2V_CUDA_TRY(launch_pdl(resize_f32_kernel<1>, grid, 256, 0, (cudaStream_t)stream, src, H, W, dst, OH, OW, clip_in, out_u8));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_fake_to_bgr_u8(const float* fake, int N, int H, int W, uint8_t* out, void* stream) {
  if (N == 0) return S2V_OK;
  if (!fake || !out || N < 0 || N > 65535 || H <= 0 || W <= 0) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(fake_to_bgr_kernel, dim3(ceil_div((long long)H * W, 256), N), 256, 0, (cudaStream_t)stream, fake, H, W, out));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_face_batch(const uint8_t* oface, const uint8_t* face, int N, int S, float* img_batch, float* img_original, void* stream) {
  if (N == 0) return S2V_OK;
  if (!oface || !face || !img_batch || !img_original || N < 0 || N > 65535 || S <= 0) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(face_batch_kernel, dim3(ceil_div((long long)S * S, 256), N), 256, 0, (cudaStream_t)stream, oface, face, S, img_batch, img_original));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_compose_pred_u8(const float* pred, const float* img_batch, const float* img_original, int N, int S, int compose,
                                   uint8_t* out, void* stream) {
  if (N == 0) return S2V_OK;
  if (!pred || !out || N < 0 || N > 65535 || S <= 0 || (compose && (!img_batch || !img_original))) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(compose_pred_kernel, dim3(ceil_div((long long)S * S, 256), N), 256, 0, (cudaStream_t)stream, pred, img_batch, img_original, S, compose, out));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
