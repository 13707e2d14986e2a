// FourierUnit transforms (reference: models/ffc.py:99-102 rfftn + re/im channel interleave,
// :116-121 de-interleave + irfftn, norm='ortho') for the 12x12 / 24x24 / 48x48 feature maps
// of LNet's FFC decoder.  Sizes are 3*2^k, so the 1-D transform is a mixed radix-2/radix-3
// Cooley-Tukey, fully unrolled on register arrays (compile-time twiddle indices into a
// constant-memory W_48 table); the 2-D transform is row pass -> shared memory -> column pass.
// Two real rows ride one complex FFT (packed real-input trick) in both directions.
// One block = one image n x CB channels; fp16 channels-last I/O, fp32 math.
#include "common.cuh"

namespace s2v {

// fft2d_mma.cu: the same transforms as dense DFT matrix products on mma.sync with the tiles moved by TMA, selected per direction and
// size with S2V_FFT_MMA.  Measured on B200 (profiles/r2c_summary.md, us at B = 128 / 256): 48 x 48 rfft2 18.9 / 33.7 against 23.6 / 53.2
// here, irfft2 27.7 / 46.7 against 37.1 / 75.9 -> the default for both 48 x 48 transforms; at 24 and 12 px the register FFT below is
// as fast or faster (11.4 / 18.2 vs 10.2 / 18.5 and 11.2 / 19.5 vs 9.7 / 13.6 for rfft2) and stays.  The price is one extra fp16 rounding
// (the twiddles): 7e-4 - 1e-3 of the output peak against 3e-4 - 6e-4 (test_fft2), LNet parity unchanged within 0.1 dB.
int fft_mma_init();
int rfft2_mma(const s2v_view* x, const s2v_view* sp, cudaStream_t st);
int irfft2_mma(const s2v_view* sp, const s2v_view* add, const s2v_view* y, cudaStream_t st);

__constant__ float2 c_tw48[48];   // exp(-2*pi*i*j/48)

// (measured on B200: the same butterflies with the sm_100 packed fp32x2 instructions - __fadd2_rn / __ffma2_rn, 22 % fewer
//  floating-point instructions - run in exactly the same time, 24.0 vs 23.7 us for rfft2 48 x 48 at B = 128: the kernels are
//  not bound by their arithmetic instruction count; -DS2V_FFT_PACKED keeps the variant for A/B runs)
#ifdef S2V_FFT_PACKED
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
template <bool INV>
__device__ __forceinline__ float2 cmul_tw(float2 a, float2 w) {   // a * w  (INV: a * conj(w))
  if (INV) return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

template <int N, bool INV>
__device__ __forceinline__ void fft(float2 (&x)[N]) {
  if constexpr (N == 1) {
    return;
  } else if constexpr (N == 3) {
    const float s = INV ? 0.86602540378443865f : -0.86602540378443865f;   // Im of W_3
    const float2 a = x[0], b = x[1], c = x[2];
    const float2 t = cadd(b, c), d = csub(b, c);
    x[0] = cadd(a, t);
    const float2 m = make_float2(a.x - 0.5f * t.x, a.y - 0.5f * t.y);
    const float2 r = make_float2(-s * d.y, s * d.x);      // i*s*d
    x[1] = cadd(m, r);
    x[2] = csub(m, r);
  } else {
    static_assert(N % 2 == 0, "radix-2 split needs even N");
    float2 e[N / 2], o[N / 2];
#pragma unroll
    for (int k = 0; k < N / 2; ++k) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
    fft<N / 2, INV>(e);
    fft<N / 2, INV>(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const float2 t = cmul_tw<INV>(o[k], c_tw48[k * (48 / N)]);
      x[k] = cadd(e[k], t);
      x[k + N / 2] = csub(e[k], t);
    }
  }
}

// x [N,S,S,C] -> spec [N,S,S/2+1,2C]; block (n, CB channels), threads (S/2+1)*CB
template <int S, int CB>
__global__ void __launch_bounds__((S / 2 + 1) * CB) rfft2_kernel(View x, View sp, int rev) {
  pdl_trigger();
  pdl_wait();
  constexpr int K = S / 2 + 1;
  extern __shared__ float2 sm[];      // [S][K][CB]
  const int c = threadIdx.x % CB, t = threadIdx.x / CB;
  // rev: images from the end of the tensor first (the part the producing conv wrote last and L2 still holds)
  const int n = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, ch = blockIdx.x * CB + c;
  if (t < S / 2) {                    // row pass: rows 2t, 2t+1 as one complex signal
    float2 z[S];
    const __half* r0 = x.p + n * x.sn + (2 * t) * x.sh + ch;
    const __half* r1 = r0 + x.sh;
#pragma unroll
    for (int w = 0; w < S; ++w) z[w] = make_float2(__half2float(r0[w * x.sw]), __half2float(r1[w * x.sw]));
    fft<S, false>(z);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float2 a = z[k], b = z[(S - k) % S];
      // R0 = (Z[k] + conj(Z[S-k]))/2 ; R1 = (Z[k] - conj(Z[S-k]))/(2i)
      sm[((2 * t) * K + k) * CB + c] = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      sm[((2 * t + 1) * K + k) * CB + c] = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
    }
  }
  __syncthreads();
  {                                   // column pass: thread (k = t, c)
    const int k = t;
    float2 col[S];
#pragma unroll
    for (int h = 0; h < S; ++h) col[h] = sm[(h * K + k) * CB + c];
    fft<S, false>(col);
    const float nrm = 1.f / (float)S;   // ortho: 1/sqrt(S*S)
    __half* o = sp.p + n * sp.sn + k * sp.sw + 2 * ch;
#pragma unroll
    for (int u = 0; u < S; ++u)
      *reinterpret_cast<__half2*>(o + u * sp.sh) = __floats2half2_rn(col[u].x * nrm, col[u].y * nrm);
  }
}

// spec [N,S,S/2+1,2C] -> y [N,S,S,C] (+ add)
template <int S, int CB>
__global__ void __launch_bounds__((S / 2 + 1) * CB) irfft2_kernel(View sp, View add, View y, int rev) {
  pdl_trigger();
  pdl_wait();
  constexpr int K = S / 2 + 1;
  extern __shared__ float2 sm[];      // [S][K][CB]
  const int c = threadIdx.x % CB, t = threadIdx.x / CB;
  // rev: images from the end of the tensor first (the part the producing conv wrote last and L2 still holds)
  const int n = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, ch = blockIdx.x * CB + c;
  {                                   // inverse along H for column k = t
    const int k = t;
    float2 col[S];
    const __half* p = sp.p + n * sp.sn + k * sp.sw + 2 * ch;
#pragma unroll
    for (int u = 0; u < S; ++u) col[u] = __half22float2(*reinterpret_cast<const __half2*>(p + u * sp.sh));
    fft<S, true>(col);
#pragma unroll
    for (int h = 0; h < S; ++h) sm[(h * K + k) * CB + c] = col[h];
  }
  __syncthreads();
  if (t < S / 2) {                    // c2r along W for rows 2t, 2t+1 packed as one complex inverse
    float2 z[S];
    const float2* y0 = sm + ((2 * t) * K) * CB + c;
    const float2* y1 = sm + ((2 * t + 1) * K) * CB + c;
    // DC and Nyquist columns: imaginary parts are ignored (torch.fft.irfftn semantics)
    z[0] = make_float2(y0[0].x, y1[0].x);
    z[S / 2] = make_float2(y0[(S / 2) * CB].x, y1[(S / 2) * CB].x);
#pragma unroll
    for (int k = 1; k < S / 2; ++k) {
      const float2 a = y0[k * CB], b = y1[k * CB];
      z[k] = make_float2(a.x - b.y, a.y + b.x);          // Y0[k] + i*Y1[k]
      z[S - k] = make_float2(a.x + b.y, b.x - a.y);      // conj(Y0[k]) + i*conj(Y1[k])
    }
    fft<S, true>(z);
    const float nrm = 1.f / (float)S;
    __half* o0 = y.p + n * y.sn + (2 * t) * y.sh + ch;
    __half* o1 = o0 + y.sh;
    if (add.p) {
      const __half* a0 = add.p + n * add.sn + (2 * t) * add.sh + ch;
      const __half* a1 = a0 + add.sh;
      // residual loads in groups of 8 ahead of the stores (add and y may alias as far as the compiler knows)
#pragma unroll
      for (int w0 = 0; w0 < S; w0 += 8) {
        __half r0[8], r1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w0 + j < S) { r0[j] = a0[(w0 + j) * add.sw]; r1[j] = a1[(w0 + j) * add.sw]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w0 + j < S) {
            o0[(w0 + j) * y.sw] = __float2half_rn(fmaf(z[w0 + j].x, nrm, __half2float(r0[j])));
            o1[(w0 + j) * y.sw] = __float2half_rn(fmaf(z[w0 + j].y, nrm, __half2float(r1[j])));
          }
        }
      }
    } else {
#pragma unroll
      for (int w = 0; w < S; ++w) {
        o0[w * y.sw] = __float2half_rn(z[w].x * nrm);
        o1[w * y.sw] = __float2half_rn(z[w].y * nrm);
      }
    }
  }
}

// Channels per block.  The fp16 channels-last rows are read / written 2 bytes per thread, so a block touches CB * 2 contiguous
// bytes per pixel: 16 channels = one full 32-byte sector (8 channels use half of every sector they fetch).  Measured on B200, LNet
// B = 128: 48 x 48 with CB 8 -> 16 (one 400-thread block per SM instead of two 200-thread ones): irfft2 45.8 -> 36.7 us, rfft2
// 24.8 -> 23.3 us; 24 x 24 with CB 16 -> 32: irfft2 12.1 -> 11.1 us.  (Rejected: an fp16 exchange buffer + 96 registers for three
// 48 x 48 blocks per SM - the spills cost more than the occupancy gives, 24.9 -> 29.1 / 45.7 -> 49.3 us.)
static int fft48_cb16() {
  static const int v = [] { const char* e = getenv("S2V_FFT48_CB16"); return e ? atoi(e) : 1; }();     // development knob
  return v;
}
static int fft24_cb32() {
  static const int v = [] { const char* e = getenv("S2V_FFT24_CB32"); return e ? atoi(e) : 1; }();     // development knob
  return v;
}
// S2V_FFT_MMA: bits 0 / 1 / 2 = rfft2 at 48 / 24 / 12 px, bits 3 / 4 / 5 = irfft2 at 48 / 24 / 12 px on the tensor-core path (fft2d_mma.cu)
static bool fft_use_mma(int s, bool inverse) {
  static const int mask = [] { const char* e = getenv("S2V_FFT_MMA"); return e ? atoi(e) : 9; }();       // default: both 48 x 48 transforms (measured faster); 0 = register FFT everywhere
  const int bit = s == 48 ? 0 : s == 24 ? 1 : s == 12 ? 2 : -1;
  return bit >= 0 && ((mask >> (bit + (inverse ? 3 : 0))) & 1);
}
static int fft_rev() {
  static const int rev = [] { const char* e = getenv("S2V_FFT_REV"); return e ? atoi(e) : 0; }();
  return rev;
}

template <int S, int CB>
static int launch_rfft2(const s2v_view* x, const s2v_view* sp, cudaStream_t st) {
  constexpr int K = S / 2 + 1;
  const size_t smem = (size_t)S * K * CB * sizeof(float2);
  static DeviceOnce attr;     // per device; idempotent attribute set (same value every time)
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  if (attr.needed(dev)) { S2V_CUDA_TRY(cudaFuncSetAttribute(rfft2_kernel<S, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr.mark(dev); }
  launch_pdl(rfft2_kernel<S, CB>, dim3(x->c / CB, x->n), K * CB, smem, st, mk(x), mk(sp), fft_rev());
  return cudaGetLastError() == cudaSuccess ? S2V_OK : S2V_ECUDA;
}
template <int S, int CB>
static int launch_irfft2(const s2v_view* sp, const s2v_view* add, const s2v_view* y, cudaStream_t st) {
  constexpr int K = S / 2 + 1;
  const size_t smem = (size_t)S * K * CB * sizeof(float2);
  static DeviceOnce attr;
  const int dev = current_device();
  if (dev < 0) return S2V_ECUDA;
  if (attr.needed(dev)) { S2V_CUDA_TRY(cudaFuncSetAttribute(irfft2_kernel<S, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr.mark(dev); }
  launch_pdl(irfft2_kernel<S, CB>, dim3(y->c / CB, y->n), K * CB, smem, st, mk(sp), mk(add && add->ptr ? add : nullptr), mk(y), fft_rev());
  return cudaGetLastError() == cudaSuccess ? S2V_OK : S2V_ECUDA;
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_fft_init(void) {
  float2 h[48];
  for (int j = 0; j < 48; ++j) {
    const double a = -2.0 * 3.14159265358979323846 * j / 48.0;
    h[j] = make_float2((float)cos(a), (float)sin(a));
  }
  if (cudaMemcpyToSymbol(c_tw48, h, sizeof(h)) != cudaSuccess) return S2V_ECUDA;
  return fft_mma_init();
}

static bool fft_shapes_ok(const s2v_view* x, const s2v_view* sp) {
  if (!view_ok(x) || !view_ok(sp)) return false;
  if (x->h != x->w || sp->h != x->h || sp->w != x->w / 2 + 1 || sp->c != 2 * x->c || sp->n != x->n) return false;
  if (x->n > 65535) return false;
  return true;
}

extern "C" int s2v_rfft2(const s2v_view* x, const s2v_view* spec, void* stream) {
  if (!fft_shapes_ok(x, spec)) return S2V_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (fft_use_mma(x->h, false)) return rfft2_mma(x, spec, st);
  if (x->h == 12 && x->c % 32 == 0) return launch_rfft2<12, 32>(x, spec, st);
  if (x->h == 12 && x->c % 8 == 0) return launch_rfft2<12, 8>(x, spec, st);
  if (x->h == 24 && x->c % 32 == 0 && fft24_cb32()) return launch_rfft2<24, 32>(x, spec, st);
  if (x->h == 24 && x->c % 16 == 0) return launch_rfft2<24, 16>(x, spec, st);
  if (x->h == 24 && x->c % 8 == 0) return launch_rfft2<24, 8>(x, spec, st);
  if (x->h == 48 && x->c % 16 == 0 && fft48_cb16()) return launch_rfft2<48, 16>(x, spec, st);
  if (x->h == 48 && x->c % 8 == 0) return launch_rfft2<48, 8>(x, spec, st);
  return S2V_EINVAL;
}

extern "C" int s2v_irfft2(const s2v_view* spec, const s2v_view* add, const s2v_view* y, void* stream) {
  if (!fft_shapes_ok(y, spec)) return S2V_EINVAL;
  if (add && add->ptr && (!view_ok(add) || add->h != y->h || add->w != y->w || add->c != y->c || add->n != y->n)) return S2V_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (fft_use_mma(y->h, true)) return irfft2_mma(spec, add && add->ptr ? add : nullptr, y, st);
  if (y->h == 12 && y->c % 32 == 0) return launch_irfft2<12, 32>(spec, add, y, st);
  if (y->h == 12 && y->c % 8 == 0) return launch_irfft2<12, 8>(spec, add, y, st);
  if (y->h == 24 && y->c % 32 == 0 && fft24_cb32()) return launch_irfft2<24, 32>(spec, add, y, st);
  if (y->h == 24 && y->c % 16 == 0) return launch_irfft2<24, 16>(spec, add, y, st);
  if (y->h == 24 && y->c % 8 == 0) return launch_irfft2<24, 8>(spec, add, y, st);
  if (y->h == 48 && y->c % 16 == 0 && fft48_cb16()) return launch_irfft2<48, 16>(spec, add, y, st);
  if (y->h == 48 && y->c % 8 == 0) return launch_irfft2<48, 8>(spec, add, y, st);
  return S2V_EINVAL;
}
