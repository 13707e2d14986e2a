// Mel front end (reference: futils/audio.py:45-51 with futils/hparams.py:21-61 and
// librosa 0.9.2 stft/filters.mel semantics) as one fused kernel:
//   pre-emphasis -> centre pad -> frame(800, hop 200) -> periodic Hann -> 800-point DFT
//   -> |.| -> 80-band mel projection -> 20*log10(max(1e-5,.)) - 20 -> clip(8*(S+100)/100-4, +-4)
//
// n_fft = 800 = 32 * 25 is not a power of two, so the transform is a mixed-radix
// Cooley-Tukey: one WARP per STFT column; lane n1 holds the 25 samples x[25*n1 + n2];
// the radix-32 stage is 25 simultaneous 32-point FFTs done with warp-shuffle (xor)
// butterflies, then the W_800^{n2*k1} twiddles, then a register-resident 25-point DFT
// (constant-memory twiddles) for the 13 values of k2 that land in bins 0..400.
// Memory-bound by design: 200 new samples in + 80 floats out per column (1120 B).
#include "common.cuh"

namespace s2v {

constexpr int kNfft = 800, kHop = 200, kBins = 401, kMels = 80;
constexpr int kWarpsPerBlock = 8;

__constant__ float2 c_w25[25];   // exp(-2*pi*i*j/25)
// W_800^{j * bitrev5(lane)} for j = 1..24, indexed [j - 1][lane]: the inter-stage twiddles of the 32 x 25 decomposition, computed
// once in double on the host (s2v_mel_init) instead of 24 sincospif per lane and column; 6 KB, L1 resident, lane-contiguous reads
__device__ float2 g_tw800[24 * 32];

struct W25Init {
  float2 v[25];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float preemph_at(const float* __restrict__ wav, long long n, long long orig, int reflect) {
  if (orig < 0 || orig >= n) {
    if (!reflect) return 0.f;
    orig = orig < 0 ? -orig : 2 * (n - 1) - orig;      // np.pad(mode='reflect')
  }
  float v = wav[orig];
  if (orig > 0) v = __fmaf_rn(-0.97f, wav[orig - 1], v);   // lfilter([1,-0.97],[1])
  return v;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) mel_kernel(const float* __restrict__ wav, long long n, int T,
                                                                 const float* __restrict__ basis,
                                                                 const int* __restrict__ band_range,
                                                                 float* __restrict__ out, int reflect) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_hann[kNfft];
  __shared__ float s_x[kWarpsPerBlock][kNfft];
  __shared__ float s_mag[kWarpsPerBlock][kBins + 15];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) s_hann[i] = 0.5f - 0.5f * cospif(2.f * (float)i / (float)kNfft);
  __syncthreads();
  const int t = blockIdx.x * kWarpsPerBlock + warp;
  if (t >= T) return;            // no block-level barrier below

  // 1. frame t of the centre-padded, pre-emphasised signal, times the window
  const long long base = (long long)t * kHop - kNfft / 2;
  for (int i = lane; i < kNfft; i += 32) s_x[warp][i] = preemph_at(wav, n, base + i, reflect) * s_hann[i];
  __syncwarp();

  // 2. radix-32 stage across lanes (DIF, xor-shuffle butterflies); lane = n1, 25 columns n2
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) v[j] = make_float2(s_x[warp][25 * lane + j], 0.f);
  float2 tw[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
    float sn, cs;
    sincospif(-2.f * (float)((lane & (half - 1)) * (16 / half)) / 32.f, &sn, &cs);
    tw[s] = make_float2(cs, sn);
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
    const bool bot = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < 25; ++j) {
      float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[j].x, half), __shfl_xor_sync(0xffffffffu, v[j].y, half));
      float2 r = bot ? make_float2(o.x - v[j].x, o.y - v[j].y) : make_float2(v[j].x + o.x, v[j].y + o.y);
      v[j] = bot ? cmul(r, tw[s]) : r;
    }
  }
  const int k1 = (int)(__brev((unsigned)lane) >> 27);    // lane holds output k1 = bitrev5(lane)

  // 3. twiddles W_800^{n2*k1}
#pragma unroll
  for (int j = 1; j < 25; ++j) v[j] = cmul(v[j], __ldg(&g_tw800[(j - 1) * 32 + lane]));

  // 4. 25-point DFT over n2 for k2 = 0..12  ->  bin k = k1 + 32*k2 (<= 400 kept)
#pragma unroll
  for (int k2 = 0; k2 < 13; ++k2) {
    float2 acc = v[0];
#pragma unroll
    for (int j = 1; j < 25; ++j) {
      const float2 w = c_w25[(j * k2) % 25];
      acc.x = fmaf(v[j].x, w.x, fmaf(-v[j].y, w.y, acc.x));
      acc.y = fmaf(v[j].x, w.y, fmaf(v[j].y, w.x, acc.y));
    }
    const int k = k1 + 32 * k2;
    if (k < kBins) s_mag[warp][k] = sqrtf(acc.x * acc.x + acc.y * acc.y);
  }
  __syncwarp();

  // 5. mel projection + amp_to_db - ref_level_db + symmetric normalise/clip
  for (int m = lane; m < kMels; m += 32) {
    int lo = 0, hi = kBins;
    if (band_range) { lo = band_range[2 * m]; hi = band_range[2 * m + 1]; }
    const float* row = basis + m * kBins;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(row[k], s_mag[warp][k], acc);
    float s = 20.f * log10f(fmaxf(1e-5f, acc)) - 20.f;
    float r = 8.f * ((s + 100.f) / 100.f) - 4.f;
    out[(size_t)m * T + t] = fminf(fmaxf(r, -4.f), 4.f);
  }
}

// start column of window i (inference.py:209-216): int(i * (80./fps)) in float64, tail rule
__host__ __device__ inline long long window_start(long long i, double mult, long long T, bool* is_tail) {
  long long s = (long long)((double)i * mult);
  *is_tail = (s + 16 > T);
  return *is_tail ? T - 16 : s;
}

__global__ void mel_windows_kernel(const float* __restrict__ mel, long long T, double mult, long long first,
                                   long long count, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * kMels * 16) return;
  const int col = (int)(idx & 15);
  const int m = (int)((idx >> 4) % kMels);
  const long long wi = idx / (kMels * 16);
  bool tail;
  const long long s = window_start(first + wi, mult, T, &tail);
  out[idx] = mel[(size_t)m * T + s + col];
}

}  // namespace s2v

using namespace s2v;

static cudaError_t upload_w25() {
  // exp(-2*pi*i*j/25) in double, rounded once; idempotent (same bytes every call)
  float2 h[25];
  for (int j = 0; j < 25; ++j) {
    double a = -2.0 * 3.14159265358979323846 * j / 25.0;
    h[j] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t e = cudaMemcpyToSymbol(c_w25, h, sizeof(h));
  if (e != cudaSuccess) return e;
  static float2 tw[24 * 32];
  for (int j = 1; j < 25; ++j)
    for (int lane = 0; lane < 32; ++lane) {
      int k1 = 0;
      for (int b = 0; b < 5; ++b) k1 |= ((lane >> b) & 1) << (4 - b);          // lane holds output k1 = bitrev5(lane)
      const double a = -2.0 * 3.14159265358979323846 * (double)((j * k1) % 800) / 800.0;
      tw[(j - 1) * 32 + lane] = make_float2((float)cos(a), (float)sin(a));
    }
  return cudaMemcpyToSymbol(g_tw800, tw, sizeof(tw));
}

extern "C" int s2v_mel_num_frames(int64_t n_samples) { return n_samples < 0 ? S2V_EINVAL : (int)(1 + n_samples / kHop); }

extern "C" int s2v_mel_init(void) { return upload_w25() == cudaSuccess ? S2V_OK : S2V_ECUDA; }

extern "C" int s2v_melspectrogram_f32(const float* wav, int64_t n_samples, const float* basis, const int32_t* band_range,
                                      float* mel_out, int pad_reflect, void* stream) {
  if (n_samples < 0 || !basis || !mel_out || (n_samples > 0 && !wav)) return S2V_EINVAL;
  if (pad_reflect && n_samples <= kNfft / 2) return S2V_EINVAL;
  const int T = 1 + (int)(n_samples / kHop);
  launch_pdl(mel_kernel, ceil_div(T, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream, 
      wav, n_samples, T, basis, band_range, mel_out, pad_reflect);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int64_t s2v_mel_window_count(int64_t n_cols, double fps) {
  if (n_cols < 16 || !(fps > 0)) return S2V_EINVAL;
  const double mult = 80.0 / fps;
  // first i with int(i*mult) + 16 > T; walk from a safe lower bound
  long long i = (long long)((double)(n_cols - 16) / mult) - 2;
  if (i < 0) i = 0;
  for (;; ++i) {
    bool tail;
    window_start(i, mult, n_cols, &tail);
    if (tail) return i + 1;
  }
}

extern "C" int s2v_mel_window_starts_host(int64_t n_cols, double fps, int32_t* starts_host, int64_t count) {
  if (n_cols < 16 || !(fps > 0) || !starts_host) return S2V_EINVAL;
  const double mult = 80.0 / fps;
  for (int64_t i = 0; i < count; ++i) {
    bool tail;
    starts_host[i] = (int32_t)window_start(i, mult, n_cols, &tail);
  }
  return S2V_OK;
}

extern "C" int s2v_mel_windows_f32(const float* mel, int64_t n_cols, double fps, int64_t first, int64_t count,
                                   float* out, void* stream) {
  if (count == 0) return S2V_OK;
  if (!mel || !out || n_cols < 16 || !(fps > 0) || first < 0 || count < 0) return S2V_EINVAL;
  if (first + count > s2v_mel_window_count(n_cols, fps)) return S2V_EINVAL;
  const long long total = count * kMels * 16;
  launch_pdl(mel_windows_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, mel, n_cols, 80.0 / fps, first, count, out);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
