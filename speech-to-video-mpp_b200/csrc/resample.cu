// load_wav (reference: futils/audio.py:9-10 -> librosa.core.load(path, sr)): PCM decode to float32 mono and the band-limited
// 'kaiser_best' resampler librosa 0.9.2 delegates to resampy (resampy/interpn.py resample_f, J. O. Smith's algorithm):
//   y[t] = sum_i (win[off_l + i*step] + eta_l * delta[off_l + i*step]) * x[n - i]          (left wing, i < min(n+1, (nwin-off_l)/step))
//        + sum_k (win[off_r + k*step] + eta_r * delta[off_r + k*step]) * x[n + k + 1]      (right wing)
// with tau = t * sr_orig/sr_new, n = int(tau), frac = scale*(tau - n), off = int(frac*num_table), eta its fraction.
// One thread per output sample, float64 arithmetic in resampy's operation order (separate multiply and add, so the result is
// what the numpy restatement computes), the 32 769-tap half window + its differences stay L1/L2 resident.  The filter table is
// built on the host (futils/audio.py of this package) and passed in, like the mel basis.
#include "common.cuh"

namespace s2v {

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long n_orig, const double* __restrict__ win,
                                                      const double* __restrict__ delta, int nwin, int num_table, double ratio,
                                                      float* __restrict__ y, long long n_out) {
  pdl_trigger();
  pdl_wait();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const double scale = ratio < 1.0 ? ratio : 1.0;
  const int step = (int)(scale * (double)num_table);
  const double tau = __dmul_rn((double)t, 1.0 / ratio);
  const long long n = (long long)tau;
  double acc = 0.0;
  // left wing
  double frac = __dmul_rn(scale, tau - (double)n);
  double index_frac = __dmul_rn(frac, (double)num_table);
  int offset = (int)index_frac;
  double eta = index_frac - (double)offset;
  long long cnt = (nwin - offset) / step;
  if (cnt > n + 1) cnt = n + 1;
  for (long long i = 0; i < cnt; ++i) {
    const int idx = offset + (int)i * step;
    const double w = __dadd_rn(win[idx], __dmul_rn(eta, delta[idx]));
    acc = __dadd_rn(acc, __dmul_rn(w, (double)x[n - i]));
  }
  // right wing
  frac = scale - frac;
  index_frac = __dmul_rn(frac, (double)num_table);
  offset = (int)index_frac;
  eta = index_frac - (double)offset;
  cnt = (nwin - offset) / step;
  if (cnt > n_orig - n - 1) cnt = n_orig - n - 1;
  for (long long k = 0; k < cnt; ++k) {
    const int idx = offset + (int)k * step;
    const double w = __dadd_rn(win[idx], __dmul_rn(eta, delta[idx]));
    acc = __dadd_rn(acc, __dmul_rn(w, (double)x[n + k + 1]));
  }
  y[t] = (float)acc;
}

// soundfile's float32 decode + librosa.to_mono: kind 0 = int16 / 2^15, 1 = int32 / 2^31, 2 = uint8 (x - 128) / 2^7, 3 = float32
__global__ void __launch_bounds__(256) pcm_to_mono_kernel(const void* __restrict__ pcm, int kind, int channels, long long n,
                                                         float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int c = 0; c < channels; ++c) {
    const long long j = i * channels + c;
    float v;
    if (kind == 0) v = (float)reinterpret_cast<const short*>(pcm)[j] / 32768.f;
    else if (kind == 1) v = (float)((double)reinterpret_cast<const int*>(pcm)[j] / 2147483648.0);
    else if (kind == 2) v = ((float)reinterpret_cast<const unsigned char*>(pcm)[j] - 128.f) / 128.f;
    else v = reinterpret_cast<const float*>(pcm)[j];
    acc = channels == 1 ? v : __fadd_rn(acc, v);            // np.mean over the channel axis: sequential float32 sum / count
  }
  out[i] = channels == 1 ? acc : acc / (float)channels;
}

}  // namespace s2v

using namespace s2v;

extern "C" int64_t s2v_resample_out_len(int64_t n_in, int sr_orig, int sr_new) {
  if (n_in < 0 || sr_orig <= 0 || sr_new <= 0) return -1;
  return (int64_t)((double)n_in * ((double)sr_new / (double)sr_orig));      // int(n * ratio), resampy/core.py
}

extern "C" int s2v_resample_f32(const float* x, int64_t n_in, int sr_orig, int sr_new, const double* win, const double* delta,
                                int nwin, int num_table, float* y, int64_t n_out, void* stream) {
  if (!x || !y || !win || !delta || n_in <= 0 || n_out < 0 || sr_orig <= 0 || sr_new <= 0 || nwin <= 0 || num_table <= 0) return S2V_EINVAL;
  if (n_out == 0) return S2V_OK;
  const double ratio = (double)sr_new / (double)sr_orig;
  if (n_out > s2v_resample_out_len(n_in, sr_orig, sr_new)) return S2V_EINVAL;
  if ((int)((ratio < 1.0 ? ratio : 1.0) * (double)num_table) < 1) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(resample_kernel, ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream, x, (long long)n_in, win, delta, nwin, num_table,
                          ratio, y, (long long)n_out));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}

extern "C" int s2v_pcm_to_mono_f32(const void* pcm, int kind, int channels, int64_t n_frames, float* out, void* stream) {
  if (n_frames == 0) return S2V_OK;
  if (!pcm || !out || kind < 0 || kind > 3 || channels <= 0 || n_frames < 0) return S2V_EINVAL;
  S2V_CUDA_TRY(launch_pdl(pcm_to_mono_kernel, ceil_div(n_frames, 256), 256, 0, (cudaStream_t)stream, pcm, kind, channels, (long long)n_frames, out));
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
