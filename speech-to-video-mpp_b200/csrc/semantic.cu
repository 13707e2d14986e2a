// 3DMM coefficient windows for DNet's driving_source (reference: futils/inference_utils.py:73-91
// obtain_seq_index + transform_semantic, called per frame at preprocessing/facing.py:184).
// One launch builds the [73, 26] window of every frame of a batch: a clamped 26-row gather of the
// per-frame coefficient table [T, D] (D >= 262: id 80 | exp 64 | tex 80 | angle 3 | gamma 27 | trans 3 | crop 5),
// the column selection exp 80:144 | angle 224:227 | trans 254:257 | crop 259:262, the optional
// crop[:, 0] *= ratio, the float32 conversion of torch.Tensor(...) and the permute(1, 0).
// Pure index + one-multiply work: bit-exact against the reference; memory-bound (7.6 KB out per frame).
#include "common.cuh"

namespace s2v {

constexpr int kSemRows = 73, kSemWin = 26;

__device__ __forceinline__ int sem_src_col(int r) {
  return r < 64 ? 80 + r : r < 67 ? 224 + (r - 64) : r < 70 ? 254 + (r - 67) : 259 + (r - 70);
}

// grid (ceil(73*26/256), n_frames): thread -> (r, j) of frame first + blockIdx.y; output index r*26 + j is contiguous
template <typename T>
__global__ void __launch_bounds__(256) semantic_windows_kernel(const T* __restrict__ sem, int n_rows, int D,
                                                               const int* __restrict__ frame_idx, int first,
                                                               double ratio, int use_ratio, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kSemRows * kSemWin) return;
  const int r = e / kSemWin, j = e - r * kSemWin;
  const int f = frame_idx ? frame_idx[blockIdx.y] : first + (int)blockIdx.y;
  const int row = min(max(f - 13 + j, 0), n_rows - 1);             // obtain_seq_index: range(i-13, i+13) clamped
  T v = sem[(size_t)row * D + sem_src_col(r)];
  if (use_ratio && r == 70) v = v * (T)ratio;                       // crop[:, -3] *= crop_norm_ratio, in the table's dtype
  out[(size_t)blockIdx.y * (kSemRows * kSemWin) + e] = (float)v;
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_semantic_windows(const void* semantic, int is_f64, int n_rows, int D, const int32_t* frame_idx,
                                    int first, int count, double ratio, int use_ratio, float* out, void* stream) {
  if (count == 0) return S2V_OK;
  if (!semantic || !out || n_rows <= 0 || D < 262 || count < 0 || count > 65535) return S2V_EINVAL;
  const dim3 grid(ceil_div(kSemRows * kSemWin, 256), count);
  if (is_f64)
    launch_pdl(semantic_windows_kernel<double>, grid, 256, 0, (cudaStream_t)stream, (const double*)semantic, n_rows, D,
               (const int*)frame_idx, first, ratio, use_ratio, out);
  else
    launch_pdl(semantic_windows_kernel<float>, grid, 256, 0, (cudaStream_t)stream, (const float*)semantic, n_rows, D,
               (const int*)frame_idx, first, ratio, use_ratio, out);
  S2V_CHECK_LAUNCH();
  return S2V_OK;
}
