// Shared helpers for libs2v kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/s2v.h"

#include <atomic>

namespace s2v {
// The failing cudaError_t of the calling thread's last S2V_ECUDA return (api.cu); s2v_last_cuda_error() reports it.
void note_cuda_error(cudaError_t e);
}  // namespace s2v

#define S2V_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) { s2v::note_cuda_error(e__); return S2V_ECUDA; } \
  } while (0)
// for runtime calls that return their own status (cudaFuncSetAttribute, cudaLaunchKernelEx, ...)
#define S2V_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t e__ = (expr);                                \
    if (e__ != cudaSuccess) { s2v::note_cuda_error(e__); (void)cudaGetLastError(); return S2V_ECUDA; } \
  } while (0)

namespace s2v {

// ---- per-device one-time state ----------------------------------------------------------------------------
// Function attributes (MaxDynamicSharedMemorySize) and the SM count belong to a DEVICE, not to the process: a library
// used on several GPUs from one process must set / query them once per device.  Lock-free, idempotent (two threads racing
// on the first call both set the same attribute value).
constexpr int kMaxDevices = 64;
static inline int current_device() {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return -1;
  return d;
}
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool needed(int dev) const { return !(done.load(std::memory_order_acquire) >> dev & 1ull); }
  void mark(int dev) { done.fetch_or(1ull << dev, std::memory_order_release); }
};
static inline int sm_count(int dev) {
  static std::atomic<int> n_sm[kMaxDevices];
  int v = n_sm[dev].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return -1;
    n_sm[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// A forward is ~570 small dependent kernels replayed from a CUDA graph.  Every kernel of the library is launched
// with programmaticStreamSerialization: it calls pdl_trigger() first (the NEXT kernel's CTAs may be scheduled as soon
// as all CTAs of this one are resident, so its launch latency and prologue overlap this kernel's execution / tail)
// and pdl_wait() before its first access to global memory an earlier kernel may have written or may still read.
// pdl_wait() returns only when the preceding kernel has COMPLETED and flushed its memory, so ordering is transitive
// as long as every kernel executes it.  S2V_PDL=0 disables the launch attribute (the instructions become no-ops).
// Measured on B200 (LNet B=128, graph replay): memory-bound kernels triggering at their start: 15.66 -> 15.55 ms/step;
// conv_tc triggering (at its start OR after its last load): 16.1 / 15.95 ms/step, i.e. dependents scheduled under a
// running persistent GEMM cost more than they hide - so conv_tc only WAITS (its prologue and resident-weight fetch
// overlap the previous kernel's tail) and never triggers early (S2V_PDL_TRIG_CONV = 0).
#ifndef S2V_PDL_TRIG_OTHER
#define S2V_PDL_TRIG_OTHER 1
#endif
#ifndef S2V_PDL_TRIG_CONV
#define S2V_PDL_TRIG_CONV 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if S2V_PDL_TRIG_OTHER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger_conv() {
#if S2V_PDL_TRIG_CONV == 1
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline bool pdl_enabled() {
  static int v = -1;                       // resolved once; immutable afterwards
  if (v < 0) { const char* e = getenv("S2V_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct View {           // device-side copy of s2v_view with typed pointer
  __half* p;
  int n, h, w, c;
  long long sn, sh, sw;
};

static inline View mk(const s2v_view* v) {
  View r;
  r.p = v ? (__half*)v->ptr : nullptr;
  if (v) { r.n = v->n; r.h = v->h; r.w = v->w; r.c = v->c; r.sn = v->sn; r.sh = v->sh; r.sw = v->sw; }
  else { r.n = r.h = r.w = r.c = 0; r.sn = r.sh = r.sw = 0; }
  return r;
}
static inline View mk(const s2v_view& v) { return mk(&v); }

static inline bool view_ok(const s2v_view* v) {
  if (!v || !v->ptr) return false;
  if (v->n <= 0 || v->h <= 0 || v->w <= 0 || v->c <= 0) return false;
  if ((v->c & 7) || (v->sw & 7) || (v->sh & 7) || (v->sn & 7)) return false;
  if (((uintptr_t)v->ptr) & 15) return false;
  return true;
}

__device__ __forceinline__ float act_apply(float v, int act, float p) {
  switch (act) {
    case S2V_ACT_RELU: return fmaxf(v, 0.f);
    case S2V_ACT_LRELU: return v > 0.f ? v : v * p;
    case S2V_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case S2V_ACT_TANH: return tanhf(v);
    case S2V_ACT_GELU: {
      float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      return 0.5f * v * (1.f + tanhf(u));
    }
    default: return v;
  }
}

struct __align__(16) H8 { __half2 v[4]; };

__device__ __forceinline__ void h8_to_f(const H8& h, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ H8 f_to_h8(const float* f) {
  H8 h;
#pragma unroll
  for (int i = 0; i < 4; ++i) h.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return h;
}
// one 128-bit load (a struct copy may be split into four 32-bit loads when the source is shared memory)
__device__ __forceinline__ H8 ld_h8(const __half* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  H8 h;
  *reinterpret_cast<uint32_t*>(&h.v[0]) = u.x;
  *reinterpret_cast<uint32_t*>(&h.v[1]) = u.y;
  *reinterpret_cast<uint32_t*>(&h.v[2]) = u.z;
  *reinterpret_cast<uint32_t*>(&h.v[3]) = u.w;
  return h;
}
// one 128-bit store: a struct copy of H8 is emitted as four 32-bit stores when the destination is shared memory (the conv
// epilogue's staging tile: 4-way bank-conflicted STS.32 instead of one conflict-free STS.128)
__device__ __forceinline__ void st_h8(__half* p, const H8& v) {
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&v.v[0]);
  u.y = *reinterpret_cast<const uint32_t*>(&v.v[1]);
  u.z = *reinterpret_cast<const uint32_t*>(&v.v[2]);
  u.w = *reinterpret_cast<const uint32_t*>(&v.v[3]);
  *reinterpret_cast<uint4*>(p) = u;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace s2v
