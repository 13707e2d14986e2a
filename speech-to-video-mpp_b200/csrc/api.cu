// Library-level entry points: error strings, version, device check.
#include "common.cuh"

extern "C" const char* s2v_strerror(int code) {
  switch (code) {
    case S2V_OK: return "ok";
    case S2V_EINVAL: return "invalid argument or unsupported shape";
    case S2V_ECUDA: return "CUDA runtime/driver call failed";
    case S2V_EUNSUPPORTED: return "device or driver does not support the sm_100a kernels";
    default: return "unknown s2v error";
  }
}

extern "C" int s2v_version(void) { return 100; }

extern "C" int s2v_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return S2V_ECUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return S2V_ECUDA;
  return major == 10 ? S2V_OK : S2V_EUNSUPPORTED;
}

// The cudaError_t behind the calling thread's most recent S2V_ECUDA return.  (S2V_CHECK_LAUNCH consumes the runtime's
// sticky-free error with cudaGetLastError, so it is kept here; cudaPeekAtLastError alone would read "no error".)
static thread_local cudaError_t t_last_error = cudaSuccess;
void s2v::note_cuda_error(cudaError_t e) { t_last_error = e; }

extern "C" const char* s2v_last_cuda_error(void) {
  const cudaError_t e = t_last_error != cudaSuccess ? t_last_error : cudaPeekAtLastError();
  return cudaGetErrorString(e);
}
