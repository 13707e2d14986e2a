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

// last CUDA error text of this library's runtime instance (diagnostics for the Python wrapper)
extern "C" const char* s2v_last_cuda_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }
