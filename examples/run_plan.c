/*
 * run_plan.c - a host WITHOUT Python running LNet.forward / DNet.forward (reference models/LNet.py:122-139,
 * models/DNet.py:20-28) through the C ABI of include/s2v.h: load a plan file written by s2v_b200.plan_export, bind it
 * to two cudaMalloc'ed buffers, fill the input slots from raw float32 files, replay, write the output slots.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/run_plan.c -o examples/run_plan \
 *       -Lspeech-to-video-mpp_b200 -ls2v -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../speech-to-video-mpp_b200'
 *   examples/run_plan lnet_b8.s2vplan out_prefix mel=mel.f32 face=face.f32      ->  out_prefix.out.f32
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "s2v.h"

#define CK(x) do { int rc__ = (x); if (rc__ != 0) { fprintf(stderr, "%s failed: %d (%s; %s)\n", #x, rc__, s2v_strerror(rc__), s2v_last_cuda_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s plan out_prefix [slot=file.f32 ...]\n", argv[0]); return 2; }
  CK(s2v_device_ok());
  CK(s2v_mel_init());
  CK(s2v_fft_init());
  s2v_plan* plan = NULL;
  CK(s2v_plan_load(argv[1], &plan));
  void *cdev = NULL, *wdev = NULL;
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  CU(cudaMalloc(&cdev, (size_t)s2v_plan_const_bytes(plan) + 256));
  CU(cudaMalloc(&wdev, (size_t)s2v_plan_workspace_bytes(plan) + 256));
  CK(s2v_plan_bind(plan, cdev, wdev, st));
  const int n_io = s2v_plan_num_io(plan);
  for (int i = 0; i < n_io; ++i) {                       /* inputs: slot=file arguments */
    const char* name; int64_t off, bytes; int is_out;
    CK(s2v_plan_io_info(plan, i, &name, &off, &bytes, &is_out));
    if (is_out) continue;
    const char* path = NULL;
    for (int a = 3; a < argc; ++a) {
      const size_t k = strlen(name);
      if (strncmp(argv[a], name, k) == 0 && argv[a][k] == '=') path = argv[a] + k + 1;
    }
    if (!path) { fprintf(stderr, "no file given for input slot '%s' (%lld bytes)\n", name, (long long)bytes); return 2; }
    void* h = malloc((size_t)bytes);
    FILE* f = fopen(path, "rb");
    if (!h || !f || fread(h, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "cannot read %lld bytes from %s\n", (long long)bytes, path); return 2; }
    fclose(f);
    CU(cudaMemcpyAsync((char*)wdev + off, h, (size_t)bytes, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    free(h);
  }
  CK(s2v_plan_run(plan, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < n_io; ++i) {                       /* outputs: <prefix>.<slot>.f32 */
    const char* name; int64_t off, bytes; int is_out;
    CK(s2v_plan_io_info(plan, i, &name, &off, &bytes, &is_out));
    if (!is_out) continue;
    char path[1024];
    snprintf(path, sizeof path, "%s.%s.f32", argv[2], name);
    void* h = malloc((size_t)bytes);
    if (!h) return 2;
    CU(cudaMemcpy(h, (char*)wdev + off, (size_t)bytes, cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(h, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "cannot write %s\n", path); return 2; }
    fclose(f);
    free(h);
    printf("%s: %lld bytes\n", path, (long long)bytes);
  }
  printf("%d launches replayed, constants %lld B, workspace %lld B\n", s2v_plan_num_ops(plan),
         (long long)s2v_plan_const_bytes(plan), (long long)s2v_plan_workspace_bytes(plan));
  s2v_plan_free(plan);
  cudaFree(cdev); cudaFree(wdev); cudaStreamDestroy(st);
  return 0;
}
