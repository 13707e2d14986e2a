// Development / CPU test helper: writes the host-built A-fragment tables of csrc/fft2d_mma.cu (fftmma::build_tables) to a file, so
// that tools/emu_fft_mma.py can emulate the kernels lane by lane without a GPU (tests/test_fft_mma_emulation.py).
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -o dump tools/dump_fft_frags.cu && ./dump frag.bin
#include "../speech-to-video-mpp_b200/csrc/fft2d_mma.cu"
#include <stdio.h>
namespace s2v { void note_cuda_error(cudaError_t) {} }
int main(int argc, char** argv) {
  static uint4 host[s2v::fftmma::kFragsTotal * 32];
  s2v::fftmma::build_tables<48>(host);
  s2v::fftmma::build_tables<24>(host);
  s2v::fftmma::build_tables<12>(host);
  FILE* f = fopen(argc > 1 ? argv[1] : "frag.bin", "wb");
  if (!f) return 1;
  const size_t n = fwrite(host, sizeof(host), 1, f);
  fclose(f);
  return n == 1 ? 0 : 1;
}
