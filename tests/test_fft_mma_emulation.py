"""CPU check of the tensor-core DFT variant of the FourierUnit transforms (csrc/fft2d_mma.cu, reference: models/ffc.py:99-121):
the A-fragment tables the host builds and the kernels' index arithmetic (ldmatrix.trans addressing, mma.m16n8k16 register layouts,
in-place tile passes), emulated lane by lane in numpy (tools/emu_fft_mma.py), must give numpy's rfft2 / irfft2 (ortho) for the three
LNet sizes within the same 2e-3-of-max bound the GPU test uses - in both buffer layouts: the padded in-place tile of the cp.async
kernels and the dense TMA destination buffers of the default (48 x 48) kernels, bit-identical to each other.  Needs nvcc (host code
only; no GPU)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")
def test_fragment_tables_and_lane_emulation(tmp_path):
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    exe, frag = str(tmp_path / "dump"), str(tmp_path / "frag.bin")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr", "-o", exe,
                        os.path.join(ROOT, "tools", "dump_fft_frags.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert subprocess.run([exe, frag], timeout=60).returncode == 0
    assert os.path.getsize(frag) == 90 * 32 * 16
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fft_mma.py"), frag], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
