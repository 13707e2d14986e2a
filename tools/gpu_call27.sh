#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fft" --timeout 600 2>&1 | tail -2
python tools/mb_fft.py 2>&1 | tee gpurun_out/r2z_fft.txt
MB_B=256 python tools/mb_fft.py 2>&1 | tee -a gpurun_out/r2z_fft.txt
