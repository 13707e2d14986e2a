#!/bin/bash
mkdir -p gpurun_out
export MB_REPS=1 MB_WARM=1
python tools/mb_layers.py res1 > gpurun_out/r2g_mb_res1.txt 2>&1 && python tools/mb_fft.py > gpurun_out/r2g_mb_fft.txt 2>&1 || { echo plain run failed; tail gpurun_out/r2g_mb_res1.txt gpurun_out/r2g_mb_fft.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 8 -o gpurun_out/r2g_ncu_res1 -f python tools/mb_layers.py res1 > gpurun_out/r2g_ncu_res1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft2 -c 12 -o gpurun_out/r2g_ncu_fft -f python tools/mb_fft.py > gpurun_out/r2g_ncu_fft.log 2>&1
ls -la gpurun_out/*.ncu-rep
