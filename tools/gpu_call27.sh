#!/bin/bash
mkdir -p gpurun_out
for v in "" _scalar; do
echo "== libs2v$v.so"
S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v$v.so python tools/mb_fft.py 2>&1
S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v$v.so MB_B=256 python tools/mb_fft.py 2>&1
done | tee gpurun_out/r2v_fft.txt
