#!/bin/bash
mkdir -p gpurun_out
S2V_EXP=16 S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py 2>&1 | grep -v "^conv_tc:" > gpurun_out/r2t_store_latency.txt
cat gpurun_out/r2t_store_latency.txt
