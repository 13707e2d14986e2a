#!/bin/bash
mkdir -p gpurun_out
S2V_TC_DEBUG=1 S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py > gpurun_out/r2r_epi_prof.txt 2>&1
cat gpurun_out/r2r_epi_prof.txt
