#!/bin/bash
mkdir -p gpurun_out
for ts in 1 0; do
echo "== S2V_TMA_STORE=$ts"
S2V_TMA_STORE=$ts S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py 2>&1 | grep -v "^conv_tc:"
S2V_TMA_STORE=$ts MB_GRAPH=1 python tools/mb_layers.py res 2>&1
S2V_TMA_STORE=$ts python tools/mb_dnet_layers.py 2>&1
done > gpurun_out/r2t_tma_store.txt 2>&1
cat gpurun_out/r2t_tma_store.txt
