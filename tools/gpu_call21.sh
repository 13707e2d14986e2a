#!/bin/bash
# STS.128 staging stores + tree totals + two A rings: correctness (conv / LNet / DNet tests) and per-layer / per-plan times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_lnet.py tests/test_gpu_dnet.py -q -m gpu -x --timeout 600 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py > gpurun_out/r2s_epi_prof.txt 2>&1; grep -v "^conv_tc:" gpurun_out/r2s_epi_prof.txt
MB_GRAPH=1 python tools/mb_layers.py res > gpurun_out/r2s_layers.txt 2>&1; cat gpurun_out/r2s_layers.txt
python tools/mb_dnet_layers.py > gpurun_out/r2s_dnet_layers.txt 2>&1; cat gpurun_out/r2s_dnet_layers.txt
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/r2s_breakdown_$w.txt 2>&1; head -1 gpurun_out/r2s_breakdown_$w.txt; done
