#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_lnet.py tests/test_gpu_dnet.py tests/test_gpu_enet.py -q -m gpu -x --timeout 600 2>&1 | tail -2
MB_GRAPH=1 python tools/mb_layers.py res 2>&1 | tee gpurun_out/r2w_layers.txt
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/r2w_breakdown_$w.txt 2>&1; head -1 gpurun_out/r2w_breakdown_$w.txt; done
