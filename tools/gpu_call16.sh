#!/bin/bash
# round-2 call 16: conv_head with the channel pad matched to Cout (resident weights for flow_out)
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_dnet.py tests/test_gpu_lnet.py -q -m gpu -x > gpurun_out/c16_tests.log 2>&1; tail -3 gpurun_out/c16_tests.log
python tools/plan_breakdown.py dnet > gpurun_out/x_dnet.txt 2>&1; head -1 gpurun_out/x_dnet.txt; grep "head" gpurun_out/x_dnet.txt | cut -c1-140
python tools/plan_breakdown.py lnet > gpurun_out/x_lnet.txt 2>&1; head -1 gpurun_out/x_lnet.txt; grep "head" gpurun_out/x_lnet.txt | cut -c1-140
