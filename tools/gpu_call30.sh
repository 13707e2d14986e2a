#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lnet.py tests/test_gpu_dnet.py tests/test_gpu_plan.py -q -m gpu -x --timeout 600 2>&1 | tail -2
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/r2x_breakdown_$w.txt 2>&1; head -1 gpurun_out/r2x_breakdown_$w.txt; grep "finalize" gpurun_out/r2x_breakdown_$w.txt; done
