#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --no-classes > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
head -c 400 gpurun_out/r2b_bench.json; echo
S2V_PIPE_OVERLAP=0 timeout 600 python bench.py --no-extras --no-cpu-baseline --no-classes > gpurun_out/r2b_bench_noov.json 2> gpurun_out/r2b_bench_noov.err; echo "bench noov rc=$?"
head -c 400 gpurun_out/r2b_bench_noov.json; echo
