#!/bin/bash
# first GPU call of round 2: tests, bench (overlap on / off), reference arm sanity, smoke
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2a_bench.err
S2V_PIPE_OVERLAP=0 timeout 600 python bench.py --no-extras --no-cpu-baseline --no-classes > gpurun_out/r2a_bench_noov.json 2> gpurun_out/r2a_bench_noov.err; echo "bench noov rc=$?"
head -c 600 gpurun_out/r2a_bench.json; echo; head -c 600 gpurun_out/r2a_bench_noov.json
