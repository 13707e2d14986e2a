#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -3 gpurun_out/r2u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2u_bench.json; echo
