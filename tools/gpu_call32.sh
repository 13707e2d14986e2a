#!/bin/bash
# round 2, third session: evidence of the final build (N = 1): all -m gpu tests, smoke, bench, ncu launch list of one step
P=r2f
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log
tail -4 gpurun_out/${P}_pytest.log
cp gpurun_out/parity_report.txt gpurun_out/${P}_parity_report.txt 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${P}_bench.err
head -c 700 gpurun_out/${P}_bench.json; echo
