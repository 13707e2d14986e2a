#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -3 gpurun_out/r2u_pytest.log
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; head -c 300 gpurun_out/r2u_bench.json; echo
S2V_LIB=$PWD/speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py 2>&1 | grep -v "^conv_tc:" > gpurun_out/r2u_epi_prof.txt
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/r2u_breakdown_$w.txt 2>&1; head -1 gpurun_out/r2u_breakdown_$w.txt; done
