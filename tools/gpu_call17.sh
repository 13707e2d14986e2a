#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -3 gpurun_out/r2q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; head -c 250 gpurun_out/r2q_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err; echo "ref rc=$?"
python tools/profile_step.py --seconds 12 > gpurun_out/r2q_profile_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_ncu_launches_clip12s.csv python tools/profile_step.py --seconds 12 > gpurun_out/r2q_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2q_ncu_launches_clip12s.csv)"
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/r2q_breakdown_$w.txt 2>&1; head -1 gpurun_out/r2q_breakdown_$w.txt; done
python tools/plan_breakdown.py dnet 192 > gpurun_out/r2q_breakdown_dnet_b192.txt 2>&1; head -1 gpurun_out/r2q_breakdown_dnet_b192.txt
