#!/bin/bash
# evidence of the final build of this session (N = 1): tests, smoke, bench + reference arm, ncu launch list, ncu --set full of the
# changed conv_tc kernel (DNet down0 and LNet's 48 x 48 merged FFC GEMM), per-layer tables
P=${1:-r2b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log
tail -3 gpurun_out/${P}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; head -c 250 gpurun_out/${P}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo "ref rc=$?"; head -c 200 gpurun_out/${P}_bench_ref.json; echo
python tools/profile_step.py --seconds 12 > gpurun_out/${P}_profile_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_ncu_launches_clip12s.csv python tools/profile_step.py --seconds 12 > gpurun_out/${P}_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${P}_ncu_launches_clip12s.csv)"
python tools/mb_dnet_layers.py > gpurun_out/${P}_mb_dnet_layers.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/${P}_ncu_full_dnet_down0 -f python tools/mb_dnet_layers.py > gpurun_out/${P}_ncu_full_a.log 2>&1
echo "full down0 rc=$?"
MB_GRAPH=0 MB_REPS=2 MB_WARM=1 python tools/mb_layers.py res0.all+narrow+stats > /dev/null 2>&1 && \
MB_GRAPH=0 MB_REPS=2 MB_WARM=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/${P}_ncu_full_res0_all -f python tools/mb_layers.py res0.all+narrow+stats > gpurun_out/${P}_ncu_full_b.log 2>&1
echo "full res0.all rc=$?"
python tools/mb_fft.py > /dev/null 2>&1
for w in lnet dnet; do python tools/plan_breakdown.py $w > gpurun_out/${P}_breakdown_$w.txt 2>&1; head -1 gpurun_out/${P}_breakdown_$w.txt; done
python tools/plan_breakdown.py dnet 192 > gpurun_out/${P}_breakdown_dnet_b192.txt 2>&1; head -1 gpurun_out/${P}_breakdown_dnet_b192.txt
python tools/plan_breakdown.py lnet 256 > gpurun_out/${P}_breakdown_lnet_b256.txt 2>&1; head -1 gpurun_out/${P}_breakdown_lnet_b256.txt
ls -la gpurun_out/${P}_*
