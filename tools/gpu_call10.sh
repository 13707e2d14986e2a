#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2k_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/r2k_bench_ref.json; echo
python tools/profile_step.py --seconds 12 > gpurun_out/r2k_profile_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_ncu_launches_clip12s.csv python tools/profile_step.py --seconds 12 > gpurun_out/r2k_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2k_ncu_launches_clip12s.csv)"
MB_X=1 python tools/mb_dnet_layers.py > gpurun_out/r2k_mb_dnet.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/r2k_ncu_full_dnet_down0 -f python tools/mb_dnet_layers.py > gpurun_out/r2k_ncu_full.log 2>&1
echo "full rc=$?"; cat gpurun_out/r2k_mb_dnet.txt
