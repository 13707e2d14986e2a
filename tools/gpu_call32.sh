#!/bin/bash
# round 2, third session: evidence of the final build (N = 1): all -m gpu tests, smoke, bench, ncu --set full of the shipped
# flow_warp4_kernel (bench configuration: B = 64, N(0, 3^2) flow), of one conv_head launch and one affine_act launch of the step
P=r2c
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log
tail -4 gpurun_out/${P}_pytest.log
cp gpurun_out/parity_report.txt gpurun_out/${P}_parity_report.txt 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${P}_bench.err
head -c 700 gpurun_out/${P}_bench.json; echo
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flow_warp4 -s 2 -c 1 -o gpurun_out/${P}_ncu_full_flow_warp4 -f python tools/mb_warp_once.py > gpurun_out/${P}_ncu_a.log 2>&1; echo "ncu warp rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_head -s 1 -c 1 -o gpurun_out/${P}_ncu_full_conv_head -f python tools/profile_step.py --seconds 12 > gpurun_out/${P}_ncu_b.log 2>&1; echo "ncu head rc=$?"
ls -la gpurun_out/*.ncu-rep
