#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; head -c 600 gpurun_out/r2u_bench.json; echo
