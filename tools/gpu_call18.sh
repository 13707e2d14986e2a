#!/bin/bash
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT NCCL_DEBUG_FILE=gpurun_out/r2b_nccl_n${N}_%p.log $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err; echo rc=$?
wc -l gpurun_out/r2b_bench_n$N.json; head -c 300 gpurun_out/r2b_bench_n$N.json; echo
ls gpurun_out/r2b_nccl_n${N}_*.log | head -1 | xargs -I{} cp {} gpurun_out/r2b_nccl_rank_sample_n$N.log; rm -f gpurun_out/r2b_nccl_n${N}_[0-9]*.log
$TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref_n$N.json 2>/dev/null; echo "ref rc=$? lines=$(wc -l < gpurun_out/r2b_bench_ref_n$N.json)"
if [ "$N" = "8" ]; then
$TR tools/bench_clip.py --seconds 600 --batches 128,256 --reps 2 > gpurun_out/r2b_clip600_n8.jsonl 2> gpurun_out/r2b_clip600_n8.err; echo rc=$?
cut -c1-330 gpurun_out/r2b_clip600_n8.jsonl
fi
