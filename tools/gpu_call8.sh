#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT NCCL_DEBUG_FILE=gpurun_out/r2i_nccl_%p.log $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err; echo rc=$?
head -c 400 gpurun_out/r2i_bench_n8.json; echo
grep -h -i "nranks" gpurun_out/r2i_nccl_*.log | head -3 | cut -c1-200
grep -h -i "NVLS" gpurun_out/r2i_nccl_*.log | head -2 | cut -c1-200
rm -f gpurun_out/r2i_nccl_*.log.keep; ls gpurun_out/r2i_nccl_*.log | head -1 | xargs -I{} cp {} gpurun_out/r2i_nccl_rank_sample.log; rm -f gpurun_out/r2i_nccl_[0-9]*.log
$TR bench.py --gpus 8 --steps 10 --warmup 3 --no-classes --lnet-batch 192 > gpurun_out/r2i_bench_n8_lb192.json 2>/dev/null; head -c 200 gpurun_out/r2i_bench_n8_lb192.json; echo
$TR tools/bench_clip.py --seconds 600 --batches 64,128,256,512 --reps 2 > gpurun_out/r2i_clip600_n8.jsonl 2> gpurun_out/r2i_clip600_n8.err; echo rc=$?
cat gpurun_out/r2i_clip600_n8.jsonl | cut -c1-330
