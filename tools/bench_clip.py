#!/usr/bin/env python
"""Full per-frame path on a synthetic clip, frame-sharded over the GPUs of one box (BASELINE.json configs[3] and [4]).

    python tools/bench_clip.py --seconds 60                       # configs[3] at N=1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_clip.py --seconds 600 --batches 32,64,128,256,512         # configs[4] sweep

Every rank computes the whole mel locally, takes its contiguous frame range (parallel.shard_range), runs
mel windows -> DNet -> glue -> LNet on it, and the generated frames are gathered with ONE all_gather over
NVLink (parallel.gather_frames) INSIDE the timed region.  Time = CUDA events per rank, max over ranks.
One JSON line per LNet batch size on rank 0.  Inputs are synthetic (oracle/synth.py) and resident in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--batches", default="256", help="LNet batch caps to sweep")
    ap.add_argument("--dnet-batch", type=int, default=192, help="DNet batch cap")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    import s2v_b200  # noqa: F401
    from oracle import synth, weights
    from s2v_b200 import parallel
    from s2v_b200.futils import audio
    from s2v_b200.models.DNet import DNet
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import LipSyncPipeline

    lnet = LNet().to(dev).eval()
    lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    dnet = DNet().to(dev).eval()
    dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)

    wav = torch.from_numpy(synth.wav(args.seconds, seed=0)).to(dev)
    total = audio.mel_window_count(1 + wav.numel() // 200, 25.0)
    lo, hi = parallel.shard_range(total, rank, world)
    n = hi - lo
    srcs, coeffs = synth.dnet_inputs(64, seed=1)
    # this rank's frames only: frame i uses synthetic source (i mod 64), so shards see the same per-frame inputs as N=1
    idx = (torch.arange(lo, hi) % 64).to(dev)
    srcs, coeffs = srcs.to(dev)[idx].contiguous(), coeffs.to(dev)[idx].contiguous()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for b in [int(x) for x in args.batches.split(",")]:
        pipe = LipSyncPipeline(lnet, dnet, lnet_batch=b, dnet_batch=min(b, args.dnet_batch))

        def step():
            frames = pipe.run(wav, srcs, coeffs, rank, world)
            return parallel.gather_frames(frames, total)

        for _ in range(2):                               # warm-up: plans, graph capture (tail batches too), NCCL channels
            out = step()
        assert out.shape[0] == total
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            out = step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = t.item()
            print(json.dumps({"workload": "mel -> DNet -> glue -> LNet on a %.0f s synthetic clip (%d frames), frame-sharded, final all_gather timed" % (args.seconds, total),
                              "n_gpus": world, "frames": total, "frames_per_rank": n, "lnet_batch": b, "dnet_batch": min(b, args.dnet_batch),
                              "ms": round(ms, 2), "frames_per_s": round(total / ms * 1e3, 1),
                              "gather_bytes": int(total * 3 * 96 * 96 * 4), "reps": args.reps,
                              "checksum": float(out.double().sum().item())}), flush=True)
        del pipe
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
