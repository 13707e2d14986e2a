#!/usr/bin/env python
"""Benchmark of the per-frame lip-sync hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = one LNet forward (BASELINE.json configs[1]: batch 128, 96x96 faces, 80x16 mel windows,
16-bit operands / fp32 accumulation) over a batch of synthetic frames already resident in HBM.
`value` is whole-job frames/s (all ranks); `e2e` is the same metric with pinned HOST inputs and outputs, every step's
H2D + D2H inside the timed region, through the package's host batch loop `pipeline.stream_batches(LNet, ...)` (copy-in /
forward / copy-out on three streams) - the plain one-stream loop around `LNet.forward` is reported beside it.
N > 1 (torchrun): frames are independent, every rank runs its own batch (weak scaling, no data-path
collective); time = max over ranks.  --impl reference times the oracle port of the reference's
PyTorch path on the box's host cores (the reference tree itself cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAME_GFLOP = 56.10            # LNet conv+linear FLOPs per frame (SURVEY A.2 / BASELINE.md section 3)
METRIC = "generated frames/sec (LNet, 96 px, batch 128)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sus=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(sample_frames: int, reps: int = 1):
    """The reference's PyTorch fp32 eager path (oracle port, oracle/nets.py) on the host cores."""
    import torch
    from oracle import nets, synth, weights
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict("lnet", 0)
    mel, face = synth.lnet_inputs(sample_frames, seed=0)
    with torch.no_grad():
        nets.lnet_forward(sd, mel[:1], face[:1])          # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        for _ in range(reps):
            nets.lnet_forward(sd, mel, face)
        dt = (time.perf_counter() - t0) / reps
    return {"value": sample_frames / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/nets.py lnet_forward fp32 eager, batch %d of the 128-frame step, %d rep(s), %.2f s/rep" % (sample_frames, reps, dt)}


def _time(fn, reps, dev):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / reps


def measure_extras(dev, lnet):
    """The other rows of the hot path, each timed on the device with inputs resident in
    HBM: DNet B=64 (configs[2]), the fused warp kernel against the HBM roofline, the mel front end, and
    the chained full path (configs[3]: 60 s clip, 1497 frames)."""
    import torch
    from oracle import synth, weights
    from s2v_b200.futils import audio, flow_util
    from s2v_b200.models.DNet import DNet
    from s2v_b200.pipeline import LipSyncPipeline
    peaks = _peaks()
    out = {}
    dnet = DNet().to(dev).eval()
    dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
    src, coeff = synth.dnet_inputs(64, seed=0)
    src, coeff = src.to(dev), coeff.to(dev)
    ms = _time(lambda: dnet(src, coeff), 5, dev)
    out["dnet_b64"] = {"frames_per_s": round(64 / ms * 1e3, 1), "ms": round(ms, 3),
                       "tflops_useful": round(64 * 101.45 / ms, 1), "note": "DNet full forward, 256x256, batch 64 (configs[2]); 101.45 useful GFLOP/frame"}
    ms = _time(lambda: dnet(src, coeff, stage="warp"), 5, dev)
    out["dnet_b64_stage_warp"] = {"frames_per_s": round(64 / ms * 1e3, 1), "ms": round(ms, 3)}
    s, fl = synth.warp_inputs(64, seed=0)
    s, fl = s.to(dev), fl.to(dev)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def warp_flushed():
        scratch.zero_()                       # 256 MiB write: evicts the 126 MB L2 between timed launches
        flow_util.warp_flow(s, fl)
    ms_both = _time(warp_flushed, 10, dev)
    ms_flush = _time(lambda: scratch.zero_(), 10, dev)
    ms = max(ms_both - ms_flush, 1e-3)
    gb = 64 * 1605632 / 1e9
    out["warp_kernel_b64"] = {"us": round(ms * 1e3, 2), "GBps_algorithmic": round(gb / (ms * 1e-3), 1),
                              "frac_of_hbm_peak": round(gb / (ms * 1e-3) / peaks["hbm"], 3), "hbm_peak_GBps": peaks["hbm"],
                              "alg_bytes_per_frame": 1605632, "note": "fused flow_to_deformation+resize+grid_sample, fp32, L2 flushed between launches"}
    wav = torch.from_numpy(synth.wav(60.0, seed=0)).to(dev)
    ms = _time(lambda: audio.mel_windows(audio.melspectrogram_device(wav)), 10, dev)
    out["mel_60s_clip"] = {"ms": round(ms, 4), "stft_columns": 4801, "windows": 1497,
                           "GBps_algorithmic": round((4801 * 1120 + 1497 * 5120) / 1e9 / (ms * 1e-3), 2),
                           "note": "melspectrogram + 80x16 window gather; launch/latency-bound at this size"}
    n = 1497
    srcs, coeffs = synth.dnet_inputs(64, seed=1)
    srcs = srcs.to(dev).repeat((n + 63) // 64, 1, 1, 1)[:n]
    coeffs = coeffs.to(dev).repeat((n + 63) // 64, 1, 1)[:n]
    pipe = LipSyncPipeline(lnet, dnet)
    for _ in range(2):                        # warm-up: the tail-batch plans (25 / 89 frames) are built and graph-captured here
        pipe.run(wav, srcs, coeffs)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pipe.run(wav, srcs, coeffs)
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b)
    out["full_path_60s_clip"] = {"frames": n, "ms": round(ms, 2), "frames_per_s": round(n / ms * 1e3, 1),
                                 "note": "mel -> DNet(B=64) -> glue -> LNet(B=128) on one GPU, configs[3] at N=1"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.ref_frames, reps=1)
    # each "step" = the bounded sample; K steps + W warm-ups would repeat it; one rep is already ~10-30 s of CPU work
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.ref_frames / cb["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LNet forward, 96x96, 80x16 mel windows; bounded sample of %d frames per step on host cores" % args.ref_frames},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--ref-frames", type=int, default=32)
    ap.add_argument("--cpu-frames", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the DNet / warp / mel / full-path side measurements")
    ap.add_argument("--breakdown", default="", help="write the per-kernel-class time table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep NCCL's banner ("NCCL version ...", printed to stdout at NCCL_DEBUG >= VERSION) out of the one-line JSON stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("S2V_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    import s2v_b200  # noqa: F401
    from oracle import synth, weights           # synthetic inputs + seeded weights (not on the timed path)
    from s2v_b200.models.LNet import LNet

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    net = LNet().to(dev).eval()
    net.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    mel, face = synth.lnet_inputs(B, seed=rank)
    mel_d, face_d = mel.to(dev), face.to(dev)
    eng = net.engine()
    ent = eng.plan_for(B)
    ent["io"]["mel"].copy_(mel_d)
    ent["io"]["face"].copy_(face_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(W):
        eng._run(ent)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        eng._run(ent)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / K
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public API: pinned host -> device -> LNet.forward -> host ---------
    mel_h, face_h = mel.pin_memory(), face.pin_memory()
    out_h = torch.empty(B, 3, 96, 96, dtype=torch.float32).pin_memory()
    for _ in range(2):
        out_h.copy_(net(mel_h.to(dev, non_blocking=True), face_h.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    e0.record()
    for _ in range(K):
        out_h.copy_(net(mel_h.to(dev, non_blocking=True), face_h.to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_seq = world * B / (tms.item() / K * 1e-3)
    # the same K host batches through the package's batch loop (pipeline.stream_batches = the reference's generation loop,
    # inference.py:259-267, with copy-in / forward / copy-out on three streams): every step still copies its inputs from pinned
    # host memory and its frames back inside the timed region; the copies of neighbouring steps overlap the forward
    from s2v_b200.pipeline import stream_batches
    gen = lambda k: (((mel_h, face_h), out_h) for _ in range(k))
    stream_batches(net, gen(3))
    barrier()
    e0.record()
    stream_batches(net, gen(K))
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_pipe = world * B / (tms.item() / K * 1e-3)
    e2e_value = max(e2e_seq, e2e_pipe)
    e2e_api = ("s2v_b200.pipeline.stream_batches(LNet, pinned host batches): H2D, LNet.forward, D2H per step on three streams" if e2e_pipe >= e2e_seq
               else "out_h.copy_(LNet.forward(mel_h.to(dev), face_h.to(dev))) per step, pinned host tensors, one stream")

    # ---- per-kernel-class device times --------------------------------------------------------------
    # Every class's launches (in plan order) are captured into their OWN CUDA graph and its replay is timed with CUDA
    # events: pure device time of that class's kernels, without the host cost of issuing ~500 launches one by one (an
    # eager s2v_conv_tc call costs ~12 us of CPU - more than many of the kernels run).  The per-op table of --breakdown
    # still uses eager events around every launch (host-inflated for short kernels; use it for ranking only).
    roof, table = None, None
    if rank == 0:
        ops_l = ent["plan"].ops
        cls_of = lambda op: ("conv_tc" if op.name.endswith("[tc]") else "conv_head" if op.name.endswith("[head]")
                             else "conv_simt" if op.name.endswith("[simt]") else op.name)
        classes = {}
        for op in ops_l:
            classes.setdefault(cls_of(op), []).append(op)
        acc = {}
        side = torch.cuda.Stream(device=dev)
        for cls, cops in classes.items():
            with torch.cuda.stream(side):
                for op in cops:
                    op.run()
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for op in cops:
                        op.run()
                g.replay()
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                a.record(side)
                for _ in range(reps):
                    g.replay()
                b.record(side)
                torch.cuda.synchronize(dev)
            acc[cls] = [a.elapsed_time(b) / reps, len(cops), sum(getattr(op, "alg_flops", 0.0) for op in cops),
                        sum(getattr(op, "alg_bytes", 0.0) for op in cops)]
        total = sum(v[0] for v in acc.values())
        peaks = _peaks()
        # per-layer roofline of the tensor-core class: a layer cannot run faster than max(flops / tensor peak, bytes / HBM peak);
        # many of LNet's layers (1x1 spectral convs, 48x48 / 96x96 levels with 48-128 channels) are HBM-bound by that measure
        lay_ms = sum(max(op.alg_flops / (peaks["tf_sus"] * 1e12), op.io_bytes / (peaks["hbm"] * 1e9)) * 1e3 for op in classes.get("conv_tc", []))
        lay_hbm = sum(1 for op in classes.get("conv_tc", []) if op.io_bytes / (peaks["hbm"] * 1e9) > op.alg_flops / (peaks["tf_sus"] * 1e12))
        table = {}
        for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            table[k] = {"ms": round(v[0], 4), "launches": v[1], "share": round(v[0] / total, 4)}
            if v[3] > 0:       # memory-bound class: algorithmic bytes (each tensor read / written once) against the measured HBM peak
                gbs = v[3] / (v[0] * 1e-3) / 1e9
                table[k].update({"GBps_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)})
        tc = acc.get("conv_tc")
        if tc:
            ach = tc[2] / (tc[0] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit GEMM, %d launches/step)" % tc[1],
                    "achieved": round(ach, 2), "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sus"], 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of the step's dominant conv_tc shape (12x12 level, 3x3, K = 9216:
                    # 36 of the 245 launches, ~23 % of the class time) from profiles/r1c_ncu_full_conv_tc_res2_raw.csv (56.2 MB read +
                    # 1.3 MB written to DRAM; the 9.4 MB output is still in L2 when the kernel ends); its algorithmic bytes are 51.8 MB
                    # (37.7 in + 4.7 weights + 9.4 out), the rest is the 148 CTAs' weight tiles missing L2.  Other shapes: profiles/r1c_summary.md
                    "traffic": 57.5e6, "traffic_of": "one 3x3 K=9216 launch (ncu --set full, DRAM bytes per launch; algorithmic 51.8e6)",
                    "peak_source": peaks["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "alg_gflop_per_step": round(tc[2] / 1e9, 1), "kernel_ms_per_step": round(tc[0], 3),
                    "avg_launch_us": round(1e3 * tc[0] / tc[1], 2),
                    "share_of_step": round(tc[0] / total, 4), "sum_of_classes_ms": round(total, 3),
                    "per_layer_roofline": {"min_ms_per_step": round(lay_ms, 3), "frac": round(lay_ms / tc[0], 4), "hbm_bound_launches": lay_hbm,
                                           "how": "sum over the class's launches of max(flops / tensor peak, min HBM bytes / HBM peak) / measured class time"},
                    "how": "all conv_tc launches of one step replayed from their own CUDA graph, CUDA events on that stream"}
        if args.breakdown:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops_l]
            reps = 3
            per_op = {}
            for r in range(reps):
                torch.cuda.synchronize(dev)
                for op, (a, b) in zip(ops_l, evs):
                    a.record()
                    op.run()
                    b.record()
                torch.cuda.synchronize(dev)
                if r == 0:
                    continue                      # first eager pass = warm-up
                for op, (a, b) in zip(ops_l, evs):
                    po = per_op.setdefault(op.name, [0.0, 0, 0.0])
                    po[0] += a.elapsed_time(b) / (reps - 1)
                    po[1] += 1 if r == 1 else 0
                    po[2] += getattr(op, "alg_flops", 0.0) if r == 1 else 0.0
            with open(args.breakdown, "w") as f:
                import re
                grouped = {}
                for name, v in per_op.items():          # fold the 9 blocks x 2 convs of a decoder level together
                    key = re.sub(r"res(\d)\.res\d\.conv\d", r"res\1.*", name)
                    key = re.sub(r"layers\.\d", "layers.*", key)
                    key = re.sub(r"audio_encoder\.\d+", "audio_encoder.*", key)
                    g = grouped.setdefault(key, [0.0, 0, 0.0])
                    g[0] += v[0]; g[1] += v[1]; g[2] += v[2]
                rows = [{"op": k, "ms": round(v[0], 4), "launches": v[1], "us_per_launch": round(1e3 * v[0] / max(v[1], 1), 2),
                         "tflops": round(v[2] / (v[0] * 1e-3) / 1e12, 1) if v[2] and v[0] > 0 else None}
                        for k, v in sorted(grouped.items(), key=lambda kv: -kv[1][0])]
                json.dump({"per_class": table, "sum_ms": sum(v[0] for v in per_op.values()), "ms_per_step_graph": ms_step, "per_op_group": rows}, f, indent=1)
        eng._run(ent)                             # leave the workspace in a consistent state again
        torch.cuda.synchronize(dev)

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = measure_extras(dev, net)
    if rank == 0:
        cb = None if args.no_cpu_baseline else cpu_baseline(args.cpu_frames)
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[1]: LNet forward, batch %d per GPU, 96x96 faces, 80x16 mel windows, "
                                       "seeded random-init weights (oracle/weights.py seed 0)" % B,
                           "frames_per_step_per_gpu": B, "l2": "per-step working set (>1 GB of activations + 255 MB weights) exceeds the 126 MB L2; no explicit flush",
                           "parallelism": "frame-sharded x%d, no data-path collective" % world,
                           "gflop_per_frame": FRAME_GFLOP},
                "tflops_algorithmic": round(value * FRAME_GFLOP / 1e3, 1),
                "clocks": clocks,
                "e2e": {"value": round(e2e_value, 1), "unit": "frames/s",
                        "h2d_bytes_per_step": int(mel.numel() * 4 + face.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4),
                        "api": e2e_api, "stream_batches_value": round(e2e_pipe, 1), "sequential_loop_value": round(e2e_seq, 1)},
                "gpu_launches": K * len(ent["plan"]),
                "launches_per_step": len(ent["plan"]),
                "roofline": roof, "cpu_baseline": cb, "kernel_classes": table, "other_rows": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
