#!/usr/bin/env python
"""Benchmark of the per-frame lip-sync hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seconds S]

The metric is BASELINE.json's: generated frames/s of the FULL per-frame path (mel -> DNet 256x256 -> glue -> LNet 96x96).
A step = one pass over BASELINE.json configs[3]: a 60 s synthetic clip (960 000 samples -> 4 801 STFT columns -> 1 497
frames, 25 fps), DNet in batches of <= 192, LNet in batches of <= 256 (`pipeline.LipSyncPipeline`).

* `value`: whole-job frames/s with the clip's inputs (wav, 256x256 sources, 3DMM windows) resident in HBM.  N > 1 (torchrun):
  STRONG scaling - the same clip is frame-sharded (`parallel.shard_range`: rank r owns a contiguous frame range, full weight
  replica, whole mel computed locally, no collective on the compute path) and the generated frames are gathered over NVLink
  with one all_gather (`parallel.gather_frames`) INSIDE the timed region; time = max over ranks.
* `e2e`: the same step through the same public call with PINNED HOST inputs and outputs: every rank copies its shard's
  sources / windows / wav to the device and its frames back inside the timed region (copy streams overlap the networks).
* `roofline`: the dominant kernel class (conv_tc, tcgen05 implicit GEMM) over every plan the step replays, each class timed
  from its own CUDA graph with CUDA events; `other_rows` carries configs[1] (LNet B=128) and configs[2] (DNet B=64) with their
  own kernel-class tables, the warp / mel kernels against the HBM roofline and stock PyTorch eager on the same GPU.
* `--impl reference`: the reference's PyTorch fp32 eager path (oracle port - the reference tree cannot travel to the GPU box)
  on the box's host cores, same metric, each step a bounded sample of the clip's frames; every declared step really runs.
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

LNET_GFLOP, DNET_GFLOP = 56.10, 101.45          # conv + linear FLOPs per frame
FRAME_GFLOP = LNET_GFLOP + DNET_GFLOP
METRIC = "generated frames/sec (LNet+DNet, 96/256 px)"


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


# ------------------------------------------------------------------------------------------------ CPU arm
class CpuPath:
    """The reference's per-frame path on the host cores (TEST INFRASTRUCTURE used as the baseline): oracle/mel.py (fp64 numpy
    restatement of futils/audio.py) + oracle/nets.py (functional port of models/DNet.py, models/LNet.py, futils/flow_util.py,
    fp32 eager) with the reference's own batching: DNet one frame at a time (preprocessing/facing.py:176-194), LNet one
    frame per forward here (batch 1).  Inputs: the 5 s seeded clip of BASELINE.json configs[0]."""

    def __init__(self):
        import torch
        from oracle import mel as omel, nets, synth, weights
        torch.set_num_threads(os.cpu_count() or 1)
        self.torch, self.omel, self.nets = torch, omel, nets
        self.sd_l, self.sd_d = weights.make_state_dict("lnet", 0), weights.make_state_dict("dnet", 0)
        self.wav = synth.wav(5.0, seed=0)
        self.n_clip = len(omel.mel_window_starts(1 + len(self.wav) // 200))      # 122
        self.src, self.coeff = synth.dnet_inputs(8, seed=1)
        self.cores = torch.get_num_threads()
        t0 = time.perf_counter()
        self.windows = torch.from_numpy(omel.mel_windows(omel.melspectrogram(self.wav)))
        self.mel_s = time.perf_counter() - t0                                    # whole-clip mel + window time (122 frames)
        with torch.no_grad():                                                    # thread pool / allocator warm-up
            self.frames(1)

    def frames(self, f):
        """f frames through DNet -> glue -> LNet, one at a time; returns seconds."""
        torch, nets = self.torch, self.nets
        t0 = time.perf_counter()
        with torch.no_grad():
            for i in range(f):
                j = i % self.src.shape[0]
                fake = nets.dnet_forward(self.sd_d, self.src[j:j + 1], self.coeff[j:j + 1])["fake_image"]
                face = nets.glue_dnet_to_lnet(fake)
                nets.lnet_forward(self.sd_l, self.windows[i % self.n_clip:i % self.n_clip + 1], face)
        return time.perf_counter() - t0

    def step(self, f):
        """One bounded sample: the mel of f frames' worth of audio (timed: a fresh melspectrogram of f/122 of the clip,
        at least 16 columns) + f frames of the networks.  Returns seconds."""
        n = max(int(len(self.wav) * f / self.n_clip), 16 * 200)
        t0 = time.perf_counter()
        self.omel.mel_windows(self.omel.melspectrogram(self.wav[:n]))
        return (time.perf_counter() - t0) + self.frames(f)


def cpu_baseline(sample_frames: int):
    cp = CpuPath()
    dt = cp.step(sample_frames)
    return {"value": sample_frames / dt, "unit": "frames/s", "cores": cp.cores, "kind": "port",
            "sample": "oracle port (oracle/mel.py + oracle/nets.py, fp32 eager) of mel -> DNet(B=1) -> glue -> LNet(B=1): %d frames of the "
                      "5 s seeded clip of configs[0] (whole-clip mel %.3f s for 122 frames), %.2f s" % (sample_frames, cp.mel_s, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cp = CpuPath()
    f = args.ref_frames
    W, K = max(args.warmup, 0), max(args.steps, 1)
    for _ in range(W):
        cp.step(f)
    t0 = time.perf_counter()
    for _ in range(K):
        cp.step(f)
    dt = (time.perf_counter() - t0) / K
    val = f / dt
    cb = {"value": val, "unit": "frames/s", "cores": cp.cores, "kind": "port",
          "sample": "each step = %d frames of the 5 s seeded clip through the oracle port (mel restatement + DNet B=1 -> glue -> LNet B=1, "
                    "fp32 eager, all host threads); %d warm-up + %d timed steps really run" % (f, W, K)}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (mel f64)", "data": "synthetic",
            "config": {"workload": "full per-frame path mel -> DNet 256x256 -> glue -> LNet 96x96 (BASELINE.json configs[3] metric); "
                                   "bounded sample of %d frames per step on the host cores, reference batching (one frame per forward)" % f,
                       "frames_per_step": f},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU helpers
def _time(fn, reps, dev, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / reps


def _cls_of(op):
    return ("conv_tc" if op.name.endswith("[tc]") else "conv_head" if op.name.endswith("[head]")
            else "conv_simt" if op.name.endswith("[simt]") else op.name)


def class_times(dev, ent, reps=3):
    """Device time of every kernel class of one plan: the class's launches (in plan order) are captured into their OWN CUDA
    graph and its replay is timed with CUDA events on that stream - pure device time, without the host cost of issuing
    hundreds of launches one by one.  Returns {class: [ms, launches, alg_flops, alg_bytes, per-layer-roofline ms, hbm-bound launches]}."""
    import torch
    peaks = _peaks()
    classes = {}
    for op in ent["plan"].ops:
        classes.setdefault(_cls_of(op), []).append(op)
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
            a.record(side)
            for _ in range(reps):
                g.replay()
            b.record(side)
            torch.cuda.synchronize(dev)
        tc_like = cls in ("conv_tc", "conv_head")
        lay = sum(max(getattr(op, "alg_flops", 0.0) / (peaks["tf_sus"] * 1e12), op.io_bytes / (peaks["hbm"] * 1e9)) * 1e3 for op in cops) if tc_like else 0.0
        nh = sum(1 for op in cops if op.io_bytes / (peaks["hbm"] * 1e9) > getattr(op, "alg_flops", 0.0) / (peaks["tf_sus"] * 1e12)) if tc_like else 0
        acc[cls] = [a.elapsed_time(b) / reps, len(cops), sum(getattr(op, "alg_flops", 0.0) for op in cops),
                    sum(getattr(op, "alg_bytes", 0.0) for op in cops), lay, nh]
    return acc


def class_table(acc_list):
    """[(acc, uses)] -> (table, totals) summed over plans."""
    peaks = _peaks()
    tot = {}
    for acc, uses in acc_list:
        for k, v in acc.items():
            t = tot.setdefault(k, [0.0, 0, 0.0, 0.0, 0.0, 0])
            for i in range(6):
                t[i] += v[i] * uses
    total = sum(v[0] for v in tot.values()) or 1.0
    table = {}
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        row = {"ms": round(v[0], 4), "launches": v[1], "share": round(v[0] / total, 4)}
        if v[3] > 0 and v[0] > 0:          # memory-bound class: algorithmic bytes (each tensor read / written once) vs the measured HBM peak
            gbs = v[3] / (v[0] * 1e-3) / 1e9
            row.update({"GBps_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)})
        if v[2] > 0 and v[0] > 0:
            tf = v[2] / (v[0] * 1e-3) / 1e12
            row.update({"TFLOPs_algorithmic": round(tf, 1), "frac_of_tensor_peak": round(tf / peaks["tf_sus"], 3)})
        table[k] = row
    return table, tot, total


def roofline_of(tot, total, what):
    peaks = _peaks()
    tc = tot.get("conv_tc")
    if not tc or tc[0] <= 0:
        return None
    ach = tc[2] / (tc[0] * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 / TMEM / TMA implicit GEMM), %d launches per %s" % (tc[1], what),
            "achieved": round(ach, 2), "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sus"], 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of the class's largest launch on this metric (DNet editing_net.encoder.down0,
            # 256 x 256, 64 -> 128, 3x3, batch 64) from one ncu --set full capture of the current kernel
            # (profiles/r2b_ncu_full_conv_tc_dnet_down0_raw.csv): 538.1 MB read + 1019.2 MB written; algorithmic 1610.8 MB
            # (536.9 in + 0.15 weights + 1073.7 out) - no re-reads.  LNet 48 x 48 merged FFC GEMM: 153.9 MB vs 185.7 algorithmic
            # (profiles/r2b_ncu_full_conv_tc_res0_all_raw.csv; part of the input still in L2)
            "traffic": 1557.3e6, "traffic_of": "one DNet down0 launch (256x256, 64->128, 3x3, B=64; ncu --set full, DRAM bytes per launch; algorithmic 1610.8e6); other shapes: profiles/",
            "peak_source": peaks["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
            "alg_gflop": round(tc[2] / 1e9, 1), "kernel_ms": round(tc[0], 3), "avg_launch_us": round(1e3 * tc[0] / tc[1], 2),
            "share_of_step_kernels": round(tc[0] / total, 4), "sum_of_classes_ms": round(total, 3),
            "per_layer_roofline": {"min_ms": round(tc[4], 3), "frac": round(tc[4] / tc[0], 4), "hbm_bound_launches": tc[5],
                                   "how": "sum over the class's launches of max(flops / tensor peak, min HBM bytes / HBM peak) / measured class time"},
            "how": "every conv_tc launch of every plan the step replays, replayed from its own CUDA graph, CUDA events on that stream, x uses per step"}


def measure_extras(dev, lnet, dnet, args):
    """configs[1] / configs[2] and the memory-bound kernels, each timed on the device with inputs resident in HBM; plus the
    stock-PyTorch-eager row (oracle/nets.py on the same GPU: cuDNN / cuBLAS / cuFFT / ATen - SURVEY 2.2's per-kernel bar)."""
    import torch
    from oracle import nets, synth, weights
    from s2v_b200.futils import audio, flow_util
    peaks = _peaks()
    out = {}
    # ---- configs[1]: LNet B=128 ---------------------------------------------------------------------------------------
    mel, face = synth.lnet_inputs(128, seed=0)
    mel, face = mel.to(dev), face.to(dev)
    leng = lnet.engine()
    ent = leng.plan_for(128)
    ent["io"]["mel"].copy_(mel)
    ent["io"]["face"].copy_(face)
    ms = _time(lambda: leng._run(ent), 20, dev)
    table, tot, total = class_table([(class_times(dev, ent), 1)])
    leng._run(ent)
    out["lnet_b128"] = {"frames_per_s": round(128 / ms * 1e3, 1), "ms": round(ms, 3), "tflops_algorithmic": round(128 * LNET_GFLOP / ms, 1),
                        "frac_of_tensor_peak": round(128 * LNET_GFLOP / ms / peaks["tf_sus"], 3), "launches": len(ent["plan"]),
                        "roofline": roofline_of(tot, total, "forward"), "kernel_classes": table,
                        "note": "BASELINE.json configs[1]: LNet forward, batch 128, 96x96, graph replay, inputs resident"}
    # ---- configs[2]: DNet B=64 ----------------------------------------------------------------------------------------
    src, coeff = synth.dnet_inputs(64, seed=0)
    src, coeff = src.to(dev), coeff.to(dev)
    deng = dnet.engine()
    dnet(src, coeff)
    dent = deng._plans[(64, 26, "full")]
    ms = _time(lambda: deng._run(dent), 10, dev)
    table, tot, total = class_table([(class_times(dev, dent), 1)])
    deng._run(dent)
    out["dnet_b64"] = {"frames_per_s": round(64 / ms * 1e3, 1), "ms": round(ms, 3), "tflops_useful": round(64 * DNET_GFLOP / ms, 1),
                       "frac_of_tensor_peak": round(64 * DNET_GFLOP / ms / peaks["tf_sus"], 3), "launches": len(dent["plan"]),
                       "roofline": roofline_of(tot, total, "forward"), "kernel_classes": table,
                       "note": "BASELINE.json configs[2]: DNet full forward, 256x256, batch 64; 101.45 useful GFLOP/frame"}
    ms = _time(lambda: dnet(src, coeff, stage="warp"), 5, dev)
    out["dnet_b64_stage_warp"] = {"frames_per_s": round(64 / ms * 1e3, 1), "ms": round(ms, 3)}
    # ---- the fused warp kernel against the HBM roofline (L2 flushed between launches) -------------------------------------
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms_flush = _time(lambda: scratch.zero_(), 10, dev)
    gb = 64 * 1605632 / 1e9
    for name, (s, fl) in (("warp_kernel_b64", synth.warp_inputs(64, seed=0)),
                          ("warp_kernel_b64_dnet_flow", (src.cpu(), dnet(src, coeff, stage="warp")["flow_field"].cpu()))):
        s, fl = s.to(dev), fl.to(dev)

        def warp_flushed():
            scratch.zero_()                       # 256 MiB write: evicts the 126 MB L2 between timed launches
            flow_util.warp_flow(s, fl)
        ms = max(_time(warp_flushed, 10, dev) - ms_flush, 1e-3)
        out[name] = {"us": round(ms * 1e3, 2), "GBps_algorithmic": round(gb / (ms * 1e-3), 1),
                     "frac_of_hbm_peak": round(gb / (ms * 1e-3) / peaks["hbm"], 3), "hbm_peak_GBps": peaks["hbm"], "alg_bytes_per_frame": 1605632,
                     "note": "fused flow_to_deformation + resize + grid_sample, fp32, " + ("N(0,3^2) synthetic flow (configs[2] iii)" if name == "warp_kernel_b64" else "the flow DNet itself predicts")}
    # ---- mel front end, 60 s and 600 s clips -----------------------------------------------------------------------------
    for sec in (60.0, 600.0):
        wav = torch.from_numpy(synth.wav(sec, seed=0)).to(dev)
        t_cols = 1 + wav.numel() // 200
        n_win = audio.mel_window_count(t_cols, 25.0)

        def mel_flushed():
            scratch.zero_()
            audio.mel_windows(audio.melspectrogram_device(wav))
        ms = max(_time(mel_flushed, 10, dev) - ms_flush, 1e-3)
        gbm = (t_cols * 1120 + n_win * 5120) / 1e9
        out["mel_%ds_clip" % sec] = {"us": round(ms * 1e3, 2), "stft_columns": t_cols, "windows": n_win, "GBps_algorithmic": round(gbm / (ms * 1e-3), 1),
                                     "frac_of_hbm_peak": round(gbm / (ms * 1e-3) / peaks["hbm"], 4),
                                     "note": "melspectrogram + 80x16 window gather (2 launches), L2 flushed; 1120 B per STFT column + 5120 B per window"}
    # ---- stock PyTorch eager on the same GPU (the oracle port on CUDA) ------------------------------------------------------
    if not args.no_torch_eager:
        sdl = {k: v.to(dev) for k, v in weights.make_state_dict("lnet", 0).items()}
        sdd = {k: v.to(dev) for k, v in weights.make_state_dict("dnet", 0).items()}
        rows = {}
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ms_l = _time(lambda: nets.lnet_forward(sdl, mel, face), 3, dev, warm=2)
                ms_d = _time(lambda: nets.dnet_forward(sdd, src, coeff), 3, dev, warm=2)
            rows[tag] = {"lnet_b128_frames_per_s": round(128 / ms_l * 1e3, 1), "dnet_b64_frames_per_s": round(64 / ms_d * 1e3, 1),
                         "full_path_frames_per_s": round(1e3 / (ms_l / 128 + ms_d / 64), 1)}
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        rows["note"] = ("oracle/nets.py (functional port of the reference modules) in PyTorch eager on this GPU: cuDNN convs, cuBLAS linears, cuFFT, "
                        "ATen elementwise; full_path = 1 / (LNet B=128 time per frame + DNet B=64 time per frame)")
        out["torch_eager_b200"] = rows
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=60.0, help="clip length (60 = configs[3], 600 = configs[4])")
    ap.add_argument("--lnet-batch", type=int, default=256)
    ap.add_argument("--dnet-batch", type=int, default=192)
    ap.add_argument("--ref-frames", type=int, default=2)
    ap.add_argument("--cpu-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the LNet / DNet / warp / mel / torch-eager side measurements")
    ap.add_argument("--no-torch-eager", action="store_true")
    ap.add_argument("--no-classes", action="store_true", help="skip the per-kernel-class graphs (no roofline object)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE line (the JSON): everything else any library prints at the C level (NCCL's version banner and its
    # NCCL_DEBUG=INFO init lines go to stdout by default) is sent to stderr by pointing fd 1 at fd 2 for the whole run; the JSON
    # line is written to the saved descriptor at the end.  NCCL's init lines therefore stay visible (ranks, NVLS, channels).
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)

    import s2v_b200  # noqa: F401
    from oracle import synth, weights           # synthetic inputs + seeded weights (not on the timed path)
    from s2v_b200 import parallel
    from s2v_b200.futils import audio
    from s2v_b200.models.DNet import DNet
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import LipSyncPipeline, balanced_batches

    K, W = max(args.steps, 1), max(args.warmup, 3)
    lnet = LNet().to(dev).eval()
    lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    dnet = DNet().to(dev).eval()
    dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)

    wav_np = synth.wav(args.seconds, seed=0)
    total = audio.mel_window_count(1 + len(wav_np) // 200, 25.0)
    lo, hi = parallel.shard_range(total, rank, world)
    n = hi - lo
    # this rank's frames only: frame i uses synthetic source (i mod 64), so shards see the same per-frame inputs as N=1
    src64, co64 = synth.dnet_inputs(64, seed=1)
    idx = torch.arange(lo, hi) % 64
    srcs_h, coeffs_h = src64[idx].contiguous().pin_memory(), co64[idx].contiguous().pin_memory()
    wav_h = torch.from_numpy(wav_np).pin_memory()
    out_h = torch.empty(n, 3, 96, 96, dtype=torch.float32).pin_memory()
    wav, srcs, coeffs = wav_h.to(dev), srcs_h.to(dev), coeffs_h.to(dev)
    pipe = LipSyncPipeline(lnet, dnet, lnet_batch=args.lnet_batch, dnet_batch=args.dnet_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step):
        for _ in range(W):
            out = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(K):
            out = step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / K, out

    # ---- device-resident throughput (the clip's inputs already in HBM) ------------------------------------------------------
    def step_dev():
        return parallel.gather_frames(pipe.run(wav, srcs, coeffs, rank, world), total)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, out = timed(step_dev)
    clocks = sampler.stop() if rank == 0 else None
    assert out.shape[0] == total
    checksum = float(out.double().sum().item())
    value = total / (ms_step * 1e-3)

    # ---- end to end: pinned host inputs -> device -> path -> pinned host frames, copies inside the timed region ----------------
    def step_e2e():
        return parallel.gather_frames(pipe.run(wav_h, srcs_h, coeffs_h, rank, world, out_host=out_h), total)
    ms_e2e, out2 = timed(step_e2e)
    e2e_value = total / (ms_e2e * 1e-3)
    h2d = torch.tensor([float(wav_h.numel() * 4 + srcs_h.numel() * 4 + coeffs_h.numel() * 4)], device=dev)
    if world > 1:
        dist.all_reduce(h2d)
    e2e_ok = bool(torch.equal(out2, out)) and bool(torch.equal(out_h, out[lo:hi].cpu()))

    # ---- launches per step + per-kernel-class device times of every plan the step replays ---------------------------------------
    leng, deng = lnet.engine(), dnet.engine()
    d_sizes, l_sizes = balanced_batches(n, args.dnet_batch), balanced_batches(n, args.lnet_batch)
    pad8 = lambda b: b if (b < 8 or b % 8 == 0) else (b + 7) // 8 * 8
    uses = []                                       # (engine, plan key, uses per step)
    for b in sorted(set(pad8(x) for x in l_sizes)):
        uses.append((leng, b, sum(1 for x in l_sizes if pad8(x) == b)))
    for b in sorted(set(pad8(x) for x in d_sizes)):
        uses.append((deng, (b, 26, "full"), sum(1 for x in d_sizes if pad8(x) == b)))
    launches = 2 + len(d_sizes) + sum(len(e._plans[k]["plan"]) * u for e, k, u in uses)      # mel + windows, one glue per DNet batch
    lt = torch.tensor([float(launches)], device=dev)
    if world > 1:
        dist.all_reduce(lt)
    roof = table = None
    if rank == 0 and not args.no_classes:
        acc_list = [(class_times(dev, e._plans[k]), u) for e, k, u in uses]
        table, tot, total_ms = class_table(acc_list)
        roof = roofline_of(tot, total_ms, "step (rank 0's shard)")
        for e, k, u in uses:                         # leave the workspaces in a consistent state again
            e._run(e._plans[k])
        torch.cuda.synchronize(dev)

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = measure_extras(dev, lnet, dnet, args)
    if rank == 0:
        cb = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args.cpu_frames)
        peaks = _peaks()
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (mel, warp, statistics f32)", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[3]: full per-frame path mel -> DNet(256x256, batches <= %d) -> glue -> LNet(96x96, batches <= %d) "
                                       "on a %.0f s synthetic clip (%d frames), seeded random-init weights (oracle/weights.py seed 0)"
                                       % (args.dnet_batch, args.lnet_batch, args.seconds, total),
                           "frames_per_step": total, "frames_per_rank": n,
                           "parallelism": "frame-sharded x%d (contiguous ranges), no data-path collective, one final all_gather of the frames over NVLink inside the timed region" % world,
                           "l2": "inputs larger than L2: %.2f GB of sources per rank and > 1 GB of activations per batch against the 126 MB L2; no explicit flush" % (srcs.numel() * 4 / 1e9),
                           "gflop_per_frame": FRAME_GFLOP, "dnet_lnet_overlap": bool(pipe.overlap)},
                "tflops_algorithmic": round(value * FRAME_GFLOP / 1e3, 1),
                "frac_of_tensor_peak_whole_step": round(value * FRAME_GFLOP / 1e3 / (peaks["tf_sus"] * world), 4),
                "clocks": clocks,
                "e2e": {"value": round(e2e_value, 1), "unit": "frames/s", "ms_per_step": round(ms_e2e, 3),
                        "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(total * 3 * 96 * 96 * 4),
                        "api": "s2v_b200.pipeline.LipSyncPipeline.run(pinned host wav / sources / windows, out_host=pinned frames): every rank stages its "
                               "shard batch by batch on a copy stream and copies its frames back on another; + parallel.gather_frames",
                        "matches_device_resident_run": e2e_ok},
                "gpu_launches": int(K * lt.item()), "launches_per_step": int(lt.item()),
                "frame_checksum": checksum,
                "roofline": roof, "cpu_baseline": cb, "kernel_classes": table, "other_rows": extras}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
