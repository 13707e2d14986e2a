#!/usr/bin/env python
"""One step of bench.py's workload (60 s clip, mel -> DNet -> glue -> LNet on one GPU) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists (development tool).  --seconds shortens the clip."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import s2v_b200  # noqa
from oracle import synth, weights
from s2v_b200.futils import audio
from s2v_b200.models.DNet import DNet
from s2v_b200.models.LNet import LNet
from s2v_b200.pipeline import LipSyncPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lnet = LNet().to(dev).eval(); lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
dnet = DNet().to(dev).eval(); dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
wav = torch.from_numpy(synth.wav(args.seconds, 0)).to(dev)
n = audio.mel_window_count(1 + wav.numel() // 200, 25.0)
s64, c64 = synth.dnet_inputs(64, 1)
idx = torch.arange(n) % 64
src, co = s64[idx].to(dev), c64[idx].to(dev)
pipe = LipSyncPipeline(lnet, dnet)
for _ in range(3):
    out = pipe.run(wav, src, co)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = pipe.run(wav, src, co)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("frames", tuple(out.shape), "checksum", float(out.double().sum()))
