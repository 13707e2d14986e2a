#!/usr/bin/env python
"""Per-layer-group device times of the real LNet / DNet plans (development tool).  Every group's launches are captured into
their own CUDA graph and replayed (pure device time, plan order inside the group).

    python tools/plan_breakdown.py [lnet|dnet|enet] [B]
"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import s2v_b200  # noqa
from oracle import synth, weights

which = sys.argv[1] if len(sys.argv) > 1 else "lnet"
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
if which == "lnet":
    from s2v_b200.models.LNet import LNet
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    net = LNet().to(dev).eval(); net.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    mel, face = synth.lnet_inputs(B, 0)
    net(mel.to(dev), face.to(dev)); eng = net.engine(); ent = eng._plans[B]
elif which == "dnet":
    from s2v_b200.models.DNet import DNet
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    net = DNet().to(dev).eval(); net.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
    s, c = synth.dnet_inputs(B, 0)
    net(s.to(dev), c.to(dev)); eng = net.engine(); ent = eng._plans[(B, 26, "full")]
else:
    from oracle import enet as oenet
    from s2v_b200.models.ENet import ENet
    from s2v_b200.models.LNet import LNet
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    net = ENet(lnet=LNet()).to(dev).eval(); net.load_state_dict(oenet.make_state_dict(0), strict=True)
    mel, _ = synth.lnet_inputs(B, 0)
    face, gt = torch.rand(B, 6, 384, 384), torch.rand(B, 3, 384, 384)
    net(mel.to(dev), face.to(dev), gt.to(dev)); eng = net.engine(); ent = list(eng._plans.values())[0]


def key(name):
    k = re.sub(r"res(\d)\.res\d\.conv\d", r"res\1.*", name)
    k = re.sub(r"layers\.\d", "layers.*", k)
    k = re.sub(r"audio_encoder\.\d+", "audio_encoder.*", k)
    k = re.sub(r"\.ph\d\d", ".ph*", k)
    k = re.sub(r"res(\d)\.res\d", r"res\1.*", k)
    return k


groups = {}
for op in ent["plan"].ops:
    groups.setdefault(key(op.name), []).append(op)
# memory-bound ops share a name: split them by their algorithmic byte count so that the big ones are visible
split = {}
for k, ops_ in groups.items():
    if all(getattr(o, "alg_flops", 0.0) == 0.0 for o in ops_) and len(ops_) > 4:
        for o in ops_:
            split.setdefault("%s[%5.0fMB]" % (k, o.alg_bytes / 1e6), []).append(o)
    else:
        split[k] = ops_
side = torch.cuda.Stream(device=dev)
rows = []
for k, ops_ in split.items():
    with torch.cuda.stream(side):
        for o in ops_:
            o.run()
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for o in ops_:
                o.run()
        g.replay(); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        for _ in range(5):
            g.replay()
        b.record(side); torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / 5
    fl = sum(getattr(o, "alg_flops", 0.0) for o in ops_)
    by = sum(getattr(o, "alg_bytes", 0.0) or getattr(o, "io_bytes", 0.0) for o in ops_)
    ideal = sum(max(getattr(o, "alg_flops", 0.0) / (peaks["bf16_tflops_sustained"] * 1e12), (getattr(o, "alg_bytes", 0.0) or getattr(o, "io_bytes", 0.0)) / (peaks["hbm_gbs"] * 1e9)) for o in ops_) * 1e3
    rows.append((ms, k, len(ops_), fl, by, ideal))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("%s B=%d: %d launches, sum of groups %.3f ms, ideal (per-op max(flops/peak, bytes/hbm)) %.3f ms" % (which, B, len(ent["plan"]), tot, sum(r[5] for r in rows)))
for ms, k, n, fl, by, ideal in rows:
    print("%-58s %8.3f ms %4d x %8.1f us  %6.0f TF/s %6.0f GB/s  ideal %7.3f ms  eff %.2f" % (k, ms, n, 1e3 * ms / n, fl / ms / 1e9, by / ms / 1e6, ideal, ideal / ms))
