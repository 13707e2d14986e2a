"""Development tool: LNet / DNet forward time vs batch size (graph replay, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth, weights
from s2v_b200.models.DNet import DNet
from s2v_b200.models.LNet import LNet
dev = torch.device("cuda", 0)
lnet = LNet().to(dev).eval(); lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
dnet = DNet().to(dev).eval(); dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)


def t(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    for B in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "25,32,60,64,89,96,120,128").split(",")]:
        mel, face = synth.lnet_inputs(B, seed=0)
        mel, face = mel.to(dev), face.to(dev)
        ms = t(lambda: lnet(mel, face))
        bd = min(B, int(os.environ.get("MB_DNET_MAX", "64")))
        src, coeff = synth.dnet_inputs(bd, seed=0)
        src, coeff = src.to(dev), coeff.to(dev)
        md = t(lambda: dnet(src, coeff))
        print("B=%3d  LNet %.2f ms (%.1f us/frame)   DNet(B=%d) %.2f ms (%.1f us/frame)" % (B, ms, 1e3 * ms / B, bd, md, 1e3 * md / bd), flush=True)
