#!/usr/bin/env python
"""Does the hardware overlap the LNet graph (one stream) with the DNet graph (another stream)?  Times K x (LNet B=128 + 2 x DNet B=64)
sequentially on one stream and concurrently on two (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import s2v_b200  # noqa
from oracle import synth, weights
from s2v_b200.models.DNet import DNet
from s2v_b200.models.LNet import LNet

dev = torch.device("cuda", 0)
lnet = LNet().to(dev).eval(); lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
dnet = DNet().to(dev).eval(); dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
mel, face = synth.lnet_inputs(128, 0)
src, co = synth.dnet_inputs(64, 0)
lnet(mel.to(dev), face.to(dev)); dnet(src.to(dev), co.to(dev))
lnet(mel.to(dev), face.to(dev)); dnet(src.to(dev), co.to(dev))
le, de = lnet.engine(), dnet.engine()
lent, dent = le._plans[128], de._plans[(64, 26, "full")]
K = 10
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def timed(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


def seq():
    for _ in range(K):
        lent["graph"].replay(); dent["graph"].replay(); dent["graph"].replay()


def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        for _ in range(K):
            lent["graph"].replay()
    with torch.cuda.stream(s2):
        for _ in range(2 * K):
            dent["graph"].replay()
    cur.wait_stream(s1); cur.wait_stream(s2)


def only_l():
    for _ in range(K):
        lent["graph"].replay()


def only_d():
    for _ in range(2 * K):
        dent["graph"].replay()


tl, td, ts, tc = timed(only_l), timed(only_d), timed(seq), timed(conc)
print("LNet x%d %.2f ms | DNet x%d %.2f ms | sequential %.2f ms | two streams %.2f ms (%.3f of sequential)" % (K, tl, 2 * K, td, ts, tc, tc / ts))
