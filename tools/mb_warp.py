"""Development micro-benchmark of the fused flow-warp kernel (L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth
from s2v_b200.futils import flow_util
dev = torch.device("cuda", 0)
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


from s2v_b200 import _lib as L, ops
lib = L.require_device(0)


def t(fn, reps=10):
    """fn is captured into a CUDA graph (reps copies) so that host launch overhead cannot hide in the number"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for sigma in (3.0, 1.0, 0.0):
    s, fl = synth.warp_inputs(64, seed=0)
    s, fl = s.to(dev), (fl * (sigma / 3.0)).to(dev)
    out = torch.empty_like(s)
    op = ops.op_flow_warp(lib, s, fl, out)
    both = t(lambda: (scratch.zero_(), op.run()))
    flush = t(lambda: scratch.zero_())
    ms = both - flush
    print("flow sigma %.1f: %.1f us  %.0f GB/s algorithmic (flushed L2)" % (sigma, ms * 1e3, 64 * 1605632 / 1e9 / (ms * 1e-3)))
    hot = t(lambda: op.run())
    print("            hot L2: %.1f us  %.0f GB/s" % (hot * 1e3, 64 * 1605632 / 1e9 / (hot * 1e-3)))
