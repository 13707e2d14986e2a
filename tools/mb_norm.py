"""Development micro-benchmark of the statistics / apply kernels on the LNet shapes (graph-timed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops
lib = L.require_device(0)
B = 128


def t(fn, reps=20):
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
    return a.elapsed_time(b) / reps * 1e3


for (S, C, name) in ((12, 1024, "res2"), (24, 256, "res1"), (48, 128, "res0"), (96, 64, "ln96x64"), (96, 128, "ln96x128"), (48, 256, "ln48x256")):
    x = torch.randn(B, S, S, C, device="cuda").half()
    mb = x.numel() * 2 / 1e6
    chunks = ops.stats_chunks(B, S * S, C)
    partial = torch.empty(B, chunks, C, 2, device="cuda")
    a, b = torch.rand(B, C, device="cuda"), torch.rand(B, C, device="cuda")
    gb = torch.randn(B, 2 * C, device="cuda")
    yp = torch.zeros(B, S + 2, S + 2, C, device="cuda", dtype=torch.float16)
    res = torch.randn(B, S, S, C, device="cuda").half()
    o1 = ops.op_chan_stats(lib, x, chunks, partial)
    o2 = ops.op_adain_finalize(lib, partial, B, chunks, C, S * S, gb[:, :C], gb[:, C:], gb.stride(0), a, b)
    o3 = ops.op_affine_act(lib, x, a, b, yp[:, 1:-1, 1:-1, :], act=L.ACT_LRELU, act_param=0.01, res=res, reflect1=1)
    o4 = ops.op_affine_act(lib, x, a, b, yp[:, 1:-1, 1:-1, :], act=L.ACT_LRELU, act_param=0.01, reflect1=1)
    u1, u2, u3, u4 = t(o1.run), t(o2.run), t(o3.run), t(o4.run)
    print("%-9s %6.1f MB chunks=%2d | stats %6.1f us (%.2f TB/s) | finalize %5.1f us | affine+res+reflect %6.1f us (%.2f TB/s) | affine %6.1f us (%.2f TB/s)"
          % (name, mb, chunks, u1, mb / u1, u2, u3, 3 * mb / u3, u4, 2 * mb / u4))
