"""Development micro-benchmark: DNet's large 3x3 layers at B=64 through s2v_conv_tc (us/launch); run with S2V_CTA2 / S2V_TC_DEBUG."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
torch.manual_seed(0)
B = 64


def bench(name, x, w, y, reps=10, **kw):
    op = ops.op_conv(lib, x, w, y, name=name, **kw)
    for _ in range(2):
        op.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            op.run()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    print("%-40s %8.1f us  %7.1f TFLOP/s  %6.0f GB/s min-bytes" % (name, us, op.alg_flops / us / 1e6, op.io_bytes / us / 1e3), flush=True)


for (s, ci, co) in ((256, 64, 128), (256, 64, 64), (128, 128, 256), (128, 128, 128), (64, 256, 256)):
    x = torch.randn(B, s, s, ci, device="cuda").half()
    w = ops.pack_w_tc(torch.randn(co, ci, 3, 3, device="cuda") * 0.03)
    y = torch.empty(B, s, s, co, device="cuda", dtype=torch.float16)
    bench("3x3 %d->%d @%d" % (ci, co, s), x, w, y, k=(3, 3), pad=(1, 1))
    del x, w, y
