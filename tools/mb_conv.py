"""Micro-benchmark of s2v_conv_tc on single layer shapes (development tool; prints us/launch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
torch.manual_seed(0)


def bench(name, x, w, y, reps=30, **kw):
    op = ops.op_conv(lib, x, w, y, name=name, **kw)
    for _ in range(3):
        op.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            op.run()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    print("%-46s %8.1f us  %7.1f TFLOP/s" % (name, us, op.alg_flops / us / 1e6))
    return us


B = 128
# res2 st1: M=18432 K=768 N=384
xp = torch.randn(B, 14, 14, 1024, device="cuda").half()
wt = ops.pack_w_tc(torch.randn(384, 768, 1, 1, device="cuda") * 0.03)
y = torch.empty(B, 12, 12, 384, device="cuda", dtype=torch.float16)
sc, bi = torch.rand(384, device="cuda"), torch.rand(384, device="cuda")
bench("st1 res2 slice-view box448 +scale/bias/relu", xp[:, 1:-1, 1:-1, 256:], wt, y, scale=sc, bias=bi, act=L.ACT_RELU)
bench("st1 res2 slice-view box448 plain", xp[:, 1:-1, 1:-1, 256:], wt, y)
xc = torch.randn(B, 12, 12, 768, device="cuda").half()
bench("st1 res2 contiguous box448", xc, wt, y)
bench("st1 res2 contiguous flat", xc.reshape(1, 1, -1, 768), wt, y.reshape(1, 1, -1, 384))
# fu res2: M=10752 K=768 N=768 flat
xf = torch.randn(1, 1, B * 84, 768, device="cuda").half()
wf = ops.pack_w_tc(torch.randn(768, 768, 1, 1, device="cuda") * 0.03)
yf = torch.empty(1, 1, B * 84, 768, device="cuda", dtype=torch.float16)
bench("fu res2 flat N=768", xf, wf, yf)
# st2 res0: M=294912 K=48 N=96, + res2
x0 = torch.randn(B, 48, 48, 48, device="cuda").half()
w0 = ops.pack_w_tc(torch.randn(96, 48, 1, 1, device="cuda") * 0.1)
R = torch.randn(B, 48, 48, 128, device="cuda").half()
bench("st2 res0 K=48 N=96 res2 in-place", x0, w0, R[..., 32:], res2=R[..., 32:])
bench("st2 res0 K=48 N=96 no res", x0, w0, R[..., 32:])
y0 = torch.empty(B, 48, 48, 96, device="cuda", dtype=torch.float16)
bench("st2 res0 K=48 N=96 contiguous out", x0, w0, y0)
# 3x3 res2 to_l: K=9216 N=256
wl = ops.pack_w_tc(torch.randn(256, 1024, 3, 3, device="cuda") * 0.01)
yl = torch.empty(B, 12, 12, 256, device="cuda", dtype=torch.float16)
bench("to_l res2 3x3 K=9216 N=256", xp, wl, yl, k=(3, 3))
# encoder down0: 96x96 64->128
xe = torch.randn(B, 96, 96, 64, device="cuda").half()
we = ops.pack_w_tc(torch.randn(128, 64, 3, 3, device="cuda") * 0.04)
ye = torch.empty(B, 96, 96, 128, device="cuda", dtype=torch.float16)
bench("down0 3x3 64->128 @96", xe, we, ye, k=(3, 3), pad=(1, 1))
# big square GEMM for reference: M=16384 K=4096 N=4096?  (N<=... per launch N tiles of 256)
xg = torch.randn(1, 1, 16384, 4096, device="cuda").half()
wg = ops.pack_w_tc(torch.randn(4096, 4096, 1, 1, device="cuda") * 0.01)
yg = torch.empty(1, 1, 16384, 4096, device="cuda", dtype=torch.float16)
bench("gemm 16384x4096x4096", xg, wg, yg)
