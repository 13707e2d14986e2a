"""Development micro-benchmark: affine_act on DNet's dense shapes (S2V_AFFINE_FLAT = 0 / 1), graph-timed, L2 flushed by the size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops
lib = L.require_device(0)
B = int(os.environ.get("MB_B", "64"))


def t(fn, reps=10):
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


for (S, C) in ((256, 64), (128, 128), (64, 256)):
    x = torch.randn(B, S, S, C, device="cuda").half()
    r = torch.randn(B, S, S, C, device="cuda").half()
    y = torch.empty_like(x)
    a, b, ra, rb = (torch.rand(B, C, device="cuda") for _ in range(4))
    mb = x.numel() * 2 / 1e6
    o1 = ops.op_affine_act(lib, x, a, b, y, act=L.ACT_LRELU, act_param=0.1)
    o2 = ops.op_affine_act(lib, x, a, b, y, act=L.ACT_NONE, res=r)
    o3 = ops.op_affine_act(lib, x, a, b, y, act=L.ACT_LRELU, act_param=0.1, res=r, res_ab=(ra, rb))
    u1, u2, u3 = t(o1.run), t(o2.run), t(o3.run)
    print("%dx%dx%d B=%d %6.0f MB | x->y %6.1f us (%.2f TB/s) | x+res->y %6.1f us (%.2f TB/s) | affine2 %6.1f us (%.2f TB/s)"
          % (S, S, C, B, mb, u1, 2 * mb / u1, u2, 3 * mb / u2, u3, 3 * mb / u3))
