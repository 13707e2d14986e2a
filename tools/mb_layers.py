"""Development micro-benchmark: the LNet FFC layer shapes at B=128 through s2v_conv_tc (us/launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
torch.manual_seed(0)
B = int(os.environ.get("MB_B", "128"))
only = sys.argv[1] if len(sys.argv) > 1 else ""


def bench(name, x, w, y, reps=int(os.environ.get('MB_REPS', '20')), **kw):
    if only and only not in name:
        return
    op = ops.op_conv(lib, x, w, y, name=name, **kw)
    for _ in range(int(os.environ.get('MB_WARM', '3'))):
        op.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("MB_GRAPH", "0") == "1":      # GPU-side cost only: the launches are replayed from a CUDA graph
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                op.run()
        g.replay(); torch.cuda.synchronize()
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    else:
        a.record()
        for _ in range(reps):
            op.run()
        b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    print("%-40s %8.1f us  %7.1f TFLOP/s" % (name, us, op.alg_flops / us / 1e6), flush=True)


def level(tag, S, C):
    cg = C * 3 // 4; cl = C - cg; ch = cg // 2
    xp = torch.randn(B, S + 2, S + 2, C, device="cuda").half()
    R = torch.empty(B, S, S, C, device="cuda", dtype=torch.float16)
    s2 = torch.randn(B, S, S, ch, device="cuda").half()
    wl = ops.pack_w_tc(torch.randn(cl, C, 3, 3, device="cuda") * 0.02)
    wg = torch.cat([ops.pack_w_tc(torch.randn(cg, cl, 3, 3, device="cuda") * 0.02), ops.pack_w_tc(torch.randn(cg, ch, 1, 1, device="cuda") * 0.02)], 1).contiguous()
    bench(f"{tag}.to_l", xp, wl, R[..., :cl], k=(3, 3))
    bench(f"{tag}.l2g+st2", xp[..., :cl], wg, R[..., cl:], k=(3, 3), x2=s2)
    s1 = torch.empty(B, S, S, ch, device="cuda", dtype=torch.float16)
    w1 = ops.pack_w_tc(torch.randn(ch, cg, 1, 1, device="cuda") * 0.02)
    sc, bi = torch.rand(ch, device="cuda"), torch.rand(ch, device="cuda")
    bench(f"{tag}.st1", xp[:, 1:-1, 1:-1, cl:], w1, s1, scale=sc, bias=bi, act=L.ACT_RELU)
    F1 = torch.randn(1, 1, B * S * (S // 2 + 1), cg, device="cuda").half()
    F2 = torch.empty_like(F1)
    wf = ops.pack_w_tc(torch.randn(cg, cg, 1, 1, device="cuda") * 0.02)
    sc2, bi2 = torch.rand(cg, device="cuda"), torch.rand(cg, device="cuda")
    bench(f"{tag}.fu", F1, wf, F2, scale=sc2, bias=bi2, act=L.ACT_RELU)


level("res0", 48, 128)
level("res1", 24, 256)
level("res2", 12, 1024)
x = torch.randn(B, 96, 96, 64, device="cuda").half()
wf = ops.pack_w_tc(torch.randn(3, 64, 7, 7, device="cuda") * 0.02)
yf = torch.empty(B, 3, 96, 96, device="cuda")
bench("final7x7", x, wf, None, k=(7, 7), pad=(3, 3), act=L.ACT_SIGMOID, y_f32=yf, out_shape=(B, 3, 96, 96))
xr = torch.randn(B, 12, 12, 256, device="cuda").half()
X = torch.randn(B, 14, 14, 1024, device="cuda").half()[:, 1:-1, 1:-1, :512]
wo = ops.pack_w_tc(torch.randn(512, 256, 1, 1, device="cuda") * 0.02)
bench("ca.out+res2", xr, wo, X, res2=X, bias=torch.rand(512, device="cuda"))
xd = torch.randn(B, 96, 96, 64, device="cuda").half()
wd = ops.pack_w_tc(torch.randn(128, 64, 3, 3, device="cuda") * 0.02)
yd = torch.empty(B, 96, 96, 128, device="cuda", dtype=torch.float16)
bench("enc.down0", xd, wd, yd, k=(3, 3), pad=(1, 1), bias=torch.rand(128, device="cuda"))


def level_all(tag, S, C):
    cg = C * 3 // 4; cl = C - cg; ch = cg // 2
    xp = torch.randn(B, S + 2, S + 2, C, device="cuda").half()
    R = torch.empty(B, S, S, C, device="cuda", dtype=torch.float16)
    s2 = torch.randn(B, S, S, ch, device="cuda").half()
    wa = torch.cat([ops.pack_w_tc(torch.randn(C, C, 3, 3, device="cuda") * 0.02), ops.pack_w_tc(torch.randn(C, ch, 1, 1, device="cuda") * 0.02)], 1).contiguous()
    bench(f"{tag}.all", xp, wa, R, k=(3, 3), x2=s2)
    tiles = ops.box_tiles(S, S, B, (3, 3))
    partial = torch.zeros(B, tiles, C, 2, device="cuda")
    bench(f"{tag}.all+stats", xp, wa, R, k=(3, 3), x2=s2, stats=(partial, 0, 0))
    nfrom = -(-cl // 64) * 64           # the net's configuration: structural zero block hint (models/LNet.py:254-257)
    if nfrom < C and cl % 32 == 0:
        bench(f"{tag}.all+narrow", xp, wa, R, k=(3, 3), x2=s2, narrow=(nfrom, cl))
        bench(f"{tag}.all+narrow+stats", xp, wa, R, k=(3, 3), x2=s2, narrow=(nfrom, cl), stats=(partial, 0, 0))
        tot = torch.zeros(B, tiles, 4, 2, device="cuda")
        bench(f"{tag}.all+narrow+totals", xp, wa, R, k=(3, 3), x2=s2, narrow=(nfrom, cl), stats=(tot, 0, 0, "totals"))


level_all("res0", 48, 128)
