"""Development: per-phase clock breakdown of the conv_tc epilogue / MMA warps (needs the S2V_EPI_PROF library variant:
python tools/build_variant.py prof -DS2V_EPI_PROF; S2V_LIB=speech-to-video-mpp_b200/libs2v_prof.so python tools/mb_epi_prof.py)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
prof = lib.s2v_dbg_epi_prof
prof.restype, prof.argtypes = C.c_int, (C.POINTER(C.c_ulonglong), C.c_int)
torch.manual_seed(0)
EPI = ["wait tfull", "wait store-read + bar", "TMEM->cvt->smem", "tempty arrive", "fence + bar", "TMA store issue (+stats to end)"]


def run(name, x, w, y, reps=10, **kw):
    op = ops.op_conv(lib, x, w, y, name=name, **kw)
    for _ in range(2):
        op.run()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    prof(buf, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        op.run()
    b.record(); torch.cuda.synchronize()
    prof(buf, 1)
    us = a.elapsed_time(b) * 1e3 / reps
    nt, nm = max(buf[8], 1), max(buf[12], 1)
    print("%-34s %7.1f us/launch | epilogue thread: %d tiles/launch, clk per tile:" % (name, us, nt // reps) +
          " ".join(" [%s] %d" % (EPI[i], buf[i] // nt) for i in range(6)))
    print("%-34s                   | MMA warp: %d tiles/launch, clk per tile: [wait tempty] %d  [wait A] %d  [issue+commit] %d" %
          ("", nm // reps, buf[9] // nm, buf[10] // nm, buf[11] // nm) +
          ("  | bulk store issue -> staging read: %d clk" % (buf[13] // buf[14]) if buf[14] else ""), flush=True)


B = 128
S, Cc = 48, 128
cg = Cc * 3 // 4; cl = Cc - cg; ch = cg // 2
xp = torch.randn(B, S + 2, S + 2, Cc, device="cuda").half()
R = torch.empty(B, S, S, Cc, device="cuda", dtype=torch.float16)
s2 = torch.randn(B, S, S, ch, device="cuda").half()
wa = torch.cat([ops.pack_w_tc(torch.randn(Cc, Cc, 3, 3, device="cuda") * 0.02), ops.pack_w_tc(torch.randn(Cc, ch, 1, 1, device="cuda") * 0.02)], 1).contiguous()
tiles = ops.box_tiles(S, S, B, (3, 3))
partial = torch.zeros(B, tiles, Cc, 2, device="cuda")
nfrom = -(-cl // 64) * 64
run("res0.all+narrow+stats", xp, wa, R, k=(3, 3), x2=s2, narrow=(nfrom, cl), stats=(partial, 0, 0))
run("res0.all+narrow", xp, wa, R, k=(3, 3), x2=s2, narrow=(nfrom, cl))
S, Cc = 24, 256
cg = Cc * 3 // 4; cl = Cc - cg; ch = cg // 2
xp = torch.randn(B, S + 2, S + 2, Cc, device="cuda").half()
R = torch.empty(B, S, S, Cc, device="cuda", dtype=torch.float16)
s2 = torch.randn(B, S, S, ch, device="cuda").half()
wg = torch.cat([ops.pack_w_tc(torch.randn(cg, cl, 3, 3, device="cuda") * 0.02), ops.pack_w_tc(torch.randn(cg, ch, 1, 1, device="cuda") * 0.02)], 1).contiguous()
run("res1.l2g+st2", xp[..., :cl], wg, R[..., cl:], k=(3, 3), x2=s2)
del xp, R, s2
B = 64
for (s, ci, co) in ((256, 64, 128), (256, 64, 64), (128, 128, 256)):
    x = torch.randn(B, s, s, ci, device="cuda").half()
    w = ops.pack_w_tc(torch.randn(co, ci, 3, 3, device="cuda") * 0.03)
    y = torch.empty(B, s, s, co, device="cuda", dtype=torch.float16)
    tl = ops.box_tiles(s, s, B, (3, 3))
    tot = torch.zeros(B, tl, 4, 2, device="cuda")
    run("dnet 3x3 %d->%d @%d totals" % (ci, co, s), x, w, y, k=(3, 3), pad=(1, 1), stats=(tot, 0, 0, "totals"),
        bias=torch.rand(co, device="cuda"))
    run("dnet 3x3 %d->%d @%d plain" % (ci, co, s), x, w, y, k=(3, 3), pad=(1, 1))
    del x, w, y
