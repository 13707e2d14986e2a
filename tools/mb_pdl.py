"""Development micro-benchmark: per-launch cost of a chain of tiny dependent kernels replayed from a CUDA graph,
with and without the programmatic-dependent-launch attribute (S2V_PDL=0/1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
for shape in ((4, 12, 12, 64), (128, 48, 48, 128)):
    a = torch.randn(*shape, device="cuda").half()
    b = torch.randn_like(a)
    y = torch.empty_like(a)
    sc, sh = torch.ones(shape[0], shape[3], device="cuda"), torch.zeros(shape[0], shape[3], device="cuda")
    op = ops.op_affine_act(lib, a, sc, sh, y, act=L.ACT_LRELU, act_param=0.1)
    for _ in range(3):
        op.run()
    torch.cuda.synchronize()
    reps = 200
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            op.run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("affine_act %s: %.2f us per launch (graph of %d), PDL=%s" % (shape, e0.elapsed_time(e1) * 1e3 / reps, reps, os.environ.get("S2V_PDL", "1")))
