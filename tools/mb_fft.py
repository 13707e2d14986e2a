"""Development micro-benchmark: the FourierUnit transforms at LNet's B=128 shapes (us / launch, GB/s algorithmic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
B = int(os.environ.get("MB_B", "128"))
reps, warm = int(os.environ.get("MB_REPS", "20")), int(os.environ.get("MB_WARM", "3"))
for S, ch in ((48, 48), (24, 96), (12, 384)):
    x = torch.randn(B, S, S, ch, device="cuda").half()
    F = torch.empty(B, S, S // 2 + 1, 2 * ch, device="cuda", dtype=torch.float16)
    y = torch.empty_like(x)
    for name, op in (("rfft2", ops.op_rfft2(lib, x, F)), ("irfft2", ops.op_irfft2(lib, F, x, y))):
        for _ in range(warm):
            op.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get("MB_GRAPH", "1") == "1":      # device time only: the launches are replayed from a CUDA graph
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
        print("%-7s S=%d ch=%d  %7.1f us  %6.0f GB/s" % (name, S, ch, us, op.alg_bytes / us / 1e3), flush=True)
