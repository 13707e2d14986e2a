import os, sys, torch
sys.path.insert(0, "/root/repo")
import s2v_b200
from s2v_b200 import _lib as L, ops
lib = L.require_device(0)
torch.manual_seed(0)
M, K, N = 512, 128, 64
x = torch.randn(1, 1, M, K, device="cuda").half()
w = torch.randn(N, K, 1, 1, device="cuda") * 0.1
ref = (x.float().reshape(M, K) @ w.half().float().reshape(N, K).t())
for mode in (0, 1):
    for sft in (8, 16, 1, 2, 3, 4, 5, 7, 9, 13):
        os.environ["S2V_DBG_SHIFT"] = str(sft); os.environ["S2V_DBG_MODE"] = str(mode)
        y = torch.zeros(1, 1, M, N, device="cuda", dtype=torch.float16)
        ops.op_conv(lib, x, ops.pack_w_tc(w), y, box=(128, 1, 1)).run()
        torch.cuda.synchronize()
        err = (y.float().reshape(M, N) - ref).abs().max().item()
        print("mode", mode, "shift", sft, "max err %.4f" % err, "OK" if err < 0.02 else "BAD")
