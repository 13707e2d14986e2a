"""Development micro-benchmark: the cross-attention core at LNet's B=128 shape (us / launch); S2V_ATTN = 0 tcgen05, 1 mma.sync, 2 SIMT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import _lib as L, ops

lib = L.require_device(0)
torch.manual_seed(5)
n, t, heads, dh = int(os.environ.get("MB_B", "128")), 144, 4, 64
qk = torch.randn(n, 1, t, 2 * heads * dh, device="cuda").half()
v = torch.randn(n, 1, t, heads * dh, device="cuda").half()
o = torch.empty(n, 1, t, heads * dh, dtype=torch.float16, device="cuda")
op = ops.op_attention(lib, qk[..., :256], qk[..., 256:], v, o, heads, dh ** -0.5)
for _ in range(3):
    op.run()
torch.cuda.synchronize()
q, k = (qk[..., i * 256:(i + 1) * 256].float().reshape(n, t, heads, dh).permute(0, 2, 1, 3) for i in (0, 1))
vv = v.float().reshape(n, t, heads, dh).permute(0, 2, 1, 3)
ref = ((torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5).softmax(-1) @ vv).permute(0, 2, 1, 3).reshape(n, 1, t, heads * dh)
err = (o.float() - ref).abs().max().item()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    op.run()
b.record(); torch.cuda.synchronize()
print("S2V_ATTN=%s  %.1f us / launch   max_abs err %.2e" % (os.environ.get("S2V_ATTN", "0"), a.elapsed_time(b) * 50, err))
