import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth
from s2v_b200 import _lib as L, ops
lib = L.require_device(0)
s, fl = synth.warp_inputs(64, seed=0)
s, fl = s.cuda(), fl.cuda()
out = torch.empty_like(s)
op = ops.op_flow_warp(lib, s, fl, out)
for _ in range(5):
    op.run()
torch.cuda.synchronize()
print("ok")
