#!/bin/bash
mkdir -p gpurun_out
python tools/mb_overlap.py > gpurun_out/r2f_overlap.txt 2>&1
S2V_PDL=0 python tools/mb_overlap.py >> gpurun_out/r2f_overlap.txt 2>&1
cat gpurun_out/r2f_overlap.txt | tail -3
S2V_TC_DEBUG=1 python - > gpurun_out/r2f_tc_debug_lnet.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import torch, s2v_b200
from oracle import synth, weights
from s2v_b200.models.LNet import LNet
net = LNet().cuda().eval(); net.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
eng = net.engine(); eng.use_graph = False
ent = eng.plan_for(128)
for op in ent["plan"].ops:
    if op.name.endswith("[tc]"):
        print("##", op.name, flush=True); sys.stdout.flush()
        op.run(); torch.cuda.synchronize()
PY
S2V_TC_DEBUG=1 python - > gpurun_out/r2f_tc_debug_dnet.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import torch, s2v_b200
from oracle import synth, weights
from s2v_b200.models.DNet import DNet
net = DNet().cuda().eval(); net.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
eng = net.engine(); eng.use_graph = False
s, c = synth.dnet_inputs(64, 0)
ent = eng._get_plan((64, 26, "full"), eng._build(64, 26, "full"))
for op in ent["plan"].ops:
    if op.name.endswith("[tc]"):
        print("##", op.name, flush=True); sys.stdout.flush()
        op.run(); torch.cuda.synchronize()
PY
grep -c conv_tc gpurun_out/r2f_tc_debug_lnet.txt gpurun_out/r2f_tc_debug_dnet.txt
