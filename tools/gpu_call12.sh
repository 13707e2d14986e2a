#!/bin/bash
mkdir -p gpurun_out
for db in 64 128 192; do
  timeout 600 python bench.py --dnet-batch $db --no-extras --no-cpu-baseline --no-classes --steps 5 > gpurun_out/r2n_bench_db$db.json 2> gpurun_out/r2n_bench_db$db.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_db$db.json')); print('dnet_batch $db', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, s2v_b200
from oracle import synth, weights
from s2v_b200.models.DNet import DNet
net = DNet().cuda().eval(); net.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
for b in (64, 128):
    s, c = synth.dnet_inputs(b, 0)
    net(s.cuda(), c.cuda())
    print(b, net.engine().plan_cache_info()["bytes"] / 1e9, "GB cumulative")
PY
