#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1 2 4; do
  for w in lnet dnet; do
    S2V_PF=$pf python tools/plan_breakdown.py $w > gpurun_out/r2j_bd_${w}_pf$pf.txt 2>&1
    echo "pf=$pf $(head -1 gpurun_out/r2j_bd_${w}_pf$pf.txt)"
  done
done
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_lnet.py tests/test_gpu_kernels.py tests/test_gpu_enet.py -q -m gpu --timeout 900 2>&1 | tail -4
python tools/plan_breakdown.py enet > gpurun_out/r2j_bd_enet.txt 2>&1; head -8 gpurun_out/r2j_bd_enet.txt
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, s2v_b200
from oracle import synth
from s2v_b200.futils import audio
for sec in (60.0, 600.0):
    wav = torch.from_numpy(synth.wav(sec, 0)).cuda()
    for _ in range(3): audio.melspectrogram_device(wav)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): audio.melspectrogram_device(wav)
    b.record(); torch.cuda.synchronize()
    print("mel %ds: %.1f us" % (sec, a.elapsed_time(b)*100))
PY
