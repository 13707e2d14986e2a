#!/bin/bash
# where does the time of the epilogue-heavy LNet FFC GEMMs go?  S2V_EXP: 1 = epilogue does nothing, 2 = no MMAs, 4 = no stats walk
mkdir -p gpurun_out
for e in 0 1 2 3 4; do
  echo "== S2V_EXP=$e"
  S2V_EXP=$e MB_GRAPH=1 python tools/mb_layers.py res 2>&1 | grep -v "^$"
done > gpurun_out/r2r_exp_layers.txt 2>&1
cat gpurun_out/r2r_exp_layers.txt
