#!/bin/bash
mkdir -p gpurun_out
for c in "" 1; do echo "== S2V_CTA2=$c"; S2V_CTA2=$c python tools/mb_dnet_layers.py 2>&1; done | tee gpurun_out/r2y_cta2.txt
