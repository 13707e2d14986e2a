#!/bin/bash
# round-2 call 15: A-slot planner knob (S2V_MIN_ASLOTS) on the FFC layer micro-benchmarks and on the real plans
for v in "0 0" "3 0" "4 0" "3 1" "4 1"; do set -- $v
  echo "== MIN_ASLOTS=$1 B1=$2"
  S2V_MIN_ASLOTS=$1 S2V_MIN_ASLOTS_B1=$2 MB_GRAPH=1 python tools/mb_layers.py res 2>&1 | grep -v "conv_tc:" | grep "res0.all+narrow+stats\|res1\|res0.st1\|res0.fu"
  S2V_MIN_ASLOTS=$1 S2V_MIN_ASLOTS_B1=$2 python tools/plan_breakdown.py lnet 2>/dev/null | head -1
  S2V_MIN_ASLOTS=$1 S2V_MIN_ASLOTS_B1=$2 python tools/plan_breakdown.py dnet 2>/dev/null | head -1
done
