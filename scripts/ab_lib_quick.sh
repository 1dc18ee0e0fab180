#!/bin/bash
# same-box A/B of two builds of the library on the training step only: scripts/ab_lib_quick.sh <libA> <libB> [rounds] [RN_FLAGS]
A=$1; B=$2; R=${3:-2}; F=${4:-}
for i in $(seq 1 $R); do
  for L in $A $B; do
    RN_FLAGS="$F" RN_B200_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.read()); pm = d["roofline"]["per_mode"]
print(sys.argv[1], round(d["ms_per_step"], 3), "eager", round(d["config"]["eager_ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in pm.items()})' "$L"
  done
done
