#!/bin/bash
# A/B timing of two builds of librnerf_b200.so on the same GPU box: scripts/ab.sh <libA> <libB> [rounds]
A=$1; B=$2; R=${3:-2}
for i in $(seq 1 $R); do
  for L in $A $B; do
    RN_B200_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pm=d['roofline']['per_mode']
r=(d.get('render') or {}).get('value')
po=(d.get('pose_opt') or {}).get('ms_per_step')
print('$L', round(d['ms_per_step'],3), 'fwd', round(pm['nt_forward']['ms_per_step'],3), 'dgrad', round(pm['nn_dgrad']['ms_per_step'],3), 'wgrad', round(pm['tn_wgrad']['ms_per_step'],3), 'render', r and round(r,3), 'pose_opt_ms', po and round(po,3))"
  done
done
