#!/bin/bash
# A/B of rn_set_flag variants of the SAME library on one GPU box: scripts/ab_flags.sh "<flags A>" "<flags B>" [rounds]
A=$1; B=$2; R=${3:-2}
for i in $(seq 1 $R); do
  for F in "$A" "$B"; do
    RN_FLAGS="$F" python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pm=d['roofline']['per_mode']
print('flags[$F]', round(d['ms_per_step'],3), 'fwd', round(pm['nt_forward']['ms_per_step'],3), 'dgrad', round(pm['nn_dgrad']['ms_per_step'],3), 'wgrad', round(pm['tn_wgrad']['ms_per_step'],3), 'render', round(d['render']['value'],3), 'pose_opt_ms', round(d['pose_opt']['ms_per_step'],3))"
  done
done
