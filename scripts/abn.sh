#!/bin/bash
# timing of N builds of librnerf_b200.so on the same GPU box, interleaved: scripts/abn.sh <rounds> <lib>...
R=$1; shift
for i in $(seq 1 $R); do
  for L in "$@"; do
    RN_B200_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pm=d['roofline']['per_mode']
print('$L', round(d['ms_per_step'],3), 'fwd', round(pm['nt_forward']['ms_per_step'],3), 'dgrad', round(pm['nn_dgrad']['ms_per_step'],3), 'wgrad', round(pm['tn_wgrad']['ms_per_step'],3), 'render', round(d['extra']['render_mrays_per_s_1gpu'],3))"
  done
done
