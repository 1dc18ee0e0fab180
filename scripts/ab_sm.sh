#!/bin/bash
# per-family GEMM times of the training step under rn_set_flag variants (no render / pose-opt blocks): scripts/ab_sm.sh "<flags>" ...
for F in "$@"; do
  RN_FLAGS="$F" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.read()); pm = d["roofline"]["per_mode"]
print("flags[%s]" % sys.argv[1], round(d["ms_per_step"], 3), "eager", round(d["config"]["eager_ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3),
      {k: round(v["ms_per_step"], 3) for k, v in pm.items()}, "loss", round(d["config"]["loss"], 6))' "$F"
done
