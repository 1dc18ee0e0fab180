#!/bin/bash
# data-parallel A/B on N GPUs: scripts/ab_dp.sh N "ENV=.. ENV=.." "ENV=.." ...   (each argument = one variant's environment)
N=$1; shift
port=29600
for V in "$@"; do
  port=$((port+1))
  env $V python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.read())
print("variant[%s]" % sys.argv[1], "N", d["n_gpus"], "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"], 3), "eager", round(d["config"]["eager_ms_per_step"], 3), "mhz", d["clocks"]["sm_mhz"])' "$V"
done
