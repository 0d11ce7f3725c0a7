#!/bin/sh
# Developer tool (GPU box): A/B of an environment switch on the default workload, alternating runs on one box.
#   gpurun -- sh tools/gpu_ab.sh CPB200_REFIT_UNFUSED [workload]
V=$1; W=${2:-pile1m}
for k in 1 2 3; do
  for on in 0 1; do
    if [ $on = 1 ]; then export $V=1; else unset $V; fi
    python bench.py --steps 30 --warmup 5 --no-sub --workload $W 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V', $on, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],3))"
  done
done
