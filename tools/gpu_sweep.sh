#!/bin/sh
# Developer tool (GPU box): sweep one environment variable over values on the default workload, two rounds.
#   gpurun -- sh tools/gpu_sweep.sh CPB200_PACK_CTAS "1 2 4" [extra env assignment]
V=$1; VALS=$2; [ -n "$3" ] && export $3
for k in 1 2; do
  for x in $VALS; do
    export $V=$x
    python bench.py --steps 30 --warmup 5 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V', '$x', '$3', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],3))"
  done
done
