#!/bin/bash
# kernel time of one bench workload under an environment switch:  exp_variants.sh <workload> <ENV_NAME> <values...>
w=$1; name=$2; shift 2
for v in "$@"; do
  env $name=$v python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', '$name=$v', 'kernel_ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))
"
done
