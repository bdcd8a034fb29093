#!/bin/bash
# nearest k=1 (50M x 5M, sorted): device-resident kernel time per variant library / thread count
for v in "" $@; do
    lib=""; [ -n "$v" ] && lib="$PWD/variants/$v.so"
    BRG_LIB=$lib python bench.py --workload nearest --steps 10 --no-extras --no-cpu --no-e2e --no-parity 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nearest', '${v:-default}', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))
"
done
