#!/bin/bash
# e2e (host buffers -> H2D -> kernel -> D2H) throughput of the headline workload as a function of the staging chunk size
for c in 524288 1048576 2097152 4194304 8388608 16777216; do
  BRG_CHUNK_ROWS=$c python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk_rows', $c, 'e2e_ms', round(d['e2e']['ms_per_step'], 3), 'e2e_probes_per_s', '%.3e' % d['e2e']['value'], 'kernel_ms', round(d['ms_per_step'], 4))
"
done
