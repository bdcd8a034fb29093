for t in 8 12 16; do
BRG_HOST_THREADS=$t timeout 600 python bench.py --no-extras --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('threads $t', 'e2e %.3g (%.1f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'pageable %.3g (%.1f ms)' % (d['e2e_pageable']['value'], d['e2e_pageable']['ms_per_step']), 'parity', d.get('parity_vs_oracle_sample'))
"
done
nproc
