#!/bin/bash
# headline kernel time as a function of the CTA size of count_kernel_tiled (2 CTAs per SM; registers per thread follow)
for t in 256 320; do
  BRG_COUNT_THREADS=$t python bench.py --steps 10 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('count_threads', $t, 'kernel_ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))
"
done
