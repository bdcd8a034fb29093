for lib in variants/r2_44.so variants/r2_60.so; do for w in count coverage; do
BRG_LIB=$lib BRG_COUNT_WIN=1 BRG_COV_WIN=1 timeout 300 python bench.py --workload $w --steps 10 --no-extras --no-cpu --no-e2e --no-parity 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w $lib', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))
"
done; done
