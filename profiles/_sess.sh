for lib in "" variants/nw_nopf.so; do
BRG_LIB=$lib timeout 300 python bench.py --workload nearest --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nearest lib=$lib', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
done
