for lib in "" variants/ld128.so; do for nb in 0 1000000; do BRG_LIB=$lib BUILD_ROWS=$nb timeout 300 python profiles/exp_shuffled.py 2>&1 | tail -1 | sed "s#^#lib=$lib #"; done; done
BRG_LIB=variants/nw_w7.so timeout 300 python bench.py --workload nearest --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nearest w7', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
