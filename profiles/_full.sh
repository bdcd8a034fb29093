mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; echo "tests exit $?" >> gpurun_out/gputests.log; tail -4 gpurun_out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', d.get('parity_vs_oracle_sample'))
        for k,v in d['extras'].items():
            if isinstance(v,dict) and 'ms_per_step' in v: print(k, round(v['ms_per_step'],4), round(v.get('roofline_frac', v.get('roofline',{}).get('frac',0)),3), v.get('parity_all'), v.get('pairs_path',''), (v.get('e2e') or {}).get('ms_per_step'))
            else: print(k, str(v)[:200])
PY
