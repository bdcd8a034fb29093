mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_session_api.py -x -q -m gpu > gpurun_out/near_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/near_tests.log; tail -3 gpurun_out/near_tests.log
timeout 600 python -m pytest tests/test_gpu_full_configs.py -x -q -m gpu -k nearest >> gpurun_out/near_tests.log 2>&1; echo "full exit $?" >> gpurun_out/near_tests.log; tail -2 gpurun_out/near_tests.log
timeout 300 python bench.py --workload nearest --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nearest', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
BRG_COUNT_WIN=1 timeout 300 python bench.py --workload count --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('count WIN', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
