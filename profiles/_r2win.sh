mkdir -p gpurun_out
BRG_COUNT_WIN=1 BRG_COV_WIN=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_session_api.py tests/test_gpu_ingest.py -x -q -m gpu > gpurun_out/r2win_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2win_tests.log; tail -3 gpurun_out/r2win_tests.log
BRG_COUNT_WIN=1 BRG_COV_WIN=1 timeout 900 python -m pytest tests/test_gpu_full_configs.py -x -q -m gpu -k "coverage or count" >> gpurun_out/r2win_tests.log 2>&1; echo "full exit $?" >> gpurun_out/r2win_tests.log; tail -2 gpurun_out/r2win_tests.log
for w in count coverage; do for v in 0 1; do
BRG_COUNT_WIN=$v BRG_COV_WIN=$v timeout 300 python bench.py --workload $w --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w WIN=$v', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
done; done
