mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_join.py -x -q -m gpu > gpurun_out/near_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/near_tests.log; tail -3 gpurun_out/near_tests.log
for lib in "" variants/nw64.so variants/nw_c3.so; do
BRG_LIB=$lib timeout 300 python bench.py --workload nearest --steps 10 --no-extras --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nearest lib=$lib', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'parity', d.get('parity_vs_oracle_sample'))
"
done
timeout 300 python profiles/exp_join.py --no-chunked --sorted-only --check | cut -c1-330
