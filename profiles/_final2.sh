mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; echo "tests exit $?" >> gpurun_out/gputests.log; tail -4 gpurun_out/gputests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 150 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_final.err
