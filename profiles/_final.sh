mkdir -p gpurun_out
timeout 330 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; echo "tests exit $?" >> gpurun_out/gputests.log; tail -4 gpurun_out/gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_final.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:nearest_k1_kernel_todo --launch-skip 7 --launch-count 2 -f -o gpurun_out/near_todo_bins python profiles/exp_bins.py nearest > gpurun_out/ncu_near_todo.log 2>&1; echo "ncu rc $?"
