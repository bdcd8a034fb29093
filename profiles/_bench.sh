mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 300 python profiles/exp_build_time.py 2>&1 | tail -6
