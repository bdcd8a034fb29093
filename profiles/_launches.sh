mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_count.csv python bench.py --no-cpu --no-extras --no-parity --steps 5 --warmup 3 > gpurun_out/ncu_b1.log 2>&1; echo "rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_pairs.csv python bench.py --workload pairs --no-cpu --no-extras --no-parity --no-e2e --steps 5 --warmup 3 > gpurun_out/ncu_b2.log 2>&1; echo "rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_nearest.csv python bench.py --workload nearest --no-cpu --no-extras --no-parity --no-e2e --steps 5 --warmup 3 > gpurun_out/ncu_b3.log 2>&1; echo "rc $?"
wc -l gpurun_out/r2_launches_*.csv
