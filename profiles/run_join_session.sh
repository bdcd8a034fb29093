#!/bin/bash
# one GPU session for the one-launch pair join: parity tests, then timings of the shipped library and of the variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join.py -x -q > gpurun_out/join_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/join_tests.log
tail -3 gpurun_out/join_tests.log
timeout 600 python profiles/exp_join.py --check --sorted-only > gpurun_out/join_exp.log 2>&1; echo "exit $?" >> gpurun_out/join_exp.log
for v in variants/fj_*.so; do
  BRG_LIB=$v timeout 300 python profiles/exp_join.py --no-chunked --sorted-only --no-ordered >> gpurun_out/join_exp.log 2>&1; echo "exit $?" >> gpurun_out/join_exp.log
done
cut -c1-330 gpurun_out/join_exp.log
