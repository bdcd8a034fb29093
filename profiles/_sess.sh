mkdir -p gpurun_out; rm -f gpurun_out/join_exp2.log
for v in variants/fj_*.so; do
  BRG_LIB=$v timeout 300 python profiles/exp_join.py --no-chunked --sorted-only --prof >> gpurun_out/join_exp2.log 2>&1; echo "exit $?" >> gpurun_out/join_exp2.log
done
cat gpurun_out/join_exp2.log
