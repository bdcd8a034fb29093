mkdir -p gpurun_out; rm -f gpurun_out/join_exp2.log
for v in variants/fj_*.so; do
  BRG_LIB=$v timeout 300 python profiles/exp_join.py --no-chunked --sorted-only --no-ordered >> gpurun_out/join_exp2.log 2>&1; echo "exit $?" >> gpurun_out/join_exp2.log
done
cut -c1-200 gpurun_out/join_exp2.log
