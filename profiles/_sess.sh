for lib in "" variants/rs_nopf.so; do echo "lib=$lib"; BRG_LIB=$lib timeout 300 python profiles/exp_build_time.py 2>&1 | tail -4; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_merge.py -x -q -m gpu 2>&1 | tail -2
