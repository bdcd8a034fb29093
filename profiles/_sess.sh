BRG_BUILD_TRACE=1 timeout 300 python profiles/exp_build_once.py count 2>&1 | tail -40
