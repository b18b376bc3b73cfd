timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_12.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_12.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/r2_tests_12.log | head -20
timeout 600 python tools/bench_configs.py lti8 lti16 lti32 > gpurun_out/r2_cfg5_lti.log 2>&1; tail -12 gpurun_out/r2_cfg5_lti.log
TACD_LTI_KERNEL=0 timeout 600 python tools/bench_configs.py lti8 lti16 lti32 > gpurun_out/r2_cfg5_generic.log 2>&1; tail -12 gpurun_out/r2_cfg5_generic.log
