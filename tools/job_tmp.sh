timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_14.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_14.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/r2_tests_14.log | head -30
timeout 300 python -m pytest tests/test_gpu_lti.py -q -s 2>&1 | grep -E "^lti |^   |^nu=|passed|failed" > gpurun_out/r2_tests_lti2.log; cat gpurun_out/r2_tests_lti2.log
timeout 600 python tools/bench_configs.py lti8 lti16 lti32 > gpurun_out/r2_cfg5_lti4.log 2>&1; tail -12 gpurun_out/r2_cfg5_lti4.log
TACD_LTI_KERNEL=0 timeout 600 python tools/bench_configs.py lti8 lti16 lti32 > gpurun_out/r2_cfg5_generic2.log 2>&1; tail -12 gpurun_out/r2_cfg5_generic2.log
