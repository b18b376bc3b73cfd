timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests_34.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_34.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/r2_tests_34.log | head -20
echo "== new"; timeout 300 python tools/bench_configs.py ppo512 ppo4096 2>&1 | grep "B=" | cut -c1-112
echo "== prev"; TACD_LIB=$PWD/tacdmpc_b200/variants/prev.so timeout 300 python tools/bench_configs.py ppo512 ppo4096 2>&1 | grep "B=" | cut -c1-112
echo "== r1 tree"; (cd gpurun_scratch_r1 && timeout 300 python tools/bench_configs.py ppo512 ppo4096 2>&1 | grep "B=" | cut -c1-112)
