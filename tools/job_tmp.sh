timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_26.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_26.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/r2_tests_26.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_2.log 2>&1; tail -24 gpurun_out/r2_smoke_2.log
