# scratch job for `gpurun -- 'bash tools/job_tmp.sh'` (developer tool): the full GPU suite, then the contract bench line
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; grep -E "passed|failed" gpurun_out/tests.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/tests.log | head -20
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json
