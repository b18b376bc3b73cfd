# scratch job for `gpurun -- 'bash tools/job_tmp.sh'` (developer tool): the full GPU suite, smoke, then the contract bench line
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; grep -E "passed|failed" gpurun_out/tests.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/tests.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; head -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; l=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); t=l['timing']
print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM frac %.3f launches %s clocks %s'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['frac'],l['gpu_launches'],l['clocks']))"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
