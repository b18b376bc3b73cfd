run() { python bench.py --no-extra --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); t=l['timing']; print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM iters %.3f launches %s'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['mean_iters'],l['gpu_launches']))"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_33.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_33.log | tail -2; grep -E "^FAILED|^E  " gpurun_out/r2_tests_33.log | head -20
echo "== new shared"; run
echo "== new"; timeout 300 python tools/bench_configs.py sweep unicycle unicycle_tight 2>&1 | grep "B="
