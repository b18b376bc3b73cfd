run() { python bench.py --no-extra --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); t=l['timing']; print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM iters %.3f'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['mean_iters']))"; }
export TACD_HYBRID_CAP=0
echo "== default (MAXW 24)"; run
for w in 12 16 20; do echo "== variant w$w"; TACD_LIB=$PWD/tacdmpc_b200/variants/w$w.so run; done
for lay in per_element diag; do echo "== layout $lay (default lib)"; run --layout $lay; echo "== layout $lay (w16)"; TACD_LIB=$PWD/tacdmpc_b200/variants/w16.so run --layout $lay; done
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests_8.log 2>&1; grep -E "passed|failed" gpurun_out/r2_tests_8.log | tail -2; grep -E "^FAILED" gpurun_out/r2_tests_8.log | sed "s/ - .*//" | head -20
