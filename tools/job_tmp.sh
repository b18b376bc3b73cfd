run() { python bench.py --no-extra --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); t=l['timing']; print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM iters %.3f'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['mean_iters']))"; }
echo "== base"; run; run
echo "== noref"; TACD_LIB=$PWD/tacdmpc_b200/variants/noref.so run; TACD_LIB=$PWD/tacdmpc_b200/variants/noref.so run
