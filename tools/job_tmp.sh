run() { python bench.py --no-extra --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); t=l['timing']; print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM iters %.3f'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['mean_iters']))"; }
echo "== v9"; run; echo "== v9 w18"; TACD_LIB=$PWD/tacdmpc_b200/variants/w18.so run; echo "== v8"; TACD_LIB=$PWD/tacdmpc_b200/variants/v8.so run; echo "== v9"; run; echo "== v9 w18"; TACD_LIB=$PWD/tacdmpc_b200/variants/w18.so run
export TACD_BENCH_NO_GRAPH=1
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:ilqr_forward_group -s 3 -c 1 -f -o gpurun_out/r2g_fwd_group_v9 $B > gpurun_out/prof_ncu_v9.log 2>&1; tail -2 gpurun_out/prof_ncu_v9.log | head -1 | cut -c1-100
