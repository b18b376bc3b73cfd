python bench.py > gpurun_out/r2_bench_4.json 2> gpurun_out/r2_bench_4.err; tail -c 300 gpurun_out/r2_bench_4.err
for lay in per_element diag; do python bench.py --no-extra --no-cpu-baseline --layout $lay > gpurun_out/r2_bench_4_$lay.json 2>/dev/null; done
python - <<'PY'
import json
for f in ("r2_bench_4","r2_bench_4_per_element","r2_bench_4_diag"):
    l=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); t=l['timing']
    print(f, 'value %.2fM fwd %.4f bwd %.4f e2e %.2fM frac %.3f'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['frac']))
    for e in l.get("extra_configs") or []: print("   ", e["config"], '%.3gM fwd %.3f bwd %.3f iters %.2f'%(e["solves_per_s"]/1e6,e["fwd_ms"],e["bwd_ms"],e["mean_iters"]))
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
