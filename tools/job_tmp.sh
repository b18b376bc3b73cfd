python bench.py > gpurun_out/r2_bench_5.json 2> gpurun_out/r2_bench_5.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_bench_5.json").read().strip().splitlines()[-1]); t=l['timing']
print('value %.2fM fwd %.4f bwd %.4f e2e %.2fM frac %.3f launches %s'%(l['value']/1e6,t['fwd_ms'],t['bwd_ms'],l['e2e']['value']/1e6,l['roofline']['frac'], l['gpu_launches']))
print(json.dumps(l['roofline'])[:900]); print(l['cpu_baseline']); print(l['clocks'])
for e in l.get("extra_configs") or []: print("   ", e["config"], '%.3gM fwd %.3f bwd %.3f iters %.2f'%(e["solves_per_s"]/1e6,e["fwd_ms"],e["bwd_ms"],e["mean_iters"]))
PY
