TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --no-extra --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2>/dev/null
for n in 2 4 8; do timeout 300 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --no-extra --no-cpu-baseline > gpurun_out/r2_scale_n$n.json 2>/dev/null; done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    l=json.loads(open(f"gpurun_out/r2_scale_n{n}.json").read().strip().splitlines()[-1]); t=l["timing"]
    if n==1: base=(l["value"], l["e2e"]["value"])
    print(n, "value %.1fM (%.3f) e2e %.1fM (%.3f) fwd %.4f bwd %.4f"%(l["value"]/1e6, l["value"]/base[0]/n, l["e2e"]["value"]/1e6, l["e2e"]["value"]/base[1]/n, t["fwd_ms"], t["bwd_ms"]))
PY
