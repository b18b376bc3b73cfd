"""Print the key numbers of one bench.py JSON line read from stdin (developer convenience)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    b = r.get("backward") or {}
    print(tag, f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.4f}  fwd {r.get('launch_ms', 0):.4f} ms  "
               f"bwd {b.get('launch_ms', 0):.4f} ms  e2e {d.get('e2e', {}).get('value', 0):.4g}  "
               f"fp32 {100 * (r.get('fp32_pipe') or {}).get('frac_of_nominal', 0):.2f}%  hbm {100 * r.get('frac', 0):.2f}%  "
               f"bwd hbm {100 * b.get('frac', 0):.2f}%  clk {d.get('clocks', {}).get('sm_mhz')}")
